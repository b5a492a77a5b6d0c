// Post-path transform of the reference's evaluation graph:
//   volumetric.py:277-282  to_orig_cam: x' = R x per joint, left/right joints swapped when det(R) <= 0
//   volumetric.py:221-222  matmul_joint_coords = einsum('Bij,BCj->BCi')
//   datasets.py:76-79      JointInfo.mirror_mapping
// [n, J, 3] float32 in and out; one thread per (crop, joint).  The work is a few kilobytes: the kernel exists so
// that the skeletons can stay on the device between the decode and whatever consumes them, not for speed.
#include "common.h"

namespace metro {

namespace {

struct ToOrigCamParams {
  const float *poses;
  const float *rot;
  float *out;
  int n, j;
  int mirror[kMaxJointsOut];
};

__global__ void to_orig_cam_kernel(const ToOrigCamParams p) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.n * p.j) return;
  const int b = t / p.j, c = t - b * p.j;
  float r[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) r[i] = p.rot[size_t(b) * 9 + i];
  const float det = r[0] * (r[4] * r[8] - r[5] * r[7]) - r[1] * (r[3] * r[8] - r[5] * r[6]) + r[2] * (r[3] * r[7] - r[4] * r[6]);
  const int src = det > 0.f ? c : p.mirror[c];     // tf.where(det > 0, x, gather(x, mirror_mapping))
  const float *x = p.poses + (size_t(b) * p.j + src) * 3;
  const float x0 = x[0], x1 = x[1], x2 = x[2];
  float *y = p.out + size_t(t) * 3;
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] = fmaf(r[3 * i + 2], x2, fmaf(r[3 * i + 1], x1, r[3 * i] * x0));
}

}  // namespace

metro_status to_orig_cam_launch(const float *poses, const float *rot, const int32_t *mirror, int n, int j, float *out,
                                cudaStream_t stream) {
  ToOrigCamParams p{};
  p.poses = poses; p.rot = rot; p.out = out; p.n = n; p.j = j;
  for (int i = 0; i < j; ++i) p.mirror[i] = mirror[i];
  const int threads = 128, blocks = (n * j + threads - 1) / threads;
  to_orig_cam_kernel<<<blocks, threads, 0, stream>>>(p);
  METRO_CUDA(cudaGetLastError());
  return METRO_OK;
}

}  // namespace metro
