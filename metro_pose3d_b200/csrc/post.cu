// Post-path transform of the reference's evaluation graph:
//   volumetric.py:277-282  to_orig_cam: x' = R x per joint, left/right joints swapped when det(R) <= 0
//   volumetric.py:221-222  matmul_joint_coords = einsum('Bij,BCj->BCi')
//   datasets.py:76-79      JointInfo.mirror_mapping
// [n, J, 3] float32 in and out; one thread per (crop, joint).  The work is a few kilobytes: the kernel exists so
// that the skeletons can stay on the device between the decode and whatever consumes them, not for speed.
#include <cuda_fp16.h>

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

// 'true-root-depth' back-projection (volumetric.py:190-198,285): heatmap coordinates -> image pixels
// (heatmap_to_image, :288-295) -> homogeneous -> inverse intrinsics -> rays x (depth relative to the root joint, the LAST
// one, in mm + z offset).  One thread per (crop, joint); float64 inside (a dozen operations), float32 in and out.
__global__ void back_project_kernel(const float *coords01, const float *inv_k, const float *z_off, float *out, int n, int j,
                                    double lrc, double add_xy, double box) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * j) return;
  const int b = t / j;
  const float *c = coords01 + size_t(t) * 3, *root = coords01 + (size_t(b) * j + (j - 1)) * 3, *k = inv_k + size_t(b) * 9;
  const double u = double(c[0]) * lrc + add_xy, v = double(c[1]) * lrc + add_xy;
  const double depth = (double(c[2]) - double(root[2])) * box + double(z_off[b]);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    out[size_t(t) * 3 + i] = float((double(k[3 * i]) * u + double(k[3 * i + 1]) * v + double(k[3 * i + 2])) * depth);
}

// t.heatmap_pred_z (volumetric.py:165): softmax over (H, W, D) of one joint, summed over H and W -> [D] per (crop, joint).
// One block per (crop, joint); not on the hot path (the evaluation graph's extra fetch), so a plain two-pass reduction.
template <typename T>
__global__ void __launch_bounds__(256) heatmap_z_kernel(const T *head, float *out, int side, int J, int D) {
  constexpr int kMaxD = 16;
  __shared__ double red[256];
  __shared__ double acc[kMaxD];
  const int j = blockIdx.x % J, img = blockIdx.x / J, tid = threadIdx.x;
  const int P = side * side, C = D * J;
  const T *base = head + size_t(img) * P * C;
  double mx = -1e300;
  for (int i = tid; i < P * D; i += 256) mx = fmax(mx, double(float(base[size_t(i / D) * C + (i % D) * J + j])));
  red[tid] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[tid] = fmax(red[tid], red[tid + s]);
    __syncthreads();
  }
  mx = red[0];
  __syncthreads();
  double part[kMaxD];
  for (int d = 0; d < kMaxD; ++d) part[d] = 0.0;
  // thread t owns pixels t, t + 256, ...: all depths of a pixel, so no cross-depth traffic
  for (int px = tid; px < P; px += 256)
    for (int d = 0; d < D; ++d) part[d] += exp(double(float(base[size_t(px) * C + d * J + j])) - mx);
  double total = 0.0;
  for (int d = 0; d < D; ++d) {
    red[tid] = part[d];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (tid < s) red[tid] += red[tid + s];
      __syncthreads();
    }
    if (tid == 0) acc[d] = red[0];
    total += red[0];
    __syncthreads();
  }
  if (tid < D) out[(size_t(img) * J + j) * D + tid] = float(acc[tid] / total);
}

}  // namespace

metro_status back_project_launch(const float *coords01, const float *inv_k, const float *z_off, int n, int j, double lrc,
                                 double add_xy, double box, float *out, cudaStream_t stream) {
  const int threads = 128, blocks = (n * j + threads - 1) / threads;
  back_project_kernel<<<blocks, threads, 0, stream>>>(coords01, inv_k, z_off, out, n, j, lrc, add_xy, box);
  METRO_CUDA(cudaGetLastError());
  return METRO_OK;
}

metro_status heatmap_z_launch(const void *head, bool f16, int n, int side, int j, int depth, float *out, cudaStream_t stream) {
  if (depth > 16) return fail(METRO_ERR_VALUE, "heatmap_z: depth %d > 16", depth);
  if (f16) heatmap_z_kernel<__half><<<unsigned(n * j), 256, 0, stream>>>(static_cast<const __half *>(head), out, side, j, depth);
  else heatmap_z_kernel<float><<<unsigned(n * j), 256, 0, stream>>>(static_cast<const float *>(head), out, side, j, depth);
  METRO_CUDA(cudaGetLastError());
  return METRO_OK;
}

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
