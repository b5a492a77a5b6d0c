// Crop extraction on the GPU: the step that precedes the hot path in the reference's loader,
// cameralib.reproject_image_fast (src/cameralib.py:406-429; call site src/data/data_loading.py:93) --
//   coords = homography @ (x, y, 1);  map = coords[:2] / coords[2];  cv2.remap(image, map, INTER_LINEAR, BORDER_CONSTANT)
// for uint8 RGB frames -> uint8 [n, side, side, 3] crops that feed metro_infer_u8 without a host round trip.
//
// Arithmetic is OpenCV's (imgproc/src/imgwarp.cpp, remapBilinear<FixedPtCast<int, uchar, 15>>), restated and pinned
// bit for bit against cv2.remap in oracle/crop_oracle.py: source coordinates rounded to 1/32 pixel (round half to
// even), 15-bit integer weights (32 - fx)(32 - fy) * 32 ... (the weight of an exact integer coordinate saturates to
// 32767 like OpenCV's int16 table), (sum + 2^14) >> 15, constant border.  The three float32 multiply-adds of the
// homography are evaluated as ((h0 * x + h1 * y) + h2), each operation rounded, no fused multiply-add (the
// reference leaves this order to its BLAS).
// One thread per output pixel; HBM-bound (3 bytes written, <= 12 gathered per pixel), nowhere near the hot path's cost.
#include "common.h"

namespace metro {

namespace {

constexpr int kMaxCropsPerLaunch = 32;    // sources travel in the kernel parameters: no device-side table, no allocation

struct CropSrc {
  const unsigned char *frame;
  int height, width, row_stride;
  float h[9];
};
struct CropParams {
  CropSrc src[kMaxCropsPerLaunch];
  unsigned char *out;
  int side, border;
};

__global__ void __launch_bounds__(256) extract_crops_kernel(const __grid_constant__ CropParams p) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= p.side * p.side) return;
  const CropSrc &s = p.src[blockIdx.y];
  const int oy = px / p.side, ox = px - oy * p.side;
  const float x = float(ox), y = float(oy);
  const float cx = __fadd_rn(__fadd_rn(__fmul_rn(s.h[0], x), __fmul_rn(s.h[1], y)), s.h[2]);
  const float cy = __fadd_rn(__fadd_rn(__fmul_rn(s.h[3], x), __fmul_rn(s.h[4], y)), s.h[5]);
  const float cw = __fadd_rn(__fadd_rn(__fmul_rn(s.h[6], x), __fmul_rn(s.h[7], y)), s.h[8]);
  const float mx = __fdiv_rn(cx, cw), my = __fdiv_rn(cy, cw);
  // cvRound(v * 32): round half to even, saturating like the float -> int32 conversion
  // (x86's conversion, which OpenCV runs on, yields INT_MIN for NaN and for values beyond the int32 range)
  const float tx = __fmul_rn(mx, 32.0f), ty = __fmul_rn(my, 32.0f);
  const int sx = (tx >= -2147483648.0f && tx < 2147483648.0f) ? __float2int_rn(tx) : int(0x80000000);
  const int sy = (ty >= -2147483648.0f && ty < 2147483648.0f) ? __float2int_rn(ty) : int(0x80000000);
  const int fx = sx & 31, fy = sy & 31;
  const int ix = max(-32768, min(32767, sx >> 5)), iy = max(-32768, min(32767, sy >> 5));
  const int w00 = min((32 - fx) * (32 - fy) * 32, 32767), w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  int acc[3] = {0, 0, 0};
  auto tap = [&](int yy, int xx, int w) {
    if (w == 0) return;
    if (yy >= 0 && yy < s.height && xx >= 0 && xx < s.width) {
      const unsigned char *q = s.frame + size_t(yy) * s.row_stride + size_t(xx) * 3;
      acc[0] += int(q[0]) * w; acc[1] += int(q[1]) * w; acc[2] += int(q[2]) * w;
    } else {
      acc[0] += p.border * w; acc[1] += p.border * w; acc[2] += p.border * w;
    }
  };
  tap(iy, ix, w00); tap(iy, ix + 1, w01); tap(iy + 1, ix, w10); tap(iy + 1, ix + 1, w11);
  unsigned char *o = p.out + (size_t(blockIdx.y) * p.side * p.side + px) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = (unsigned char)max(0, min(255, (acc[c] + (1 << 14)) >> 15));
}

}  // namespace

metro_status extract_crops_launch(const metro_crop_src *srcs, int n, int side, int border, unsigned char *out, cudaStream_t stream) {
  for (int lo = 0; lo < n; lo += kMaxCropsPerLaunch) {
    const int cnt = n - lo < kMaxCropsPerLaunch ? n - lo : kMaxCropsPerLaunch;
    CropParams p{};
    for (int i = 0; i < cnt; ++i) {
      const metro_crop_src &s = srcs[lo + i];
      p.src[i].frame = s.frame_dev; p.src[i].height = s.height; p.src[i].width = s.width; p.src[i].row_stride = s.row_stride_bytes;
      for (int k = 0; k < 9; ++k) p.src[i].h[k] = s.homography[k];
    }
    p.out = out + size_t(lo) * side * side * 3; p.side = side; p.border = border;
    dim3 grid(unsigned((side * side + 255) / 256), unsigned(cnt));
    extract_crops_kernel<<<grid, 256, 0, stream>>>(p);
  }
  METRO_CUDA(cudaGetLastError());
  return METRO_OK;
}

}  // namespace metro
