// Root of the network: conv1 7x7/2 + bias (no BN / activation) and pool1 3x3/2 with ZERO padding,
// fused with the first unit's pre-activation BN+ReLU.
//   resnet_v2.py:219-224, resnet_utils.py:124-135 (explicit pad (3,3) then VALID),
//   resnet_utils.py:177-185 (pad with zeros, not -inf: border maxima are clamped at >= 0).
// conv1 has Cin = 3: the image is re-packed (space-to-depth 2x2, fp16) so that the 7x7/2 conv becomes
// a 4x4 stride-1 conv with 16-channel pixels that the tcgen05 implicit-GEMM kernel (conv_gemm.cu)
// consumes with K = 4 row taps x 64; see metro_api.cu.
#include <cuda_fp16.h>

#include "common.h"
#include "root_pool.h"

namespace metro {

namespace {

// Space-to-depth pack: fp32 / uint8 NHWC [n,256,256,3] -> fp16 [n, hp, wp, win*16].
// thread = one 16-channel group (32 bytes): the 2x2 input pixels of s2d pixel (h2, w2) as
// [p=0: (q0: r,g,b) (q1: r,g,b)] [p=1: ...] + 4 zeros; values are rounded to fp16 exactly like the
// reference's cast to FLAGS.dtype (architectures.py:29); the uint8 variant fuses the /255 of
// improc.py:56-61.  win = 4 additionally replicates each pixel into its 4 sliding-window slots.
template <bool U8>
__global__ void __launch_bounds__(256) s2d_pack_kernel(const void *__restrict__ img, __half *__restrict__ out, int n,
                                                       int in_side, int hp, int wp, int win) {
  const size_t total = size_t(n) * hp * wp * win;
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int slot = int(i % win);
  size_t r = i / win;
  const int w = int(r % wp); r /= wp;
  const int h = int(r % hp);
  const int b = int(r / hp);
  const int h2 = h - 2, w2 = w + slot - 2;             // s2d coordinates of the source pixel
  float v[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) v[k] = 0.f;
  const int half = in_side / 2;
  if (h2 >= 0 && h2 < half && w2 >= 0 && w2 < half) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const size_t base = ((size_t(b) * in_side + (2 * h2 + p)) * in_side + 2 * w2) * 3;
      if (U8) {
        const unsigned char *s = static_cast<const unsigned char *>(img) + base;
#pragma unroll
        for (int k = 0; k < 6; ++k) v[p * 6 + k] = float(s[k]) * (1.0f / 255.0f);
      } else {
        const float2 *s = reinterpret_cast<const float2 *>(static_cast<const float *>(img) + base);
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float2 t = s[k]; v[p * 6 + 2 * k] = t.x; v[p * 6 + 2 * k + 1] = t.y; }
      }
    }
  }
  uint4 o0, o1;
  __half2 *a = reinterpret_cast<__half2 *>(&o0), *c = reinterpret_cast<__half2 *>(&o1);
#pragma unroll
  for (int k = 0; k < 4; ++k) a[k] = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
  c[0] = __floats2half2_rn(v[8], v[9]);
  c[1] = __floats2half2_rn(v[10], v[11]);
  c[2] = __floats2half2_rn(0.f, 0.f);
  c[3] = c[2];
  uint4 *dst = reinterpret_cast<uint4 *>(out + i * 16);
  dst[0] = o0;
  dst[1] = o1;
}

// pool1 + first pre-activation.  thread = 8 channels of one output pixel.
__global__ void __launch_bounds__(256) pool_preact_kernel(const __half *__restrict__ in, __half *__restrict__ raw,
                                                          __half *__restrict__ pre, const float *__restrict__ scale,
                                                          const float *__restrict__ shift, int n, int in_side,
                                                          int out_side, int c) {
  const int c8 = c / 8;
  const size_t total = size_t(n) * out_side * out_side * c8;
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cg = int(i % c8);
  size_t r = i / c8;
  const int ow = int(r % out_side); r /= out_side;
  const int oh = int(r % out_side);
  const int b = int(r / out_side);
  __half2 m[4];
  bool first = true;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int ih = oh * 2 + kh - 1, iw = ow * 2 + kw - 1;
      uint4 v = make_uint4(0, 0, 0, 0);                      // zero padding takes part in the max (Q6)
      if (ih >= 0 && ih < in_side && iw >= 0 && iw < in_side)
        v = *reinterpret_cast<const uint4 *>(in + ((size_t(b) * in_side + ih) * in_side + iw) * c + cg * 8);
      const __half2 *vh = reinterpret_cast<const __half2 *>(&v);
      if (first) { for (int k = 0; k < 4; ++k) m[k] = vh[k]; first = false; }
      else { for (int k = 0; k < 4; ++k) m[k] = __hmax2(m[k], vh[k]); }
    }
  const size_t off = ((size_t(b) * out_side + oh) * out_side + ow) * c + cg * 8;
  uint4 o;
  __half2 *oh2 = reinterpret_cast<__half2 *>(&o);
  for (int k = 0; k < 4; ++k) oh2[k] = m[k];
  if (raw) *reinterpret_cast<uint4 *>(raw + off) = o;
  uint4 o2;
  __half2 *p2 = reinterpret_cast<__half2 *>(&o2);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 y = __half22float2(m[k]);
    const int ch = cg * 8 + 2 * k;
    p2[k] = __floats2half2_rn(fmaxf(fmaf(y.x, scale[ch], shift[ch]), 0.f),
                              fmaxf(fmaf(y.y, scale[ch + 1], shift[ch + 1]), 0.f));
  }
  *reinterpret_cast<uint4 *>(pre + off) = o2;
}

}  // namespace

metro_status s2d_pack_launch(const void *img, bool u8, __half *out, int n, int in_side, int hp, int wp, int win,
                             cudaStream_t stream) {
  if (n == 0) return METRO_OK;
  const size_t total = size_t(n) * hp * wp * win;
  const unsigned blocks = unsigned((total + 255) / 256);
  if (u8) s2d_pack_kernel<true><<<blocks, 256, 0, stream>>>(img, out, n, in_side, hp, wp, win);
  else s2d_pack_kernel<false><<<blocks, 256, 0, stream>>>(img, out, n, in_side, hp, wp, win);
  METRO_CUDA(cudaGetLastError());
  return METRO_OK;
}

metro_status pool_preact_launch(const __half *in, __half *raw, __half *pre, const float *scale, const float *shift,
                                int n, int in_side, int out_side, int c, cudaStream_t stream) {
  if (n == 0) return METRO_OK;
  const size_t total = size_t(n) * out_side * out_side * (c / 8);
  const unsigned blocks = unsigned((total + 255) / 256);
  pool_preact_kernel<<<blocks, 256, 0, stream>>>(in, raw, pre, scale, shift, n, in_side, out_side, c);
  METRO_CUDA(cudaGetLastError());
  return METRO_OK;
}

}  // namespace metro
