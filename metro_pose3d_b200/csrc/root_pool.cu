// Root of the network: conv1 7x7/2 + bias (no BN / activation) and pool1 3x3/2 with ZERO padding,
// fused with the first unit's pre-activation BN+ReLU.
//   resnet_v2.py:219-224, resnet_utils.py:124-135 (explicit pad (3,3) then VALID),
//   resnet_utils.py:177-185 (pad with zeros, not -inf: border maxima are clamped at >= 0).
// conv1 has Cin = 3 (K = 147): it is computed on the CUDA cores from an fp32 / uint8 NHWC image
// whose values are rounded to fp16 first (the reference casts the input to FLAGS.dtype,
// architectures.py:29), fp16-rounded filters, fp32 accumulation, fp16 NHWC output.
#include <cuda_fp16.h>

#include "common.h"
#include "root_pool.h"

namespace metro {

namespace {

constexpr int kRootTileH = 8, kRootTileW = 32;          // output pixels per CTA (256 threads)
constexpr int kPatchH = (kRootTileH - 1) * 2 + 7;       // 21
constexpr int kPatchW = (kRootTileW - 1) * 2 + 7;       // 69
constexpr int kPatchWPad = kPatchW * 3 + 1;             // floats per patch row (+1: bank spread)
constexpr int kRootSmem = (147 * 64 + kPatchH * kPatchWPad) * 4;

template <bool U8>
__global__ void __launch_bounds__(256) root_conv_kernel(const void *__restrict__ img, const float *__restrict__ w,
                                                        const float *__restrict__ bias, __half *__restrict__ out,
                                                        int in_side, int out_side) {
  extern __shared__ float sm[];
  float *s_w = sm;                       // [147][64]  (kh,kw,ci) x cout, already fp16-rounded values
  float *s_x = sm + 147 * 64;            // [21][69*3 + 1]
  const int tid = threadIdx.x;
  const int n = blockIdx.z;
  const int oh0 = blockIdx.y * kRootTileH, ow0 = blockIdx.x * kRootTileW;
  for (int i = tid; i < 147 * 64; i += 256) s_w[i] = w[i];
  const int ih0 = oh0 * 2 - 3, iw0 = ow0 * 2 - 3;
  for (int i = tid; i < kPatchH * kPatchW * 3; i += 256) {
    const int r = i / (kPatchW * 3), cc = i - r * (kPatchW * 3);
    const int ih = ih0 + r, iw = iw0 + cc / 3, ch = cc % 3;
    float v = 0.f;
    if (ih >= 0 && ih < in_side && iw >= 0 && iw < in_side) {
      const size_t idx = ((size_t(n) * in_side + ih) * in_side + iw) * 3 + ch;
      if (U8) v = float(static_cast<const unsigned char *>(img)[idx]) * (1.0f / 255.0f);
      else v = static_cast<const float *>(img)[idx];
      v = __half2float(__float2half_rn(v));
    }
    s_x[r * kPatchWPad + cc] = v;
  }
  __syncthreads();
  const int ty = tid / kRootTileW, tx = tid % kRootTileW;
  float acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = bias[i];
#pragma unroll 1
  for (int kh = 0; kh < 7; ++kh) {
    const float *xrow = s_x + (ty * 2 + kh) * kPatchWPad + tx * 6;
#pragma unroll 1
    for (int t = 0; t < 21; ++t) {       // (kw, ci) flattened: contiguous in the patch row
      const float x = xrow[t];
      const float4 *wr = reinterpret_cast<const float4 *>(s_w + (kh * 21 + t) * 64);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 wv = wr[j];
        acc[4 * j] = fmaf(x, wv.x, acc[4 * j]);
        acc[4 * j + 1] = fmaf(x, wv.y, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(x, wv.z, acc[4 * j + 2]);
        acc[4 * j + 3] = fmaf(x, wv.w, acc[4 * j + 3]);
      }
    }
  }
  const int oh = oh0 + ty, ow = ow0 + tx;
  if (oh < out_side && ow < out_side) {
    uint4 *dst = reinterpret_cast<uint4 *>(out + ((size_t(n) * out_side + oh) * out_side + ow) * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint4 o;
      __half2 *oh2 = reinterpret_cast<__half2 *>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) oh2[i] = __floats2half2_rn(acc[8 * j + 2 * i], acc[8 * j + 2 * i + 1]);
      dst[j] = o;
    }
  }
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

metro_status root_conv_launch(const void *img, bool u8, const float *w, const float *bias, __half *out, int n,
                              int in_side, int out_side, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    METRO_CUDA(cudaFuncSetAttribute(root_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRootSmem));
    METRO_CUDA(cudaFuncSetAttribute(root_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRootSmem));
    configured = true;
  }
  if (n == 0) return METRO_OK;
  const dim3 grid((out_side + kRootTileW - 1) / kRootTileW, (out_side + kRootTileH - 1) / kRootTileH, n);
  if (u8) root_conv_kernel<true><<<grid, 256, kRootSmem, stream>>>(img, w, bias, out, in_side, out_side);
  else root_conv_kernel<false><<<grid, 256, kRootSmem, stream>>>(img, w, bias, out, in_side, out_side);
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
