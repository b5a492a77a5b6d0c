#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.h"

namespace metro {
// img: float32 (or uint8 when u8) NHWC [n,in,in,3]; w: device float [7*7*3][64] (fp16-rounded values);
// out: fp16 NHWC [n,out,out,64].
metro_status root_conv_launch(const void *img, bool u8, const float *w, const float *bias, __half *out, int n,
                              int in_side, int out_side, cudaStream_t stream);
// in: fp16 NHWC [n,in,in,c] -> raw (may be null) and pre = relu(scale*raw+shift), both [n,out,out,c].
metro_status pool_preact_launch(const __half *in, __half *raw, __half *pre, const float *scale, const float *shift,
                                int n, int in_side, int out_side, int c, cudaStream_t stream);
}  // namespace metro
