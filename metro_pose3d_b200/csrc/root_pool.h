#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.h"

namespace metro {
// img: float32 (or uint8 when u8) NHWC [n,in,in,3] -> fp16 [n,hp,wp,win*16] space-to-depth (zero border
// 2 before / 1 after).
metro_status s2d_pack_launch(const void *img, bool u8, __half *out, int n, int in_side, int hp, int wp, int win,
                             cudaStream_t stream);
// in: fp16 NHWC [n,in,in,c] -> raw (may be null) and pre = relu(scale*raw+shift), both [n,out,out,c].
metro_status pool_preact_launch(const __half *in, __half *raw, __half *pre, const float *scale, const float *shift,
                                int n, int in_side, int out_side, int c, cudaStream_t stream);
}  // namespace metro
