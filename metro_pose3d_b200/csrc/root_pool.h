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
// ---- fused root (root_fused.cu): image pack -> conv1 7x7/2 + bias -> zero-padded pool1 -> first pre-activation ----
size_t root_packed_image_elems();            // fp16 elements per crop of the packed image
size_t root_packed_weight_elems();
void root_pack_weights(const float *w_hwio, __half *dst);
// map_out: 128 bytes, 64-byte aligned (a CUtensorMap) over `n` crops of the packed image
metro_status root_make_image_map(void *map_out, const __half *packed, int n);
metro_status img_pack_launch(const void *img, bool u8, __half *out, int n, cudaStream_t stream);
// raw (may be null) / pre: fp16 NHWC [*,64,64,64]; conv_dbg (may be null): conv1 output [*,128,128,64];
// works on crops n_base .. n_base + n of all buffers
metro_status root_fused_launch(const void *image_map, const __half *wpack, const float *bias, const float *pscale,
                               const float *pshift, __half *raw, __half *pre, __half *conv_dbg, int n, int n_base,
                               int num_sms, cudaStream_t stream);
}  // namespace metro
