#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "common.h"

namespace metro {
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
                               int num_sms, cudaStream_t stream, long long *prof = nullptr,
                               unsigned int *sig_flags = nullptr);
// ---- version 2: the image pack folded into the kernel (loader warps convert the caller's float32 / uint8 rows into a
// shared-memory ring), two conv rows per N = 128 accumulator.  `images` points at crop 0 of this call's input. ----
size_t root2_packed_weight_elems();
void root2_pack_weights(const float *w_hwio, __half *dst);
metro_status root_fused2_launch(const void *images, bool u8, const __half *wpack2, const float *bias, const float *pscale,
                                const float *pshift, __half *raw, __half *pre, __half *conv_dbg, int n, int n_base, int num_sms,
                                cudaStream_t stream, unsigned long long *tstamp = nullptr);
constexpr unsigned int kRootBandsPerCrop = 8;   // what a crop's counter reaches when sig_flags is given
}  // namespace metro
