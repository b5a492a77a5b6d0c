// Fused implicit-GEMM convolution on tcgen05 tensor cores (sm_100a): host-side description.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <string>

#include "common.h"

namespace metro {

constexpr int kTileM = 128;     // output pixels per tile (UMMA M)
constexpr int kTileK = 64;      // fp16 channels per K block = one 128-byte swizzle row
constexpr int kMaxTaps = 9;
constexpr int kMaxStages = 8;
constexpr int kWarpStageBytes = 32 * 64;    // one epilogue warp's staging box: 32 rows x 32 fp16 columns (2 KB)
constexpr int kEpilogueWarps = 8;

// Kernel parameters (passed by value as a __grid_constant__; tensor maps must be 64-byte aligned).
struct alignas(64) ConvGemmParams {
  CUtensorMap amap[4];   // source 0, NHWC fp16, viewed as [C, W', H', N]; 4 = (row,col) parity views
                         // of a stride-2 conv (entry 0 only for stride 1)
  CUtensorMap a2map;     // optional source 1 (1x1 on the output grid): the projection shortcut's input, or
                         // the identity shortcut's raw tensor (sub-sampled / shifted view) with diag2 = 1
  CUtensorMap bmap;      // packed weights [cout_pad][K_total] fp16, K-major
  CUtensorMap bidmap;    // the same matrix with 32-row boxes: one CTA's half of a 64x64 identity block (diag2)
  CUtensorMap o1map;     // output 1  [M][cout] fp16 (TMA store), unused on the direct (fp32) path
  CUtensorMap o2map;     // output 2  [M][cout] fp16
  // K loop: taps x cblk0 blocks from source 0, then cblk1 blocks from source 1
  int taps, cblk0, cblk1, diag2;
  signed char tap_map[12], tap_dh[12], tap_dw[12];
  // M tiling: a tile is 128 consecutive output pixels = th full rows of nb images
  int m_total, wo, ho, th, nb, tiles_per_img, m_tiles, n_tiles, cout;
  int n_base, m_base;    // first crop / first output row of the batch slice this launch works on
  int reverse;           // walk the tile list backwards: consecutive layers alternate, so a layer starts on the
                         // rows its producer wrote last (still resident in L2)
  // "tall" staging of 3x3 stride-1 convolutions: amap[1] is the (th + 2*rate)-row box, a stage = that box +
  // the three weight boxes of one kernel column; tall_row_step = rate * W * 128 bytes between kernel rows
  int tall, tall_a_bytes, tall_stage_bytes, tall_row_step;
  // epilogue: y = acc*scale + shift (+ res) ; relu? ; store y (fp16|fp32) ;
  //           y2 = relu(fp16(y)*scale2 + shift2) -> fp16 (the consumer's pre-activation)
  const float *scale, *shift, *scale2, *shift2;
  // optional A-operand transform (1x1 convs): x -> relu(x * ascale[c] + ashift[c]) applied in shared memory
  const float *ascale, *ashift;
  void *out1;            // direct-store path only (logits head)
  int out1_f32;
  int has_out1, has_out2, relu1;
  // shared-memory plan (bytes from the 1024-aligned base)
  int stages;
  int off_stage, off_par, off_apar, off_bar, smem_bytes;
  long long *prof;       // optional [grid][8] per-CTA role timers (cycles); nullptr = off
  // cross-kernel dataflow (ptx.cuh): instead of waiting for the whole previous grid (griddepcontrol.wait) a tile
  // waits until the PRODUCER layer has completed the crops it reads, and every epilogue warp reports its 32 rows
  // of a crop once their stores are complete.  Counters are indexed by absolute crop; null = off.
  const unsigned int *dep_flags;   // producer layer's counters
  unsigned int dep_expected;       // pieces per crop the producer reports: (its output pixels per crop / 32) x its N tiles
  unsigned int *sig_flags;         // this layer's counters
  // whole-layer shortcut: every CTA adds 1 to its layer's word on exit; once a consumer has seen the producer's word
  // reach the producer's grid size it stops polling per-crop counters (the steady state of a launch)
  const unsigned int *dep_done;
  unsigned int dep_ctas;
  unsigned int *sig_done;
  // conv_chain.cu (conv3 of a unit + conv1 of the next in one kernel): the second convolution's weights [cout1][cout]
  // K-major, its folded BN, and the shared-memory offset of the pre-activation operand buffer; o2map = its output
  CUtensorMap w1map;
  const float *scale1c, *shift1c;
  int cout1, off_a2;
  // optional device time stamps of THIS launch (metro_profile): [0] = earliest CTA start, [1] = latest CTA end, in
  // %globaltimer nanoseconds -- the kernel's duration inside the real, overlapped pipeline (no events between launches)
  unsigned long long *tstamp;
};

struct ConvGemmLaunch {
  ConvGemmParams prm;
  int block_n = 128;       // 64 | 128 | 160 | 256
  bool direct = false;     // direct-store epilogue (logits head: cout not a multiple of 64)
  std::string name;
  double flops_per_img = 0;
  unsigned int sig_expected = 0;   // what this layer's counters reach per crop: (ho * wo / 32) x n_tiles
  bool signals = false;            // long tiles (a whole crop or more per CTA tile): worth a report per tile
  bool chain = false;              // conv_chain.cu kernel (prm.cout1 > 0)
};

// Tensor-map helpers (driver entry point resolved at run time; no link-time libcuda dependency).
metro_status make_act_tensor_map(CUtensorMap *map, const void *base, int n, int h, int w, int c,
                                 int sub, int ph, int pw, int box_w, int box_h, int box_n);
metro_status make_weight_tensor_map(CUtensorMap *map, const void *base, int cout_pad, int k_total, int block_n);
metro_status make_out_tensor_map(CUtensorMap *map, const void *base, long long m_rows, int cout);

int conv_gemm_pick_block_n(int cout, bool direct, long long m_rows);
int conv_gemm_cout_pad(int cout, int block_n);
// Lays out shared memory (stage count, staging buffers) once block_n and the has_* flags are set.
metro_status conv_gemm_plan_smem(ConvGemmLaunch &L, int k_blocks);
// Fills the M-tiling fields for `n` images (crops n_base .. n_base + n of the buffers) of an out_side x out_side output.
metro_status conv_gemm_set_batch(ConvGemmParams &p, int n, int n_base = 0);
metro_status conv_gemm_geometry(ConvGemmParams &p, int out_side);
// `prm` = L.prm with this call's batch slice / direction / profiling pointer filled in (L itself is never mutated
// after it is built, so launches are capture-safe).
metro_status conv_gemm_launch(const ConvGemmLaunch &L, const ConvGemmParams &prm, int num_sms, cudaStream_t stream);
// CTAs the launch of `prm` runs (what its sig_done word reaches)
int conv_gemm_grid(const ConvGemmParams &prm, int num_sms);

// ---- conv_chain.cu: conv3 (+ shortcut) of unit u and conv1 of unit u + 1 in one kernel ----
metro_status conv_chain_plan_smem(ConvGemmParams &p);
metro_status conv_chain_launch(const ConvGemmParams &prm, int num_sms, cudaStream_t stream);
int conv_chain_grid(const ConvGemmParams &prm, int num_sms);
// Packs HWIO float32 filters into [cout_pad][K] fp16 in the kernel's K-block order; `w2` (1x1,
// [cin2][cout]) is appended along K.
void conv_gemm_pack_weights(const float *w_hwio, int k, int cin, int cout, const float *w2, int cin2,
                            int cout_pad, __half *dst);
// Tap table for a k x k conv with the given stride / rate / leading pad.
metro_status conv_gemm_set_taps(ConvGemmParams &p, int k, int stride, int rate, int pad_lo);

}  // namespace metro
