// Root of the network in one kernel: conv1 7x7/2 + bias (resnet_v2.py:219-220, explicit pad (3,3) then
// VALID, resnet_utils.py:124-135), pool1 3x3/2 with ZERO padding (resnet_v2.py:222-224,
// resnet_utils.py:177-185: border maxima are clamped at >= 0) and the first unit's pre-activation
// BN+ReLU (resnet_v2.py:119).  The 128x128x64 conv1 tensor (2 MB per crop) never reaches HBM.
//
// conv1 on the tensor core without an im2col expansion anywhere:
//   * the image is re-packed once (img_pack_kernel) to fp16 "pixel pairs": P[n][h][q'][8] holds input
//     columns (2q'-3, 2q'-2) x (r,g,b,0), q' in [0,132): 16 bytes per pair, zero pairs at both ends;
//   * a tile is one conv1 output row (128 pixels = the UMMA M) of one crop.  Output pixel wo, kernel row
//     kh reads pairs wo .. wo+3 of input row 2*ho+kh-3: for CONSECUTIVE wo these windows start 16 bytes
//     apart, which is exactly the row pitch of the un-swizzled K-major UMMA operand layout
//     ((8,m),(8,2)):((16 B, SBO),(2 B, LBO)) with SBO = 128 B and LBO = 16 B.  So the seven input rows are
//     staged ONCE by one TMA box (14.8 KB) and 14 MMAs (7 kernel rows x 2 pair-pairs, K = 16 each) read
//     overlapping windows of them straight from shared memory;
//   * the packed weights (14 x [64 cout][16 k] fp16, zero for kw = 7 and the 4th channel) stay resident
//     in shared memory for the whole kernel.
// A CTA owns a band of pool rows of one crop and walks its conv rows top to bottom; the pool is separable:
// each epilogue thread keeps the previous two rows of its own column in registers and writes the vertical
// maximum of three rows to shared memory every second row, from which the pooled row (raw and pre-activated)
// is the horizontal stride-2 maximum.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <vector>

#include "common.h"
#include "ptx.cuh"
#include "root_fused.h"

namespace metro {

namespace {

constexpr int kSide = 256, kConvW = 128, kPoolW = 64, kC = 64;
constexpr int kPairs = 132;                          // pairs per packed input row
constexpr int kRowBytes = kPairs * 16;               // 2112
constexpr int kStageRows = 7;
constexpr int kStageBytes = 15360;                   // 7 * 2112 = 14784, padded to a multiple of 1 KB
constexpr int kStages = 4;
constexpr int kMmas = 14;                            // 7 kernel rows x 2 K=16 steps
constexpr int kWBytes = kMmas * 2048;                // [mma][k-chunk 2][cout-group 8][8 rows][16 B]
constexpr int kHistBytes = kConvW * 128;             // one fp16 conv row: 128 px x 64 ch
constexpr int kAccStages = 4;                        // TMEM accumulator ring: the epilogue drains rows in pairs
constexpr int kThreads = 384;                        // 4 control warps + 8 epilogue warps (2 per TMEM lane quarter)

struct alignas(64) RootParams {
  CUtensorMap pmap;        // P as [176 fp16][6 segments][256 rows][n] (a row = 132 pairs x 8 fp16)
  const __half *wpack;     // kWBytes, already in the shared-memory operand layout
  const float *bias;       // [64] conv1 bias
  const float *pscale, *pshift;   // [64] first unit's pre-activation
  __half *raw, *pre;       // [n][64][64][64]; raw may be null
  __half *conv_dbg;        // optional [n][128][128][64]: conv1 output (keep_activations)
  int n, n_base, bands_per_img, pool_rows_per_band;   // crops n_base .. n_base + n of the buffers
  long long *prof;         // optional [grid][16] role timers (cycles)
  unsigned int *sig_flags; // optional per-crop completion counters (ptx.cuh dataflow): +1 per finished band
};

constexpr int kOffW = kStages * kStageBytes;                 // 61440
constexpr int kOffHist = kOffW + kWBytes;                    // 90112
constexpr int kOffBias = kOffHist + 2 * kHistBytes;          // two slots of column-wise maxima
constexpr int kOffBar = kOffBias + 256;
constexpr int kSmemBytes = kOffBar + 256;

// un-swizzled K-major operand: rows 16 B apart inside an 8-row core matrix, `sbo` between 8-row groups,
// `lbo` between the two 8-element K chunks (cute/atom/mma_traits_sm100.hpp, LayoutType::INTERLEAVE)
__device__ __forceinline__ uint64_t make_nosw_kmajor_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3FFF);
  d |= uint64_t((lbo >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;                         // descriptor version (sm_100)
  return d;                                       // layout type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

__global__ void __launch_bounds__(kThreads, 1) root_fused_kernel(const __grid_constant__ RootParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBar);
  uint64_t *full = bars, *empty = bars + kStages, *tfull = bars + 2 * kStages, *tempty = tfull + kAccStages;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 2 * kAccStages);
  float *s_bias = reinterpret_cast<float *>(smem + kOffBias);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_bands = p.n * p.bands_per_img;
  const int PB = p.pool_rows_per_band;

  if (warp == 0 && lane == 0) ptx::prefetch_tensormap(&p.pmap);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { ptx::mbar_init(full + i, 1); ptx::mbar_init(empty + i, 1); }
    for (int i = 0; i < kAccStages; ++i) { ptx::mbar_init(tfull + i, 1); ptx::mbar_init(tempty + i, 8); }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(s_tmem, kAccStages * kC);
    ptx::tmem_relinquish();
  }
  // resident weights (already in operand layout) and the bias
  for (int i = threadIdx.x; i < kWBytes / 16; i += kThreads)
    reinterpret_cast<uint4 *>(smem + kOffW)[i] = reinterpret_cast<const uint4 *>(p.wpack)[i];
  if (threadIdx.x < kC) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  ptx::fence_proxy_async();                       // generic-proxy writes of the weights -> tensor-core reads
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  ptx::griddep_wait();                            // the packed image comes from the previous kernel
  ptx::griddep_launch_dependents();

  // conv rows of band b: r = 2*p0 - 1 .. 2*(p0 + PB) - 1 (row -1 = zero padding of the pool, not computed)
  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      long long tw = 0;
      const long long tstart = clock64();
      for (int band = blockIdx.x; band < n_bands; band += gridDim.x) {
        const int img = p.n_base + band / p.bands_per_img, p0 = (band % p.bands_per_img) * PB;
        for (int r = max(2 * p0 - 1, 0); r <= 2 * (p0 + PB) - 1; ++r) {
          const long long t0 = p.prof ? clock64() : 0;
          ptx::mbar_wait(empty + stage, phase ^ 1);
          if (p.prof) tw += clock64() - t0;
          ptx::mbar_arrive_expect_tx(full + stage, kStageRows * kRowBytes);
          ptx::tma_load_4d(smem + stage * kStageBytes, &p.pmap, full + stage, 0, 0, 2 * r - 3, img);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (p.prof) { p.prof[blockIdx.x * 16 + 0] = clock64() - tstart; p.prof[blockIdx.x * 16 + 1] = tw; }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(128, kC);
      const uint64_t db0 = make_nosw_kmajor_desc(ptx::smem_u32(smem + kOffW), 1024, 128);
      int stage = 0;
      uint32_t phase = 0, it = 0;
      long long t_acc = 0, t_full = 0;
      const long long tstart = clock64();
      for (int band = blockIdx.x; band < n_bands; band += gridDim.x) {
        const int p0 = (band % p.bands_per_img) * PB;
        for (int r = max(2 * p0 - 1, 0); r <= 2 * (p0 + PB) - 1; ++r, ++it) {
          const int acc = it % kAccStages;
          const long long t0 = p.prof ? clock64() : 0;
          ptx::mbar_wait(tempty + acc, ((it / kAccStages) & 1) ^ 1);
          const long long t1 = p.prof ? clock64() : 0;
          ptx::mbar_wait(full + stage, phase);
          if (p.prof) { t_acc += t1 - t0; t_full += clock64() - t1; }
          ptx::tc_fence_after();
          // descriptors differ only in their 16-byte-granular start address: one 64-bit add per operand
          const uint64_t da0 = make_nosw_kmajor_desc(ptx::smem_u32(smem + stage * kStageBytes), 16, 128);
#pragma unroll
          for (int t = 0; t < kMmas; ++t) {
            const int kh = t >> 1, jp = t & 1;
            // A row m = output column wo: pairs wo + 2*jp, wo + 2*jp + 1 of packed input row kh
            ptx::umma_f16(tmem_base + acc * kC, da0 + uint64_t((kh * kRowBytes + jp * 32) >> 4), db0 + uint64_t(t * (2048 >> 4)),
                          idesc, t != 0);
          }
          ptx::umma_commit(empty + stage);
          ptx::umma_commit(tfull + acc);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (p.prof) { p.prof[blockIdx.x * 16 + 2] = clock64() - tstart; p.prof[blockIdx.x * 16 + 3] = t_acc; p.prof[blockIdx.x * 16 + 4] = t_full; p.prof[blockIdx.x * 16 + 5] = it; }
    }
  } else if (warp >= 4) {
    // ---- epilogue: thread = (conv output column wo = TMEM lane, half of the 64 channels) ----
    // The 3x3/2 max-pool is separable.  Vertical: a thread sees every conv row of its own column, so it keeps
    // the previous odd and even rows in registers (fp16, 16 + 16 registers) and, on every odd row, writes the
    // column-wise maximum of the three rows to shared memory.  Horizontal: after one barrier per pooled row the
    // 256 threads take the maximum of three neighbouring columns (stride 2) and emit the pooled row.
    const int e = warp - 4, q = e & 3, hf = e >> 2;           // q == warp % 4: the TMEM lane quarter
    const int wo = q * 32 + lane;
    const int et = threadIdx.x - 128;                         // 0..255
    const uint32_t hist_a = ptx::smem_u32(smem + kOffHist);
    const uint32_t taddr0 = tmem_base + (uint32_t(q * 32) << 16) + hf * 32;
    // horizontal role of this thread: channel chunk (8 channels) and pooled columns pw0 + 32 i
    const int chunk = et & 7, pw0 = et >> 3;
    float ps[8], pf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ps[i] = p.pscale[chunk * 8 + i]; pf[i] = p.pshift[chunk * 8 + i]; }
    __half2 odd[16];                                          // previous odd conv row of this column
    float bias[32];                                           // this thread's 32 channels (constant for the kernel)
#pragma unroll
    for (int i = 0; i < 32; ++i) bias[i] = s_bias[hf * 32 + i];
    uint32_t it = 0, pooled = 0;
    long long t_epi_wait = 0;
    const long long t_epi_start = clock64();
    // TMEM row -> fp16 conv1 values (+ bias, the rounding point of the stored tensor) of this thread's 32 channels
    auto to_half = [&](const uint32_t (&v)[32], __half2 (&h)[16], __half *dbg) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float *b = bias + 8 * j;
        uint4 o;
        o.x = pack2(__uint_as_float(v[8 * j + 0]) + b[0], __uint_as_float(v[8 * j + 1]) + b[1]);
        o.y = pack2(__uint_as_float(v[8 * j + 2]) + b[2], __uint_as_float(v[8 * j + 3]) + b[3]);
        o.z = pack2(__uint_as_float(v[8 * j + 4]) + b[4], __uint_as_float(v[8 * j + 5]) + b[5]);
        o.w = pack2(__uint_as_float(v[8 * j + 6]) + b[6], __uint_as_float(v[8 * j + 7]) + b[7]);
        h[4 * j + 0] = *reinterpret_cast<__half2 *>(&o.x); h[4 * j + 1] = *reinterpret_cast<__half2 *>(&o.y);
        h[4 * j + 2] = *reinterpret_cast<__half2 *>(&o.z); h[4 * j + 3] = *reinterpret_cast<__half2 *>(&o.w);
        if (dbg) reinterpret_cast<uint4 *>(dbg)[j] = o;
      }
    };
    auto wait_row = [&](uint32_t i) {
      const long long t0 = p.prof ? clock64() : 0;
      ptx::mbar_wait(tfull + i % kAccStages, (i / kAccStages) & 1);
      if (p.prof) t_epi_wait += clock64() - t0;
    };
    for (int band = blockIdx.x; band < n_bands; band += gridDim.x) {
      const int img = p.n_base + band / p.bands_per_img, p0 = (band % p.bands_per_img) * PB;
      __half *dbg0 = p.conv_dbg ? p.conv_dbg + (size_t(img) * kConvW * kConvW + wo) * kC + hf * 32 : nullptr;
      // halo row 2*p0 - 1 (the zero padding row above the image takes part in the max, Q6)
      if (p0 == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k) odd[k] = __floats2half2_rn(0.f, 0.f);
      } else {
        wait_row(it);
        ptx::tc_fence_after();
        uint32_t v[32];
        __syncwarp();
        ptx::tmem_ld_32x32(taddr0 + (it % kAccStages) * kC, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty + it % kAccStages);
        ++it;
        to_half(v, odd, dbg0 ? dbg0 + size_t(2 * p0 - 1) * kConvW * kC : nullptr);
      }
      for (int pr = p0; pr < p0 + PB; ++pr) {
        // conv rows 2*pr (even) and 2*pr + 1 (odd) together: one wake-up, two TMEM loads in flight
        wait_row(it);
        wait_row(it + 1);
        ptx::tc_fence_after();
        uint32_t va[32], vb[32];
        __syncwarp();
        ptx::tmem_ld_32x32(taddr0 + (it % kAccStages) * kC, va);
        ptx::tmem_ld_32x32(taddr0 + ((it + 1) % kAccStages) * kC, vb);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) { ptx::mbar_arrive(tempty + it % kAccStages); ptx::mbar_arrive(tempty + (it + 1) % kAccStages); }
        it += 2;
        __half2 ev[16], cur[16];
        to_half(va, ev, dbg0 ? dbg0 + size_t(2 * pr) * kConvW * kC : nullptr);
        to_half(vb, cur, dbg0 ? dbg0 + size_t(2 * pr + 1) * kConvW * kC : nullptr);
        // vertical maximum of rows 2*pr - 1, 2*pr, 2*pr + 1 of this column -> shared memory
        const uint32_t slot = hist_a + (pooled & 1) * kHistBytes + uint32_t(wo) * 128u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
          for (int k = 0; k < 4; ++k) oh[k] = __hmax2(__hmax2(odd[4 * j + k], ev[4 * j + k]), cur[4 * j + k]);
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slot + (uint32_t((4 * hf + j) ^ (wo & 7)) << 4)),
                       "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                       : "memory");
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) odd[k] = cur[k];
        // one barrier per pooled row: it also orders this row's reads before the write two rows later, which
        // goes to the same slot (a thread arrives here only after its reads of the previous pooled row)
        ptx::named_bar_sync(1, 256);
        const uint32_t vrow = hist_a + (pooled & 1) * kHistBytes;
        ++pooled;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int pw = pw0 + 32 * i;
          __half2 m[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) m[k] = __floats2half2_rn(0.f, 0.f);     // column -1 is zero padding
          bool first = pw > 0;
#pragma unroll
          for (int dc = -1; dc <= 1; ++dc) {
            const int col = 2 * pw + dc;
            if (col < 0) continue;
            const uint4 x = ptx::lds_v4u(vrow + uint32_t(col) * 128u + (uint32_t(chunk ^ (col & 7)) << 4));
            const __half2 *xh = reinterpret_cast<const __half2 *>(&x);
            if (first) {
#pragma unroll
              for (int k = 0; k < 4; ++k) m[k] = xh[k];
              first = false;
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) m[k] = __hmax2(m[k], xh[k]);
            }
          }
          const size_t o = ((size_t(img) * kPoolW + pr) * kPoolW + pw) * kC + chunk * 8;
          uint4 ro;
          __half2 *rh = reinterpret_cast<__half2 *>(&ro);
#pragma unroll
          for (int k = 0; k < 4; ++k) rh[k] = m[k];
          if (p.raw) *reinterpret_cast<uint4 *>(p.raw + o) = ro;
          uint4 po;
          uint32_t *pw32 = reinterpret_cast<uint32_t *>(&po);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 y = __half22float2(m[k]);
            pw32[k] = pack2_relu(fmaf(y.x, ps[2 * k], pf[2 * k]), fmaf(y.y, ps[2 * k + 1], pf[2 * k + 1]));
          }
          *reinterpret_cast<uint4 *>(p.pre + o) = po;
        }
      }
      if (p.sig_flags) {
        // this band's pooled rows are stored: every thread publishes its stores (GPU scope), one reports the band
        __threadfence();
        ptx::named_bar_sync(1, 256);
        if (et == 0) ptx::flag_signal(p.sig_flags + img);
      }
    }
    if (p.prof && e == 0 && lane == 0) { p.prof[blockIdx.x * 16 + 6] = clock64() - t_epi_start; p.prof[blockIdx.x * 16 + 7] = t_epi_wait; }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kAccStages * kC);
  }
}

// float32 / uint8 NHWC [n,256,256,3] -> fp16 pairs [n][256][132][8]; thread = one pair.  Values are rounded
// to fp16 exactly like the reference's cast to FLAGS.dtype (architectures.py:29); the uint8 variant fuses the
// /255 of improc.py:56-61.
template <bool U8>
__global__ void __launch_bounds__(256) img_pack_kernel(const void *__restrict__ img, __half *__restrict__ out, int n) {
  const size_t total = size_t(n) * kSide * kPairs;
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int q = int(i % kPairs);
  const size_t row = i / kPairs;                      // n * 256 + h
  float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int col = 2 * q - 3 + e;
    if (col < 0 || col >= kSide) continue;
    const size_t base = (row * kSide + col) * 3;
    if (U8) {
      const unsigned char *s = static_cast<const unsigned char *>(img) + base;
      for (int c = 0; c < 3; ++c) v[3 * e + c] = float(s[c]) * (1.0f / 255.0f);
    } else {
      const float *s = static_cast<const float *>(img) + base;
      for (int c = 0; c < 3; ++c) v[3 * e + c] = s[c];
    }
  }
  uint4 o;
  o.x = pack2(v[0], v[1]); o.y = pack2(v[2], 0.f); o.z = pack2(v[3], v[4]); o.w = pack2(v[5], 0.f);
  reinterpret_cast<uint4 *>(out)[i] = o;
}

// =====================================================================================================================
// Version 2 of the root: the image pack is folded into the kernel and two conv rows share one MMA.
//
//  * No packed image in HBM: four loader warps read the float32 (or uint8) crop rows straight from the caller's
//    buffer, convert to fp16 (the cast of architectures.py:29; uint8: x * (1/255f), bit-identical after the cast to
//    improc.py:56-61's division) and write the "pixel pair" rows of the layout above into a ring of 32 input rows in
//    shared memory.  A conv row pair needs 4 new input rows; the 138 MB write + read of the packed image and one
//    launch per step disappear.
//  * conv rows (a, a + 1) are ONE accumulator of 128 columns: the 9 input rows 2a-3 .. 2a+5 are the K dimension
//    (window position kpos), the B operand of position kpos holds the filters of kernel row kpos for conv row a
//    (zero for kpos > 6) in its first 64 rows and of kernel row kpos - 2 for conv row a + 1 (zero for kpos < 2) in the
//    other 64: 18 MMAs of N = 128 (64 cycles each) per pair instead of 28 of N = 64 (56 cycles each).
//  * A CTA owns a band of 16 pool rows of one crop = 16 conv row pairs plus the pair above them whose odd row is the
//    pool's halo; the epilogue is the one of version 1 with both rows of a pair coming from one accumulator.
constexpr int kRingRows = 32;                       // input rows resident in shared memory (a pair's window is 9)
constexpr int kPairMmas = 18;                       // 9 window positions x 2 K = 16 steps
constexpr int kW2Bytes = kPairMmas * 4096;          // [mma][k-chunk 2][row-group 16][8 rows][16 B]
constexpr int kPairDepth = 4;                       // pairs in flight: accumulators (4 x 128 columns) and row hand-over
constexpr int kBand2 = 16;                          // pool rows per band
constexpr int kLoaderWarps = 4;
constexpr int kThreads2 = kThreads + kLoaderWarps * 32;   // 4 control + 8 epilogue + 4 loader warps

struct alignas(16) Root2Params {
  const void *images;      // [n][256][256][3] float32 or uint8, crop 0 of THIS call
  int u8;
  const __half *wpack;     // kW2Bytes, operand layout
  const float *bias, *pscale, *pshift;
  __half *raw, *pre, *conv_dbg;
  int n, n_base;
  long long *prof;
  unsigned long long *tstamp;
};

constexpr int k2OffW = kRingRows * kRowBytes;                 // 67584
constexpr int k2OffHist = k2OffW + kW2Bytes;                  // + 73728
constexpr int k2OffBias = k2OffHist + 2 * kHistBytes;
constexpr int k2OffBar = k2OffBias + 256;
constexpr int k2SmemBytes = k2OffBar + 256;

__global__ void __launch_bounds__(kThreads2, 1) root_fused2_kernel(const __grid_constant__ Root2Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + k2OffBar);
  uint64_t *rfull = bars, *rempty = bars + kPairDepth, *tfull = bars + 2 * kPairDepth, *tempty = bars + 3 * kPairDepth;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + 4 * kPairDepth);
  float *s_bias = reinterpret_cast<float *>(smem + k2OffBias);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kBandsPerImg = kPoolW / kBand2;
  const int n_bands = p.n * kBandsPerImg;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kPairDepth; ++i) {
      ptx::mbar_init(rfull + i, kLoaderWarps); ptx::mbar_init(rempty + i, 1);
      ptx::mbar_init(tfull + i, 1); ptx::mbar_init(tempty + i, 8);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(s_tmem, kPairDepth * 128);
    ptx::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kW2Bytes / 16; i += kThreads2)
    reinterpret_cast<uint4 *>(smem + k2OffW)[i] = reinterpret_cast<const uint4 *>(p.wpack)[i];
  // the ring starts zeroed: the two zero pairs in front of a row and the 2.5 behind it are never written again
  for (int i = threadIdx.x; i < kRingRows * kRowBytes / 16; i += kThreads2) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < kC) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  ptx::griddep_wait();                            // the caller's images / whatever ran before in the stream
  ptx::griddep_launch_dependents();
  ptx::stamp_begin(p.tstamp);

  // Pairs of a band: j = 0 .. 16, conv rows (2 p0 - 2 + 2 j, 2 p0 - 1 + 2 j); pair 0 only contributes the halo row and
  // does not exist for the top band (the pool's zero padding).  Every role walks the same sequence.
  if (warp >= 4 + 8) {
    // ================================ loaders ================================
    const int lt = threadIdx.x - (4 + 8) * 32;     // 0..127: two rows at a time, 64 threads (4 pixels each) per row
    const int half = lt >> 6, t = lt & 63;
    uint32_t seq = 0;                              // pairs issued so far
    uint32_t ring = 0;                             // ring slot of the next row to write (rows are loaded in window order)
    for (int band = blockIdx.x; band < n_bands; band += gridDim.x) {
      const int img = band / kBandsPerImg, p0 = (band % kBandsPerImg) * kBand2;
      const unsigned char *src = static_cast<const unsigned char *>(p.images) + size_t(img) * kSide * kSide * 3 * (p.u8 ? 1 : 4);
      for (int j = p0 == 0 ? 1 : 0; j <= kBand2; ++j, ++seq) {
        const int a = 2 * p0 - 2 + 2 * j;          // even conv row of the pair
        const bool first = j == (p0 == 0 ? 1 : 0);
        const int r_lo = first ? 2 * a - 3 : 2 * a + 2, r_hi = 2 * a + 5;     // input rows this pair adds to the ring
        // rows overwritten now were last read by the pair 4 back (kRingRows = 32 makes that hold at a band start too)
        bool slot_free = seq < kPairDepth;         // waited for just before the first store: the loads do not need the slot
        // L2 prefetch of the rows two pairs ahead (the crops come from HBM: this turns the DRAM latency of the loads
        // below into an L2 hit by the time they are issued); one 128-byte line per thread
        {
          const int pr_lo = r_hi + 5, pr_rows = 4;                      // rows 2(a+4)+2 .. +5 of pair j + 2
          const int line = lt;                                          // 0..127
          const int bytes_per_row = kSide * 3 * (p.u8 ? 1 : 4);
          const int off = line * 128;
          if (off < pr_rows * bytes_per_row && pr_lo >= 0 && pr_lo + pr_rows <= kSide && j + 2 <= kBand2)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(src + size_t(pr_lo) * bytes_per_row + off));
        }
        // all loads of up to three rounds (two rows each) are issued before the first conversion: the loader is bound by
        // global-load latency, not by bytes
        for (int rb = r_lo; rb <= r_hi; rb += 6) {
          float4 f[3][3];
          uint32_t w[3][3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int r = rb + 2 * k + half;
            if (r <= r_hi && r >= 0 && r < kSide) {
              if (p.u8) {
                const uint32_t *q = reinterpret_cast<const uint32_t *>(src + (size_t(r) * kSide + 4 * t) * 3);
                w[k][0] = q[0]; w[k][1] = q[1]; w[k][2] = q[2];
              } else {
                const float4 *q = reinterpret_cast<const float4 *>(src + (size_t(r) * kSide + 4 * t) * 12);
                f[k][0] = q[0]; f[k][1] = q[1]; f[k][2] = q[2];
              }
            }
          }
          if (!slot_free) { ptx::mbar_wait(rempty + seq % kPairDepth, ((seq / kPairDepth) & 1) ^ 1); slot_free = true; }
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int r = rb + 2 * k + half;
            if (r > r_hi) continue;
            uint2 o[4];
            if (r < 0 || r >= kSide) {
#pragma unroll
              for (int i = 0; i < 4; ++i) o[i] = make_uint2(0u, 0u);
            } else if (p.u8) {
              const uint32_t w0 = w[k][0], w1 = w[k][1], w2 = w[k][2];
              const unsigned char b[12] = {(unsigned char)w0, (unsigned char)(w0 >> 8), (unsigned char)(w0 >> 16), (unsigned char)(w0 >> 24),
                                           (unsigned char)w1, (unsigned char)(w1 >> 8), (unsigned char)(w1 >> 16), (unsigned char)(w1 >> 24),
                                           (unsigned char)w2, (unsigned char)(w2 >> 8), (unsigned char)(w2 >> 16), (unsigned char)(w2 >> 24)};
#pragma unroll
              for (int i = 0; i < 4; ++i)
                o[i] = make_uint2(pack2(float(b[3 * i]) * (1.0f / 255.0f), float(b[3 * i + 1]) * (1.0f / 255.0f)),
                                  pack2(float(b[3 * i + 2]) * (1.0f / 255.0f), 0.f));
            } else {
              const float4 f0 = f[k][0], f1 = f[k][1], f2 = f[k][2];
              o[0] = make_uint2(pack2(f0.x, f0.y), pack2(f0.z, 0.f));
              o[1] = make_uint2(pack2(f0.w, f1.x), pack2(f1.y, 0.f));
              o[2] = make_uint2(pack2(f1.z, f1.w), pack2(f2.x, 0.f));
              o[3] = make_uint2(pack2(f2.y, f2.z), pack2(f2.w, 0.f));
            }
            // column c sits in pair (c + 3) / 2, half (c + 3) % 2: columns 4t .. 4t+3 are 32 contiguous bytes from 32 t + 24
            const uint32_t slot = (ring + uint32_t(r - r_lo)) % kRingRows;
            const uint32_t dst = ptx::smem_u32(smem) + slot * kRowBytes + 32u * t + 24u;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(dst + 8u * i), "r"(o[i].x), "r"(o[i].y) : "memory");
          }
        }
        ring = (ring + uint32_t(r_hi - r_lo + 1)) % kRingRows;
        ptx::fence_proxy_async();                  // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(rfull + seq % kPairDepth);
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(128, 128);
      const uint64_t db0 = make_nosw_kmajor_desc(ptx::smem_u32(smem + k2OffW), 2048, 128);
      uint32_t seq = 0, win = 0;                   // win: ring slot of the current pair's window position 0
      for (int band = blockIdx.x; band < n_bands; band += gridDim.x) {
        const int p0 = (band % kBandsPerImg) * kBand2;
        for (int j = p0 == 0 ? 1 : 0; j <= kBand2; ++j, ++seq) {
          const bool first = j == (p0 == 0 ? 1 : 0);
          // the window slides by 4 rows per pair; a band's first pair starts on 9 fresh rows
          if (seq > 0) win = (win + (first ? 9u : 4u)) % kRingRows;
          const uint32_t acc = seq % kPairDepth, ph = (seq / kPairDepth) & 1;
          ptx::mbar_wait(tempty + acc, ph ^ 1);
          ptx::mbar_wait(rfull + acc, ph);
          ptx::tc_fence_after();
#pragma unroll
          for (int tt = 0; tt < kPairMmas; ++tt) {
            const int kpos = tt >> 1, jp = tt & 1;
            const uint32_t row = (win + uint32_t(kpos)) % kRingRows;
            const uint64_t da = make_nosw_kmajor_desc(ptx::smem_u32(smem) + row * kRowBytes + jp * 32, 16, 128);
            ptx::umma_f16(tmem_base + acc * 128, da, db0 + uint64_t(tt * (4096 >> 4)), idesc, tt != 0);
          }
          ptx::umma_commit(rempty + acc);
          ptx::umma_commit(tfull + acc);
        }
      }
    }
  } else if (warp >= 4) {
    // ---- epilogue: version 1's, with the even row of a pair in accumulator columns 0..63 and the odd row in 64..127 ----
    const int e = warp - 4, q = e & 3, hf = e >> 2;
    const int wo = q * 32 + lane;
    const int et = threadIdx.x - 128;
    const uint32_t hist_a = ptx::smem_u32(smem + k2OffHist);
    const uint32_t taddr0 = tmem_base + (uint32_t(q * 32) << 16) + hf * 32;
    const int chunk = et & 7, pw0 = et >> 3;
    float ps[8], pf[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ps[i] = p.pscale[chunk * 8 + i]; pf[i] = p.pshift[chunk * 8 + i]; }
    __half2 odd[16];
    float bias[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) bias[i] = s_bias[hf * 32 + i];
    uint32_t seq = 0, pooled = 0;
    auto to_half = [&](const uint32_t (&v)[32], __half2 (&h)[16], __half *dbg) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float *b = bias + 8 * j;
        uint4 o;
        o.x = pack2(__uint_as_float(v[8 * j + 0]) + b[0], __uint_as_float(v[8 * j + 1]) + b[1]);
        o.y = pack2(__uint_as_float(v[8 * j + 2]) + b[2], __uint_as_float(v[8 * j + 3]) + b[3]);
        o.z = pack2(__uint_as_float(v[8 * j + 4]) + b[4], __uint_as_float(v[8 * j + 5]) + b[5]);
        o.w = pack2(__uint_as_float(v[8 * j + 6]) + b[6], __uint_as_float(v[8 * j + 7]) + b[7]);
        h[4 * j + 0] = *reinterpret_cast<__half2 *>(&o.x); h[4 * j + 1] = *reinterpret_cast<__half2 *>(&o.y);
        h[4 * j + 2] = *reinterpret_cast<__half2 *>(&o.z); h[4 * j + 3] = *reinterpret_cast<__half2 *>(&o.w);
        if (dbg) reinterpret_cast<uint4 *>(dbg)[j] = o;
      }
    };
    for (int band = blockIdx.x; band < n_bands; band += gridDim.x) {
      const int img = p.n_base + band / kBandsPerImg, p0 = (band % kBandsPerImg) * kBand2;
      __half *dbg0 = p.conv_dbg ? p.conv_dbg + (size_t(img) * kConvW * kConvW + wo) * kC + hf * 32 : nullptr;
      if (p0 == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k) odd[k] = __floats2half2_rn(0.f, 0.f);
      } else {
        // the halo: odd row 2 p0 - 1 of pair 0 (its even row belongs to the band above)
        const uint32_t acc = seq % kPairDepth;
        ptx::mbar_wait(tfull + acc, (seq / kPairDepth) & 1);
        ptx::tc_fence_after();
        uint32_t v[32];
        __syncwarp();
        ptx::tmem_ld_32x32(taddr0 + acc * 128 + 64, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty + acc);
        ++seq;
        to_half(v, odd, dbg0 ? dbg0 + size_t(2 * p0 - 1) * kConvW * kC : nullptr);
      }
      for (int pr = p0; pr < p0 + kBand2; ++pr, ++seq) {
        const uint32_t acc = seq % kPairDepth;
        ptx::mbar_wait(tfull + acc, (seq / kPairDepth) & 1);
        ptx::tc_fence_after();
        uint32_t va[32], vb[32];
        __syncwarp();
        ptx::tmem_ld_32x32(taddr0 + acc * 128, va);
        ptx::tmem_ld_32x32(taddr0 + acc * 128 + 64, vb);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty + acc);
        __half2 ev[16], cur[16];
        to_half(va, ev, dbg0 ? dbg0 + size_t(2 * pr) * kConvW * kC : nullptr);
        to_half(vb, cur, dbg0 ? dbg0 + size_t(2 * pr + 1) * kConvW * kC : nullptr);
        const uint32_t slot = hist_a + (pooled & 1) * kHistBytes + uint32_t(wo) * 128u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          __half2 *oh = reinterpret_cast<__half2 *>(&o);
#pragma unroll
          for (int k = 0; k < 4; ++k) oh[k] = __hmax2(__hmax2(odd[4 * j + k], ev[4 * j + k]), cur[4 * j + k]);
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slot + (uint32_t((4 * hf + j) ^ (wo & 7)) << 4)),
                       "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                       : "memory");
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) odd[k] = cur[k];
        ptx::named_bar_sync(1, 256);
        const uint32_t vrow = hist_a + (pooled & 1) * kHistBytes;
        ++pooled;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int pw = pw0 + 32 * i;
          __half2 m[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) m[k] = __floats2half2_rn(0.f, 0.f);
          bool first = pw > 0;
#pragma unroll
          for (int dc = -1; dc <= 1; ++dc) {
            const int col = 2 * pw + dc;
            if (col < 0) continue;
            const uint4 x = ptx::lds_v4u(vrow + uint32_t(col) * 128u + (uint32_t(chunk ^ (col & 7)) << 4));
            const __half2 *xh = reinterpret_cast<const __half2 *>(&x);
            if (first) {
#pragma unroll
              for (int k = 0; k < 4; ++k) m[k] = xh[k];
              first = false;
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) m[k] = __hmax2(m[k], xh[k]);
            }
          }
          const size_t o = ((size_t(img) * kPoolW + pr) * kPoolW + pw) * kC + chunk * 8;
          uint4 ro;
          __half2 *rh = reinterpret_cast<__half2 *>(&ro);
#pragma unroll
          for (int k = 0; k < 4; ++k) rh[k] = m[k];
          if (p.raw) *reinterpret_cast<uint4 *>(p.raw + o) = ro;
          uint4 po;
          uint32_t *pw32 = reinterpret_cast<uint32_t *>(&po);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 y = __half22float2(m[k]);
            pw32[k] = pack2_relu(fmaf(y.x, ps[2 * k], pf[2 * k]), fmaf(y.y, ps[2 * k + 1], pf[2 * k + 1]));
          }
          *reinterpret_cast<uint4 *>(p.pre + o) = po;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kPairDepth * 128);
  }
  ptx::stamp_end(p.tstamp);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

size_t root2_packed_weight_elems() { return kW2Bytes / 2; }

void root2_pack_weights(const float *w_hwio, __half *dst) {
  // HWIO [7][7][3][64] -> 18 x [k-chunk 2][row-group 16][8 rows][8 k] fp16 (un-swizzled K-major operand of 128 rows):
  // MMA t = 2 * kpos + jp; rows 0..63 = conv row a (kernel row kpos), rows 64..127 = conv row a + 1 (kernel row kpos - 2)
  for (size_t i = 0; i < size_t(kW2Bytes / 2); ++i) dst[i] = __float2half_rn(0.f);
  for (int t = 0; t < kPairMmas; ++t) {
    const int kpos = t >> 1, jp = t & 1;
    for (int k = 0; k < 16; ++k) {
      const int kw = 2 * (2 * jp + (k >> 3)) + ((k >> 2) & 1), ch = k & 3;
      if (kw > 6 || ch > 2) continue;
      for (int row = 0; row < 128; ++row) {
        const int kh = row < 64 ? kpos : kpos - 2, o = row & 63;
        if (kh < 0 || kh > 6) continue;
        dst[size_t(t) * 2048 + (k >> 3) * 1024 + (row >> 3) * 64 + (row & 7) * 8 + (k & 7)] =
            __float2half_rn(w_hwio[((size_t(kh) * 7 + kw) * 3 + ch) * kC + o]);
      }
    }
  }
}

metro_status root_fused2_launch(const void *images, bool u8, const __half *wpack2, const float *bias, const float *pscale,
                                const float *pshift, __half *raw, __half *pre, __half *conv_dbg, int n, int n_base, int num_sms,
                                cudaStream_t stream, unsigned long long *tstamp) {
  if (n == 0) return METRO_OK;
  static PerDeviceOnce configured;     // function attributes are per device
  metro_status cst = configured.run([] {
    METRO_CUDA(cudaFuncSetAttribute(root_fused2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemBytes));
    return METRO_OK;
  });
  if (cst != METRO_OK) return cst;
  Root2Params p{};
  p.images = images; p.u8 = u8 ? 1 : 0; p.wpack = wpack2; p.bias = bias; p.pscale = pscale; p.pshift = pshift;
  p.raw = raw; p.pre = pre; p.conv_dbg = conv_dbg; p.n = n; p.n_base = n_base; p.prof = nullptr; p.tstamp = tstamp;
  const int n_bands = n * (kPoolW / kBand2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(n_bands < num_sms ? n_bands : num_sms)); cfg.blockDim = dim3(kThreads2);
  cfg.dynamicSmemBytes = k2SmemBytes; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = getenv("METRO_NO_PDL") != nullptr;
  cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
  METRO_CUDA(cudaLaunchKernelEx(&cfg, root_fused2_kernel, p));
  return METRO_OK;
}

size_t root_packed_image_elems() { return size_t(kSide) * kPairs * 8; }
size_t root_packed_weight_elems() { return kWBytes / 2; }

void root_pack_weights(const float *w_hwio, __half *dst) {
  // HWIO [7][7][3][64] -> 14 x [k-chunk 2][cout-group 8][8 rows][8 k] fp16 (un-swizzled K-major operand):
  // MMA t = 2*kh + jp, k = 16 values = pairs (2jp, 2jp+1) x (2 pixels) x (r,g,b,0): kw = 2*(2jp + k/8) + (k/4)%2
  for (size_t i = 0; i < size_t(kWBytes / 2); ++i) dst[i] = __float2half_rn(0.f);
  for (int t = 0; t < kMmas; ++t) {
    const int kh = t >> 1, jp = t & 1;
    for (int k = 0; k < 16; ++k) {
      const int kw = 2 * (2 * jp + (k >> 3)) + ((k >> 2) & 1), ch = k & 3;
      if (kw > 6 || ch > 2) continue;
      for (int o = 0; o < kC; ++o)
        dst[size_t(t) * 1024 + (k >> 3) * 512 + (o >> 3) * 64 + (o & 7) * 8 + (k & 7)] =
            __float2half_rn(w_hwio[((size_t(kh) * 7 + kw) * 3 + ch) * kC + o]);
    }
  }
}

metro_status root_make_image_map(void *map_out, const __half *packed, int n) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return fail(METRO_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  // a packed row (132 pairs x 8 fp16 = 2112 B) is described as 6 segments of 176 fp16: the TMA engine works
  // through a box one innermost row at a time, and 16-byte innermost rows (one pair) made the copy of a
  // 14.8 KB tile cost ~1800 cycles
  const cuuint64_t dims[4] = {176, 6, cuuint64_t(kSide), cuuint64_t(n)};
  const cuuint64_t strides[3] = {352, cuuint64_t(kRowBytes), cuuint64_t(kRowBytes) * kSide};
  const cuuint32_t box[4] = {176, 6, cuuint32_t(kStageRows), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(static_cast<CUtensorMap *>(map_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half *>(packed),
                        dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(METRO_ERR_CUDA, "cuTensorMapEncodeTiled(packed image, n=%d) -> %d", n, int(r));
  return METRO_OK;
}

metro_status img_pack_launch(const void *img, bool u8, __half *out, int n, cudaStream_t stream) {
  if (n == 0) return METRO_OK;
  const size_t total = size_t(n) * kSide * kPairs;
  const unsigned blocks = unsigned((total + 255) / 256);
  if (u8) img_pack_kernel<true><<<blocks, 256, 0, stream>>>(img, out, n);
  else img_pack_kernel<false><<<blocks, 256, 0, stream>>>(img, out, n);
  METRO_CUDA(cudaGetLastError());
  return METRO_OK;
}

metro_status root_fused_launch(const void *image_map, const __half *wpack, const float *bias, const float *pscale,
                               const float *pshift, __half *raw, __half *pre, __half *conv_dbg, int n, int n_base,
                               int num_sms, cudaStream_t stream, long long *prof, unsigned int *sig_flags) {
  if (n == 0) return METRO_OK;
  static PerDeviceOnce configured;     // function attributes are per device
  metro_status cst = configured.run([] {
    METRO_CUDA(cudaFuncSetAttribute(root_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    return METRO_OK;
  });
  if (cst != METRO_OK) return cst;
  RootParams p;
  p.pmap = *static_cast<const CUtensorMap *>(image_map);
  p.wpack = wpack; p.bias = bias; p.pscale = pscale; p.pshift = pshift;
  p.raw = raw; p.pre = pre; p.conv_dbg = conv_dbg;
  p.prof = prof; p.sig_flags = sig_flags;
  p.n = n; p.n_base = n_base; p.pool_rows_per_band = 8; p.bands_per_img = kPoolW / 8;
  const int n_bands = n * p.bands_per_img;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(n_bands < num_sms ? n_bands : num_sms)); cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = getenv("METRO_NO_PDL") != nullptr;
  cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
  METRO_CUDA(cudaLaunchKernelEx(&cfg, root_fused_kernel, p));
  return METRO_OK;
}

}  // namespace metro
