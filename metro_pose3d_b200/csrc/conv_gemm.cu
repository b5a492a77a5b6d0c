// Fused implicit-GEMM convolution for sm_100a.
//
//   D[M = N*Ho*Wo pixels, Cout] = sum over taps (kh,kw) and 64-channel blocks of
//                                 X[n, ho*s + kh*r - pad, wo*s + kw*r - pad, c] * W[kh,kw,c,cout]
//
// One kernel template serves every convolution of the residual body and the logits head
// (resnet_v2.py:123-136,234-236); DESIGN.md section 3.1 has the measured rationale of each choice:
//   * tile = CTA pair (cluster of 2, tcgen05 cta_group::2): 256 output pixels x BLOCK_N channels; each CTA
//     stages its own 128 pixel rows of A and half of the weight rows, the leader issues M = 256 MMAs.
//   * A operand: NHWC fp16 activations fetched by TMA as 4-D boxes [64 ch, Wo, th, nb] (128 output
//     pixels = th full output rows of nb crops) at the tap's shifted coordinate; TMA's out-of-bounds
//     zero fill IS the convolution's zero padding (resnet_utils.py:120-135), dilation is a larger
//     shift, and a stride-2 conv reads four (row,col)-parity views of the input (tensor maps with
//     doubled strides), so no im2col buffer and no strided gather exists anywhere.  Narrow 3x3 stride-1
//     convolutions stage one tall column-shifted box per kernel column that serves all three kernel rows.
//   * B operand: weights pre-packed [Cout][K] fp16 K-major, 2-D TMA boxes [64, BLOCK_N / 2].
//   * both land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes.
//   * accumulators live in TMEM (fp32), double-buffered so the epilogue of tile i overlaps the
//     main loop of tile i+1; the kernel is persistent (one CTA per SM, static round-robin pair tiles,
//     walked forwards or backwards on alternating layers).
//   * warp roles: 0 = TMA producer (A), 3 = TMA producer (B), 1 = MMA issuer (one elected thread),
//     2 = TMEM allocator, 4..11 = epilogue in two groups (group g drains accumulator stage g; thread =
//     accumulator row = output pixel), 12..15 (kXform only) = A-tile pre-activation.
//   * epilogue fuses folded BN / bias, ReLU and the *next* layer's pre-activation BN+ReLU as a second
//     output (resnet_v2.py:119), so no stand-alone normalisation pass ever touches HBM.  Both kinds of
//     shortcut are extra K blocks of the same accumulator: the projection shortcut (resnet_v2.py:123-125)
//     against its 1x1 filters, the identity shortcut (optionally sub-sampled with the centred-stride
//     offset, resnet_v2.py:120-121) against identity weights.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "conv_gemm.h"
#include "ptx.cuh"

namespace metro {

namespace {

constexpr int kCtrlWarps = 4;                 // 0 = TMA producer (A), 1 = MMA issuer, 2 = TMEM allocator, 3 = TMA producer (B)
constexpr int kEpiWarps = 8;                  // two groups of four; group g drains accumulator stage g
constexpr int kThreads = (kCtrlWarps + kEpiWarps) * 32;
constexpr int kXformWarps = 4;                // optional: apply the pre-activation to the A tiles in shared memory
constexpr int kThreadsXform = kThreads + kXformWarps * 32;
constexpr int kSmemLimit = 232448;            // 227 KB opt-in maximum per CTA on sm_100

enum Mode { kSingle = 0, kDual = 1, kDirect = 2 };

template <int BLOCK_N>
struct Cfg {
  static constexpr int kABytes = kTileM * kTileK * 2;          // 16 KB: this CTA's 128 pixel rows
  static constexpr int kBBytes = (BLOCK_N / 2) * kTileK * 2;   // this CTA's half of the weight rows
  static constexpr int kBlockBytes = kABytes + kBBytes;        // one 64-channel K block
  // K blocks per pipeline stage (one barrier round trip, one commit): the per-stage cost of the two
  // single-thread roles (~250 cycles) has to stay below the stage's MMA time, which is only
  // 4 x 56 / 64 cycles per K block for narrow tiles
  static constexpr int kSub = BLOCK_N <= 160 ? 2 : 1;
  static constexpr int kStageBytes = kSub * kBlockBytes;
  static constexpr int kAccCols = BLOCK_N <= 64 ? 64 : (BLOCK_N <= 128 ? 128 : 256);  // per accumulator stage
  static constexpr int kTmemCols = 2 * kAccCols;               // power of two >= 32
};

// barrier block layout (uint64 slots): full[8] empty[8] ready[8] tfull[2] tempty[2], then the TMEM base
constexpr int kBarFull = 0, kBarEmpty = kMaxStages, kBarReady = 2 * kMaxStages, kBarTFull = 3 * kMaxStages,
              kBarTEmpty = kBarTFull + 2, kBarCount = kBarTEmpty + 2;

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack_relu_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// per-CTA role timers (cycles), written when p.prof != nullptr
enum Prof { kPTotal = 0, kPProdWait, kPMmaWaitFull, kPMmaWaitAcc, kPEpiWaitAcc, kPEpiBusy, kPEpiWaitStore, kPTiles,
            kPEpiLd, kPEpiMath, kPEpiSts, kPEpiIssue, kPEpiPar, kPMmaIssue, kPMmaCommit, kPProdIssue, kPCount = 16 };

// The kernel runs as CTA pairs (cluster of 2 = one TPC, tcgen05 cta_group::2): a pair owns a tile of
// 256 output pixels x BLOCK_N channels.  Each CTA stages its own 128 pixel rows of A and HALF of the
// weight rows per K block (so a stage is 16 KB + BLOCK_N/2 x 128 B and shared-memory write + read
// traffic per MMA halves against a one-CTA tile), the leader's elected thread issues one M=256 MMA
// for both tensor cores, and each CTA drains its own 128 accumulator rows from its own TMEM.
//
// kXform: the A operand is a unit's RAW input and the unit's pre-activation relu(bn(x)) (resnet_v2.py:119)
// is applied to each A tile in shared memory by four extra warps between the TMA landing and the MMA, so
// the producing conv3 does not have to write the pre-activated tensor to HBM at all (1x1 convs only).
template <int BLOCK_N, int kMode, bool kXform>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kXform ? kThreadsXform : kThreads, 1)
    conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  using C = Cfg<BLOCK_N>;
  constexpr int kParVecs = kMode == kDual ? 3 : 2;   // single/direct: scale, shift; dual: shift, scale2, shift2
  constexpr bool kNarrowId = BLOCK_N == 256 && kMode != kDirect && !kXform;   // one K block per stage there
  // 1024-byte alignment is required by the 128B swizzle atoms (8 rows x 128 B).  The kernel has no
  // static shared memory, so the dynamic window starts at the CTA's shared base; verified below.
  extern __shared__ __align__(1024) unsigned char smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("metro: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  unsigned char *tiles = smem;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
  uint64_t *full = bars + kBarFull, *empty = bars + kBarEmpty, *ready = bars + kBarReady;
  uint64_t *tfull = bars + kBarTFull, *tempty = bars + kBarTEmpty;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + kBarCount);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // K blocks per tile: taps x cblk0 from source 0, then source 1: all cblk1 blocks (projection
  // shortcut) or, for the identity shortcut (diag2), only the BLOCK_N/64 blocks whose identity
  // weights hit this N tile.
  const int k0 = p.taps * p.cblk0;
  const uint32_t rank = ptx::cluster_ctarank();     // 0 = leader of the pair
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_tiles_total = ((p.m_tiles + 1) >> 1) * p.n_tiles;   // pair tiles
  const int stages = p.stages;
  const bool prof = p.prof != nullptr;
  const long long t_start = prof ? clock64() : 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.amap[0]);
    ptx::prefetch_tensormap(&p.bmap);
    if (p.cblk1) ptx::prefetch_tensormap(&p.a2map);
    if (kMode != kDirect) {
      if (p.has_out1) ptx::prefetch_tensormap(&p.o1map);
      if (p.has_out2) ptx::prefetch_tensormap(&p.o2map);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      ptx::mbar_init(full + i, 1); ptx::mbar_init(empty + i, 1);
      ptx::mbar_init(ready + i, 2 * kXformWarps);    // leader's: the transform warps of both CTAs
    }
    // tempty (the leader's is the one used): the four warps of the group in BOTH CTAs
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(tfull + i, 1); ptx::mbar_init(tempty + i, kEpiWarps); }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_pair(s_tmem, C::kTmemCols);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncwarp();
  ptx::cluster_sync();                               // both CTAs' barriers exist before any remote signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) ran while the
  // previous kernel in the stream was still draining; its outputs may only be touched from here on.
  // (With dataflow counters the grid-wide wait is replaced by per-crop waits in the activation producer: this kernel's
  // first tiles start while the previous layer's last wave is still running.)
  if (!p.dep_flags) ptx::griddep_wait();
  ptx::griddep_launch_dependents();
  ptx::stamp_begin(p.tstamp);             // after the grid-wide dependency (if any): the prologue above overlapped the previous kernel
  const int n_end = p.m_total / (p.ho * p.wo);       // one past the last crop of this call's slice

  if (warp == 0 || warp == 3) {
    // ================================ TMA producers ===============================
    // Two single-thread producers walk the same K-block sequence: warp 0 fetches the activation boxes
    // (A) and arms the stage's byte count, warp 3 fetches the weight boxes (B).  One thread can retire a
    // wait -> arm -> issue chain only every ~400 cycles (tools/ubench/tma_rate.cu: 396 cycles per box
    // whatever its size), so the two operand streams must not share a thread.
    if (ptx::elect_one()) {
      const bool is_a = warp == 0;
      int stage = 0;
      uint32_t phase = 0;
      long long t_wait = 0, t_issue = 0;
      // operands of both CTAs land on the LEADER's full barrier (its A producer alone arms the byte count)
      // (with kXform every CTA has its own full barrier: its transform warps wait on it locally)
      const uint32_t full0 = ptx::mapa(ptx::smem_u32(full), kXform ? rank : 0);
      int sub = 0, left = 0;                        // K block within the stage; K blocks of the tile still to load
      int id_blocks = 0;                            // identity K blocks at the end of the current tile
      bool dep_all_done = false;                    // dataflow: the producer layer has been seen complete
      auto acquire = [&]() -> unsigned char * {
        if (sub == 0) {
          if (prof) {
            const long long t0 = clock64();
            ptx::mbar_wait(empty + stage, phase ^ 1);
            t_wait += clock64() - t0;
          } else {
            ptx::mbar_wait(empty + stage, phase ^ 1);
          }
          if (is_a && (kXform || rank == 0)) {
            // identity-shortcut K blocks of a 256-wide tile carry a 64-channel slice of the identity only
            // (narrow_id): 16 KB of activations + 32 weight rows per CTA
            const bool narrow = kNarrowId && p.diag2 && left <= id_blocks;
            ptx::mbar_arrive_expect_tx(full + stage, narrow ? 2 * (C::kABytes + 32 * kTileK * 2)
                                                            : (kXform ? 1 : 2) * C::kBlockBytes * min(C::kSub, left));
          }
        }
        return tiles + stage * C::kStageBytes + sub * C::kBlockBytes;
      };
      auto advance = [&]() {
        --left;
        if (++sub == C::kSub || left == 0) { sub = 0; if (++stage == stages) { stage = 0; phase ^= 1; } }
      };
      for (int tile = pair; tile < n_tiles_total; tile += n_pairs) {
        const int te = p.reverse ? n_tiles_total - 1 - tile : tile;   // work-list direction alternates between layers
        const int mp = te / p.n_tiles, nt = te - mp * p.n_tiles;
        const int mt = 2 * mp + int(rank);          // this CTA's 128-pixel tile (may be one past the end: zero fill)
        int n0, h0;
        if (p.nb == 1) { n0 = mt / p.tiles_per_img; h0 = (mt - n0 * p.tiles_per_img) * p.th; }
        else { n0 = mt * p.nb; h0 = 0; }
        n0 += p.n_base;
        const int ncol = nt * BLOCK_N + int(rank) * (BLOCK_N / 2);   // this CTA's half of the weight rows
        const int cb2_0 = p.diag2 ? nt * (BLOCK_N / 64) : 0;
        const int n2 = p.diag2 ? min(BLOCK_N / 64, p.cblk1 - cb2_0) : p.cblk1;
        left = k0 + n2;
        id_blocks = p.diag2 ? n2 : 0;
        if (is_a && p.dep_flags && !dep_all_done) {
          // the crops this CTA's 128 pixels belong to must be complete in the producer layer (its inputs from
          // earlier layers are then complete too: every layer waited for the same crops of its own producer)
          if (ptx::flag_load(p.dep_done) >= p.dep_ctas) {
            dep_all_done = true;                     // the whole producer grid has exited: no more polling
          } else {
            const int c_end = min(n0 + p.nb, n_end);
            for (int c = n0; c < c_end; ++c) ptx::flag_wait(p.dep_flags + c, p.dep_expected);
          }
          ptx::fence_proxy_async_all();
        }
        if (p.tall) {
          // 3x3 stride-1 convolution, "tall" staging: a stage holds, for one kernel column kw and one channel
          // block, the (th + 2*rate) input rows that serve all three kernel rows (one column-shifted box; TMA
          // zero fill pads left/right/top/bottom), plus the three weight boxes of that column.  A traffic per
          // tile drops from 9 x th rows to 3 x (th + 2*rate) rows.
          for (int kw = 0; kw < 3; ++kw)
            for (int cb = 0; cb < p.cblk0; ++cb) {
              if (prof) {
                const long long t0 = clock64();
                ptx::mbar_wait(empty + stage, phase ^ 1);
                t_wait += clock64() - t0;
              } else {
                ptx::mbar_wait(empty + stage, phase ^ 1);
              }
              unsigned char *sa = tiles + stage * p.tall_stage_bytes;
              const uint32_t fb = full0 + uint32_t(stage) * 8u;
              if (is_a) {
                if (rank == 0) ptx::mbar_arrive_expect_tx(full + stage, 2 * p.tall_stage_bytes);
                ptx::tma_load_4d_pair(sa, &p.amap[1], fb, cb * kTileK, p.tap_dw[kw], h0 + p.tap_dh[0], n0);
              } else {
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
                  ptx::tma_load_2d_pair(sa + p.tall_a_bytes + kh * C::kBBytes, &p.bmap, fb,
                                        ((kh * 3 + kw) * p.cblk0 + cb) * kTileK, ncol);
              }
              if (++stage == stages) { stage = 0; phase ^= 1; }
            }
          continue;
        }
        if (is_a) {
          for (int tap = 0; tap < p.taps; ++tap) {
            const CUtensorMap *am = &p.amap[p.tap_map[tap]];
            const int dw = p.tap_dw[tap], hh = h0 + p.tap_dh[tap];
            for (int cb = 0; cb < p.cblk0; ++cb) {
              unsigned char *sa = acquire();
              const long long i0 = prof ? clock64() : 0;
              ptx::tma_load_4d_pair(sa, am, full0 + uint32_t(stage) * 8u, cb * kTileK, dw, hh, n0);
              advance();
              if (prof) t_issue += clock64() - i0;
            }
          }
          for (int cb = cb2_0; cb < cb2_0 + n2; ++cb) {
            unsigned char *sa = acquire();
            ptx::tma_load_4d_pair(sa, &p.a2map, full0 + uint32_t(stage) * 8u, cb * kTileK, 0, h0, n0);
            advance();
          }
        } else {
          // weights: K blocks 0 .. k0-1 of source 0, then blocks k0 + cb2_0 .. of the appended source-1 range
          for (int kb = 0; kb < k0 + n2; ++kb) {
            const int kcol = (kb < k0 ? kb : kb + cb2_0) * kTileK;
            unsigned char *sa = acquire();
            if (kNarrowId && p.diag2 && kb >= k0)   // rows 64 j + 32 rank .. of this N tile: the 64x64 identity block, halved
              ptx::tma_load_2d_pair(sa + C::kABytes, &p.bidmap, full0 + uint32_t(stage) * 8u, kcol,
                                    nt * BLOCK_N + 64 * (kb - k0) + 32 * int(rank));
            else
              ptx::tma_load_2d_pair(sa + C::kABytes, &p.bmap, full0 + uint32_t(stage) * 8u, kcol, ncol);
            advance();
          }
        }
      }
      if (prof && is_a) { p.prof[blockIdx.x * kPCount + kPProdWait] = t_wait; p.prof[blockIdx.x * kPCount + kPProdIssue] = t_issue; }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(2 * kTileM, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0, it = 0;
      long long t_full = 0, t_acc = 0, t_mi = 0, t_mc = 0;
      for (int tile = pair; tile < n_tiles_total; tile += n_pairs, ++it) {
        const int nt = (p.reverse ? n_tiles_total - 1 - tile : tile) % p.n_tiles;
        const int n_kb = k0 + (p.diag2 ? min(BLOCK_N / 64, p.cblk1 - nt * (BLOCK_N / 64)) : p.cblk1);
        const int acc = it & 1;
        {
          const long long t0 = prof ? clock64() : 0;
          ptx::mbar_wait(tempty + acc, ((it >> 1) & 1) ^ 1);
          if (prof) t_acc += clock64() - t0;
        }
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C::kAccCols;
        if (p.tall) {
          const int n_st = 3 * p.cblk0;
          for (int st = 0; st < n_st; ++st) {
            {
              const long long t0 = prof ? clock64() : 0;
              ptx::mbar_wait(full + stage, phase);
              if (prof) t_full += clock64() - t0;
            }
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(tiles + stage * p.tall_stage_bytes);
            const uint64_t da = ptx::make_sw128_kmajor_desc(sa);
            const uint64_t db = ptx::make_sw128_kmajor_desc(sa + p.tall_a_bytes);
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              // kernel row kh reads the same box `rate` image rows further down (a multiple of 1 KB: the
              // swizzle phase of the operand rows is unchanged)
              const uint32_t offa = uint32_t(kh * p.tall_row_step) >> 4, offb = uint32_t(kh * C::kBBytes) >> 4;
#pragma unroll
              for (int k = 0; k < kTileK / 16; ++k)
                ptx::umma_f16_pair(d_tmem, da + offa + 2 * k, db + offb + 2 * k, idesc, (st | kh | k) != 0);
            }
            ptx::umma_commit_pair(empty + stage, 3);
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
          ptx::umma_commit_pair(tfull + acc, 3);
          continue;
        }
        for (int kb = 0; kb < n_kb; kb += C::kSub) {
          {
            const long long t0 = prof ? clock64() : 0;
            ptx::mbar_wait((kXform ? ready : full) + stage, phase);
            if (prof) t_full += clock64() - t0;
          }
          ptx::tc_fence_after();
          const long long m0 = prof ? clock64() : 0;
          const uint32_t sa = ptx::smem_u32(tiles + stage * C::kStageBytes);
          const uint64_t da = ptx::make_sw128_kmajor_desc(sa);
          const uint64_t db = ptx::make_sw128_kmajor_desc(sa + C::kABytes);
          if (kNarrowId && p.diag2 && kb >= k0) {
            // identity block j adds the shortcut's channels 64 j .. 64 j + 63: an N = 64 MMA on that column slice
            constexpr uint32_t idesc64 = ptx::make_idesc_f16(2 * kTileM, 64);
#pragma unroll
            for (int k = 0; k < kTileK / 16; ++k)
              ptx::umma_f16_pair(d_tmem + 64 * (kb - k0), da + 2 * k, db + 2 * k, idesc64, 1);
          } else
#pragma unroll
          for (int sb = 0; sb < C::kSub; ++sb) {
            if (sb > 0 && kb + sb >= n_kb) break;    // odd K-block count: the last stage is half full
#pragma unroll
            for (int k = 0; k < kTileK / 16; ++k) {
              // advance 16 fp16 = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
              const uint32_t off = uint32_t(sb * (C::kBlockBytes >> 4) + 2 * k);
              ptx::umma_f16_pair(d_tmem, da + off, db + off, idesc, (kb | sb | k) != 0);
            }
          }
          const long long m1 = prof ? clock64() : 0;
          ptx::umma_commit_pair(empty + stage, 3);   // frees the smem slot in both CTAs when these MMAs retire
          if (++stage == stages) { stage = 0; phase ^= 1; }
          if (prof) { t_mi += m1 - m0; t_mc += clock64() - m1; }
        }
        ptx::umma_commit_pair(tfull + acc, 3);       // accumulator complete -> epilogue group `acc` of both CTAs
      }
      if (prof) {
        p.prof[blockIdx.x * kPCount + kPMmaWaitFull] = t_full;
        p.prof[blockIdx.x * kPCount + kPMmaWaitAcc] = t_acc;
        p.prof[blockIdx.x * kPCount + kPTiles] = it;
        p.prof[blockIdx.x * kPCount + kPMmaIssue] = t_mi;
        p.prof[blockIdx.x * kPCount + kPMmaCommit] = t_mc;
      }
    }
  } else if (kXform && warp >= kCtrlWarps + kEpiWarps) {
    // ======================= A-operand transform: x -> relu(x * scale + shift) =======================
    // thread = one 16-byte chunk (8 channels) of 8 rows of the 128 x 64 tile: the channel vectors stay in
    // registers for the whole K block, a warp touches 4 whole 128-byte rows per access (conflict-free in
    // the 128B-swizzled layout), values are rounded to fp16 exactly like the stored pre-activation was.
    const int xt = threadIdx.x - (kCtrlWarps + kEpiWarps) * 32;     // 0..127
    const int chunk = xt & 7, row0 = xt >> 3;
    const float *apar = reinterpret_cast<const float *>(smem + p.off_apar);   // [2][cin]
    for (int i = xt; i < 2 * p.cblk0 * kTileK; i += kXformWarps * 32) {
      const int c = i % (p.cblk0 * kTileK);
      const_cast<float *>(apar)[i] = i < p.cblk0 * kTileK ? p.ascale[c] : p.ashift[c];
    }
    ptx::named_bar_sync(3, kXformWarps * 32);
    const uint32_t ready0 = ptx::mapa(ptx::smem_u32(ready), 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < n_tiles_total; tile += n_pairs) {
      for (int kb = 0; kb < k0; kb += C::kSub) {
        ptx::mbar_wait(full + stage, phase);
#pragma unroll
        for (int sb = 0; sb < C::kSub; ++sb) {
          if (sb > 0 && kb + sb >= k0) break;
          const int cb = (kb + sb) % p.cblk0;
          const float *ps = apar + cb * kTileK + chunk * 8, *pf = ps + p.cblk0 * kTileK;
          const float4 s0 = *reinterpret_cast<const float4 *>(ps), s1 = *reinterpret_cast<const float4 *>(ps + 4);
          const float4 f0 = *reinterpret_cast<const float4 *>(pf), f1 = *reinterpret_cast<const float4 *>(pf + 4);
          const uint32_t base = ptx::smem_u32(tiles + stage * C::kStageBytes + sb * C::kBlockBytes);
#pragma unroll
          for (int i = 0; i < 128 / (kXformWarps * 4); ++i) {
            const int row = row0 + kXformWarps * 4 * i;
            const uint32_t a = base + uint32_t(row) * 128u + (uint32_t(chunk ^ (row & 7)) << 4);
            uint4 v = ptx::lds_v4u(a);
            uint32_t *w = reinterpret_cast<uint32_t *>(&v);
            const float2 x0 = __half22float2(*reinterpret_cast<const __half2 *>(&w[0]));
            const float2 x1 = __half22float2(*reinterpret_cast<const __half2 *>(&w[1]));
            const float2 x2 = __half22float2(*reinterpret_cast<const __half2 *>(&w[2]));
            const float2 x3 = __half22float2(*reinterpret_cast<const __half2 *>(&w[3]));
            const float2 y0 = __ffma2_rn(x0, make_float2(s0.x, s0.y), make_float2(f0.x, f0.y));
            const float2 y1 = __ffma2_rn(x1, make_float2(s0.z, s0.w), make_float2(f0.z, f0.w));
            const float2 y2 = __ffma2_rn(x2, make_float2(s1.x, s1.y), make_float2(f1.x, f1.y));
            const float2 y3 = __ffma2_rn(x3, make_float2(s1.z, s1.w), make_float2(f1.z, f1.w));
            w[0] = pack_relu_f16x2(y0.x, y0.y); w[1] = pack_relu_f16x2(y1.x, y1.y);
            w[2] = pack_relu_f16x2(y2.x, y2.y); w[3] = pack_relu_f16x2(y3.x, y3.y);
            sts_v4(a, v);
          }
        }
        ptx::fence_proxy_async();                    // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(ready0 + uint32_t(stage) * 8u);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= kCtrlWarps) {
    // ================================ epilogue ====================================
    // Group g = tiles with local index == g (mod 2) = accumulator stage g.  A warp owns the 32
    // accumulator rows of its TMEM lane quarter: it converts them 32 columns at a time, stages the
    // 32 x 32 fp16 box in its own 2 KB of shared memory (64-byte rows, 64B swizzle) and stores it
    // with its own TMA store, so the eight warps never wait for each other inside a tile.
    const int e = warp - kCtrlWarps, g = e >> 2, q = e & 3;   // q == warp % 4: the TMEM lane quarter
    const int gt = threadIdx.x - (kCtrlWarps + 4 * g) * 32;   // 0..127 within the group
    float *par = reinterpret_cast<float *>(smem + p.off_par) + g * kParVecs * BLOCK_N;
    const uint32_t par_a = ptx::smem_u32(par);
    const uint32_t taddr0 = tmem_base + (uint32_t(q * 32) << 16) + g * C::kAccCols;
    const int n_st = (p.has_out1 ? 1 : 0) + (p.has_out2 ? 1 : 0);
    const uint32_t st1 = ptx::smem_u32(smem + p.off_stage) + uint32_t(e * n_st) * 2048u;
    const uint32_t st2 = st1 + (p.has_out1 ? 2048u : 0u);
    const uint32_t row_a = uint32_t(lane) * 64u, sw = uint32_t(lane >> 1) & 3u;
    const uint32_t tempty0 = ptx::mapa(ptx::smem_u32(tempty + g), 0);   // the leader's barrier
    int cur_nt = -1;
    unsigned int *sig_prev = nullptr;                // counter of the tile whose stores are still in flight
    long long t_acc = 0, t_busy = 0, t_store = 0, t_ld = 0, t_math = 0, t_sts = 0, t_issue = 0, t_par = 0;
    uint32_t k = 0;
    for (int tile = pair + g * n_pairs; tile < n_tiles_total; tile += 2 * n_pairs, ++k) {
      const int te = p.reverse ? n_tiles_total - 1 - tile : tile;
      const int mp = te / p.n_tiles, nt = te - mp * p.n_tiles;
      const int m0 = p.m_base + (2 * mp + int(rank)) * kTileM + q * 32;
      const long long tp0 = prof ? clock64() : 0;
      if (nt != cur_nt) {                            // per-channel epilogue vectors of this N tile
        if (cur_nt >= 0) ptx::named_bar_sync(1 + g, 128);   // the group is done with the previous ones
        for (int i = gt; i < BLOCK_N; i += 128) {
          const int c = nt * BLOCK_N + i;
          if (kMode == kDual) {
            par[i] = p.shift[c]; par[BLOCK_N + i] = p.scale2[c]; par[2 * BLOCK_N + i] = p.shift2[c];
          } else {
            par[i] = p.scale[c]; par[BLOCK_N + i] = p.shift[c];
          }
        }
        ptx::named_bar_sync(1 + g, 128);
        cur_nt = nt;
      }
      long long t0 = prof ? clock64() : 0;
      t_par += t0 - tp0;
      ptx::mbar_wait(tfull + g, k & 1);
      ptx::tc_fence_after();
      long long t1 = prof ? clock64() : 0;
      t_acc += t1 - t0;

      if constexpr (kMode == kDirect) {
        // ---- direct-store path (logits head: few columns, masked tail; fp32 or fp16) ----
        const int m = m0 + lane;
        const bool valid = m < p.m_total;
#pragma unroll 1
        for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
          const int col0 = nt * BLOCK_N + chunk * 32;
          if (col0 >= p.cout) break;                 // uniform: padded head columns
          uint32_t v[32];
          __syncwarp();                              // tcgen05.ld is .sync.aligned: reconverge first
          ptx::tmem_ld_32x32(taddr0 + chunk * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = col0 + 8 * j;
            if (col >= p.cout) break;                // uniform (cout is a multiple of 8)
            const uint32_t pa = par_a + uint32_t(chunk * 32 + 8 * j) * 4u;
            const float4 s0 = ptx::lds_v4(pa), s1 = ptx::lds_v4(pa + 16);
            const float4 b0 = ptx::lds_v4(pa + BLOCK_N * 4), b1 = ptx::lds_v4(pa + BLOCK_N * 4 + 16);
            float f[8];
            f[0] = fmaf(__uint_as_float(v[8 * j + 0]), s0.x, b0.x); f[1] = fmaf(__uint_as_float(v[8 * j + 1]), s0.y, b0.y);
            f[2] = fmaf(__uint_as_float(v[8 * j + 2]), s0.z, b0.z); f[3] = fmaf(__uint_as_float(v[8 * j + 3]), s0.w, b0.w);
            f[4] = fmaf(__uint_as_float(v[8 * j + 4]), s1.x, b1.x); f[5] = fmaf(__uint_as_float(v[8 * j + 5]), s1.y, b1.y);
            f[6] = fmaf(__uint_as_float(v[8 * j + 6]), s1.z, b1.z); f[7] = fmaf(__uint_as_float(v[8 * j + 7]), s1.w, b1.w);
            if (p.relu1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            if (valid) {
              const size_t off = size_t(m) * p.cout + col;
              if (p.out1_f32) {
                float4 *dst = reinterpret_cast<float4 *>(static_cast<float *>(p.out1) + off);
                dst[0] = make_float4(f[0], f[1], f[2], f[3]);
                dst[1] = make_float4(f[4], f[5], f[6], f[7]);
              } else {
                uint4 o;
                o.x = pack_f16x2(f[0], f[1]); o.y = pack_f16x2(f[2], f[3]);
                o.z = pack_f16x2(f[4], f[5]); o.w = pack_f16x2(f[6], f[7]);
                *reinterpret_cast<uint4 *>(static_cast<__half *>(p.out1) + off) = o;
              }
            }
          }
        }
        ptx::tc_fence_before();
        if (p.sig_flags) __threadfence();            // this lane's stores, GPU scope, before the warp's report
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive_cluster(tempty0);
          if (p.sig_flags && m0 < p.m_total) ptx::flag_signal(p.sig_flags + m0 / (p.ho * p.wo));
        }
      } else {
        // ---- fp16 path: TMEM -> registers -> swizzled 32x32 box in smem -> TMA store.
        //      (The identity-shortcut residual is not an epilogue operand: it is accumulated by the
        //      tensor core as extra K blocks against identity weights, see the producer.) ----
        constexpr int kChunks = BLOCK_N / 32;
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
          uint32_t v[32];
          const long long c0 = prof ? clock64() : 0;
          __syncwarp();
          ptx::tmem_ld_32x32(taddr0 + c * 32, v);
          ptx::tmem_ld_wait();
          const long long c1 = prof ? clock64() : 0;
          if (c == kChunks - 1) {                    // accumulator drained: hand the stage back early
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(tempty0);
          }
          uint4 o1[4], o2[4];
          const uint32_t pa = par_a + uint32_t(c * 32) * 4u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t pj = pa + uint32_t(j) * 32u;
            uint32_t *w1 = reinterpret_cast<uint32_t *>(&o1[j]);
            uint32_t *w2 = reinterpret_cast<uint32_t *>(&o2[j]);
            if constexpr (kMode == kSingle) {
              const float4 s0 = ptx::lds_v4(pj), s1 = ptx::lds_v4(pj + 16);
              const float4 b0 = ptx::lds_v4(pj + BLOCK_N * 4), b1 = ptx::lds_v4(pj + BLOCK_N * 4 + 16);
              const float2 y0 = __ffma2_rn(make_float2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1])), make_float2(s0.x, s0.y), make_float2(b0.x, b0.y));
              const float2 y1 = __ffma2_rn(make_float2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])), make_float2(s0.z, s0.w), make_float2(b0.z, b0.w));
              const float2 y2 = __ffma2_rn(make_float2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])), make_float2(s1.x, s1.y), make_float2(b1.x, b1.y));
              const float2 y3 = __ffma2_rn(make_float2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])), make_float2(s1.z, s1.w), make_float2(b1.z, b1.w));
              if (p.relu1) {
                w1[0] = pack_relu_f16x2(y0.x, y0.y); w1[1] = pack_relu_f16x2(y1.x, y1.y);
                w1[2] = pack_relu_f16x2(y2.x, y2.y); w1[3] = pack_relu_f16x2(y3.x, y3.y);
              } else {
                w1[0] = pack_f16x2(y0.x, y0.y); w1[1] = pack_f16x2(y1.x, y1.y);
                w1[2] = pack_f16x2(y2.x, y2.y); w1[3] = pack_f16x2(y3.x, y3.y);
              }
            } else {
              // y = acc + bias (raw sum, stored as fp16); y2 = relu(fp16(y) * scale2 + shift2): the next
              // unit's pre-activation computed from the fp16 value its consumer would have read
              const float4 b0 = ptx::lds_v4(pj), b1 = ptx::lds_v4(pj + 16);
              const float4 s0 = ptx::lds_v4(pj + BLOCK_N * 4), s1 = ptx::lds_v4(pj + BLOCK_N * 4 + 16);
              const float4 f0 = ptx::lds_v4(pj + BLOCK_N * 8), f1 = ptx::lds_v4(pj + BLOCK_N * 8 + 16);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
              const float ff[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 y = __fadd2_rn(make_float2(__uint_as_float(v[8 * j + 2 * i]), __uint_as_float(v[8 * j + 2 * i + 1])),
                                            make_float2(bb[2 * i], bb[2 * i + 1]));
                w1[i] = pack_f16x2(y.x, y.y);
                const float2 yh = __half22float2(*reinterpret_cast<const __half2 *>(&w1[i]));
                const float2 z = __ffma2_rn(yh, make_float2(ss[2 * i], ss[2 * i + 1]), make_float2(ff[2 * i], ff[2 * i + 1]));
                w2[i] = pack_relu_f16x2(z.x, z.y);
              }
            }
          }
          // the previous store from this warp's staging buffers must have finished reading them
          const long long c2 = prof ? clock64() : 0;
          if (lane == 0) ptx::bulk_wait_read<0>();
          __syncwarp();
          const long long c3 = prof ? clock64() : 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t soff = row_a + ((uint32_t(j) ^ sw) << 4);
            if (kMode == kSingle || p.has_out1) sts_v4(st1 + soff, o1[j]);
            if (kMode == kDual) sts_v4(st2 + soff, o2[j]);
          }
          ptx::fence_proxy_async();                  // generic-proxy smem writes -> visible to TMA
          __syncwarp();
          const long long c4 = prof ? clock64() : 0;
          if (lane == 0) {
            const int col0 = nt * BLOCK_N + c * 32;
            if (kMode == kSingle || p.has_out1) ptx::tma_store_2d_a(&p.o1map, st1, col0, m0);
            if (kMode == kDual) ptx::tma_store_2d_a(&p.o2map, st2, col0, m0);
            ptx::bulk_commit();
          }
          if (prof) {
            t_ld += c1 - c0; t_math += c2 - c1; t_store += c3 - c2; t_sts += c4 - c3; t_issue += clock64() - c4;
          }
        }
        if (p.sig_flags && lane == 0) {
          // report the PREVIOUS tile's rows: its kChunks store groups are complete once at most the kChunks groups
          // of this tile are still pending (no wait in practice: they were issued a whole tile ago)
          if (sig_prev) {
            ptx::bulk_wait<kChunks>();
            ptx::fence_proxy_async_all();
            ptx::flag_signal(sig_prev);
          }
          sig_prev = m0 < p.m_total ? p.sig_flags + m0 / (p.ho * p.wo) : nullptr;
        }
      }
      if (prof) t_busy += clock64() - t1;
    }
    if (kMode != kDirect && lane == 0) {
      ptx::bulk_wait<0>();                                     // smem must outlive the last TMA store
      if (sig_prev) { ptx::fence_proxy_async_all(); ptx::flag_signal(sig_prev); }
    }
    if (prof && e == 0 && lane == 0) {
      p.prof[blockIdx.x * kPCount + kPEpiWaitAcc] = t_acc;
      p.prof[blockIdx.x * kPCount + kPEpiBusy] = t_busy;
      p.prof[blockIdx.x * kPCount + kPEpiWaitStore] = t_store;
      p.prof[blockIdx.x * kPCount + kPEpiLd] = t_ld;
      p.prof[blockIdx.x * kPCount + kPEpiMath] = t_math;
      p.prof[blockIdx.x * kPCount + kPEpiSts] = t_sts;
      p.prof[blockIdx.x * kPCount + kPEpiIssue] = t_issue;
      p.prof[blockIdx.x * kPCount + kPEpiPar] = t_par;
    }
  }

  ptx::tc_fence_before();
  __syncwarp();                                      // the single-lane roles rejoin their warps first
  ptx::cluster_sync();                               // nobody exits (or frees TMEM) while its peer is still working
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, C::kTmemCols);
  }
  if (p.sig_done && threadIdx.x == 0) {              // every warp of this CTA has passed its final store wait
    ptx::fence_proxy_async_all();
    ptx::flag_signal(p.sig_done);
  }
  if (prof && threadIdx.x == 0) p.prof[blockIdx.x * kPCount + kPTotal] = clock64() - t_start;
  ptx::stamp_end(p.tstamp);
}

// ---- driver entry point -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int grid_for(const ConvGemmParams &prm, int num_sms) {
  const int pair_tiles = ((prm.m_tiles + 1) / 2) * prm.n_tiles;
  const int max_pairs = num_sms / 2;
  return 2 * (pair_tiles < max_pairs ? pair_tiles : max_pairs);   // CTA pairs (cluster of 2)
}

template <int BLOCK_N, int kMode, bool kXform = false>
metro_status launch_t(const ConvGemmParams &prm, int num_sms, cudaStream_t stream) {
  // function attributes are per device: one opt-in per (instantiation, device), safe across threads
  static PerDeviceOnce configured;
  metro_status cst = configured.run([] {
    METRO_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, kMode, kXform>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemLimit));
    return METRO_OK;
  });
  if (cst != METRO_OK) return cst;
  const int pair_tiles = ((prm.m_tiles + 1) / 2) * prm.n_tiles;
  if (pair_tiles == 0) return METRO_OK;
  const int grid = grid_for(prm, num_sms);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(grid)); cfg.blockDim = dim3(kXform ? kThreadsXform : kThreads);
  cfg.dynamicSmemBytes = size_t(prm.smem_bytes); cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = getenv("METRO_NO_PDL") != nullptr;
  cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
  METRO_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BLOCK_N, kMode, kXform>, prm));
  return METRO_OK;
}

}  // namespace

int conv_gemm_grid(const ConvGemmParams &prm, int num_sms) { return grid_for(prm, num_sms); }

metro_status make_out_tensor_map(CUtensorMap *map, const void *base, long long m_rows, int cout) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(METRO_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t dims[2] = {cuuint64_t(cout), cuuint64_t(m_rows)};
  const cuuint64_t strides[1] = {cuuint64_t(cout) * 2};
  const cuuint32_t box[2] = {32, 32};   // one epilogue warp's box: 32 columns (64 B) x 32 rows
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(METRO_ERR_CUDA, "cuTensorMapEncodeTiled(output rows=%lld cout=%d) -> %d", m_rows, cout, int(r));
  return METRO_OK;
}

metro_status make_act_tensor_map(CUtensorMap *map, const void *base, int n, int h, int w, int c, int sub, int ph,
                                 int pw, int box_w, int box_h, int box_n) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(METRO_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  // view [C, W/sub, H/sub, N] of the NHWC tensor, starting at pixel (ph, pw)
  const cuuint64_t dims[4] = {cuuint64_t(c), cuuint64_t(w / sub), cuuint64_t(h / sub), cuuint64_t(n)};
  const cuuint64_t strides[3] = {cuuint64_t(sub) * c * 2, cuuint64_t(sub) * w * c * 2, cuuint64_t(h) * w * c * 2};
  const cuuint32_t box[4] = {cuuint32_t(kTileK), cuuint32_t(box_w), cuuint32_t(box_h), cuuint32_t(box_n)};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  void *addr = const_cast<unsigned char *>(static_cast<const unsigned char *>(base)) + (size_t(ph) * w + pw) * c * 2;
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, addr, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(METRO_ERR_CUDA, "cuTensorMapEncodeTiled(activation n=%d h=%d w=%d c=%d sub=%d box=%d,%d,%d) -> %d", n,
                h, w, c, sub, box_w, box_h, box_n, int(r));
  return METRO_OK;
}

metro_status make_weight_tensor_map(CUtensorMap *map, const void *base, int cout_pad, int k_total, int block_n) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(METRO_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t dims[2] = {cuuint64_t(k_total), cuuint64_t(cout_pad)};
  const cuuint64_t strides[1] = {cuuint64_t(k_total) * 2};
  const cuuint32_t box[2] = {cuuint32_t(kTileK), cuuint32_t(block_n / 2)};   // one CTA's half of the rows
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(METRO_ERR_CUDA, "cuTensorMapEncodeTiled(weights cout_pad=%d k=%d block_n=%d) -> %d", cout_pad, k_total,
                block_n, int(r));
  return METRO_OK;
}

int conv_gemm_pick_block_n(int cout, bool direct, long long m_rows) {
  if (direct) return cout <= 160 ? 160 : 256;      // logits head: 136 / 152 channels padded to 160
  if (cout <= 64) return 64;
  if (cout <= 128) return 128;
  // 256-wide tiles always: 128-wide tiles quantise better over the 74 CTA pairs when a layer has few tiles
  // (block3: 3.46 waves), but measured 10-20 % slower there (twice the activation traffic per FLOP)
  (void)m_rows;
  return 256;
}

metro_status conv_gemm_plan_smem(ConvGemmLaunch &L, int k_blocks) {
  (void)k_blocks;
  ConvGemmParams &p = L.prm;
  const int k_sub = L.block_n <= 160 ? 2 : 1;      // Cfg::kSub
  const int stage_bytes = p.tall ? p.tall_stage_bytes : k_sub * (kTileM * kTileK * 2 + (L.block_n / 2) * kTileK * 2);
  const int n_out = L.direct ? 0 : (p.has_out1 ? 1 : 0) + (p.has_out2 ? 1 : 0);
  const int par_bytes = 2 * (p.has_out2 ? 3 : 2) * L.block_n * 4;     // two epilogue groups
  const int stage_out = kEpilogueWarps * n_out * kWarpStageBytes;
  const int apar_bytes = p.ascale ? 2 * p.cblk0 * kTileK * 4 : 0;      // pre-activation vectors of the A transform
  int stages = (kSmemLimit - 256 - par_bytes - apar_bytes - stage_out) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return fail(METRO_ERR_INTERNAL, "conv_gemm: shared memory plan leaves %d stages", stages);
  p.stages = stages;
  int off = stages * stage_bytes;
  p.off_stage = off; off += stage_out;
  p.off_par = off; off += par_bytes;
  p.off_apar = off; off += apar_bytes;
  p.off_bar = off; off += 256;
  p.smem_bytes = off;
  return METRO_OK;
}

int conv_gemm_cout_pad(int cout, int block_n) { return (cout + block_n - 1) / block_n * block_n; }

metro_status conv_gemm_geometry(ConvGemmParams &p, int out_side) {
  const int wo = out_side, ho = out_side;
  if (wo < 8 || wo > kTileM || (wo & (wo - 1)) != 0)
    return fail(METRO_ERR_VALUE, "conv_gemm: output side %d must be a power of two in [8,128]", wo);
  p.wo = wo; p.ho = ho;
  int th = kTileM / wo;
  if (th > ho) th = ho;
  p.th = th;
  p.nb = kTileM / (wo * th);
  p.tiles_per_img = p.nb == 1 ? (ho / th) : 1;
  return METRO_OK;
}

metro_status conv_gemm_set_batch(ConvGemmParams &p, int n, int n_base) {
  const int rows = n * p.ho * p.wo;
  p.n_base = n_base;
  p.m_base = n_base * p.ho * p.wo;
  p.m_total = p.m_base + rows;                     // end of the slice (absolute row)
  p.m_tiles = (rows + kTileM - 1) / kTileM;
  return METRO_OK;
}

metro_status conv_gemm_set_taps(ConvGemmParams &p, int k, int stride, int rate, int pad_lo) {
  if (k * k > kMaxTaps) return fail(METRO_ERR_VALUE, "conv_gemm: kernel size %d not supported", k);
  if (stride != 1 && stride != 2) return fail(METRO_ERR_VALUE, "conv_gemm: stride %d not supported", stride);
  p.taps = k * k;
  for (int kh = 0; kh < k; ++kh)
    for (int kw = 0; kw < k; ++kw) {
      const int t = kh * k + kw;
      const int oh = kh * rate - pad_lo, ow = kw * rate - pad_lo;   // input offset relative to out*stride
      if (stride == 1) {
        p.tap_map[t] = 0; p.tap_dh[t] = (signed char)oh; p.tap_dw[t] = (signed char)ow;
      } else {
        const int ph = ((oh % 2) + 2) % 2, pw = ((ow % 2) + 2) % 2;
        p.tap_map[t] = (signed char)(ph * 2 + pw);
        p.tap_dh[t] = (signed char)((oh - ph) / 2);
        p.tap_dw[t] = (signed char)((ow - pw) / 2);
      }
    }
  return METRO_OK;
}

void conv_gemm_pack_weights(const float *w, int k, int cin, int cout, const float *w2, int cin2, int cout_pad,
                            __half *dst) {
  // w2 == nullptr with cin2 > 0: identity block (the shortcut is the raw input itself)
  const size_t K = size_t(k) * k * cin + cin2;
  std::memset(dst, 0, size_t(cout_pad) * K * sizeof(__half));
  for (int t = 0; t < k * k; ++t)
    for (int c = 0; c < cin; ++c) {
      const float *src = w + (size_t(t) * cin + c) * cout;
      for (int o = 0; o < cout; ++o) dst[size_t(o) * K + size_t(t) * cin + c] = __float2half_rn(src[o]);
    }
  for (int c = 0; c < cin2; ++c) {
    if (!w2) {
      if (c < cout) dst[size_t(c) * K + size_t(k) * k * cin + c] = __float2half_rn(1.0f);
      continue;
    }
    const float *src = w2 + size_t(c) * cout;
    for (int o = 0; o < cout; ++o) dst[size_t(o) * K + size_t(k) * k * cin + c] = __float2half_rn(src[o]);
  }
}

metro_status conv_gemm_launch(const ConvGemmLaunch &L, const ConvGemmParams &prm, int num_sms, cudaStream_t stream) {
  if (L.direct) {
    switch (L.block_n) {
      case 160: return launch_t<160, kDirect>(prm, num_sms, stream);
      case 256: return launch_t<256, kDirect>(prm, num_sms, stream);
    }
  } else if (prm.has_out2) {
    switch (L.block_n) {
      case 64: return launch_t<64, kDual>(prm, num_sms, stream);
      case 128: return launch_t<128, kDual>(prm, num_sms, stream);
      case 256: return launch_t<256, kDual>(prm, num_sms, stream);
    }
  } else if (prm.ascale) {
    switch (L.block_n) {
      case 64: return launch_t<64, kSingle, true>(prm, num_sms, stream);
      case 128: return launch_t<128, kSingle, true>(prm, num_sms, stream);
      case 256: return launch_t<256, kSingle, true>(prm, num_sms, stream);
    }
  } else {
    switch (L.block_n) {
      case 64: return launch_t<64, kSingle>(prm, num_sms, stream);
      case 128: return launch_t<128, kSingle>(prm, num_sms, stream);
      case 256: return launch_t<256, kSingle>(prm, num_sms, stream);
    }
  }
  return fail(METRO_ERR_INTERNAL, "conv_gemm: unsupported BLOCK_N %d", L.block_n);
}

}  // namespace metro
