// Two consecutive 1x1 convolutions across a unit boundary in ONE kernel (sm_100a):
//
//   unit u   conv3 (+ shortcut) + bias                          -> raw sum   (resnet_v2.py:120-125,134-138)   [stored]
//   unit u+1 pre-activation relu(bn(raw))                                    (resnet_v2.py:119)               [on chip only]
//   unit u+1 conv1 1x1 -> BN -> ReLU                            -> r1        (resnet_v2.py:127-128)           [stored]
//
// Unfused, the raw sum (the widest tensor of a unit: 4 x the bottleneck width) is written by conv3 and read again by
// conv1 -- for the 64x64 / 32x32 / 16x16 layers of config B that read alone is 2 / 1 / 0.5 MB per crop and unit, at the
// HBM floor, plus a launch of its own whose tiles quantise over the 74 CTA pairs.  Here a CTA pair owns 256 pixels and
// walks the conv3 output channels in tiles of 128 (N = 128 MMAs run at the full tensor rate): each tile's accumulator
// is drained once -- raw sum to HBM by TMA (the next unit's identity shortcut needs it), pre-activation to shared
// memory in the K-major 128B-swizzled operand layout -- and immediately multiplied into the conv1 accumulator, which
// stays in tensor memory for the whole pixel tile.  Tensor memory: 2 x 128 columns (conv3, double-buffered against
// its epilogue) + up to 256 columns (conv1) = 512.
//
// Arithmetic and rounding points are those of the separate kernels (conv_gemm.cu: fp16 raw sum, pre-activation
// computed from the ROUNDED raw sum in fp32, fp16 operands, fp32 accumulation in the same K order), so the result is
// bit-identical to the unfused path (tests/test_net_gpu.py).
//
// Roles (384 threads, as conv_gemm.cu): warp 0 activation producer, warp 3 weight producer, warp 1 MMA issuer, warp 2
// TMEM allocator, warps 4-11 epilogue in two groups (group g drains the conv3 tiles nt == g mod 2 and half of the
// conv1 columns).  One ring of 32 KB stages carries, in issue order, conv3 K blocks (activations + 64 weight rows),
// identity-shortcut K blocks (activations + 32 rows of an identity block) and conv1 K blocks (weights only: their A
// operand is the pre-activation buffer).
#include <cstdlib>

#include "conv_gemm.h"
#include "ptx.cuh"

namespace metro {

namespace {

constexpr int kBN3 = 128;                     // conv3 output channels per tile
constexpr int kCtrlWarps = 4;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kCtrlWarps + kEpiWarps) * 32;
constexpr int kSmemLimit = 232448;
constexpr int kABytes = kTileM * kTileK * 2;  // 16 KB: 128 pixel rows of one 64-channel K block
constexpr int kStageBytes = kABytes + 8192;   // + 64 weight rows; a conv1 K block (weights only, <= 16 KB) starts at offset 0.
                                              // Slots are as small as the blocks allow: the kernel is bound by the bytes it
                                              // keeps in flight (role timers: both producers wait for slots while the MMA
                                              // thread waits for data), so the ring should be all payload
constexpr int kA2Bytes = 2 * kABytes;         // one 128-channel tile of the pre-activation = two K blocks of conv1

// barrier slots (uint64): full[8] empty[8] tfull3[2] tempty3[2] a2full[2] a2empty[2] tfull1 tempty1
constexpr int kBFull = 0, kBEmpty = kMaxStages, kBTFull3 = 2 * kMaxStages, kBTEmpty3 = kBTFull3 + 2, kBA2Full = kBTEmpty3 + 2,
              kBA2Empty = kBA2Full + 2, kBTFull1 = kBA2Empty + 2, kBTEmpty1 = kBTFull1 + 1, kBCount = kBTEmpty1 + 1;

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

template <int BN1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) conv_chain_kernel(const __grid_constant__ ConvGemmParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("metro: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  unsigned char *tiles = smem;
  unsigned char *a2buf = smem + p.off_a2;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
  uint64_t *full = bars + kBFull, *empty = bars + kBEmpty, *tfull3 = bars + kBTFull3, *tempty3 = bars + kBTEmpty3;
  uint64_t *a2full = bars + kBA2Full, *a2empty = bars + kBA2Empty, *tfull1 = bars + kBTFull1, *tempty1 = bars + kBTEmpty1;
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bars + kBCount);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_items = (p.m_tiles + 1) >> 1;          // pixel tiles of 256 rows
  const int stages = p.stages;
  const int k0 = p.cblk0;                            // K blocks of conv3's own input
  const int n2 = p.diag2 ? kBN3 / 64 : p.cblk1;      // shortcut K blocks per tile: identity slices, or the projection's input
  const int n1 = k0 + n2;
  const int NT3 = p.n_tiles;                         // conv3 tiles per pixel tile (even)
  const int C3 = NT3 * kBN3;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.amap[0]);
    ptx::prefetch_tensormap(&p.bmap);
    ptx::prefetch_tensormap(&p.w1map);
    if (p.cblk1) ptx::prefetch_tensormap(&p.a2map);
    if (p.has_out1) ptx::prefetch_tensormap(&p.o1map);
    ptx::prefetch_tensormap(&p.o2map);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < stages; ++i) { ptx::mbar_init(full + i, 1); ptx::mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tfull3 + i, 1); ptx::mbar_init(tempty3 + i, kEpiWarps);        // leader's: 4 warps of the group x 2 CTAs
      ptx::mbar_init(a2full + i, kEpiWarps); ptx::mbar_init(a2empty + i, 1);
    }
    ptx::mbar_init(tfull1, 1); ptx::mbar_init(tempty1, 2 * kEpiWarps);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_pair(s_tmem, 512);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncwarp();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (!p.dep_flags) ptx::griddep_wait();
  ptx::griddep_launch_dependents();
  ptx::stamp_begin(p.tstamp);             // after the grid-wide dependency (if any): the prologue above overlapped the previous kernel
  const int n_end = p.m_total / (p.ho * p.wo);
  // role timers (METRO_ROLE_PROF, same slots as conv_gemm.cu; 12 = epilogue waits for the operand buffer, 13 = MMA
  // waits for the pre-activation, 14 = MMA waits for the conv1 accumulator, 15 = weight producer waits for a slot)
  const bool prof = p.prof != nullptr;
  const long long t_start = prof ? clock64() : 0;
  long long *pr = p.prof + size_t(blockIdx.x) * 16;

  if (warp == 0 || warp == 3) {
    // ================================ TMA producers ===============================
    if (ptx::elect_one()) {
      const bool is_a = warp == 0;
      int stage = 0;
      uint32_t phase = 0;
      bool dep_all_done = false;
      long long t_wait = 0;
      const uint32_t full0 = ptx::mapa(ptx::smem_u32(full), 0);     // operands of both CTAs land on the leader's barrier
      // waits for the slot; the leader's activation producer arms the byte count of BOTH CTAs' boxes
      auto acquire = [&](uint32_t bytes_per_cta) -> unsigned char * {
        const long long t0 = prof ? clock64() : 0;
        ptx::mbar_wait(empty + stage, phase ^ 1);
        if (prof) t_wait += clock64() - t0;
        if (is_a && rank == 0) ptx::mbar_arrive_expect_tx(full + stage, 2 * bytes_per_cta);
        return tiles + stage * kStageBytes;
      };
      auto advance = [&]() { if (++stage == stages) { stage = 0; phase ^= 1; } };
      auto conv1_blocks = [&](int nt, int) {          // weights of conv1 for K range [128 nt, 128 nt + 128)
        for (int kb2 = 0; kb2 < 2; ++kb2) {
          unsigned char *sa = acquire((BN1 / 2) * kTileK * 2);
          if (!is_a) ptx::tma_load_2d_pair(sa, &p.w1map, full0 + uint32_t(stage) * 8u, (nt * 2 + kb2) * kTileK, int(rank) * (BN1 / 2));
          advance();
        }
      };
      for (int item = pair; item < n_items; item += n_pairs) {
        const int mp = p.reverse ? n_items - 1 - item : item;
        const int mt = 2 * mp + int(rank);
        int n0, h0;
        if (p.nb == 1) { n0 = mt / p.tiles_per_img; h0 = (mt - n0 * p.tiles_per_img) * p.th; }
        else { n0 = mt * p.nb; h0 = 0; }
        n0 += p.n_base;
        if (is_a && p.dep_flags && !dep_all_done) {
          if (ptx::flag_load(p.dep_done) >= p.dep_ctas) {
            dep_all_done = true;
          } else {
            const int c_end = min(n0 + p.nb, n_end);
            for (int c = n0; c < c_end; ++c) ptx::flag_wait(p.dep_flags + c, p.dep_expected);
          }
          ptx::fence_proxy_async_all();
        }
        for (int nt = 0; nt < NT3; ++nt) {
          const int ncol = nt * kBN3 + int(rank) * (kBN3 / 2);
          for (int kb = 0; kb < k0; ++kb) {
            unsigned char *sa = acquire(kABytes + (kBN3 / 2) * kTileK * 2);
            const uint32_t fb = full0 + uint32_t(stage) * 8u;
            if (is_a) ptx::tma_load_4d_pair(sa, &p.amap[0], fb, kb * kTileK, 0, h0, n0);
            else ptx::tma_load_2d_pair(sa + kABytes, &p.bmap, fb, kb * kTileK, ncol);
            advance();
          }
          for (int j = 0; j < n2; ++j) {
            if (p.diag2) {
              // identity slice j of this tile: the shortcut's channels 128 nt + 64 j .. + 63 against a 64 x 64 identity
              unsigned char *sa = acquire(kABytes + 32 * kTileK * 2);
              const uint32_t fb = full0 + uint32_t(stage) * 8u;
              const int cb = nt * (kBN3 / 64) + j;
              if (is_a) ptx::tma_load_4d_pair(sa, &p.a2map, fb, cb * kTileK, 0, h0, n0);
              else ptx::tma_load_2d_pair(sa + kABytes, &p.bidmap, fb, (k0 + cb) * kTileK, nt * kBN3 + 64 * j + 32 * int(rank));
            } else {
              unsigned char *sa = acquire(kABytes + (kBN3 / 2) * kTileK * 2);
              const uint32_t fb = full0 + uint32_t(stage) * 8u;
              if (is_a) ptx::tma_load_4d_pair(sa, &p.a2map, fb, j * kTileK, 0, h0, n0);
              else ptx::tma_load_2d_pair(sa + kABytes, &p.bmap, fb, (k0 + j) * kTileK, ncol);
            }
            advance();
          }
          if (nt >= 1) conv1_blocks(nt - 1, 0);
        }
        conv1_blocks(NT3 - 1, 0);
      }
      if (prof) pr[is_a ? 1 : 15] = t_wait;
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc3 = ptx::make_idesc_f16(2 * kTileM, kBN3);
      constexpr uint32_t idesc_id = ptx::make_idesc_f16(2 * kTileM, 64);
      constexpr uint32_t idesc1 = ptx::make_idesc_f16(2 * kTileM, BN1);
      const uint32_t tm1 = tmem_base + 2 * kBN3;
      int stage = 0;
      uint32_t phase = 0, q = 0, it = 0;             // q: conv3 tiles issued so far (all items)
      long long t_full = 0, t_acc = 0, t_a2 = 0, t_acc1 = 0;
      auto timed_wait = [&](uint64_t *bar, uint32_t parity, long long &acc) {
        const long long t0 = prof ? clock64() : 0;
        ptx::mbar_wait(bar, parity);
        if (prof) acc += clock64() - t0;
      };
      auto next_stage = [&]() { if (++stage == stages) { stage = 0; phase ^= 1; } };
      // conv1 partial product over the pre-activation of conv3 tile j (global index qj)
      auto conv1_part = [&](int j, uint32_t qj) {
        const int b = j & 1;
        timed_wait(a2full + b, (qj >> 1) & 1, t_a2);
        if (j == 0) timed_wait(tempty1, (it & 1) ^ 1, t_acc1);      // the previous pixel tile's conv1 accumulator is drained
        ptx::tc_fence_after();
        for (int kb2 = 0; kb2 < 2; ++kb2) {
          timed_wait(full + stage, phase, t_full);
          ptx::tc_fence_after();
          const uint64_t da = ptx::make_sw128_kmajor_desc(ptx::smem_u32(a2buf + b * kA2Bytes + kb2 * kABytes));
          const uint64_t db = ptx::make_sw128_kmajor_desc(ptx::smem_u32(tiles + stage * kStageBytes));
#pragma unroll
          for (int k = 0; k < kTileK / 16; ++k) ptx::umma_f16_pair(tm1, da + 2 * k, db + 2 * k, idesc1, (j | kb2 | k) != 0);
          ptx::umma_commit_pair(empty + stage, 3);
          next_stage();
        }
        ptx::umma_commit_pair(a2empty + b, 3);       // the pre-activation buffer may be overwritten (both CTAs)
      };
      for (int item = pair; item < n_items; item += n_pairs, ++it) {
        for (int nt = 0; nt < NT3; ++nt, ++q) {
          const int b = nt & 1;
          timed_wait(tempty3 + b, ((q >> 1) & 1) ^ 1, t_acc);
          ptx::tc_fence_after();
          const uint32_t d3 = tmem_base + b * kBN3;
          for (int kb = 0; kb < n1; ++kb) {
            timed_wait(full + stage, phase, t_full);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(tiles + stage * kStageBytes);
            const uint64_t da = ptx::make_sw128_kmajor_desc(sa);
            const uint64_t db = ptx::make_sw128_kmajor_desc(sa + kABytes);
            if (p.diag2 && kb >= k0) {
#pragma unroll
              for (int k = 0; k < kTileK / 16; ++k) ptx::umma_f16_pair(d3 + 64 * (kb - k0), da + 2 * k, db + 2 * k, idesc_id, 1);
            } else {
#pragma unroll
              for (int k = 0; k < kTileK / 16; ++k) ptx::umma_f16_pair(d3, da + 2 * k, db + 2 * k, idesc3, (kb | k) != 0);
            }
            ptx::umma_commit_pair(empty + stage, 3);
            next_stage();
          }
          ptx::umma_commit_pair(tfull3 + b, 3);
          if (nt >= 1) conv1_part(nt - 1, q - 1);
        }
        conv1_part(NT3 - 1, q - 1);
        ptx::umma_commit_pair(tfull1, 3);
      }
      if (prof) { pr[2] = t_full; pr[3] = t_acc; pr[13] = t_a2; pr[14] = t_acc1; pr[7] = it; }
    }
  } else if (warp >= kCtrlWarps) {
    // ================================ epilogue ====================================
    const int e = warp - kCtrlWarps, g = e >> 2, qw = e & 3;        // qw == warp % 4: the TMEM lane quarter
    const int et = threadIdx.x - kCtrlWarps * 32;                    // 0..255
    float *par = reinterpret_cast<float *>(smem + p.off_par);        // bias3[C3] scale2[C3] shift2[C3] scale1[BN1] shift1[BN1]
    for (int i = et; i < C3; i += kEpiWarps * 32) {
      par[i] = p.shift[i]; par[C3 + i] = p.scale2[i]; par[2 * C3 + i] = p.shift2[i];
    }
    for (int i = et; i < BN1; i += kEpiWarps * 32) { par[3 * C3 + i] = p.scale1c[i]; par[3 * C3 + BN1 + i] = p.shift1c[i]; }
    ptx::named_bar_sync(1, kEpiWarps * 32);
    const uint32_t par_a = ptx::smem_u32(par);
    const uint32_t taddr3 = tmem_base + (uint32_t(qw * 32) << 16) + g * kBN3;
    const uint32_t taddr1 = tmem_base + (uint32_t(qw * 32) << 16) + 2 * kBN3 + g * (BN1 / 2);
    const uint32_t st = ptx::smem_u32(smem + p.off_stage) + uint32_t(e) * 2048u;   // this warp's 32 x 32 fp16 staging box
    const uint32_t row_a = uint32_t(lane) * 64u, sw = uint32_t(lane >> 1) & 3u;
    const int row = qw * 32 + lane;                                   // this thread's pixel row of the CTA's 128
    const uint32_t a2row = ptx::smem_u32(a2buf + g * kA2Bytes) + uint32_t(row) * 128u;
    const uint32_t tempty3_0 = ptx::mapa(ptx::smem_u32(tempty3 + g), 0);
    const uint32_t a2full_0 = ptx::mapa(ptx::smem_u32(a2full + g), 0);
    const uint32_t tempty1_0 = ptx::mapa(ptx::smem_u32(tempty1), 0);
    uint32_t it = 0;
    long long t_wacc = 0, t_busy = 0, t_wa2 = 0, t_wst = 0;
    for (int item = pair; item < n_items; item += n_pairs, ++it) {
      const int mp = p.reverse ? n_items - 1 - item : item;
      const int m0 = p.m_base + (2 * mp + int(rank)) * kTileM + qw * 32;
      for (int nt = g; nt < NT3; nt += 2) {
        const uint32_t qg = it * uint32_t(NT3) + uint32_t(nt);
        const long long w0 = prof ? clock64() : 0;
        ptx::mbar_wait(tfull3 + g, (qg >> 1) & 1);
        const long long w1 = prof ? clock64() : 0;
        t_wacc += w1 - w0;
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < kBN3 / 32; ++c) {
          uint32_t v[32];
          __syncwarp();
          ptx::tmem_ld_32x32(taddr3 + c * 32, v);
          ptx::tmem_ld_wait();
          if (c == kBN3 / 32 - 1) {                  // accumulator drained: hand it back before the stores
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(tempty3_0);
          }
          uint4 o1[4], o2[4];
          const uint32_t pa = par_a + uint32_t(nt * kBN3 + c * 32) * 4u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t pj = pa + uint32_t(j) * 32u;
            uint32_t *w1 = reinterpret_cast<uint32_t *>(&o1[j]);
            uint32_t *w2 = reinterpret_cast<uint32_t *>(&o2[j]);
            // y = acc + bias (raw sum, fp16); z = relu(fp16(y) * scale2 + shift2): the next unit's pre-activation
            const float4 b0 = ptx::lds_v4(pj), b1 = ptx::lds_v4(pj + 16);
            const float4 s0 = ptx::lds_v4(pj + C3 * 4), s1 = ptx::lds_v4(pj + C3 * 4 + 16);
            const float4 f0 = ptx::lds_v4(pj + C3 * 8), f1 = ptx::lds_v4(pj + C3 * 8 + 16);
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
          if (p.has_out1) {
            const long long s0 = prof ? clock64() : 0;
            if (lane == 0) ptx::bulk_wait_read<0>();   // the previous store from this staging box has been read
            __syncwarp();
            if (prof) t_wst += clock64() - s0;
#pragma unroll
            for (int j = 0; j < 4; ++j) sts_v4(st + row_a + ((uint32_t(j) ^ sw) << 4), o1[j]);
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_2d_a(&p.o1map, st, nt * kBN3 + c * 32, m0);
              ptx::bulk_commit();
            }
          }
          // the pre-activation goes to the operand buffer of conv1: K block c / 2, 16-byte chunks 4 (c % 2) + j of the row
          if (c == 0) {
            const long long s0 = prof ? clock64() : 0;
            ptx::mbar_wait(a2empty + g, ((qg >> 1) & 1) ^ 1);
            if (prof) t_wa2 += clock64() - s0;
          }
          const uint32_t a2 = a2row + uint32_t(c >> 1) * uint32_t(kABytes);
#pragma unroll
          for (int j = 0; j < 4; ++j) sts_v4(a2 + ((uint32_t((c & 1) * 4 + j) ^ uint32_t(row & 7)) << 4), o2[j]);
        }
        ptx::fence_proxy_async();                    // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(a2full_0);
        if (prof) t_busy += clock64() - w1;
      }
      // ---- conv1 of the next unit: BN + ReLU, this group's half of the columns ----
      const long long w2 = prof ? clock64() : 0;
      ptx::mbar_wait(tfull1, it & 1);
      const long long w3 = prof ? clock64() : 0;
      t_wacc += w3 - w2;
      ptx::tc_fence_after();
      constexpr int kChunks1 = BN1 / 2 / 32;
#pragma unroll 1
      for (int c = 0; c < kChunks1; ++c) {
        uint32_t v[32];
        __syncwarp();
        ptx::tmem_ld_32x32(taddr1 + c * 32, v);
        ptx::tmem_ld_wait();
        if (c == kChunks1 - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(tempty1_0);
        }
        const int col0 = g * (BN1 / 2) + c * 32;
        const uint32_t pa = par_a + uint32_t(3 * C3 + col0) * 4u;
        uint4 o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t pj = pa + uint32_t(j) * 32u;
          uint32_t *w = reinterpret_cast<uint32_t *>(&o[j]);
          const float4 s0 = ptx::lds_v4(pj), s1 = ptx::lds_v4(pj + 16);
          const float4 b0 = ptx::lds_v4(pj + BN1 * 4), b1 = ptx::lds_v4(pj + BN1 * 4 + 16);
          const float2 y0 = __ffma2_rn(make_float2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1])), make_float2(s0.x, s0.y), make_float2(b0.x, b0.y));
          const float2 y1 = __ffma2_rn(make_float2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])), make_float2(s0.z, s0.w), make_float2(b0.z, b0.w));
          const float2 y2 = __ffma2_rn(make_float2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])), make_float2(s1.x, s1.y), make_float2(b1.x, b1.y));
          const float2 y3 = __ffma2_rn(make_float2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])), make_float2(s1.z, s1.w), make_float2(b1.z, b1.w));
          w[0] = pack_relu_f16x2(y0.x, y0.y); w[1] = pack_relu_f16x2(y1.x, y1.y);
          w[2] = pack_relu_f16x2(y2.x, y2.y); w[3] = pack_relu_f16x2(y3.x, y3.y);
        }
        if (lane == 0) ptx::bulk_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) sts_v4(st + row_a + ((uint32_t(j) ^ sw) << 4), o[j]);
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_2d_a(&p.o2map, st, col0, m0);
          ptx::bulk_commit();
        }
      }
      if (p.sig_flags && lane == 0) {
        // this warp's 32 rows of both outputs are stored: report them (dataflow counters, ptx.cuh)
        ptx::bulk_wait<0>();
        ptx::fence_proxy_async_all();
        if (m0 < p.m_total) ptx::flag_signal(p.sig_flags + m0 / (p.ho * p.wo));
      }
      if (prof) t_busy += clock64() - w3;
    }
    if (lane == 0) ptx::bulk_wait<0>();              // shared memory must outlive the last TMA store
    if (prof && e == 0 && lane == 0) { pr[4] = t_wacc; pr[5] = t_busy; pr[6] = t_wst; pr[12] = t_wa2; }
  }

  ptx::tc_fence_before();
  __syncwarp();
  ptx::cluster_sync();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 512);
  }
  if (p.sig_done && threadIdx.x == 0) {
    ptx::fence_proxy_async_all();
    ptx::flag_signal(p.sig_done);
  }
  if (prof && threadIdx.x == 0) pr[0] = clock64() - t_start;
  ptx::stamp_end(p.tstamp);
}

template <int BN1>
metro_status launch_chain_t(const ConvGemmParams &prm, int num_sms, cudaStream_t stream) {
  static PerDeviceOnce configured;
  metro_status cst = configured.run([] {
    METRO_CUDA(cudaFuncSetAttribute(conv_chain_kernel<BN1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    return METRO_OK;
  });
  if (cst != METRO_OK) return cst;
  const int grid = conv_chain_grid(prm, num_sms);
  if (grid == 0) return METRO_OK;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(grid)); cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = size_t(prm.smem_bytes); cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = getenv("METRO_NO_PDL") != nullptr;
  cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
  METRO_CUDA(cudaLaunchKernelEx(&cfg, conv_chain_kernel<BN1>, prm));
  return METRO_OK;
}

}  // namespace

int conv_chain_grid(const ConvGemmParams &prm, int num_sms) {
  const int items = (prm.m_tiles + 1) / 2;
  const int max_pairs = num_sms / 2;
  return 2 * (items < max_pairs ? items : max_pairs);
}

metro_status conv_chain_plan_smem(ConvGemmParams &p) {
  const int c3 = p.n_tiles * kBN3;
  const int par_bytes = (3 * c3 + 2 * p.cout1) * 4;
  const int stage_out = kEpiWarps * kWarpStageBytes;
  int stages = (kSmemLimit - 256 - par_bytes - stage_out - 2 * kA2Bytes) / kStageBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 3) return fail(METRO_ERR_INTERNAL, "conv_chain: shared memory plan leaves %d stages", stages);
  p.stages = stages;
  int off = stages * kStageBytes;
  p.off_a2 = off; off += 2 * kA2Bytes;
  p.off_stage = off; off += stage_out;
  p.off_par = off; off += par_bytes;
  p.off_bar = (off + 7) & ~7; off = p.off_bar + 256;
  p.smem_bytes = off;
  return METRO_OK;
}

metro_status conv_chain_launch(const ConvGemmParams &prm, int num_sms, cudaStream_t stream) {
  switch (prm.cout1) {
    case 64: return launch_chain_t<64>(prm, num_sms, stream);
    case 128: return launch_chain_t<128>(prm, num_sms, stream);
    case 256: return launch_chain_t<256>(prm, num_sms, stream);
  }
  return fail(METRO_ERR_INTERNAL, "conv_chain: unsupported conv1 width %d", prm.cout1);
}

}  // namespace metro
