// Fused volumetric-heatmap decode: one pass over the head tensor.
//
// Replaces, per crop, the chain the reference builds out of ~25 TensorFlow ops
//   volumetric.py:227-235  transpose/reshape/transpose, softmax over (H,W,D), per-axis decode
//   tfu.py:466-471         softmax = exp(x - max) / sum
//   tfu.py:474-499         marginals . linspace(0,1,n)  (x <- W, y <- H, z <- D)
//   volumetric.py:288-306  heatmap_to_metric
//   tfu3d.py:23-25         root_relative (last model joint)
//   main.py:127            gather(permutation)
// which materialise [N,J,H,W,D] and re-read it about ten times.  Here the NHWC head [N,H,W,D*J]
// (channel c = d*J + j) is read exactly once with 16-byte coalesced loads and
//   out[n,jo,:] = (E_j[w]/(W-1), E_j[h]/(H-1), E_j[d]/(D-1)) - same for the root, times mm scales.
//
// Layout of work: the head tensor of one crop is a contiguous byte range; it is cut into tiles of
// whole heatmap rows (~35 KB).  A CTA is PERSISTENT over work items (item = one crop, or one of
// `splits` row ranges of a large crop) and streams each item tile by tile through a 2-stage shared
// memory ring: an elected thread issues bulk asynchronous copies (cp.async.bulk, the TMA engine) one
// ring ahead -- also across item boundaries -- so the HBM stream never waits for arithmetic, and with
// 2-3 CTAs per SM ~150 KB per SM are in flight without costing a register.
// thread = (16-byte channel slot, pixel lane); a thread walks pixels lane, lane+LANES, ... of a tile,
// so shared-memory reads are conflict-free 16-byte words, and it keeps its channels' running sums in
// registers for the whole item: the only cross-thread work (lane shuffle, depth merge, output) happens
// once per item, not per tile.  The kernel is close to instruction bound at HBM speed (one exp per
// element), so the inner loop is 2-wide packed fp32 (FFMA2/FADD2) and carries no address arithmetic.
//
// Numerics: exp(x - max) is evaluated in base 2 against an INTEGER exponent k >= max * log2(e) kept
// per (thread, channel) and raised -- with an exact power-of-two re-scale of the running sums -- when a
// later tile holds a larger value.  Per-tile partial sums (<= 32 terms, two-level) are fp32; running
// sums and every merge (tiles -> lanes -> depth -> joint -> CTA splits) are fp64 re-scaled by exact
// powers of two: no transcendental and no rounding in any merge weight.  The only fp32 roundings are
// ex2.approx per element and the short per-tile sums, which keeps the result within 1e-3 mm of the
// float64 oracle.
// When a crop is split over several CTAs the last CTA to finish (ticket counter) merges the
// per-split records; the workspace counters are left zeroed for the next launch.
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace metro {

namespace {

constexpr int kMaxThreads = 512;
constexpr int kMaxStages = 4;        // shared-memory ring depth is a plan parameter (L.stages <= kMaxStages)
constexpr int kGroup = 8;            // pixel steps per fp32 partial sum (first level)
constexpr int kMaxSteps = 32;        // pixel steps per thread per tile
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kNone = -3.0e38f;   // exponent of an empty record

// VEC channels of one pixel from a 32-bit shared-memory address, as VEC/2 float pairs
template <int VEC, bool F16>
struct Vec;
template <>
struct Vec<4, false> {  // 4 x fp32, 16 bytes
  static __device__ __forceinline__ void load(uint32_t addr, float2 (&x)[2]) {
    const float4 v = ptx::lds_v4(addr);
    x[0] = make_float2(v.x, v.y); x[1] = make_float2(v.z, v.w);
  }
};
template <>
struct Vec<2, false> {  // 2 x fp32, 8 bytes
  static __device__ __forceinline__ void load(uint32_t addr, float2 (&x)[1]) { x[0] = ptx::lds_v2(addr); }
};
template <>
struct Vec<8, true> {  // 8 x fp16, 16 bytes
  static __device__ __forceinline__ void load(uint32_t addr, float2 (&x)[4]) {
    const uint4 r = ptx::lds_v4u(addr);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
  }
};
template <>
struct Vec<4, true> {  // 4 x fp16, 8 bytes
  static __device__ __forceinline__ void load(uint32_t addr, float2 (&x)[2]) {
    const float2 r = ptx::lds_v2(addr);
    const uint32_t w[2] = {__float_as_uint(r.x), __float_as_uint(r.y)};
#pragma unroll
    for (int i = 0; i < 2; ++i) x[i] = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
  }
};

// 2^d for an integer-valued d <= 0 (exact; flushes to zero far below the range that can matter).
__device__ __forceinline__ double pow2_neg(float d) {
  const int e = int(fmaxf(d, -1000.f));
  return __longlong_as_double((long long)(1023 + e) << 52);
}
__device__ __forceinline__ double shfl_xor_f64(double v, int o) {
  return __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(v), o),
                          __shfl_xor_sync(0xffffffffu, __double2loint(v), o));
}

struct ChanRec {  // one channel of one work item
  double s, sx, sy;
  float k;
  float pad;
};

// walks the tiles of the work items a CTA owns: item = blockIdx.x, + gridDim.x, ...; the tiles of a
// crop are split evenly over its `splits` items
struct Cursor {
  int item, t, t1;
  __device__ __forceinline__ void open(int it, int n_items, int splits, int tiles) {
    item = it;
    if (it < n_items) {
      const int sp = it % splits;
      t = sp * tiles / splits;
      t1 = (sp + 1) * tiles / splits;
    } else { t = t1 = 0; }
  }
  __device__ __forceinline__ bool done(int n_items) const { return item >= n_items; }
};

// MAXT = 192: the common shapes (<= 192 threads per CTA) with up to three CTAs resident per SM
template <int VEC, int LANES, bool F16, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT <= 192 ? 3 : 1) softargmax_kernel(const SoftargmaxLaunch p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int V2 = VEC / 2;
  constexpr int esize = F16 ? 2 : 4;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int slot = tid / LANES, lane = tid % LANES;
  const bool live = slot < p.slots;                    // the last warp is padded with idle threads
  const int C = p.C, J = p.J, P = p.H * p.W;
  const int c0 = slot * VEC;
  const int row_bytes = C * esize;
  const int tile_bytes = p.ppc * row_bytes;
  const int n_items = p.n * p.splits;

  unsigned char *s_ring = smem_raw;                                              // [kStages][tile]
  float2 *s_hw = reinterpret_cast<float2 *>(smem_raw + p.off_hw);                // [ppc] (row in tile, column)
  ChanRec *s_ch = reinterpret_cast<ChanRec *>(smem_raw + p.off_ch);              // [C]
  double *s_c01 = reinterpret_cast<double *>(s_ch + C);                          // [J][3]
  uint64_t *full = reinterpret_cast<uint64_t *>(s_c01 + 3 * J);                  // [kStages]
  __shared__ int s_is_last;

  auto issue = [&](const Cursor &c, int stage) {     // thread 0 only
    const int px0 = c.t * p.ppc;
    const uint32_t bytes = uint32_t(min(p.ppc, P - px0)) * row_bytes;
    const unsigned char *src = static_cast<const unsigned char *>(p.head) +
                               (size_t(c.item / p.splits) * P + px0) * row_bytes;
    ptx::mbar_arrive_expect_tx(full + stage, bytes);
    ptx::bulk_load_1d(s_ring + size_t(stage) * tile_bytes, src, bytes, full + stage);
  };
  auto advance = [&](Cursor &c) {
    if (++c.t >= c.t1) c.open(c.item + gridDim.x, n_items, p.splits, p.tiles);
  };

  Cursor prod, cons;
  cons.open(blockIdx.x, n_items, p.splits, p.tiles);
  prod = cons;
  if (tid == 0) {
    for (int st = 0; st < p.stages; ++st) ptx::mbar_init(full + st, 1);
    ptx::fence_mbar_init();
  }
  for (int q = tid; q < p.ppc; q += nthreads) {
    const int h = q / p.W;
    s_hw[q] = make_float2(float(h), float(q - h * p.W));
  }
  // programmatic dependent launch: the set-up above overlaps the tail of the kernel that produces the head
  // tensor (the logits convolution); nothing before this line touches global memory
  ptx::griddep_wait();
  ptx::griddep_launch_dependents();
  if (tid == 0) {
    for (int st = 0; st < p.stages && !prod.done(n_items); ++st) { issue(prod, st); advance(prod); }
  }
  __syncthreads();                       // barriers initialised and the (row, column) table written

  // running record of this thread's channels over the current item
  double S[VEC], SX[VEC], SY[VEC];
  float K[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { S[v] = SX[v] = SY[v] = 0.0; K[v] = kNone; }
  const float2 l2e = make_float2(kLog2e, kLog2e);
  const uint32_t step_bytes = uint32_t(LANES) * row_bytes;
  int stage = 0;
  uint32_t phase = 0;

  while (!cons.done(n_items)) {
    const int px0 = cons.t * p.ppc;
    const int steps = min(p.ppc, P - px0) / LANES;       // whole rows or an even part of one: a multiple of LANES
    const int hh = px0 / p.W;
    const float h0 = float(hh), w0 = float(px0 - hh * p.W);
    ptx::mbar_wait_sleep(full + stage, phase);
    if (live) {
      const uint32_t base = ptx::smem_u32(s_ring) + uint32_t(stage) * tile_bytes + uint32_t(lane) * row_bytes + uint32_t(c0) * esize;
      const int groups = steps / kGroup;
      // pass 1: this thread's maximum per channel over the tile -> integer exponent
      float2 m[V2];
#pragma unroll
      for (int v = 0; v < V2; ++v) m[v] = make_float2(-INFINITY, -INFINITY);
      {
        uint32_t a = base;
        for (int g = 0; g < groups; ++g) {
#pragma unroll
          for (int ii = 0; ii < kGroup; ++ii, a += step_bytes) {
            float2 x[V2];
            Vec<VEC, F16>::load(a, x);
#pragma unroll
            for (int v = 0; v < V2; ++v) { m[v].x = fmaxf(m[v].x, x[v].x); m[v].y = fmaxf(m[v].y, x[v].y); }
          }
        }
        for (int i = groups * kGroup; i < steps; ++i, a += step_bytes) {
          float2 x[V2];
          Vec<VEC, F16>::load(a, x);
#pragma unroll
          for (int v = 0; v < V2; ++v) { m[v].x = fmaxf(m[v].x, x[v].x); m[v].y = fmaxf(m[v].y, x[v].y); }
        }
      }
      float2 nk[V2];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float kt = ceilf(((v & 1) ? m[v >> 1].y : m[v >> 1].x) * kLog2e);
        if (kt > K[v]) {                                 // raise the exponent: exact re-scale of the running sums
          const double r = pow2_neg(K[v] - kt);
          S[v] *= r; SX[v] *= r; SY[v] *= r;
          K[v] = kt;
        }
        if (v & 1) nk[v >> 1].y = -K[v]; else nk[v >> 1].x = -K[v];
      }
      // pass 2: exp once per element; fp32 partial sums of kGroup terms, then of groups
      float2 ts[V2], tx[V2], ty[V2];
#pragma unroll
      for (int v = 0; v < V2; ++v) ts[v] = tx[v] = ty[v] = make_float2(0.f, 0.f);
      uint32_t a = base;
      uint32_t hwa = ptx::smem_u32(s_hw) + uint32_t(lane) * 8;
      auto step = [&](uint32_t xa, uint32_t ha, float2 (&gs)[V2], float2 (&gx)[V2], float2 (&gy)[V2]) {
        float2 x[V2];
        Vec<VEC, F16>::load(xa, x);
        const float2 rc = ptx::lds_v2(ha);
        const float fh = rc.x + h0, fw = rc.y + w0;
        const float2 fh2 = make_float2(fh, fh), fw2 = make_float2(fw, fw);
#pragma unroll
        for (int v = 0; v < V2; ++v) {
          const float2 t = __ffma2_rn(x[v], l2e, nk[v]);
          const float2 e = make_float2(ptx::ex2_approx(t.x), ptx::ex2_approx(t.y));
          gs[v] = __fadd2_rn(gs[v], e);
          gx[v] = __ffma2_rn(e, fw2, gx[v]);
          gy[v] = __ffma2_rn(e, fh2, gy[v]);
        }
      };
      for (int g = 0; g < groups; ++g, hwa += kGroup * LANES * 8) {
        float2 gs[V2], gx[V2], gy[V2];
#pragma unroll
        for (int v = 0; v < V2; ++v) gs[v] = gx[v] = gy[v] = make_float2(0.f, 0.f);
#pragma unroll
        for (int ii = 0; ii < kGroup; ++ii, a += step_bytes) step(a, hwa + ii * LANES * 8, gs, gx, gy);
#pragma unroll
        for (int v = 0; v < V2; ++v) {
          ts[v] = __fadd2_rn(ts[v], gs[v]); tx[v] = __fadd2_rn(tx[v], gx[v]); ty[v] = __fadd2_rn(ty[v], gy[v]);
        }
      }
      if (groups * kGroup < steps) {
        float2 gs[V2], gx[V2], gy[V2];
#pragma unroll
        for (int v = 0; v < V2; ++v) gs[v] = gx[v] = gy[v] = make_float2(0.f, 0.f);
        for (int i = groups * kGroup; i < steps; ++i, a += step_bytes, hwa += LANES * 8) step(a, hwa, gs, gx, gy);
#pragma unroll
        for (int v = 0; v < V2; ++v) {
          ts[v] = __fadd2_rn(ts[v], gs[v]); tx[v] = __fadd2_rn(tx[v], gx[v]); ty[v] = __fadd2_rn(ty[v], gy[v]);
        }
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        S[v] += double((v & 1) ? ts[v >> 1].y : ts[v >> 1].x);
        SX[v] += double((v & 1) ? tx[v >> 1].y : tx[v >> 1].x);
        SY[v] += double((v & 1) ? ty[v >> 1].y : ty[v >> 1].x);
      }
    }
    __syncthreads();                     // every thread is done with this ring stage
    if (tid == 0 && !prod.done(n_items)) { issue(prod, stage); advance(prod); }
    if (++stage == p.stages) { stage = 0; phase ^= 1; }
    const int item = cons.item;
    const bool item_done = (cons.t + 1 >= cons.t1);
    advance(cons);
    if (!item_done) continue;

    // ======================= end of a work item: merge and publish =======================
    const int img = item / p.splits, split = item - img * p.splits;
    // pixel lanes -> channel: the lanes of a slot sit in one warp; exact re-scaling, fp64 sums
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float kk = K[v];
#pragma unroll
      for (int o = LANES / 2; o >= 1; o >>= 1) kk = fmaxf(kk, __shfl_xor_sync(0xffffffffu, kk, o));
      const double wgt = pow2_neg(K[v] - kk);
      double a = wgt * S[v], ax = wgt * SX[v], ay = wgt * SY[v];
#pragma unroll
      for (int o = LANES / 2; o >= 1; o >>= 1) {
        a += shfl_xor_f64(a, o); ax += shfl_xor_f64(ax, o); ay += shfl_xor_f64(ay, o);
      }
      if (live && lane == 0) {
        ChanRec o; o.s = a; o.sx = ax; o.sy = ay; o.k = kk; o.pad = 0.f;
        s_ch[c0 + v] = o;
      }
      S[v] = SX[v] = SY[v] = 0.0; K[v] = kNone;        // reset for the next item
    }
    __syncthreads();
    // depth -> joint
    double TS = 0.0, TX = 0.0, TY = 0.0, TZ = 0.0;
    float TK = kNone;
    if (tid < J) {
      for (int d = 0; d < p.D; ++d) TK = fmaxf(TK, s_ch[d * J + tid].k);
      for (int d = 0; d < p.D; ++d) {
        const ChanRec r = s_ch[d * J + tid];
        const double wgt = pow2_neg(r.k - TK);
        TS += wgt * r.s; TX += wgt * r.sx; TY += wgt * r.sy; TZ += double(d) * (wgt * r.s);
      }
    }
    bool publish = true;
    if (p.splits > 1) {
      // publish this split's record; the last CTA of the crop merges them.  bar.sync orders the J
      // writers before thread 0, whose gpu-scope fence is cumulative over what it has observed.
      if (tid < J) {
        double *rec = p.partials + ((size_t(img) * p.splits + split) * J + tid) * 5;
        rec[0] = double(TK); rec[1] = TS; rec[2] = TX; rec[3] = TY; rec[4] = TZ;
      }
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(p.counters + img, 1u);
        s_is_last = (ticket == unsigned(p.splits - 1));
        if (s_is_last) __threadfence();
      }
      __syncthreads();
      publish = s_is_last != 0;
      if (publish && tid < J) {
        const double *recs = p.partials + (size_t(img) * p.splits) * J * 5;
        double gk = double(kNone);
        for (int sp = 0; sp < p.splits; ++sp) gk = fmax(gk, __ldcg(recs + (size_t(sp) * J + tid) * 5));
        TS = TX = TY = TZ = 0.0;
        for (int sp = 0; sp < p.splits; ++sp) {
          const double *rec = recs + (size_t(sp) * J + tid) * 5;
          const double wgt = pow2_neg(float(__ldcg(rec) - gk));
          TS += wgt * __ldcg(rec + 1); TX += wgt * __ldcg(rec + 2);
          TY += wgt * __ldcg(rec + 3); TZ += wgt * __ldcg(rec + 4);
        }
      }
      if (publish && tid == 0) p.counters[img] = 0;   // self-cleaning for the next launch
    }
    if (publish) {
      // expectation of linspace(0,1,n) along each axis == E[index]/(n-1); mul_* carry 1/(n-1) and mm
      if (tid < J) {
        const double inv = 1.0 / TS;
        s_c01[3 * tid] = TX * inv * p.mul_x;
        s_c01[3 * tid + 1] = TY * inv * p.mul_y;
        s_c01[3 * tid + 2] = TZ * inv * p.mul_z;
      }
      __syncthreads();
      for (int i = tid; i < p.n_out * 3; i += nthreads) {
        const int jo = i / 3, a = i - 3 * jo;
        p.out[(size_t(img) * p.n_out) * 3 + i] = float(s_c01[3 * p.perm[jo] + a] - s_c01[3 * p.root + a]);
      }
    }
    __syncthreads();                     // s_ch / s_c01 / s_is_last are reused by the next item
  }
}

size_t tile_bytes(const SoftargmaxLaunch &L) {
  return (size_t(L.ppc) * L.C * (L.head_f16 ? 2 : 4) + 127) & ~size_t(127);
}
size_t hw_bytes(const SoftargmaxLaunch &L) { return (size_t(L.ppc) * 8 + 127) & ~size_t(127); }

size_t smem_bytes(const SoftargmaxLaunch &L) {
  return L.stages * tile_bytes(L) + hw_bytes(L) + size_t(L.C) * sizeof(ChanRec) + size_t(3) * L.J * 8 + kMaxStages * 8 + 16;
}

template <int VEC, int LANES, bool F16, int MAXT>
metro_status launch_t(const SoftargmaxLaunch &L, cudaStream_t stream) {
  static size_t configured = 0;
  const size_t sm = smem_bytes(L);
  if (configured < sm) {
    METRO_CUDA(cudaFuncSetAttribute(softargmax_kernel<VEC, LANES, F16, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  const int n_items = L.n * L.splits;
  const dim3 grid(unsigned(n_items < L.max_ctas ? n_items : L.max_ctas)), block(unsigned((L.slots * L.lanes + 31) & ~31));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = sm; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  METRO_CUDA(cudaLaunchKernelEx(&cfg, softargmax_kernel<VEC, LANES, F16, MAXT>, L));
  return METRO_OK;
}

}  // namespace

metro_status softargmax_plan(const metro_softargmax_desc &d, int n, SoftargmaxLaunch &L) {
  if (d.side <= 0 || d.n_joints_model <= 0 || d.depth <= 0 || d.n_joints_out <= 0 || !d.permutation)
    return fail(METRO_ERR_VALUE, "softargmax: side, joints, depth must be positive and permutation non-null");
  if (d.n_joints_out > kMaxJointsOut) return fail(METRO_ERR_VALUE, "softargmax: n_joints_out > %d", kMaxJointsOut);
  if (d.stride <= 0 || d.proc_side <= 0) return fail(METRO_ERR_VALUE, "softargmax: stride and proc_side must be positive");
  if (n < 0) return fail(METRO_ERR_VALUE, "softargmax: negative batch");
  // a thread owns one 16-byte word of a pixel (4 fp32 / 8 fp16 channels); 8-byte words when the channel
  // count is not a multiple of that
  const bool f16 = d.head_dtype == METRO_F16;
  if (d.word_bytes != 0 && d.word_bytes != 8 && d.word_bytes != 16) return fail(METRO_ERR_VALUE, "softargmax: word_bytes must be 0, 8 or 16");
  int vec = (d.word_bytes == 8 ? 8 : 16) / (f16 ? 2 : 4);
  if (d.word_bytes == 0 && (d.n_joints_model * d.depth) % vec != 0) vec /= 2;
  if (d.head_dtype != METRO_F16 && d.head_dtype != METRO_F32) return fail(METRO_ERR_VALUE, "softargmax: bad head_dtype");
  L = SoftargmaxLaunch();
  L.n = n; L.H = L.W = d.side; L.J = d.n_joints_model; L.D = d.depth; L.C = L.J * L.D;
  if (L.C % vec != 0) return fail(METRO_ERR_VALUE, "softargmax: depth*joints (%d) must be a multiple of %d", L.C, vec);
  L.n_out = d.n_joints_out; L.root = L.J - 1;   // tfu3d.py:23-25: the last joint is the root
  for (int i = 0; i < L.n_out; ++i) {
    if (d.permutation[i] < 0 || d.permutation[i] >= L.J)
      return fail(METRO_ERR_VALUE, "softargmax: permutation[%d]=%d out of range [0,%d)", i, d.permutation[i], L.J);
    L.perm[i] = d.permutation[i];
  }
  // volumetric.py:288-306: xy_mm = (c*lrc + stride//2) * box/proc_side ; z_mm = c*box.  The additive
  // term cancels in the root-relative difference.
  const int last = d.proc_side - 1;
  const double lrc = double(last - (last % d.stride) - 1);
  const double xy = lrc * double(d.box_size_mm) / double(d.proc_side);
  L.mul_x = L.W > 1 ? xy / double(L.W - 1) : 0.0;
  L.mul_y = L.H > 1 ? xy / double(L.H - 1) : 0.0;
  L.mul_z = L.D > 1 ? double(d.box_size_mm) / double(L.D - 1) : 0.0;
  L.head_f16 = d.head_dtype == METRO_F16;
  L.slots = L.C / vec;
  L.vec = vec;
  if (L.slots > kMaxThreads) return fail(METRO_ERR_VALUE, "softargmax: too many head channels (%d)", L.C);
  const int P = L.H * L.W;
  // CTA shape: `slots` 16-byte channel slots x `lanes` pixel lanes (a power of two dividing W, so whole
  // rows split evenly over the lanes and the lanes of a slot share a warp).  Tiles are whole rows,
  // ~35 KB; a 2-stage ring per CTA and 2-3 CTAs per SM keep ~150 KB of copies in flight per SM.
  int lanes = d.lanes > 0 ? d.lanes : 4;
  while (lanes > 1 && (lanes > 8 || (lanes & (lanes - 1)) != 0 || L.W % lanes != 0 ||
                       ((L.slots * lanes + 31) & ~31) > kMaxThreads)) --lanes;
  if (((L.slots * lanes + 31) & ~31) > kMaxThreads) return fail(METRO_ERR_VALUE, "softargmax: too many head channels (%d)", L.C);
  L.lanes = lanes;
  const int row_bytes = L.C * (L.head_f16 ? 2 : 4);
  // tile: ~36 KB of whole rows, or an even part of one row when a row is larger than that (measured on
  // B200: larger tiles amortise the per-tile bookkeeping better than more resident CTAs hide latency)
  // ring: `stages` tiles of ~`tile_kb` KB in flight per CTA (tunable for experiments: METRO_SAM_STAGES / METRO_SAM_TILE_KB)
  int tile_kb = 40;
  L.stages = 2;
  if (const char *e = getenv("METRO_SAM_STAGES")) L.stages = atoi(e);
  if (const char *e = getenv("METRO_SAM_TILE_KB")) tile_kb = atoi(e);
  if (L.stages < 2) L.stages = 2;
  if (L.stages > kMaxStages) L.stages = kMaxStages;
  const int budget_px = (tile_kb * 1024) / row_bytes;
  int ppc;
  if (budget_px >= L.W) {
    int rows = budget_px / L.W;
    if (rows > kMaxSteps * lanes / L.W) rows = kMaxSteps * lanes / L.W;   // bounded per-thread partial sums
    if (rows > L.H) rows = L.H;
    if (rows < 1) rows = 1;
    ppc = rows * L.W;
    L.tiles = (L.H + rows - 1) / rows;
  } else {
    int parts = 1;
    while (L.W % (parts * 2) == 0 && (L.W / (parts * 2)) % lanes == 0 && L.W / parts > budget_px) parts *= 2;
    ppc = L.W / parts;
    if (size_t(ppc) * row_bytes > 90 * 1024) return fail(METRO_ERR_VALUE, "softargmax: a heatmap row of %d bytes cannot be tiled", row_bytes * L.W);
    L.tiles = L.H * parts;
  }
  L.ppc = ppc;
  // work items: one per crop, or -- for large heatmaps -- a fixed number of row ranges per crop that
  // depends on the heatmap shape ONLY, so a crop's result is bit-identical whatever batch or GPU shard it
  // is part of (merged by the last CTA of the crop to finish)
  int splits = d.splits > 0 ? d.splits : L.tiles / 8;
  if (splits > 4 && d.splits <= 0) splits = 4;
  if (splits < 1) splits = 1;
  if (splits > L.tiles) splits = L.tiles;
  L.splits = splits;
  L.max_ctas = 148 * 3;
  L.rpt = L.ppc / lanes;
  L.off_hw = int(L.stages * tile_bytes(L));
  L.off_ch = L.off_hw + int(hw_bytes(L));
  if (smem_bytes(L) > 200 * 1024) return fail(METRO_ERR_VALUE, "softargmax: shared memory budget exceeded");
  return METRO_OK;
}

size_t softargmax_workspace_bytes(const SoftargmaxLaunch &L) {
  const size_t counters = (size_t(L.n) * 4 + 255) & ~size_t(255);
  return counters + size_t(L.n) * L.splits * L.J * 5 * 8;
}

metro_status softargmax_launch(const SoftargmaxLaunch &L, cudaStream_t stream) {
  if (L.n == 0) return METRO_OK;
  const int threads = (L.slots * L.lanes + 31) & ~31;
#define METRO_SAM(V, LN, H) return (threads <= 192 ? launch_t<V, LN, H, 192>(L, stream) : launch_t<V, LN, H, kMaxThreads>(L, stream))
#define METRO_SAM_LANES(V, H)        \
  switch (L.lanes) {                 \
    case 1: METRO_SAM(V, 1, H);      \
    case 2: METRO_SAM(V, 2, H);      \
    case 4: METRO_SAM(V, 4, H);      \
    case 8: METRO_SAM(V, 8, H);      \
  }
  if (L.head_f16) {
    if (L.vec == 8) { METRO_SAM_LANES(8, true) } else { METRO_SAM_LANES(4, true) }
  } else {
    if (L.vec == 4) { METRO_SAM_LANES(4, false) } else { METRO_SAM_LANES(2, false) }
  }
#undef METRO_SAM_LANES
#undef METRO_SAM
  return fail(METRO_ERR_INTERNAL, "softargmax: %d lanes not instantiated", L.lanes);
}

}  // namespace metro
