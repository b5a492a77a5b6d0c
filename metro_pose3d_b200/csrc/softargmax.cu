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
// Layout of work: a CTA owns `ppc` consecutive pixels of one crop = ONE contiguous byte range of the
// head tensor, which a single elected thread fetches with one bulk asynchronous copy (cp.async.bulk,
// the TMA engine) into shared memory; several CTAs are resident per SM, so tens of KB per SM are in
// flight without costing a register, and one CTA's reduction tail overlaps the others' copies.
// thread = (pixel lane, 16-byte channel slot): shared-memory reads are conflict-free 16-byte words.
//
// Numerics: exp(x - max) is evaluated in base 2 against an INTEGER exponent k >= max * log2(e) taken
// per (thread, channel).  Because every partial record carries an integer exponent, all merges (pixel
// lanes -> channel -> depth -> joint -> CTA splits) re-scale by exact powers of two: there is no
// transcendental and no rounding in the merge weights, and the merged sums are carried in fp64.  The
// only fp32 roundings are ex2.approx per element and the short per-thread sums, which keeps the result
// within 1e-3 mm of the float64 oracle.
// When a crop is split over several CTAs the last CTA to finish (ticket counter) merges the
// per-split records; the workspace counters are left zeroed for the next launch.
#include <cuda_fp16.h>

#include "common.h"
#include "ptx.cuh"

namespace metro {

namespace {

constexpr int kMaxThreads = 320;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kNone = -3.0e38f;   // exponent of an empty record

template <int VEC>
struct Vec;
template <>
struct Vec<4> {  // 4 x fp32
  static __device__ __forceinline__ void load(const unsigned char *p, float (&x)[4]) {
    const float4 v = *reinterpret_cast<const float4 *>(p);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
  }
};
template <>
struct Vec<8> {  // 8 x fp16
  static __device__ __forceinline__ void load(const unsigned char *p, float (&x)[8]) {
    const uint4 r = *reinterpret_cast<const uint4 *>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
      x[2 * i] = f.x;
      x[2 * i + 1] = f.y;
    }
  }
};

// 2^d for an integer-valued d <= 0 (exact; flushes to zero far below the range that can matter).
__device__ __forceinline__ double pow2_neg(float d) {
  const int e = int(fmaxf(d, -1000.f));
  return __longlong_as_double((long long)(1023 + e) << 52);
}

struct ChanRec {  // one channel of this CTA's pixel range
  double s, sx, sy;
  float k;
  float pad;
};

template <int VEC>
__global__ void __launch_bounds__(kMaxThreads, 4) softargmax_kernel(const SoftargmaxLaunch p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int img = blockIdx.x / p.splits, split = blockIdx.x - img * p.splits;
  const int slot = tid % p.slots, lane = tid / p.slots;
  const int C = p.C, J = p.J, P = p.H * p.W;
  const int p_begin = split * p.ppc;
  const int n_px = min(P, p_begin + p.ppc) - p_begin;
  const int c0 = slot * VEC;
  constexpr int esize = (VEC == 8) ? 2 : 4;
  const int row_bytes = C * esize;

  // shared memory: [tile | records (aliases the tile once it has been consumed)] [channel records] [coords] [barrier]
  unsigned char *s_tile = smem_raw;
  float4 *s_rec = reinterpret_cast<float4 *>(smem_raw);                          // [lanes][C] (k, s, sx, sy)
  ChanRec *s_ch = reinterpret_cast<ChanRec *>(smem_raw + p.off_ch);              // [C]
  double *s_c01 = reinterpret_cast<double *>(s_ch + C);                          // [J][3]
  uint64_t *bar = reinterpret_cast<uint64_t *>(s_c01 + 3 * J);
  __shared__ int s_is_last;

  // ---- one bulk copy of this CTA's contiguous pixel range ----------------------------------------------
  if (tid == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
    const uint32_t bytes = uint32_t(n_px) * row_bytes;
    const unsigned char *src = static_cast<const unsigned char *>(p.head) + (size_t(img) * P + p_begin) * row_bytes;
    ptx::mbar_arrive_expect_tx(bar, bytes);
    ptx::bulk_load_1d(s_tile, src, bytes, bar);
  }
  __syncthreads();                       // barrier initialised before anyone polls it
  ptx::mbar_wait(bar, 0);

  // ---- per (thread, channel): integer exponent >= max*log2e, then exp once per element --------------
  const unsigned char *mine = s_tile + size_t(lane) * row_bytes + size_t(c0) * esize;
  const int step = p.lanes * row_bytes;
  const int n_mine = lane < n_px ? (n_px - lane + p.lanes - 1) / p.lanes : 0;
  float k[VEC], s[VEC], sx[VEC], sy[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { k[v] = -INFINITY; s[v] = sx[v] = sy[v] = 0.f; }
  for (int r = 0; r < n_mine; ++r) {
    float x[VEC];
    Vec<VEC>::load(mine + size_t(r) * step, x);
#pragma unroll
    for (int v = 0; v < VEC; ++v) k[v] = fmaxf(k[v], x[v]);
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) k[v] = (k[v] == -INFINITY) ? kNone : ceilf(k[v] * kLog2e);
  {
    int q = p_begin + lane;
    for (int r = 0; r < n_mine; ++r, q += p.lanes) {
      float x[VEC];
      Vec<VEC>::load(mine + size_t(r) * step, x);
      const int h = q / p.W;
      const float fh = float(h), fw = float(q - h * p.W);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float e = ptx::ex2_approx(fmaf(x[v], kLog2e, -k[v]));
        s[v] += e;
        sx[v] = fmaf(e, fw, sx[v]);
        sy[v] = fmaf(e, fh, sy[v]);
      }
    }
  }
  __syncthreads();                       // the tile is consumed; its space becomes the record array
#pragma unroll
  for (int v = 0; v < VEC; ++v) s_rec[size_t(lane) * C + c0 + v] = make_float4(k[v], s[v], sx[v], sy[v]);
  __syncthreads();

  // ---- pixel lanes -> channel (exact power-of-two re-scaling, fp64 sums) ---------------------------
  for (int c = tid; c < C; c += nthreads) {
    float kk = kNone;
    for (int l = 0; l < p.lanes; ++l) kk = fmaxf(kk, s_rec[size_t(l) * C + c].x);
    double a = 0.0, ax = 0.0, ay = 0.0;
    for (int l = 0; l < p.lanes; ++l) {
      const float4 r = s_rec[size_t(l) * C + c];
      const double wgt = pow2_neg(r.x - kk);
      a += wgt * double(r.y); ax += wgt * double(r.z); ay += wgt * double(r.w);
    }
    ChanRec o; o.s = a; o.sx = ax; o.sy = ay; o.k = kk; o.pad = 0.f;
    s_ch[c] = o;
  }
  __syncthreads();

  // ---- depth -> joint ----------------------------------------------------------------------------
  double S = 0.0, SX = 0.0, SY = 0.0, SZ = 0.0;
  float K = kNone;
  if (tid < J) {
    for (int d = 0; d < p.D; ++d) K = fmaxf(K, s_ch[d * J + tid].k);
    for (int d = 0; d < p.D; ++d) {
      const ChanRec r = s_ch[d * J + tid];
      const double wgt = pow2_neg(r.k - K);
      S += wgt * r.s; SX += wgt * r.sx; SY += wgt * r.sy; SZ += double(d) * (wgt * r.s);
    }
  }

  if (p.splits > 1) {
    // publish this split's record; the last CTA of the crop merges them.  bar.sync orders the J
    // writers before thread 0, whose gpu-scope fence is cumulative over what it has observed.
    if (tid < J) {
      double *rec = p.partials + ((size_t(img) * p.splits + split) * J + tid) * 5;
      rec[0] = double(K); rec[1] = S; rec[2] = SX; rec[3] = SY; rec[4] = SZ;
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      const unsigned int ticket = atomicAdd(p.counters + img, 1u);
      s_is_last = (ticket == unsigned(p.splits - 1));
      if (s_is_last) __threadfence();
    }
    __syncthreads();
    if (!s_is_last) return;
    if (tid < J) {
      const double *recs = p.partials + (size_t(img) * p.splits) * J * 5;
      double gk = double(kNone);
      for (int sp = 0; sp < p.splits; ++sp) gk = fmax(gk, __ldcg(recs + (size_t(sp) * J + tid) * 5));
      S = SX = SY = SZ = 0.0;
      for (int sp = 0; sp < p.splits; ++sp) {
        const double *rec = recs + (size_t(sp) * J + tid) * 5;
        const double wgt = pow2_neg(float(__ldcg(rec) - gk));
        S += wgt * __ldcg(rec + 1); SX += wgt * __ldcg(rec + 2);
        SY += wgt * __ldcg(rec + 3); SZ += wgt * __ldcg(rec + 4);
      }
    }
    if (tid == 0) p.counters[img] = 0;   // self-cleaning for the next launch
  }

  // expectation of linspace(0,1,n) along each axis == E[index]/(n-1); mul_* carry 1/(n-1) and mm
  if (tid < J) {
    const double inv = 1.0 / S;
    s_c01[3 * tid] = SX * inv * p.mul_x;
    s_c01[3 * tid + 1] = SY * inv * p.mul_y;
    s_c01[3 * tid + 2] = SZ * inv * p.mul_z;
  }
  __syncthreads();
  for (int i = tid; i < p.n_out * 3; i += nthreads) {
    const int jo = i / 3, a = i - 3 * jo;
    p.out[(size_t(img) * p.n_out) * 3 + i] = float(s_c01[3 * p.perm[jo] + a] - s_c01[3 * p.root + a]);
  }
}

size_t tile_region_bytes(const SoftargmaxLaunch &L) {
  const size_t tile = size_t(L.ppc) * L.C * (L.head_f16 ? 2 : 4);
  const size_t rec = size_t(L.lanes) * L.C * 16;
  return ((tile > rec ? tile : rec) + 127) & ~size_t(127);
}

size_t smem_bytes(const SoftargmaxLaunch &L) {
  return tile_region_bytes(L) + size_t(L.C) * sizeof(ChanRec) + size_t(3) * L.J * 8 + 16;
}

}  // namespace

metro_status softargmax_plan(const metro_softargmax_desc &d, int n, SoftargmaxLaunch &L) {
  if (d.side <= 0 || d.n_joints_model <= 0 || d.depth <= 0 || d.n_joints_out <= 0 || !d.permutation)
    return fail(METRO_ERR_VALUE, "softargmax: side, joints, depth must be positive and permutation non-null");
  if (d.n_joints_out > kMaxJointsOut) return fail(METRO_ERR_VALUE, "softargmax: n_joints_out > %d", kMaxJointsOut);
  if (d.stride <= 0 || d.proc_side <= 0) return fail(METRO_ERR_VALUE, "softargmax: stride and proc_side must be positive");
  if (n < 0) return fail(METRO_ERR_VALUE, "softargmax: negative batch");
  const int vec = d.head_dtype == METRO_F16 ? 8 : 4;
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
  if (L.slots > kMaxThreads) return fail(METRO_ERR_VALUE, "softargmax: too many head channels (%d)", L.C);
  const int P = L.H * L.W;
  // CTA shape: `lanes` pixel lanes x `slots` 16-byte channel slots; the CTA's tile is `ppc` pixels
  // (~35 KB by default, so ~5 CTAs = ~175 KB of copies are in flight per SM).
  int lanes = d.lanes > 0 ? d.lanes : 8;
  if (lanes > P) lanes = P;
  while (lanes > 1 && lanes * L.slots > kMaxThreads) --lanes;
  L.lanes = lanes;
  const int row_bytes = L.C * (L.head_f16 ? 2 : 4);
  int ppc_max = (40 * 1024) / row_bytes;                 // tile budget
  if (ppc_max < lanes) ppc_max = lanes;
  int splits = d.splits > 0 ? d.splits : (P * row_bytes + 36 * 1024 - 1) / (36 * 1024);
  if (splits < (P + ppc_max - 1) / ppc_max) splits = (P + ppc_max - 1) / ppc_max;
  if (splits > P) splits = P;
  L.ppc = (P + splits - 1) / splits;
  L.splits = (P + L.ppc - 1) / L.ppc;
  L.rpt = (L.ppc + lanes - 1) / lanes;
  L.off_ch = int(tile_region_bytes(L));
  if (smem_bytes(L) > 48 * 1024) return fail(METRO_ERR_VALUE, "softargmax: shared memory budget exceeded");
  return METRO_OK;
}

size_t softargmax_workspace_bytes(const SoftargmaxLaunch &L) {
  const size_t counters = (size_t(L.n) * 4 + 255) & ~size_t(255);
  return counters + size_t(L.n) * L.splits * L.J * 5 * 8;
}

metro_status softargmax_launch(const SoftargmaxLaunch &L, cudaStream_t stream) {
  if (L.n == 0) return METRO_OK;
  const dim3 grid(unsigned(L.n) * L.splits), block(unsigned(L.slots) * L.lanes);
  const size_t sm = smem_bytes(L);
  if (L.head_f16) softargmax_kernel<8><<<grid, block, sm, stream>>>(L);
  else softargmax_kernel<4><<<grid, block, sm, stream>>>(L);
  METRO_CUDA(cudaGetLastError());
  return METRO_OK;
}

}  // namespace metro
