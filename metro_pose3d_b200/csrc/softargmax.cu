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
// (channel c = d*J + j) is read exactly once with 16-byte coalesced streaming loads and
//   out[n,jo,:] = (E_j[w]/(W-1), E_j[h]/(H-1), E_j[d]/(D-1)) - same for the root, times mm scales.
//
// Layout of work: the head tensor of a crop is a contiguous run of 16-byte words (8-byte words for an fp16
// head), `slots` words per pixel.  A work item is one crop or -- for heatmaps above 256 pixels -- one of four
// (two above 32x32) row-major pixel ranges, a rule that depends on the heatmap shape only; one CTA per item,
// two CTAs resident per SM; the CTAs of a split crop form a thread-block cluster.  The CTA has slots x lanes
// threads (272-304 for the reference shapes); thread t owns channel slot t % slots and walks pixels
// (t / slots), + lanes, ...: consecutive threads read consecutive words, so
// every warp request is 512 contiguous bytes.  The stream goes global -> registers in chunks of 8 words per
// thread; a register word is reloaded with the next chunk's data after its last use, so the loads of chunk
// k+1 are in flight underneath the arithmetic of chunk k.  No shared-memory staging, no block-wide
// synchronisation inside the stream, ~18 warps per SM to hide the latency of the one exp per element.
// (The previous design staged tiles through a cp.async.bulk ring with 5-warp CTAs; it moved bytes as fast
// but its two shared-memory passes per tile and serial per-item merge cost 3-4 us per launch.)
// The per-thread records (4 channels) meet in shared memory once per item, where 16 threads per joint merge
// the lanes x depth records with one butterfly.  A CTA's life outside the stream is ~2500 cycles (records,
// merge, output), which is why items are long and few.
//
// Numerics: exp(x - max) is evaluated in base 2 against an INTEGER exponent k >= max * log2(e) kept
// per (thread, channel) and raised -- with an exact power-of-two re-scale of the running sums -- when a
// later chunk holds a larger value.  Per-thread sums are fp32, two-level (8 terms, then chunks); every
// merge (lanes x depth -> joint -> CTA splits) is fp64 with weights that are exact powers of two applied
// in fp32 (an exact scaling): no transcendental and no rounding in any merge weight.  The only fp32 roundings are ex2.approx per
// element and the short per-thread sums, which keeps the result within 1e-3 mm of the float64 oracle.
// When a crop is split over 2, 4 or 8 CTAs they are one cluster: every CTA stores its per-joint records into
// the first CTA's shared memory (st.shared::cluster), one cluster barrier later that CTA merges them.  Other
// split counts (tests, tuning) fall back to a global workspace and a ticket counter -- the last CTA to finish
// merges -- which is left zeroed for the next launch.
#include <cuda_fp16.h>

#include <cstdlib>
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace metro {

namespace {

constexpr int kMaxThreads = 640;
constexpr int kItemPixels = 256;     // default pixels per work item
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kNone = -3.0e38f;   // exponent of an empty record

// one global word of VEC channels = RAW 32-bit registers
template <int RAW>
struct Raw {
  uint32_t r[RAW];
};
// predicated streaming load: the word, or `fill` in every register when `on` is zero (no branch, so that the
// compiler keeps the load where it is written -- in the middle of the previous chunk's arithmetic)
__device__ __forceinline__ void ldg_stream(Raw<4> &v, const void *p, int on, uint32_t fill) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "mov.b32 %0, %6;\n\tmov.b32 %1, %6;\n\tmov.b32 %2, %6;\n\tmov.b32 %3, %6;\n\t"
      "@q ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}\n"
      : "=&r"(v.r[0]), "=&r"(v.r[1]), "=&r"(v.r[2]), "=&r"(v.r[3])
      : "l"(p), "r"(on), "r"(fill));
}
__device__ __forceinline__ void ldg_stream(Raw<2> &v, const void *p, int on, uint32_t fill) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %3, 0;\n\t"
      "mov.b32 %0, %4;\n\tmov.b32 %1, %4;\n\t"
      "@q ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];\n\t}\n"
      : "=&r"(v.r[0]), "=&r"(v.r[1])
      : "l"(p), "r"(on), "r"(fill));
}
// RAW registers -> V2 float pairs
template <int RAW, bool F16, int V2>
__device__ __forceinline__ void unpack(const Raw<RAW> &v, float2 (&x)[V2]) {
  if constexpr (F16) {
#pragma unroll
    for (int i = 0; i < V2; ++i) x[i] = __half22float2(*reinterpret_cast<const __half2 *>(&v.r[i]));
  } else {
#pragma unroll
    for (int i = 0; i < V2; ++i) x[i] = make_float2(__uint_as_float(v.r[2 * i]), __uint_as_float(v.r[2 * i + 1]));
  }
}

// 2^d for an integer-valued d <= 0 (exact; flushes to zero far below the range that can matter).
__device__ __forceinline__ double pow2_neg(float d) {
  const int e = int(fmaxf(d, -1000.f));
  return __longlong_as_double((long long)(1023 + e) << 52);
}
__device__ __forceinline__ float pow2_neg_f32(float d) {   // same in fp32, d clamped to the normal range
  const int e = int(fmaxf(d, -126.f));
  return __int_as_float((127 + e) << 23);
}
// 1/a for a sum of exponentials (0.5 <= a < 2^30: well inside fp32 range): fp32 seed, two Newton steps in
// fp64 (relative error 2^-23 -> 2^-46 -> below fp64 rounding), a short dependent chain instead of a division
__device__ __forceinline__ double recip(double a) {
  double r = double(__frcp_rn(float(a)));
  r = r * (2.0 - a * r);
  r = r * (2.0 - a * r);
  return r;
}
__device__ __forceinline__ double shfl_xor_f64(double v, int o) {
  return __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(v), o),
                          __shfl_xor_sync(0xffffffffu, __double2loint(v), o));
}

// VEC channels per word, CH words per chunk; MAXT threads at most, MINB CTAs per SM
template <int VEC, bool F16, int CH, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) softargmax_kernel(const SoftargmaxLaunch p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int V2 = VEC / 2;
  constexpr int esize = F16 ? 2 : 4;
  constexpr int RAW = VEC * esize / 4;
  const int tid = threadIdx.x;
  const int C = p.C, J = p.J, P = p.H * p.W;
  const int lane_px = tid / p.slots, slot = tid - lane_px * p.slots;
  const bool live = lane_px < p.lanes;                 // the last warp is padded with idle threads
  const int c0 = slot * VEC;
  const int LC = p.lanes * C;

  float *s_k = reinterpret_cast<float *>(smem_raw);                              // [lanes][C] exponent
  float *s_s = s_k + LC, *s_x = s_s + LC, *s_y = s_x + LC;                       // [lanes][C] sums
  double *s_c01 = reinterpret_cast<double *>(s_y + LC);                          // [J][3] coordinates in mm
  double *s_part = s_c01 + 3 * J;                                                // [splits][J][5], cluster launches
  __shared__ int s_is_last;

  const int item = blockIdx.x;
  const int img = item / p.splits, split = item - img * p.splits;
  const int px0 = split * p.ipx, px1 = min(P, px0 + p.ipx);
  const int nchunks = (px1 - px0 + p.lanes * CH - 1) / (p.lanes * CH);

  // running record of this thread's channels
  float2 K2[V2], ts[V2], tx[V2], ty[V2];
#pragma unroll
  for (int v = 0; v < V2; ++v) {
    K2[v] = make_float2(kNone, kNone);
    ts[v] = tx[v] = ty[v] = make_float2(0.f, 0.f);
  }
  int pix = px0 + lane_px;                             // pixel of the next word to LOAD
  const unsigned char *ptr = static_cast<const unsigned char *>(p.head) + ((size_t(img) * P + pix) * C + c0) * esize;
  const size_t stride = size_t(p.lanes) * C * esize;
  // (row, column) of the next word to COMPUTE, advanced by `lanes` pixels per word
  float fh, fw;
  {
    const int h = pix / p.W;
    fh = float(h); fw = float(pix - h * p.W);
  }
  const float Wf = float(p.W), dq = float(p.lanes / p.W), dr = float(p.lanes % p.W);
  const float2 l2e = make_float2(kLog2e, kLog2e);

  // one register buffer of CH words: a word is reloaded with the next chunk's data right after its last use,
  // so the loads of chunk k+1 are in flight underneath the arithmetic of chunk k
  Raw<RAW> buf[CH];
  constexpr uint32_t kFill = F16 ? 0xfc00fc00u : 0xff800000u;      // -inf: a missing pixel adds exp(-inf) = 0
  auto load_word = [&](Raw<RAW> &w, bool more) {
    ldg_stream(w, ptr, int(more && live && pix < px1), kFill);
    pix += p.lanes; ptr += stride;
  };

  // this thread's slot of the output (joint after the export permutation, axis), fetched before the stream
  const int out_jo = tid / 3, out_ax = tid - 3 * out_jo;
  const int out_src = tid < p.n_out * 3 ? 3 * p.perm[out_jo] + out_ax : 0;
  // programmatic dependent launch: everything above overlaps the tail of the kernel that produces the head
  // tensor (the logits convolution); nothing before this line touches global memory
  const bool prof = p.prof != nullptr && tid == 0;
  long long *stamp = p.prof + size_t(blockIdx.x) * 8;
  if (prof) { stamp[0] = clock64(); stamp[6] = (long long)ptx::globaltimer(); }
  if (!p.dep_flags && p.l2_prefetch) {
    // Ahead of the dependency wait the item's bytes are pulled from HBM into L2 (a prefetch reads no values, so it is
    // safe whatever the previous kernel is still writing: L2 is the point of coherence).  In a stream of launches
    // this overlaps the HBM transfer of launch k+1 with the merge / output tail and the completion latency of launch
    // k, during which the memory system would otherwise idle (~2 us of every ~10 at 256 crops).
    const unsigned char *b0 = static_cast<const unsigned char *>(p.head) + (size_t(img) * P + px0) * C * esize;
    const int bytes = (px1 - px0) * C * esize;
    for (int off = tid * 128; off < bytes; off += int(blockDim.x) * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(b0 + off));
    // The first `early` CTAs of a launch fit next to the previous launch's CTAs (the grid leaves that many residency
    // slots free), so they start while it is still streaming: they pull the WHOLE input of this launch into L2, in
    // equal shares, underneath the previous launch -- the later CTAs then find their item in L2.
    if (int(blockIdx.x) < p.early) {
      const size_t total = size_t(p.n) * P * C * esize;
      const unsigned char *base = static_cast<const unsigned char *>(p.head);
      for (size_t off = (size_t(blockIdx.x) * blockDim.x + tid) * 128; off < total; off += size_t(p.early) * blockDim.x * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
    }
  }
  if (p.dep_flags) {
    // one thread polls the crop's counter (acquire, GPU scope); the barrier extends the ordering to the CTA
    if (tid == 0) ptx::flag_wait(p.dep_flags + img, p.dep_expected);
    __syncthreads();
  } else {
    ptx::griddep_wait();
  }
  ptx::griddep_launch_dependents();
  ptx::stamp_begin(p.tstamp);
  // distributed shared memory of a peer may only be written once that CTA is known to have started: every CTA of a
  // split crop arrives here and waits just before its remote stores (found by compute-sanitizer: "block that might
  // not have entered yet"; in practice the whole stream lies in between)
  if (p.cluster && p.splits > 1) ptx::cluster_arrive();
  if (prof) stamp[1] = clock64();

#pragma unroll
  for (int i = 0; i < CH; ++i) load_word(buf[i], true);
  for (int c = 0; c < nchunks; ++c) {
    const bool more = c + 1 < nchunks;
    // this chunk's maximum per channel -> integer exponent; raising it re-scales the running sums exactly
    float2 nk[V2];
    {
      float2 m[V2];
      if constexpr (F16) {
        // maximum on the raw half pairs (one HMNMX2 per two elements), widened once
        __half2 mh[RAW];
#pragma unroll
        for (int i = 0; i < CH; ++i) {
#pragma unroll
          for (int v = 0; v < RAW; ++v) {
            const __half2 xv = *reinterpret_cast<const __half2 *>(&buf[i].r[v]);
            mh[v] = i ? __hmax2(mh[v], xv) : xv;
          }
        }
#pragma unroll
        for (int v = 0; v < V2; ++v) m[v] = __half22float2(mh[v]);
      } else {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          float2 x[V2];
          unpack<RAW, F16, V2>(buf[i], x);
#pragma unroll
          for (int v = 0; v < V2; ++v) {
            m[v].x = i ? fmaxf(m[v].x, x[v].x) : x[v].x;
            m[v].y = i ? fmaxf(m[v].y, x[v].y) : x[v].y;
          }
        }
      }
#pragma unroll
      for (int v = 0; v < V2; ++v) {
        const float2 kn = make_float2(fmaxf(K2[v].x, ceilf(m[v].x * kLog2e)), fmaxf(K2[v].y, ceilf(m[v].y * kLog2e)));
        const float2 r = make_float2(pow2_neg_f32(K2[v].x - kn.x), pow2_neg_f32(K2[v].y - kn.y));
        ts[v] = __fmul2_rn(ts[v], r); tx[v] = __fmul2_rn(tx[v], r); ty[v] = __fmul2_rn(ty[v], r);
        K2[v] = kn;
        nk[v] = make_float2(-kn.x, -kn.y);
      }
    }
    float2 gs[V2], gx[V2], gy[V2];
#pragma unroll
    for (int v = 0; v < V2; ++v) gs[v] = gx[v] = gy[v] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      float2 x[V2];
      unpack<RAW, F16, V2>(buf[i], x);
      load_word(buf[i], more);
      const float2 fh2 = make_float2(fh, fh), fw2 = make_float2(fw, fw);
#pragma unroll
      for (int v = 0; v < V2; ++v) {
        const float2 t = __ffma2_rn(x[v], l2e, nk[v]);
        const float2 e = make_float2(ptx::ex2_approx(t.x), ptx::ex2_approx(t.y));
        gs[v] = __fadd2_rn(gs[v], e);
        gx[v] = __ffma2_rn(e, fw2, gx[v]);
        gy[v] = __ffma2_rn(e, fh2, gy[v]);
      }
      fw += dr; fh += dq;
      if (fw >= Wf) { fw -= Wf; fh += 1.f; }
    }
#pragma unroll
    for (int v = 0; v < V2; ++v) {
      ts[v] = __fadd2_rn(ts[v], gs[v]); tx[v] = __fadd2_rn(tx[v], gx[v]); ty[v] = __fadd2_rn(ty[v], gy[v]);
    }
  }

  if (prof) stamp[2] = clock64();
  // ======================= per-thread records -> shared memory =======================
  if (live) {
    const int o = lane_px * C + c0;
#pragma unroll
    for (int v = 0; v < V2; ++v) {
      *reinterpret_cast<float2 *>(s_k + o + 2 * v) = K2[v];
      *reinterpret_cast<float2 *>(s_s + o + 2 * v) = ts[v];
      *reinterpret_cast<float2 *>(s_x + o + 2 * v) = tx[v];
      *reinterpret_cast<float2 *>(s_y + o + 2 * v) = ty[v];
    }
  }
  __syncthreads();
  if (p.cluster && p.splits > 1) ptx::cluster_wait();
  if (prof) stamp[3] = clock64();
  // a split crop's per-joint record goes to the first CTA of its cluster through distributed shared memory, or
  // -- when the launch has no clusters (split counts that are not 2, 4 or 8) -- to the global workspace
  auto emit_partial = [&](int j, double km, double a, double ax, double ay, double az) {
    if (p.cluster) {
      const uint32_t dst = ptx::mapa(ptx::smem_u32(s_part + (split * J + j) * 5), 0);
      ptx::st_cluster_f64(dst, km); ptx::st_cluster_f64(dst + 8, a); ptx::st_cluster_f64(dst + 16, ax);
      ptx::st_cluster_f64(dst + 24, ay); ptx::st_cluster_f64(dst + 32, az);
    } else {
      double *rec = p.partials + ((size_t(img) * p.splits + split) * J + j) * 5;
      rec[0] = km; rec[1] = a; rec[2] = ax; rec[3] = ay; rec[4] = az;
    }
  };
  // lanes x depth -> joint in one stage: `tpj` adjacent threads per joint (a power of two, all warps busy), each
  // takes its share of the joint's D x lanes records -- exact power-of-two weights, fp64 sums -- and a butterfly
  // over the tpj threads finishes the sum
  if (p.D == 8 && p.lanes == 8 && p.tpj == 16) {
    // the shape every reference configuration has (depth 8, 8 pixel lanes, 16 threads per joint): thread i of a
    // joint owns depth i & 7 and lanes (i >> 3) + {0, 2, 4, 6}; everything unrolled, one reciprocal per joint
    const int t_end = (J * 16 + 31) & ~31;
    for (int t = tid; t < t_end; t += blockDim.x) {
      const int j = t >> 4, i = t & 15, d = i & 7;
      const bool has = j < J;
      const int base = has ? (i >> 3) * C + d * J + j : 0;
      float k4[4], s4[4], x4[4], y4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int idx = base + q * 2 * C;
        k4[q] = s_k[idx]; s4[q] = s_s[idx]; x4[q] = s_x[idx]; y4[q] = s_y[idx];
      }
      float km = has ? fmaxf(fmaxf(k4[0], k4[1]), fmaxf(k4[2], k4[3])) : kNone;
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) km = fmaxf(km, __shfl_xor_sync(0xffffffffu, km, o));
      double a4[4], ax4[4], ay4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float wgt = has ? pow2_neg_f32(k4[q] - km) : 0.f;    // a scale by 2^-d is exact short of underflow
        a4[q] = double(wgt * s4[q]); ax4[q] = double(wgt * x4[q]); ay4[q] = double(wgt * y4[q]);
      }
      // (summing these four in fp32 instead saves ~170 cycles of the merge but doubles the error against the
      // float64 oracle, 1.8e-4 -> 3.4e-4 mm: not taken)
      double a = (a4[0] + a4[1]) + (a4[2] + a4[3]);
      double ax = (ax4[0] + ax4[1]) + (ax4[2] + ax4[3]);
      double ay = (ay4[0] + ay4[1]) + (ay4[2] + ay4[3]);
      double az = double(d) * a;
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) {
        a += shfl_xor_f64(a, o); ax += shfl_xor_f64(ax, o); ay += shfl_xor_f64(ay, o); az += shfl_xor_f64(az, o);
      }
      if (has && i == 0) {
        if (p.splits > 1) {
          emit_partial(j, double(km), a, ax, ay, az);
        } else {
          const double inv = recip(a);
          s_c01[3 * j] = ax * inv * p.mul_x;
          s_c01[3 * j + 1] = ay * inv * p.mul_y;
          s_c01[3 * j + 2] = az * inv * p.mul_z;
        }
      }
    }
  } else {
    const int tpj = p.tpj, nrec = p.D * p.lanes;
    const int t_end = (J * tpj + 31) & ~31;                      // whole warps: every lane of a shuffle takes part
    const int step_l = tpj / p.D, step_d = tpj - step_l * p.D;
    for (int t = tid; t < t_end; t += blockDim.x) {
      const int j = t / tpj, i = t - j * tpj;
      const bool has = j < J;
      const int l0 = i / p.D, d0 = i - l0 * p.D;
      float km = kNone;
      if (has) {
        int l = l0, d = d0;
        for (int r = i; r < nrec; r += tpj) {
          km = fmaxf(km, s_k[l * C + d * J + j]);
          l += step_l; d += step_d;
          if (d >= p.D) { d -= p.D; ++l; }
        }
      }
      for (int o = tpj >> 1; o >= 1; o >>= 1) km = fmaxf(km, __shfl_xor_sync(0xffffffffu, km, o));
      double a = 0.0, ax = 0.0, ay = 0.0, az = 0.0;
      if (has) {
        int l = l0, d = d0;
        for (int r = i; r < nrec; r += tpj) {
          const int idx = l * C + d * J + j;
          const float wgt = pow2_neg_f32(s_k[idx] - km);         // a scale by 2^-d is exact short of underflow
          const double sv = double(wgt * s_s[idx]);
          a += sv; ax += double(wgt * s_x[idx]); ay += double(wgt * s_y[idx]); az += double(d) * sv;
          l += step_l; d += step_d;
          if (d >= p.D) { d -= p.D; ++l; }
        }
      }
      for (int o = tpj >> 1; o >= 1; o >>= 1) {
        a += shfl_xor_f64(a, o); ax += shfl_xor_f64(ax, o); ay += shfl_xor_f64(ay, o); az += shfl_xor_f64(az, o);
      }
      if (has && i == 0) {
        if (p.splits > 1) {
          emit_partial(j, double(km), a, ax, ay, az);
        } else {
          // expectation of linspace(0,1,n) along each axis == E[index]/(n-1); mul_* carry 1/(n-1) and mm
          const double inv = recip(a);
          s_c01[3 * j] = ax * inv * p.mul_x;
          s_c01[3 * j + 1] = ay * inv * p.mul_y;
          s_c01[3 * j + 2] = az * inv * p.mul_z;
        }
      }
    }
  }
  __syncthreads();
  if (prof) stamp[4] = clock64();
  if (p.splits > 1 && p.cluster) {
    // the `splits` CTAs of a crop form one thread-block cluster: records were stored into CTA 0's shared memory
    // above; one cluster barrier (release / acquire) later CTA 0 merges them, the others are done
    ptx::cluster_sync();
    if (split != 0) { ptx::stamp_end(p.tstamp); return; }
    if (tid < J) {
      double gk = s_part[tid * 5];
      for (int sp = 1; sp < p.splits; ++sp) gk = fmax(gk, s_part[(sp * J + tid) * 5]);
      double TS = 0.0, TX = 0.0, TY = 0.0, TZ = 0.0;
      for (int sp = 0; sp < p.splits; ++sp) {
        const double *rec = s_part + (sp * J + tid) * 5;
        const double wgt = pow2_neg(float(rec[0] - gk));
        TS += wgt * rec[1]; TX += wgt * rec[2]; TY += wgt * rec[3]; TZ += wgt * rec[4];
      }
      const double inv = recip(TS);
      s_c01[3 * tid] = TX * inv * p.mul_x;
      s_c01[3 * tid + 1] = TY * inv * p.mul_y;
      s_c01[3 * tid + 2] = TZ * inv * p.mul_z;
    }
    __syncthreads();
  } else if (p.splits > 1) {
    // the last CTA of the crop merges the per-split records.  bar.sync above orders the J writers before
    // thread 0, whose gpu-scope fence is cumulative over what it has observed.
    if (tid == 0) {
      __threadfence();
      const unsigned int ticket = atomicAdd(p.counters + img, 1u);
      s_is_last = (ticket == unsigned(p.splits - 1));
      if (s_is_last) {
        __threadfence();
        p.counters[img] = 0;                           // self-cleaning for the next launch
      }
    }
    __syncthreads();
    if (!s_is_last) { ptx::stamp_end(p.tstamp); return; }
    if (tid < J) {
      const double *recs = p.partials + (size_t(img) * p.splits) * J * 5;
      double TS = 0.0, TX = 0.0, TY = 0.0, TZ = 0.0;
      if (p.splits <= 4) {
        // all records in flight at once (one L2 round trip instead of two dependent passes)
        double r[4][5];
#pragma unroll
        for (int sp = 0; sp < 4; ++sp) {
          const double *rec = recs + (size_t(sp < p.splits ? sp : 0) * J + tid) * 5;
#pragma unroll
          for (int q = 0; q < 5; ++q) r[sp][q] = __ldcg(rec + q);
        }
        double gk = r[0][0];
#pragma unroll
        for (int sp = 1; sp < 4; ++sp) gk = sp < p.splits ? fmax(gk, r[sp][0]) : gk;
#pragma unroll
        for (int sp = 0; sp < 4; ++sp) {
          const double wgt = sp < p.splits ? pow2_neg(float(r[sp][0] - gk)) : 0.0;
          TS += wgt * r[sp][1]; TX += wgt * r[sp][2]; TY += wgt * r[sp][3]; TZ += wgt * r[sp][4];
        }
      } else {
        double gk = double(kNone);
        for (int sp = 0; sp < p.splits; ++sp) gk = fmax(gk, __ldcg(recs + (size_t(sp) * J + tid) * 5));
        for (int sp = 0; sp < p.splits; ++sp) {
          const double *rec = recs + (size_t(sp) * J + tid) * 5;
          const double wgt = pow2_neg(float(__ldcg(rec) - gk));
          TS += wgt * __ldcg(rec + 1); TX += wgt * __ldcg(rec + 2);
          TY += wgt * __ldcg(rec + 3); TZ += wgt * __ldcg(rec + 4);
        }
      }
      const double inv = recip(TS);
      s_c01[3 * tid] = TX * inv * p.mul_x;
      s_c01[3 * tid + 1] = TY * inv * p.mul_y;
      s_c01[3 * tid + 2] = TZ * inv * p.mul_z;
    }
    __syncthreads();
  }
  if (p.out) {
    if (tid < p.n_out * 3)                // n_out <= 64 and the CTA has >= 192 threads... or loops below
      p.out[(size_t(img) * p.n_out) * 3 + tid] = float(s_c01[out_src] - s_c01[3 * p.root + out_ax]);
    for (int i = tid + blockDim.x; i < p.n_out * 3; i += blockDim.x) {
      const int jo = i / 3, a = i - 3 * jo;
      p.out[(size_t(img) * p.n_out) * 3 + i] = float(s_c01[3 * p.perm[jo] + a] - s_c01[3 * p.root + a]);
    }
  }
  if (p.coords01) {                       // heatmap coordinates in [0,1], model joint order (volumetric.py:234)
    for (int i = tid; i < J * 3; i += blockDim.x) {
      const int a = i % 3;
      p.coords01[size_t(img) * J * 3 + i] = float(s_c01[i] * (a == 0 ? p.unmul_x : (a == 1 ? p.unmul_y : p.unmul_z)));
    }
  }
  if (prof) { stamp[5] = clock64(); stamp[7] = (long long)ptx::globaltimer(); }
  ptx::stamp_end(p.tstamp);
}

size_t smem_bytes(const SoftargmaxLaunch &L) {
  return size_t(4) * L.lanes * L.C * sizeof(float) + size_t(3) * L.J * sizeof(double) +
         (L.cluster ? size_t(L.splits) * L.J * 5 * sizeof(double) : 0) + 16;
}

template <int VEC, bool F16, int CH, int MAXT, int MINB>
metro_status launch_t(const SoftargmaxLaunch &L, cudaStream_t stream) {
  static PerDeviceOnce configured;     // function attributes are per device
  const size_t sm = smem_bytes(L);     // softargmax_plan caps it at 100 KB
  metro_status cst = configured.run([] {
    METRO_CUDA(cudaFuncSetAttribute(softargmax_kernel<VEC, F16, CH, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    return METRO_OK;
  });
  if (cst != METRO_OK) return cst;
  const dim3 grid(unsigned(L.n) * unsigned(L.splits)), block(unsigned((L.slots * L.lanes + 31) & ~31));
  SoftargmaxLaunch LL = L;
  {
    // L2 prefetch ahead of the dependency wait: only where it can pay -- one CTA per crop (a split crop's cluster
    // cannot start early) and an input that fits L2 several times over (measured: 318 MB inputs thrash, -20 %)
    static PerDeviceOnce sm_once;
    static int sms[kMaxDevices];
    int dev = 0;
    cudaGetDevice(&dev);
    sm_once.run([&] { METRO_CUDA(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev)); return METRO_OK; });
    const size_t total = size_t(L.n) * L.H * L.W * L.C * (L.head_f16 ? 2 : 4);
    if (L.splits != 1 || total > size_t(48) << 20) LL.l2_prefetch = 0;
    const int slots = MINB * sms[dev];
    // (whole-input prefetch by the early CTAs: measured no better than each CTA prefetching its own item -- 9.35 against
    // 9.13 us at 256 crops -- so it is opt-in)
    static const bool no_early = std::getenv("METRO_SAM_EARLY") == nullptr;
    LL.early = (LL.l2_prefetch && !no_early && int(grid.x) < slots && int(grid.x) > slots / 2) ? slots - int(grid.x) : 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = sm; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  static const bool no_pdl = std::getenv("METRO_NO_PDL") != nullptr;   // A/B switch shared with the conv launches
  if (!no_pdl) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  if (L.cluster) {                       // the CTAs of one crop are consecutive in the grid
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = unsigned(L.splits);
    attr[n_attr].val.clusterDim.y = 1;
    attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  }
  cfg.attrs = attr; cfg.numAttrs = unsigned(n_attr);
  static const bool want_prof = std::getenv("METRO_SAM_PROF") != nullptr;
  if (!want_prof) {
    METRO_CUDA(cudaLaunchKernelEx(&cfg, softargmax_kernel<VEC, F16, CH, MAXT, MINB>, LL));
    return METRO_OK;
  }
  // debug: per-CTA phase durations in SM clocks (launch -> dependency wait -> stream -> records -> merge -> output)
  SoftargmaxLaunch P = LL;
  const size_t n_ctas = size_t(grid.x);
  METRO_CUDA(cudaMalloc(&P.prof, n_ctas * 8 * sizeof(long long)));
  METRO_CUDA(cudaMemsetAsync(P.prof, 0, n_ctas * 8 * sizeof(long long), stream));
  METRO_CUDA(cudaLaunchKernelEx(&cfg, softargmax_kernel<VEC, F16, CH, MAXT, MINB>, P));
  METRO_CUDA(cudaStreamSynchronize(stream));
  std::vector<long long> hst(n_ctas * 8);
  METRO_CUDA(cudaMemcpy(hst.data(), P.prof, hst.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  METRO_CUDA(cudaFree(P.prof));
  double sum[5] = {0, 0, 0, 0, 0}, mx[5] = {0, 0, 0, 0, 0};
  for (size_t c = 0; c < n_ctas; ++c)
    for (int i = 0; i < 5; ++i) {
      const double d = double(hst[c * 8 + i + 1] - hst[c * 8 + i]);
      sum[i] += d; if (d > mx[i]) mx[i] = d;
    }
  long long t_first = 0, t_last = 0, t_first_end = 0;
  for (size_t c = 0; c < n_ctas; ++c) {
    const long long b = hst[c * 8 + 6], e = hst[c * 8 + 7];
    if (b == 0 || e == 0) continue;                      // a CTA that left early (split crop, not the finisher)
    if (t_first == 0 || b < t_first) t_first = b;
    if (e > t_last) t_last = e;
    if (t_first_end == 0 || e < t_first_end) t_first_end = e;
  }
  std::fprintf(stderr, "[metro sam prof] first CTA start -> last CTA end %.2f us (first CTA end after %.2f us)\n",
               double(t_last - t_first) * 1e-3, double(t_first_end - t_first) * 1e-3);
  std::fprintf(stderr, "[metro sam prof] %zu CTAs x %u threads, cycles avg/max: wait %.0f/%.0f stream %.0f/%.0f records %.0f/%.0f merge %.0f/%.0f out %.0f/%.0f\n",
               n_ctas, block.x, sum[0] / n_ctas, mx[0], sum[1] / n_ctas, mx[1], sum[2] / n_ctas, mx[2], sum[3] / n_ctas, mx[3],
               sum[4] / n_ctas, mx[4]);
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
  // fp16 heads default to 8-byte words: twice the threads per byte, the decode of a half-width head being
  // bound by the exp per element rather than by HBM
  int vec = (d.word_bytes == 8 || (d.word_bytes == 0 && f16) ? 8 : 16) / (f16 ? 2 : 4);
  if (d.word_bytes == 0 && (d.n_joints_model * d.depth) % vec != 0 && vec * (f16 ? 2 : 4) == 16) vec /= 2;
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
  L.unmul_x = L.mul_x != 0.0 ? 1.0 / (L.mul_x * double(L.W - 1)) : 0.0;
  L.unmul_y = L.mul_y != 0.0 ? 1.0 / (L.mul_y * double(L.H - 1)) : 0.0;
  L.unmul_z = L.mul_z != 0.0 ? 1.0 / (L.mul_z * double(L.D - 1)) : 0.0;
  L.head_f16 = d.head_dtype == METRO_F16;
  L.slots = L.C / vec;
  L.vec = vec;
  if (L.slots > kMaxThreads) return fail(METRO_ERR_VALUE, "softargmax: too many head channels (%d)", L.C);
  const int P = L.H * L.W;
  // CTA shape: `slots` channel slots x `lanes` pixel lanes (a power of two): the smallest that gives the CTA
  // at least 256 threads, so that two resident CTAs put ~18 warps on an SM
  int lanes = 1;
  if (d.lanes > 0) {
    while (lanes * 2 <= d.lanes && lanes < 32) lanes *= 2;   // rounded down to a power of two
  } else {
    while (lanes < 32 && L.slots * lanes < 256) lanes *= 2;
  }
  while (lanes > 1 && ((L.slots * lanes + 31) & ~31) > kMaxThreads) lanes /= 2;
  if (((L.slots * lanes + 31) & ~31) > kMaxThreads) return fail(METRO_ERR_VALUE, "softargmax: too many head channels (%d)", L.C);
  L.lanes = lanes;
  {
    // threads per joint in the merge: the largest power of two <= 32 that the CTA's threads cover for all joints
    const int threads = (L.slots * lanes + 31) & ~31;
    L.tpj = 32;
    while (L.tpj > 1 && L.J * L.tpj > threads) L.tpj /= 2;
  }
  // work items: one per crop, or -- for heatmaps above 256 pixels -- row-major pixel ranges: four per crop up to
  // 32x32, two above (a CTA lives ~2500 cycles beyond its streaming time, so items must be long and one wave of
  // CTAs is the target: 64 crops x 4 at stride 8, 128 crops x 2 at stride 4 are the per-GPU shards of the
  // reference configurations; measured 10.7 us against 13.1 with two ranges at 32x32, 54.9 against 57.9 with
  // four at 64x64).  The number of ranges depends on the heatmap shape ONLY, so a crop's result is bit-identical
  // whatever batch or GPU shard it is part of.  The CTAs of a crop form a thread-block cluster and merge through
  // distributed shared memory.
  int item_px = kItemPixels, max_splits = P > 1024 ? 2 : 4;
  if (const char *e = getenv("METRO_SAM_ITEM_PX")) item_px = atoi(e) > 0 ? atoi(e) : item_px;
  if (const char *e = getenv("METRO_SAM_MAX_SPLITS")) max_splits = atoi(e) > 0 ? atoi(e) : max_splits;
  int splits = d.splits > 0 ? d.splits : (P + item_px - 1) / item_px;
  if (d.splits <= 0 && splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > P) splits = P;
  L.ipx = (P + splits - 1) / splits;
  L.splits = (P + L.ipx - 1) / L.ipx;              // no empty item
  // 2, 4 or 8 CTAs per crop run as one thread-block cluster and merge through distributed shared memory
  L.cluster = (L.splits == 2 || L.splits == 4 || L.splits == 8) && !getenv("METRO_SAM_NO_CLUSTER") ? 1 : 0;
  L.l2_prefetch = getenv("METRO_SAM_NO_PREFETCH") ? 0 : 1;
  if (smem_bytes(L) > 100 * 1024) return fail(METRO_ERR_VALUE, "softargmax: shared memory budget exceeded");
  return METRO_OK;
}

size_t softargmax_workspace_bytes(const SoftargmaxLaunch &L) {
  const size_t counters = (size_t(L.n) * 4 + 255) & ~size_t(255);
  return counters + size_t(L.n) * L.splits * L.J * 5 * 8;
}

metro_status softargmax_launch(const SoftargmaxLaunch &L, cudaStream_t stream) {
  if (L.n == 0) return METRO_OK;
  const int threads = (L.slots * L.lanes + 31) & ~31;
  // two CTAs per SM for the common shapes (<= 304 threads: 19 joints x 8 depths, 16-byte words, 8 lanes), one for wider ones
#define METRO_SAM(V, H, CH) return (threads <= 304 ? launch_t<V, H, CH, 304, 2>(L, stream) : launch_t<V, H, CH, kMaxThreads, 1>(L, stream))
  if (L.head_f16) {
    if (L.vec == 8) { METRO_SAM(8, true, 4); } else { METRO_SAM(4, true, 8); }
  } else {
    if (L.vec == 4) { METRO_SAM(4, false, 8); } else { METRO_SAM(2, false, 8); }
  }
#undef METRO_SAM
  return fail(METRO_ERR_INTERNAL, "softargmax: no kernel for this word size");
}

}  // namespace metro
