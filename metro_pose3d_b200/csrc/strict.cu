// Strict-precision path: the exported graph evaluated in float64 on CUDA cores.
//
// The reference exports its graph in FLAGS.dtype = float16 (default) or float32 (src/options.py:73,
// src/init.py:54-59, src/model/architectures.py:29).  The tensor-core path (conv_gemm.cu) is the float16 graph;
// this file is the precision reference INSIDE the product: the same layer plan (plan.cpp), every tensor float64
// (or, quant = 1, float64 arithmetic with the float16 roundings of the default graph applied at its storage
// points), so that |strict - exact graph| is bounded by float64 summation order (~1e-10 mm) instead of the
// ~5e-3 mm a float32 evaluation leaves (measured with the oracle: |fp32 - fp64| = 3-5e-3 mm end to end).
// It is ~50x slower than the tensor-core path and exists for verification and for callers who need the
// float32 export's accuracy or better.
//
// Layers follow resnet_v2.py:142-241 / resnet_utils.py:64-185 exactly as csrc/metro_api.cu builds them:
//   image cast -> conv1 7x7/2 + bias -> zero-padded 3x3/2 max-pool -> units { preact BN+ReLU; shortcut;
//   1x1 BN ReLU; 3x3 (stride / rate / centred) BN ReLU; 1x1 + bias; add } -> postnorm -> logits -> decode
//   (volumetric.py:227-235, tfu.py:466-499, volumetric.py:288-306, tfu3d.py:23-25, main.py:127).
#include <cuda_fp16.h>

#include <cmath>
#include <cstring>
#include <map>
#include <memory>

#include "strict.h"

namespace metro {

namespace {

constexpr double kBnEpsStrict = 1e-5;   // architectures.py:10
constexpr int BM = 64, BN = 64, BK = 16;

struct SConv {
  const double *x; int in_side, cin;
  const double *w;                       // [k*k*cin][cout], HWIO order
  int k, stride, rate, pad_lo, out_side, cout;
  const double *scale, *shift;           // y = acc * scale + shift (scale == nullptr: y = acc + shift)
  const double *res; int res_side, res_stride, res_shift;   // + res[n, oh*rs + sh, ow*rs + sh, c]
  int relu, quant;                       // quant: 1 = round to float16, 2 = round to float32
  double *y;
  int n;
};

__device__ __forceinline__ double quantize(double v, int quant) {
  // float64 -> float32 -> float16: the route of the tensor-core path (float32 accumulator, then cvt.rn.f16.f32) and
  // of the oracle's 'half' mode (torch converts double to half through float)
  if (quant == 1) return double(__half2float(__float2half_rn(float(v))));
  if (quant == 2) return double(float(v));
  return v;
}

// Tiled direct convolution as a GEMM over the flattened (tap, channel) index: 64 pixels x 64 channels per
// block, 4 x 4 outputs per thread.
__global__ void __launch_bounds__(256) strict_conv_kernel(const SConv p) {
  __shared__ double As[BK][BM + 2];
  __shared__ double Bs[BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long M = (long long)p.n * p.out_side * p.out_side;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int K = p.k * p.k * p.cin;
  // the four pixels whose A elements this thread loads (fixed for the whole K loop)
  int a_img[4], a_oh[4], a_ow[4];
  bool a_ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + (tid >> 4) + 16 * i;
    a_ok[i] = m < M;
    const long long mm = a_ok[i] ? m : 0;
    a_img[i] = int(mm / (p.out_side * p.out_side));
    const int r = int(mm - (long long)a_img[i] * p.out_side * p.out_side);
    a_oh[i] = r / p.out_side; a_ow[i] = r - a_oh[i] * p.out_side;
  }
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int k0 = 0; k0 < K; k0 += BK) {
    {
      const int kk = k0 + (tid & 15);
      int kh = 0, kw = 0, ci = 0;
      const bool kok = kk < K;
      if (kok) {
        const int tap = kk / p.cin;
        ci = kk - tap * p.cin; kh = tap / p.k; kw = tap - kh * p.k;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double v = 0.0;
        if (kok && a_ok[i]) {
          const int ih = a_oh[i] * p.stride + kh * p.rate - p.pad_lo, iw = a_ow[i] * p.stride + kw * p.rate - p.pad_lo;
          if (ih >= 0 && ih < p.in_side && iw >= 0 && iw < p.in_side)
            v = p.x[(((long long)a_img[i] * p.in_side + ih) * p.in_side + iw) * p.cin + ci];
        }
        As[tid & 15][(tid >> 4) + 16 * i] = v;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + 256 * i, nl = e & 63, kl = e >> 6;
        const int c = n0 + nl, kq = k0 + kl;
        Bs[kl][nl] = (c < p.cout && kq < K) ? p.w[(long long)kq * p.cout + c] : 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kl = 0; kl < BK; ++kl) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kl][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kl][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int img = int(m / (p.out_side * p.out_side));
    const int r = int(m - (long long)img * p.out_side * p.out_side);
    const int oh = r / p.out_side, ow = r - oh * p.out_side;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= p.cout) continue;
      double v = acc[i][j];
      v = p.scale ? v * p.scale[c] + p.shift[c] : v + p.shift[c];
      if (p.res)
        v += p.res[(((long long)img * p.res_side + oh * p.res_stride + p.res_shift) * p.res_side + ow * p.res_stride +
                    p.res_shift) * p.cout + c];
      if (p.relu) v = fmax(v, 0.0);
      p.y[m * p.cout + c] = quantize(v, p.quant);
    }
  }
}

// y = relu(x * scale[c] + shift[c])   (pre-activation / postnorm, resnet_v2.py:119,229)
__global__ void strict_bn_relu_kernel(const double *x, const double *scale, const double *shift, double *y, long long total,
                                      int c, int quant) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = int(i % c);
  y[i] = quantize(fmax(x[i] * scale[ch] + shift[ch], 0.0), quant);
}

// zero-padded 3x3 / 2 max-pool (resnet_utils.py:177-185: the pad row / column of zeros takes part in the maximum)
__global__ void strict_pool_kernel(const double *x, double *y, int n, int in_side, int out_side, int c) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * out_side * out_side * c;
  if (i >= total) return;
  const int ch = int(i % c);
  long long r = i / c;
  const int pw = int(r % out_side); r /= out_side;
  const int ph = int(r % out_side);
  const int img = int(r / out_side);
  double m = -INFINITY;
  for (int dh = -1; dh <= 1; ++dh)
    for (int dw = -1; dw <= 1; ++dw) {
      const int h = 2 * ph + dh, w = 2 * pw + dw;
      double v = 0.0;                                      // the padding value
      if (h >= 0 && h < in_side && w >= 0 && w < in_side) v = x[(((long long)img * in_side + h) * in_side + w) * c + ch];
      else if (h >= in_side || w >= in_side) continue;     // VALID on the far side: never reached for even sides
      m = fmax(m, v);
    }
  y[i] = m;
}

// float32 / uint8 crops -> float64 (uint8: float32(k) / 255 as improc.py:56-61 computes it)
__global__ void strict_image_kernel(const void *img, int u8, double *y, long long total, int quant) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float v;
  if (u8) v = __fdiv_rn(float(static_cast<const unsigned char *>(img)[i]), 255.0f);
  else v = static_cast<const float *>(img)[i];
  y[i] = quantize(double(v), quant);
}

// softmax over (H, W, D) jointly and the three marginal expectations, one block per (crop, joint)
// (volumetric.py:227-235, tfu.py:466-499); head NHWC with channel c = d * J + j
__global__ void __launch_bounds__(256) strict_decode_kernel(const double *head, double *coords, int H, int W, int D, int J) {
  __shared__ double red[4][256];
  const int j = blockIdx.x % J, img = blockIdx.x / J;
  const int tid = threadIdx.x, total = H * W * D, C = D * J;
  const double *base = head + (long long)img * H * W * C;
  double mx = -INFINITY;
  for (int i = tid; i < total; i += 256) {
    const int d = i % D, px = i / D;
    mx = fmax(mx, base[(long long)px * C + d * J + j]);
  }
  red[0][tid] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) red[0][tid] = fmax(red[0][tid], red[0][tid + s]);
    __syncthreads();
  }
  mx = red[0][0];
  __syncthreads();
  double s0 = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
  for (int i = tid; i < total; i += 256) {
    const int d = i % D, px = i / D, h = px / W, w = px - h * W;
    const double e = exp(base[(long long)px * C + d * J + j] - mx);
    s0 += e; sx += e * w; sy += e * h; sz += e * d;
  }
  red[0][tid] = s0; red[1][tid] = sx; red[2][tid] = sy; red[3][tid] = sz;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s)
      for (int q = 0; q < 4; ++q) red[q][tid] += red[q][tid + s];
    __syncthreads();
  }
  if (tid == 0) {
    double *o = coords + ((long long)img * J + j) * 3;
    // expectation of linspace(0, 1, n) along each axis
    o[0] = W > 1 ? red[1][0] / red[0][0] / double(W - 1) : 0.0;
    o[1] = H > 1 ? red[2][0] / red[0][0] / double(H - 1) : 0.0;
    o[2] = D > 1 ? red[3][0] / red[0][0] / double(D - 1) : 0.0;
  }
}

// heatmap_to_metric (volumetric.py:288-306), root_relative (tfu3d.py:23-25), gather (main.py:127)
__global__ void strict_metric_kernel(const double *coords, float *out, int n, int J, int n_out, const int *perm, double lrc,
                                     double add_xy, double box, double proc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * n_out * 3) return;
  const int ax = i % 3, jo = (i / 3) % n_out, img = i / (3 * n_out);
  auto metric = [&](int j) {
    const double c = coords[((long long)img * J + j) * 3 + ax];
    return ax < 2 ? (c * lrc + add_xy) * box / proc : c * box;
  };
  out[i] = float(metric(perm[jo]) - metric(J - 1));
}

struct Arena {
  std::vector<void *> ptrs;
  size_t total = 0;
  ~Arena() { for (void *p : ptrs) cudaFree(p); }
  metro_status alloc(double **out, size_t elems) {
    void *p = nullptr;
    const size_t bytes = (elems ? elems : 1) * sizeof(double);
    const cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(METRO_ERR_NOMEM, "strict: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    ptrs.push_back(p); total += bytes; *out = static_cast<double *>(p);
    return METRO_OK;
  }
  metro_status upload(double **out, const std::vector<double> &h) {
    metro_status st = alloc(out, h.size());
    if (st != METRO_OK) return st;
    METRO_CUDA(cudaMemcpy(*out, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    return METRO_OK;
  }
};

}  // namespace

struct StrictNet {
  NetPlan plan;
  int max_batch = 0, quant = 0;
  double lrc = 0, add_xy = 0, box = 0;
  Arena arena;
  struct Step {
    int kind = 0;          // 0 conv, 1 bn_relu, 2 pool
    SConv conv{};          // kind 0 (n filled per run; pointers are for crop 0)
    const double *x = nullptr, *scale = nullptr, *shift = nullptr; double *y = nullptr;
    size_t elems_per_crop = 0; int c = 0, in_side = 0, out_side = 0, quant = 0;
    size_t x_per_crop = 0, y_per_crop = 0, res_per_crop = 0;
  };
  std::vector<Step> steps;
  double *img = nullptr, *head = nullptr, *coords = nullptr;
  int *d_perm = nullptr;
  int n_out = 0;
  std::map<std::string, std::pair<const double *, size_t>> debug;
};

namespace {

// gamma, beta, mean, var -> scale, shift.  quant: the tensor-core path holds them as float32 (computed in double)
void bn_affine64(const float *bn, int c, bool f32, std::vector<double> &scale, std::vector<double> &shift) {
  scale.resize(c); shift.resize(c);
  const float *g = bn, *b = bn + c, *m = bn + 2 * c, *v = bn + 3 * c;
  for (int i = 0; i < c; ++i) {
    const double s = double(g[i]) / std::sqrt(double(v[i]) + kBnEpsStrict);
    const double f = double(b[i]) - double(m[i]) * s;
    scale[i] = f32 ? double(float(s)) : s;
    shift[i] = f32 ? double(float(f)) : f;
  }
}

std::vector<double> weights64(const float *w, size_t n, bool f16) {
  std::vector<double> o(n);
  for (size_t i = 0; i < n; ++i) o[i] = f16 ? double(__half2float(__float2half_rn(w[i]))) : double(w[i]);
  return o;
}

}  // namespace

metro_status strict_build(const NetPlan &pl, const float *blob, int max_batch, int quant, float box_size_mm,
                          const std::vector<int32_t> &perm, bool keep, StrictNet **out) {
  std::unique_ptr<StrictNet> net(new StrictNet());
  net->plan = pl; net->max_batch = max_batch; net->quant = quant;
  Arena &A = net->arena;
  const size_t N = size_t(max_batch);
  const bool q = quant != 0;
  metro_status st;
  auto add_conv = [&](const ConvGeom &c, const double *x, const double *scale, const double *shift, bool relu, int qmode,
                      double *y, const double *res, int res_side, int res_stride, int res_shift) -> metro_status {
    StrictNet::Step s;
    s.kind = 0;
    double *w = nullptr;
    metro_status r = A.upload(&w, weights64(blob + c.w_off, size_t(c.k) * c.k * c.cin * c.cout, q));
    if (r != METRO_OK) return r;
    s.conv.x = x; s.conv.in_side = c.in_side; s.conv.cin = c.cin; s.conv.w = w; s.conv.k = c.k; s.conv.stride = c.stride;
    s.conv.rate = c.rate; s.conv.pad_lo = c.pad_lo; s.conv.out_side = c.out_side; s.conv.cout = c.cout;
    s.conv.scale = scale; s.conv.shift = shift; s.conv.res = res; s.conv.res_side = res_side;
    s.conv.res_stride = res_stride; s.conv.res_shift = res_shift; s.conv.relu = relu ? 1 : 0; s.conv.quant = qmode;
    s.conv.y = y;
    s.x_per_crop = size_t(c.in_side) * c.in_side * c.cin;
    s.y_per_crop = size_t(c.out_side) * c.out_side * c.cout;
    s.res_per_crop = size_t(res_side) * res_side * c.cout;
    net->steps.push_back(s);
    return METRO_OK;
  };
  auto add_bn = [&](const double *x, int c, size_t elems_per_crop, const float *bn, double *y) -> metro_status {
    std::vector<double> sc, sf;
    bn_affine64(bn, c, q, sc, sf);
    double *dsc = nullptr, *dsf = nullptr;
    metro_status r;
    if ((r = A.upload(&dsc, sc)) != METRO_OK || (r = A.upload(&dsf, sf)) != METRO_OK) return r;
    StrictNet::Step s;
    s.kind = 1; s.x = x; s.scale = dsc; s.shift = dsf; s.y = y; s.c = c; s.elems_per_crop = elems_per_crop; s.quant = q ? 1 : 0;
    net->steps.push_back(s);
    return METRO_OK;
  };
  auto vec64 = [&](const float *v, int c, double **d) -> metro_status {
    std::vector<double> h(v, v + c);
    return A.upload(d, h);
  };

  // ---- root ----
  const size_t img_e = size_t(pl.proc_side) * pl.proc_side * 3;
  if ((st = A.alloc(&net->img, img_e * N)) != METRO_OK) return st;
  double *conv1 = nullptr, *pool = nullptr, *pre0 = nullptr;
  const size_t conv1_e = size_t(pl.pool_in) * pl.pool_in * 64, pool_e = size_t(pl.pool_out) * pl.pool_out * 64;
  if ((st = A.alloc(&conv1, conv1_e * N)) != METRO_OK) return st;
  double *b = nullptr;
  if ((st = vec64(blob + pl.root.b_off, 64, &b)) != METRO_OK) return st;
  if ((st = add_conv(pl.root, net->img, nullptr, b, false, q ? 1 : 0, conv1, nullptr, 0, 0, 0)) != METRO_OK) return st;
  net->debug["conv1"] = {conv1, conv1_e};
  // ---- rotating buffers ----
  size_t raw_e = pool_e, r1_e = 0, r2_e = 0;
  for (const auto &u : pl.units) {
    raw_e = std::max(raw_e, size_t(u.out_side) * u.out_side * u.depth);
    r1_e = std::max(r1_e, size_t(u.in_side) * u.in_side * u.cb);
    r2_e = std::max(r2_e, size_t(u.out_side) * u.out_side * u.cb);
  }
  double *raw[2] = {nullptr, nullptr}, *pre[2] = {nullptr, nullptr}, *r1 = nullptr, *r2 = nullptr, *tmp = nullptr;
  if (!keep) {
    for (int i = 0; i < 2; ++i)
      if ((st = A.alloc(&raw[i], raw_e * N)) != METRO_OK || (st = A.alloc(&pre[i], raw_e * N)) != METRO_OK) return st;
    if ((st = A.alloc(&r1, r1_e * N)) != METRO_OK || (st = A.alloc(&r2, r2_e * N)) != METRO_OK) return st;
  }
  if ((st = A.alloc(&tmp, raw_e * N)) != METRO_OK) return st;
  if (keep) {
    if ((st = A.alloc(&pool, pool_e * N)) != METRO_OK || (st = A.alloc(&pre0, pool_e * N)) != METRO_OK) return st;
  } else { pool = raw[0]; pre0 = pre[0]; }
  {
    StrictNet::Step s;
    s.kind = 2; s.x = conv1; s.y = pool; s.c = 64; s.in_side = pl.pool_in; s.out_side = pl.pool_out;
    net->steps.push_back(s);
    net->debug["pool1"] = {pool, pool_e};
  }
  if ((st = add_bn(pool, 64, pool_e, blob + pl.units[0].preact_off, pre0)) != METRO_OK) return st;
  double *cur_raw = pool, *cur_pre = pre0;
  // ---- units ----
  for (size_t i = 0; i < pl.units.size(); ++i) {
    const UnitPlan &u = pl.units[i];
    const bool last = i + 1 == pl.units.size();
    double *b1 = r1, *b2 = r2, *nraw = raw[(i + 1) & 1], *npre = pre[(i + 1) & 1];
    const size_t e1 = size_t(u.in_side) * u.in_side * u.cb, e2 = size_t(u.out_side) * u.out_side * u.cb;
    const size_t eo = size_t(u.out_side) * u.out_side * u.depth;
    if (keep) {
      if ((st = A.alloc(&b1, e1 * N)) != METRO_OK || (st = A.alloc(&b2, e2 * N)) != METRO_OK ||
          (st = A.alloc(&nraw, eo * N)) != METRO_OK || (st = A.alloc(&npre, eo * N)) != METRO_OK)
        return st;
    }
    std::vector<double> sc, sf;
    double *dsc = nullptr, *dsf = nullptr;
    // conv1: 1x1 -> BN -> ReLU (resnet_v2.py:127-128)
    bn_affine64(blob + u.conv1.bn_off, u.cb, q, sc, sf);
    if ((st = A.upload(&dsc, sc)) != METRO_OK || (st = A.upload(&dsf, sf)) != METRO_OK) return st;
    if ((st = add_conv(u.conv1, cur_pre, dsc, dsf, true, q ? 1 : 0, b1, nullptr, 0, 0, 0)) != METRO_OK) return st;
    net->debug[u.name + "/conv1"] = {b1, e1};
    // conv2: 3x3 -> BN -> ReLU (:130-132)
    bn_affine64(blob + u.conv2.bn_off, u.cb, q, sc, sf);
    if ((st = A.upload(&dsc, sc)) != METRO_OK || (st = A.upload(&dsf, sf)) != METRO_OK) return st;
    if ((st = add_conv(u.conv2, b1, dsc, dsf, true, q ? 1 : 0, b2, nullptr, 0, 0, 0)) != METRO_OK) return st;
    net->debug[u.name + "/conv2"] = {b2, e2};
    // shortcut (:120-125) and conv3 + bias + add (:134-138); rounded once, after the sum
    double *b3 = nullptr;
    if ((st = vec64(blob + u.conv3.b_off, u.depth, &b3)) != METRO_OK) return st;
    if (u.proj) {
      double *bs = nullptr;
      if ((st = vec64(blob + u.shortcut.b_off, u.depth, &bs)) != METRO_OK) return st;
      if ((st = add_conv(u.shortcut, cur_pre, nullptr, bs, false, 0, tmp, nullptr, 0, 0, 0)) != METRO_OK) return st;
      if ((st = add_conv(u.conv3, b2, nullptr, b3, false, q ? 1 : 0, nraw, tmp, u.out_side, 1, 0)) != METRO_OK) return st;
    } else {
      if ((st = add_conv(u.conv3, b2, nullptr, b3, false, q ? 1 : 0, nraw, cur_raw, u.in_side, u.stride, u.shift)) != METRO_OK)
        return st;
    }
    net->debug[u.name + "/out"] = {nraw, eo};
    const int64_t next_bn = last ? pl.postnorm_off : pl.units[i + 1].preact_off;
    if ((st = add_bn(nraw, u.depth, eo, blob + next_bn, npre)) != METRO_OK) return st;
    net->debug[u.name + "/pre"] = {npre, eo};
    cur_raw = nraw; cur_pre = npre;
  }
  // ---- logits (:234-236); the float16 graph casts the head to float32 (architectures.py:34) ----
  const size_t head_e = size_t(pl.feat_side) * pl.feat_side * pl.logits.cout;
  if ((st = A.alloc(&net->head, head_e * N)) != METRO_OK) return st;
  double *bl = nullptr;
  if ((st = vec64(blob + pl.logits.b_off, pl.logits.cout, &bl)) != METRO_OK) return st;
  if ((st = add_conv(pl.logits, cur_pre, nullptr, bl, false, q ? 2 : 0, net->head, nullptr, 0, 0, 0)) != METRO_OK) return st;
  net->debug["head"] = {net->head, head_e};
  if ((st = A.alloc(&net->coords, size_t(pl.n_joints) * 3 * N)) != METRO_OK) return st;
  {
    void *p = nullptr;
    METRO_CUDA(cudaMalloc(&p, perm.size() * sizeof(int)));
    A.ptrs.push_back(p);
    METRO_CUDA(cudaMemcpy(p, perm.data(), perm.size() * sizeof(int), cudaMemcpyHostToDevice));
    net->d_perm = static_cast<int *>(p);
    net->n_out = int(perm.size());
  }
  const int last_px = pl.proc_side - 1;                      // volumetric.py:288-291
  net->lrc = double(last_px - (last_px % pl.stride) - 1);
  net->add_xy = pl.centered ? double(pl.stride / 2) : 0.0;   // cancels in the root-relative difference
  net->box = double(box_size_mm);
  *out = net.release();
  return METRO_OK;
}

__global__ void strict_coords_kernel(const double *coords, float *out, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) out[i] = float(coords[i]);
}

metro_status strict_run(StrictNet *net, const void *images_dev, bool u8, int n, float *poses_dev, cudaStream_t s,
                        float *coords01_dev) {
  if (n == 0) return METRO_OK;
  const NetPlan &pl = net->plan;
  {
    const long long total = (long long)n * pl.proc_side * pl.proc_side * 3;
    strict_image_kernel<<<unsigned((total + 255) / 256), 256, 0, s>>>(images_dev, u8 ? 1 : 0, net->img, total, net->quant ? 1 : 0);
  }
  for (const auto &st : net->steps) {
    if (st.kind == 0) {
      SConv c = st.conv;
      c.n = n;
      const long long M = (long long)n * c.out_side * c.out_side;
      dim3 grid(unsigned((M + BM - 1) / BM), unsigned((c.cout + BN - 1) / BN));
      strict_conv_kernel<<<grid, 256, 0, s>>>(c);
    } else if (st.kind == 1) {
      const long long total = (long long)n * st.elems_per_crop;
      strict_bn_relu_kernel<<<unsigned((total + 255) / 256), 256, 0, s>>>(st.x, st.scale, st.shift, st.y, total, st.c, st.quant);
    } else {
      const long long total = (long long)n * st.out_side * st.out_side * st.c;
      strict_pool_kernel<<<unsigned((total + 255) / 256), 256, 0, s>>>(st.x, st.y, n, st.in_side, st.out_side, st.c);
    }
  }
  strict_decode_kernel<<<unsigned(n * pl.n_joints), 256, 0, s>>>(net->head, net->coords, pl.feat_side, pl.feat_side, pl.depth,
                                                                  pl.n_joints);
  const int tot = n * net->n_out * 3;
  if (poses_dev)
    strict_metric_kernel<<<unsigned((tot + 127) / 128), 128, 0, s>>>(net->coords, poses_dev, n, pl.n_joints, net->n_out, net->d_perm,
                                                                     net->lrc, net->add_xy, net->box, double(pl.proc_side));
  if (coords01_dev) {
    const int tc = n * pl.n_joints * 3;
    strict_coords_kernel<<<unsigned((tc + 127) / 128), 128, 0, s>>>(net->coords, coords01_dev, tc);
  }
  METRO_CUDA(cudaGetLastError());
  return METRO_OK;
}

void strict_destroy(StrictNet *net) { delete net; }
size_t strict_bytes(const StrictNet *net) { return net->arena.total; }

bool strict_debug(const StrictNet *net, const std::string &name, const double **ptr, size_t *elems_per_crop) {
  auto it = net->debug.find(name);
  if (it == net->debug.end()) return false;
  *ptr = it->second.first; *elems_per_crop = it->second.second;
  return true;
}

}  // namespace metro
