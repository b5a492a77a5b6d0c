// C-ABI of libmetro.so (include/metro.h): plan -> device weights + arena + launch list -> infer.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "common.h"
#include "conv_gemm.h"
#include "plan.h"
#include "root_fused.h"
#include "strict.h"

namespace metro {

static thread_local std::string g_last_error;

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

metro_status fail(metro_status st, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return st;
}

namespace {

constexpr double kBnEps = 1e-5;   // architectures.py:10

// gamma, beta, mean, var (c each) -> scale, shift (computed in double, stored float)
void bn_affine(const float *bn, int c, std::vector<float> &scale, std::vector<float> &shift) {
  scale.resize(c); shift.resize(c);
  const float *g = bn, *b = bn + c, *m = bn + 2 * c, *v = bn + 3 * c;
  for (int i = 0; i < c; ++i) {
    const double s = double(g[i]) / std::sqrt(double(v[i]) + kBnEps);
    scale[i] = float(s);
    shift[i] = float(double(b[i]) - double(m[i]) * s);
  }
}

struct DeviceArena {
  std::vector<void *> ptrs;
  size_t total = 0, uploaded = 0;   // uploaded: constants (weights, per-channel vectors), independent of the batch
  ~DeviceArena() { for (void *p : ptrs) cudaFree(p); }
  metro_status alloc(void **out, size_t bytes) {
    if (bytes == 0) bytes = 16;
    void *p = nullptr;
    const cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(METRO_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    ptrs.push_back(p); total += bytes; *out = p;
    return METRO_OK;
  }
  template <typename T>
  metro_status upload(T **out, const std::vector<T> &host, size_t pad_to = 0) {
    const size_t n = pad_to > host.size() ? pad_to : host.size();
    void *p = nullptr;
    metro_status st = alloc(&p, n * sizeof(T));
    if (st != METRO_OK) return st;
    METRO_CUDA(cudaMemset(p, 0, n * sizeof(T)));
    METRO_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    uploaded += (n ? n : 1) * sizeof(T);
    *out = static_cast<T *>(p);
    return METRO_OK;
  }
};

struct GemmSpec {
  std::string name;
  int n_max = 0;
  const __half *src = nullptr; int in_side = 0, cin = 0;
  int k = 1, stride = 1, rate = 1, pad_lo = 0, out_side = 0, cout = 0;
  const float *w = nullptr;                 // host HWIO
  const __half *src2 = nullptr; int cin2 = 0; const float *w2 = nullptr;
  std::vector<float> scale, shift, scale2, shift2;
  std::vector<float> ascale, ashift;        // optional pre-activation applied to source 0 inside the kernel (1x1 only)
  bool relu = false;
  const __half *res = nullptr; int res_stride = 0, res_shift = 0, res_side = 0;
  void *out1 = nullptr; bool out1_f32 = false;
  __half *out2 = nullptr;
  // chained second convolution (conv_chain.cu): 1x1 on y2 = relu(fp16(y) * scale2 + shift2), BN + ReLU, -> out_c
  const float *w_c = nullptr; int cout_c = 0;
  std::vector<float> scale_c, shift_c;
  __half *out_c = nullptr;
};

metro_status build_gemm(DeviceArena &arena, const GemmSpec &g, ConvGemmLaunch &L) {
  if (g.cin % kTileK != 0 || g.cin2 % kTileK != 0)
    return fail(METRO_ERR_VALUE, "%s: input channels (%d, %d) must be multiples of %d", g.name.c_str(), g.cin, g.cin2, kTileK);
  if (g.cout % 8 != 0) return fail(METRO_ERR_VALUE, "%s: cout %d must be a multiple of 8", g.name.c_str(), g.cout);
  if (g.stride == 2 && (g.in_side % 2 != 0)) return fail(METRO_ERR_VALUE, "%s: stride 2 needs an even input side", g.name.c_str());
  L = ConvGemmLaunch();
  L.name = g.name;
  ConvGemmParams &p = L.prm;
  std::memset(&p, 0, sizeof p);
  L.direct = g.out1_f32 || (g.cout % 64 != 0);   // the logits head (136 / 152 channels, fp32 or fp16)
  L.block_n = conv_gemm_pick_block_n(g.cout, L.direct, (long long)g.n_max * g.out_side * g.out_side);
  L.chain = g.cout_c > 0;
  if (L.chain) {
    if (L.direct || g.k != 1 || g.stride != 1 || g.out2 || g.cout % 256 != 0 || (g.cout_c != 64 && g.cout_c != 128 && g.cout_c != 256))
      return fail(METRO_ERR_VALUE, "%s: this geometry cannot be chained", g.name.c_str());
    L.block_n = 128;                               // conv3 tiles of 128 channels (conv_chain.cu)
  }
  if (L.direct && (g.out2 || g.res)) return fail(METRO_ERR_VALUE, "%s: this output shape excludes a residual / second output", g.name.c_str());
  if (g.res && g.cin2) return fail(METRO_ERR_VALUE, "%s: identity and projection shortcuts are exclusive", g.name.c_str());
  if (g.res) {
    for (float sc : g.scale)
      if (sc != 1.0f) return fail(METRO_ERR_VALUE, "%s: the residual is accumulated before the scale; scale must be 1", g.name.c_str());
  }
  const int res_c = g.res ? g.cout : 0;           // identity shortcut rides the K loop as identity weights
  const int cout_pad = conv_gemm_cout_pad(g.cout, L.block_n);
  p.cout = g.cout;
  p.n_tiles = cout_pad / L.block_n;
  metro_status st = conv_gemm_geometry(p, g.out_side);
  if (st != METRO_OK) return st;
  int K;
  std::vector<__half> packed;
  st = conv_gemm_set_taps(p, g.k, g.stride, g.rate, g.pad_lo);
  if (st != METRO_OK) return st;
  p.cblk0 = g.cin / kTileK;
  p.cblk1 = (g.cin2 + res_c) / kTileK;
  p.diag2 = g.res ? 1 : 0;
  K = g.k * g.k * g.cin + g.cin2 + res_c;
  packed.resize(size_t(cout_pad) * K);
  conv_gemm_pack_weights(g.w, g.k, g.cin, g.cout, g.res ? nullptr : g.w2, g.cin2 + res_c, cout_pad, packed.data());
  __half *d_w = nullptr;
  st = arena.upload(&d_w, packed);
  if (st != METRO_OK) return st;
  st = make_weight_tensor_map(&p.bmap, d_w, cout_pad, K, L.block_n);
  if (st != METRO_OK) return st;
  if (g.res && (st = make_weight_tensor_map(&p.bidmap, d_w, cout_pad, K, 64)) != METRO_OK) return st;   // 32-row boxes
  // activations
  if (g.stride == 1) {
    st = make_act_tensor_map(&p.amap[0], g.src, g.n_max, g.in_side, g.in_side, g.cin, 1, 0, 0, p.wo, p.th, p.nb);
    if (st != METRO_OK) return st;
    // 3x3 SAME convolutions whose tile is whole rows of one crop: stage column-shifted tall boxes
    static const bool no_tall = getenv("METRO_NO_TALL") != nullptr;
    const int tall_rows = p.th + 2 * g.rate;
    const int tall_bytes = tall_rows * p.wo * 128 + 3 * (L.block_n / 2) * 128;
    if (!no_tall && g.k == 3 && p.nb == 1 && g.pad_lo == g.rate && !g.cin2 && !g.res && tall_rows <= 256 &&
        L.block_n <= 128 && tall_bytes <= 72 * 1024) {   // narrow tiles are the ones bound by L2->SMEM traffic
                                                         // (measured: no gain at 256 wide); >= 3 stages must fit
      p.tall = 1;
      p.tall_a_bytes = tall_rows * p.wo * 128;
      p.tall_row_step = g.rate * p.wo * 128;
      p.tall_stage_bytes = p.tall_a_bytes + 3 * (L.block_n / 2) * 128;
      st = make_act_tensor_map(&p.amap[1], g.src, g.n_max, g.in_side, g.in_side, g.cin, 1, 0, 0, p.wo, tall_rows, 1);
      if (st != METRO_OK) return st;
    }
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        st = make_act_tensor_map(&p.amap[ph * 2 + pw], g.src, g.n_max, g.in_side, g.in_side, g.cin, 2, ph, pw, p.wo,
                                 p.th, p.nb);
        if (st != METRO_OK) return st;
      }
  }
  if (g.cin2) {
    st = make_act_tensor_map(&p.a2map, g.src2, g.n_max, g.out_side, g.out_side, g.cin2, 1, 0, 0, p.wo, p.th, p.nb);
    if (st != METRO_OK) return st;
  } else if (g.res) {
    // identity shortcut: res[:, shift::stride, shift::stride, :] on the output grid (resnet_v2.py:120-121)
    st = make_act_tensor_map(&p.a2map, g.res, g.n_max, g.res_side, g.res_side, g.cout, g.res_stride, g.res_shift,
                             g.res_shift, p.wo, p.th, p.nb);
    if (st != METRO_OK) return st;
  }
  // epilogue vectors (padded to cout_pad so the kernel never reads out of range)
  float *d = nullptr;
  st = arena.upload(&d, g.scale, cout_pad); if (st != METRO_OK) return st; p.scale = d;
  st = arena.upload(&d, g.shift, cout_pad); if (st != METRO_OK) return st; p.shift = d;
  if (g.out2 || L.chain) {
    st = arena.upload(&d, g.scale2, cout_pad); if (st != METRO_OK) return st; p.scale2 = d;
    st = arena.upload(&d, g.shift2, cout_pad); if (st != METRO_OK) return st; p.shift2 = d;
  }
  if (L.chain) {
    // second convolution: weights [cout_c][cout] K-major (K order = conv3's output channels), boxes of cout_c / 2 rows
    std::vector<__half> packed_c(size_t(g.cout_c) * g.cout);
    conv_gemm_pack_weights(g.w_c, 1, g.cout, g.cout_c, nullptr, 0, g.cout_c, packed_c.data());
    __half *d_wc = nullptr;
    if ((st = arena.upload(&d_wc, packed_c)) != METRO_OK) return st;
    if ((st = make_weight_tensor_map(&p.w1map, d_wc, g.cout_c, g.cout, g.cout_c)) != METRO_OK) return st;
    st = arena.upload(&d, g.scale_c); if (st != METRO_OK) return st; p.scale1c = d;
    st = arena.upload(&d, g.shift_c); if (st != METRO_OK) return st; p.shift1c = d;
    p.cout1 = g.cout_c;
  }
  if (!g.ascale.empty()) {
    if (g.k != 1 || g.stride != 1 || g.cin2 || g.res || g.out2 || L.direct || L.block_n > 256)
      return fail(METRO_ERR_VALUE, "%s: the in-kernel pre-activation needs a plain 1x1 convolution with <= 128 outputs", g.name.c_str());
    st = arena.upload(&d, g.ascale); if (st != METRO_OK) return st; p.ascale = d;
    st = arena.upload(&d, g.ashift); if (st != METRO_OK) return st; p.ashift = d;
  }
  p.relu1 = g.relu ? 1 : 0;
  p.has_out1 = g.out1 ? 1 : 0; p.has_out2 = g.out2 ? 1 : 0;
  p.out1 = g.out1; p.out1_f32 = g.out1_f32 ? 1 : 0;
  const long long m_rows = (long long)g.n_max * g.out_side * g.out_side;
  if (!L.direct) {
    if (g.out1 && (st = make_out_tensor_map(&p.o1map, g.out1, m_rows, g.cout)) != METRO_OK) return st;
    if (g.out2 && (st = make_out_tensor_map(&p.o2map, g.out2, m_rows, g.cout)) != METRO_OK) return st;
    if (L.chain && (st = make_out_tensor_map(&p.o2map, g.out_c, m_rows, g.cout_c)) != METRO_OK) return st;
  }
  if (L.chain) {
    if ((st = conv_chain_plan_smem(p)) != METRO_OK) return st;
  } else {
    const int kb_tile = p.taps * p.cblk0 + (p.diag2 ? L.block_n / 64 : p.cblk1);
    if ((st = conv_gemm_plan_smem(L, kb_tile)) != METRO_OK) return st;
  }
  L.flops_per_img = 2.0 * g.out_side * g.out_side * double(g.cout) * K;
  if (L.chain) L.flops_per_img += 2.0 * g.out_side * g.out_side * double(g.cout_c) * g.cout;
  L.sig_expected = unsigned(g.out_side * g.out_side / 32) * unsigned(L.chain ? 1 : p.n_tiles);   // one report per epilogue warp (32 rows) and N tile (chain: per pixel tile)
  // Reports cost a GPU-scope release per epilogue warp and tile: affordable where a CTA tile is a whole crop or more
  // (maps of 16x16 and below: tiles of 5-50 us), not for the ~2 us tiles of the 64x64 / 32x32 layers (measured: +9 %
  // on the step with every layer reporting)
  L.signals = g.out_side * g.out_side <= 2 * kTileM;
  conv_gemm_set_batch(p, g.n_max);
  return METRO_OK;
}

}  // namespace
}  // namespace metro

using namespace metro;

struct metro_handle {
  int device = 0, num_sms = 0, max_batch = 0;
  metro_spec spec{};
  std::vector<int32_t> perm;
  NetPlan plan;
  DeviceArena arena;
  // root: image pack -> fused conv1 + pool1 + first pre-activation
  float *d_pool_scale = nullptr, *d_pool_shift = nullptr, *d_root_bias = nullptr;
  __half *buf_packed = nullptr, *buf_root = nullptr, *d_root_w = nullptr, *pool_raw = nullptr, *pool_pre = nullptr;
  __half *d_root_w2 = nullptr;   // version 2 of the root kernel (image pack folded in, paired conv rows)
  int root_v2 = 1;               // METRO_ROOT_V1 switches back to img_pack + root_fused
  alignas(64) unsigned char image_map[128];
  std::vector<ConvGemmLaunch> gemms;
  void *buf_head = nullptr;
  SoftargmaxLaunch sam{};
  void *sam_ws = nullptr;
  std::map<std::string, std::pair<const void *, size_t>> debug;   // name -> (buffer, elems per crop)
  // host-buffer path
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr};
  float *stage_img = nullptr, *stage_pose = nullptr;
  bool host_ready = false;  // every resource of the host-buffer path exists (set last, so a failed first call retries)
  StrictNet *strict = nullptr;   // precision != METRO_PREC_F16: the float64 CUDA-core evaluation (strict.cu)
  std::string joint_names;       // the graph's constant fetches (main.py:128,140-141), '\n'-separated
  std::vector<int32_t> joint_edges;
  size_t weight_bytes = 0;       // part of arena.total that does not scale with max_batch
  float *coords01_out = nullptr; // metro_infer_coords: second fetch of the call in flight
  unsigned int *flags = nullptr; // dataflow counters [1 + gemms][max_batch]: row 0 = fused root, row 1 + i = gemms[i]
  // CUDA-graph executor for small batches (launch-bound: 52 launches of a few microseconds each): the launch
  // sequence of a call is captured once per (batch, dtype, buffers) and replayed
  struct GraphEntry {
    int n = 0; bool u8 = false; const void *images = nullptr; float *poses = nullptr;
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    int seen = 0; bool failed = false; unsigned long long last_use = 0;
  };
  std::vector<GraphEntry> graphs;
  cudaStream_t cap_stream = nullptr;
  int graph_max_batch = 32;      // METRO_GRAPH_MAX_BATCH (0 = never capture)
  unsigned long long graph_clock = 0, graph_replays = 0;
  std::vector<unsigned int> df_launched;   // per layer: CTAs launched since the counters were last zeroed (this call)
  int df_forward = 1;            // METRO_DF_KEEP_ALTERNATE: dataflow layers keep the alternating tile order
  int xform_max_cb = 256;        // METRO_XFORM_MAX_CB: widest bottleneck whose conv1 applies the pre-activation in-kernel
  int dataflow = 1;              // METRO_NO_DATAFLOW switches back to grid-wide dependencies (griddepcontrol.wait)
  int stem_chunk = 0;            // METRO_STEM_CHUNK: metro_infer runs the stem in slices of this many crops (0 = whole batch)
  int alternate = 1;      // consecutive convolutions walk their tile lists in opposite directions (METRO_NO_ALTERNATE)
  int host_chunk = 64;    // crops per PCIe slice of metro_infer_host (METRO_HOST_CHUNK)
  int host_slice_u8 = 128;  // stem and tail slice of metro_infer_host_u8: a quarter of the bytes per crop, so larger
                            // slices arrive as quickly and cost fewer launches (measured: 5.06 ms against 5.43 at 64)
  int host_tail = 64;     // crops per slice of the deep blocks (METRO_HOST_TAIL; 0 = whole batch)
  int stem_gemms = 0;     // tensor-core convolutions that run per slice (up to the last 32x32-or-larger block)
};

namespace {

metro_status check_spec(const metro_spec *spec) {
  if (!spec) return fail(METRO_ERR_VALUE, "spec is null");
  if (spec->n_joints_out <= 0 || !spec->permutation) return fail(METRO_ERR_VALUE, "permutation is required");
  if (spec->n_joints_out > kMaxJointsOut) return fail(METRO_ERR_VALUE, "n_joints_out > %d", kMaxJointsOut);
  for (int i = 0; i < spec->n_joints_out; ++i)
    if (spec->permutation[i] < 0 || spec->permutation[i] >= spec->n_joints_model)
      return fail(METRO_ERR_VALUE, "permutation[%d]=%d out of range", i, spec->permutation[i]);
  if (spec->head_dtype != METRO_F32 && spec->head_dtype != METRO_F16) return fail(METRO_ERR_VALUE, "bad head_dtype");
  if (spec->precision < METRO_PREC_F16 || spec->precision > METRO_PREC_STRICT_F16) return fail(METRO_ERR_VALUE, "bad precision");
  if (spec->n_joint_edges < 0 || (spec->n_joint_edges > 0 && !spec->joint_edges)) return fail(METRO_ERR_VALUE, "bad joint_edges");
  for (int i = 0; i < 2 * spec->n_joint_edges; ++i)
    if (spec->joint_edges[i] < 0 || spec->joint_edges[i] >= spec->n_joints_out)
      return fail(METRO_ERR_VALUE, "joint_edges[%d]=%d out of range [0,%d)", i, spec->joint_edges[i], spec->n_joints_out);
  if (spec->joint_names) {
    int lines = 1;
    for (const char *c = spec->joint_names; *c; ++c) lines += *c == '\n';
    if (lines != spec->n_joints_out) return fail(METRO_ERR_VALUE, "joint_names holds %d names for %d output joints", lines, spec->n_joints_out);
  }
  return METRO_OK;
}

metro_status build_handle(metro_handle &h, const float *blob) {
  const NetPlan &pl = h.plan;
  const int N = h.max_batch;
  DeviceArena &A = h.arena;
  const bool keep = h.spec.keep_activations != 0;
  metro_status st;
  auto alloc_half = [&](__half **p, size_t elems_per_img) -> metro_status {
    void *q = nullptr;
    metro_status s = A.alloc(&q, elems_per_img * N * sizeof(__half));
    *p = static_cast<__half *>(q);
    return s;
  };

  // ---- root: image pack -> fused conv1 7x7/2 + bias -> zero-padded pool1 -> first pre-activation (root_fused.cu) ----
  {
    if (pl.proc_side != 256 || pl.pool_in != 128 || pl.pool_out != 64)
      return fail(METRO_ERR_VALUE, "the root kernel is built for 256x256 crops (FLAGS.proc_side)");
    void *q = nullptr;
    if ((st = A.alloc(&q, size_t(N) * root_packed_image_elems() * sizeof(__half))) != METRO_OK) return st;
    h.buf_packed = static_cast<__half *>(q);
    if ((st = root_make_image_map(h.image_map, h.buf_packed, N)) != METRO_OK) return st;
    std::vector<__half> wp(root_packed_weight_elems());
    root_pack_weights(blob + pl.root.w_off, wp.data());
    if ((st = A.upload(&h.d_root_w, wp)) != METRO_OK) return st;
    std::vector<__half> wp2(root2_packed_weight_elems());
    root2_pack_weights(blob + pl.root.w_off, wp2.data());
    if ((st = A.upload(&h.d_root_w2, wp2)) != METRO_OK) return st;
    if (getenv("METRO_ROOT_V1")) h.root_v2 = 0;
    std::vector<float> bias(blob + pl.root.b_off, blob + pl.root.b_off + 64);
    if ((st = A.upload(&h.d_root_bias, bias)) != METRO_OK) return st;
    if (keep) {
      if ((st = alloc_half(&h.buf_root, size_t(pl.pool_in) * pl.pool_in * 64)) != METRO_OK) return st;
      h.debug["conv1"] = {h.buf_root, size_t(pl.pool_in) * pl.pool_in * 64};
    }
  }
  // ---- sizes of the rotating buffers ----
  size_t raw_elems = size_t(pl.pool_out) * pl.pool_out * 64, r1_elems = 0, r2_elems = 0;
  for (const auto &u : pl.units) {
    raw_elems = std::max(raw_elems, size_t(u.out_side) * u.out_side * u.depth);
    r1_elems = std::max(r1_elems, size_t(u.in_side) * u.in_side * u.cb);
    r2_elems = std::max(r2_elems, size_t(u.out_side) * u.out_side * u.cb);
  }
  __half *raw[2] = {nullptr, nullptr}, *pre[2] = {nullptr, nullptr}, *r1 = nullptr, *r2 = nullptr;
  if (const char *e = getenv("METRO_XFORM_MAX_CB")) h.xform_max_cb = atoi(e);
  if (!keep) {
    for (int i = 0; i < 2; ++i) {
      if ((st = alloc_half(&raw[i], raw_elems)) != METRO_OK) return st;
      if ((st = alloc_half(&pre[i], raw_elems)) != METRO_OK) return st;
    }
    if ((st = alloc_half(&r1, r1_elems)) != METRO_OK) return st;
    if ((st = alloc_half(&r2, r2_elems)) != METRO_OK) return st;
  }
  // ---- pool1 + first pre-activation ----
  std::vector<float> sc, sf;
  bn_affine(blob + pl.units[0].preact_off, pl.units[0].cin, sc, sf);
  if ((st = A.upload(&h.d_pool_scale, sc)) != METRO_OK) return st;
  if ((st = A.upload(&h.d_pool_shift, sf)) != METRO_OK) return st;
  const size_t pool_elems = size_t(pl.pool_out) * pl.pool_out * 64;
  __half *cur_raw, *cur_pre;
  if (keep) {
    if ((st = alloc_half(&cur_raw, pool_elems)) != METRO_OK) return st;
    if ((st = alloc_half(&cur_pre, pool_elems)) != METRO_OK) return st;
  } else { cur_raw = raw[0]; cur_pre = pre[0]; }
  h.pool_raw = cur_raw; h.pool_pre = cur_pre;
  h.debug["pool1"] = {cur_raw, pool_elems};

  // ---- residual units ----
  // The leading units whose maps are 32x32 or larger form the "stem" that metro_infer_host runs slice by
  // slice (a 64-crop slice still gives >= 256 pair tiles per layer).  The outputs of the last stem unit
  // are what the whole-batch tail reads after every slice has passed, so they get buffers of their own:
  // the rotating buffers are re-used, in other layouts, by the next slice's early layers.
  size_t stem_units = 0;
  while (stem_units < pl.units.size() && pl.units[stem_units].out_side >= 32) ++stem_units;
  h.stem_gemms = 0;
  static const bool no_chain = getenv("METRO_NO_CHAIN") != nullptr;
  bool conv1_done = false;       // this unit's conv1 ran inside the previous unit's chained kernel (conv_chain.cu)
  for (size_t i = 0; i < pl.units.size(); ++i) {
    const UnitPlan &u = pl.units[i];
    const bool last = (i + 1 == pl.units.size());
    if (u.proj && u.stride != 1)
      return fail(METRO_ERR_INTERNAL, "%s: strided projection shortcut is not part of ResNet-50/101", u.name.c_str());
    __half *b1 = r1, *b2 = r2, *nraw = raw[(i + 1) & 1], *npre = pre[(i + 1) & 1];
    const size_t e1 = size_t(u.in_side) * u.in_side * u.cb, e2 = size_t(u.out_side) * u.out_side * u.cb;
    const size_t eo = size_t(u.out_side) * u.out_side * u.depth;
    if (keep) {
      if ((st = alloc_half(&b1, e1)) != METRO_OK) return st;
      if ((st = alloc_half(&b2, e2)) != METRO_OK) return st;
    }
    if (keep || i + 1 == stem_units) {
      if ((st = alloc_half(&nraw, eo)) != METRO_OK) return st;
      if ((st = alloc_half(&npre, eo)) != METRO_OK) return st;
    }
    ConvGemmLaunch L;
    // An identity unit of block1-3 (<= 256 bottleneck channels; measured: block4's K = 2048 conv1 loses more to the
    // extra shared-memory pass than its conv3 gains) reads its RAW input and
    // applies its own pre-activation inside conv1 (conv_gemm kXform), so the previous unit's conv3 writes one
    // tensor instead of two.  Same arithmetic and rounding as the stored pre-activation, hence bit-identical.
    auto reads_raw = [&](size_t k) {
      static const bool no_xform = getenv("METRO_NO_XFORM") != nullptr;
      return !no_xform && !keep && k > 0 && k < pl.units.size() && !pl.units[k].proj && pl.units[k].cb <= h.xform_max_cb;
    };
    // conv1: 1x1, BN, ReLU on the pre-activation (resnet_v2.py:127-128)
    if (!conv1_done) {
      GemmSpec g; g.name = u.conv1.name; g.n_max = N; g.src = cur_pre; g.in_side = u.in_side; g.cin = u.cin;
      if (reads_raw(i)) {
        g.src = cur_raw;
        bn_affine(blob + u.preact_off, u.cin, g.ascale, g.ashift);
      }
      g.out_side = u.in_side; g.cout = u.cb; g.w = blob + u.conv1.w_off;
      bn_affine(blob + u.conv1.bn_off, u.cb, g.scale, g.shift);
      g.relu = true; g.out1 = b1;
      if ((st = build_gemm(A, g, L)) != METRO_OK) return st;
      h.gemms.push_back(L);
      h.debug[u.name + "/conv1"] = {b1, e1};
    }
    // conv2: 3x3 stride/rate/centred, BN, ReLU (resnet_v2.py:130-132)
    {
      GemmSpec g; g.name = u.conv2.name; g.n_max = N; g.src = b1; g.in_side = u.in_side; g.cin = u.cb;
      g.k = 3; g.stride = u.stride; g.rate = u.rate; g.pad_lo = u.conv2.pad_lo;
      g.out_side = u.out_side; g.cout = u.cb; g.w = blob + u.conv2.w_off;
      bn_affine(blob + u.conv2.bn_off, u.cb, g.scale, g.shift);
      g.relu = true; g.out1 = b2;
      if ((st = build_gemm(A, g, L)) != METRO_OK) return st;
      h.gemms.push_back(L);
      h.debug[u.name + "/conv2"] = {b2, e2};
    }
    // conv3 (+ projection shortcut as a second K range) + bias(es) + identity residual -> raw sum and
    // the next unit's pre-activation / postnorm (resnet_v2.py:119-125,134-138,229)
    {
      GemmSpec g; g.name = u.conv3.name; g.n_max = N; g.src = b2; g.in_side = u.out_side; g.cin = u.cb;
      g.out_side = u.out_side; g.cout = u.depth; g.w = blob + u.conv3.w_off;
      g.scale.assign(u.depth, 1.0f);
      g.shift.assign(blob + u.conv3.b_off, blob + u.conv3.b_off + u.depth);
      if (u.proj) {
        g.src2 = cur_pre; g.cin2 = u.cin; g.w2 = blob + u.shortcut.w_off;
        for (int c = 0; c < u.depth; ++c) g.shift[c] = float(double(g.shift[c]) + double(blob[u.shortcut.b_off + c]));
      } else {
        g.res = cur_raw; g.res_stride = u.stride; g.res_shift = u.shift; g.res_side = u.in_side;
      }
      const bool need_raw = keep || (!last && !pl.units[i + 1].proj);
      g.out1 = need_raw ? nraw : nullptr;
      g.out2 = reads_raw(i + 1) ? nullptr : npre;
      const int64_t next_bn = last ? pl.postnorm_off : pl.units[i + 1].preact_off;
      // Chain this conv3 with the NEXT unit's conv1 (conv_chain.cu) when that unit has an identity shortcut (it needs
      // the raw sum in HBM, never the pre-activation) and a bottleneck of <= 256 channels (the second accumulator has
      // to fit tensor memory next to the double-buffered first): the pre-activation stays on chip and the widest
      // tensor of the unit is not read back.  Not across the stem boundary of the sliced host path.
      const bool chain = !no_chain && !keep && !last && !pl.units[i + 1].proj && pl.units[i + 1].cb <= 256 &&
                         u.depth % 256 == 0 && u.depth <= 1024 && i + 1 != stem_units;
      if (chain) {
        const UnitPlan &v = pl.units[i + 1];
        g.name = u.conv3.name + "+" + v.conv1.name;
        g.out2 = nullptr;
        bn_affine(blob + next_bn, u.depth, g.scale2, g.shift2);
        g.w_c = blob + v.conv1.w_off; g.cout_c = v.cb; g.out_c = r1;
        bn_affine(blob + v.conv1.bn_off, v.cb, g.scale_c, g.shift_c);
      } else if (g.out2) {
        bn_affine(blob + next_bn, u.depth, g.scale2, g.shift2);
      }
      if ((st = build_gemm(A, g, L)) != METRO_OK) return st;
      h.gemms.push_back(L);
      h.debug[u.name + "/out"] = {need_raw ? nraw : nullptr, eo};
      h.debug[u.name + "/pre"] = {g.out2 ? npre : nullptr, eo};
      conv1_done = chain;
    }
    if (i + 1 == stem_units) h.stem_gemms = int(h.gemms.size());
    cur_raw = nraw; cur_pre = npre;
  }
  if (getenv("METRO_NO_ALTERNATE")) h.alternate = 0;
  if (const char *e = getenv("METRO_HOST_CHUNK")) h.host_chunk = h.host_slice_u8 = atoi(e);
  if (const char *e = getenv("METRO_HOST_TAIL")) h.host_tail = atoi(e);
  if (const char *e = getenv("METRO_STEM_CHUNK")) h.stem_chunk = atoi(e);
  if (const char *e = getenv("METRO_GRAPH_MAX_BATCH")) h.graph_max_batch = atoi(e);
  // ---- logits (resnet_v2.py:234-236) -> head tensor ----
  {
    const size_t eh = size_t(pl.feat_side) * pl.feat_side * pl.logits.cout;
    const bool f16 = h.spec.head_dtype == METRO_F16;
    if ((st = A.alloc(&h.buf_head, eh * N * (f16 ? 2 : 4))) != METRO_OK) return st;
    GemmSpec g; g.name = "logits"; g.n_max = N; g.src = cur_pre; g.in_side = pl.feat_side; g.cin = pl.feat_channels;
    g.out_side = pl.feat_side; g.cout = pl.logits.cout; g.w = blob + pl.logits.w_off;
    g.scale.assign(g.cout, 1.0f);
    g.shift.assign(blob + pl.logits.b_off, blob + pl.logits.b_off + g.cout);
    g.out1 = h.buf_head; g.out1_f32 = !f16;
    ConvGemmLaunch L;
    if ((st = build_gemm(A, g, L)) != METRO_OK) return st;
    h.gemms.push_back(L);
    h.debug["head"] = {h.buf_head, eh};
  }
  // ---- soft-argmax ----
  {
    metro_softargmax_desc d{};
    d.side = pl.feat_side; d.n_joints_model = pl.n_joints; d.depth = pl.depth; d.stride = pl.stride;
    d.centered_stride = pl.centered; d.proc_side = pl.proc_side; d.box_size_mm = h.spec.box_size_mm;
    d.n_joints_out = int(h.perm.size()); d.permutation = h.perm.data(); d.head_dtype = h.spec.head_dtype;
    if ((st = softargmax_plan(d, N, h.sam)) != METRO_OK) return st;
    const size_t ws = softargmax_workspace_bytes(h.sam);
    if ((st = A.alloc(&h.sam_ws, ws)) != METRO_OK) return st;
    METRO_CUDA(cudaMemset(h.sam_ws, 0, ws));
  }
  h.weight_bytes = A.uploaded;
  if (getenv("METRO_NO_DATAFLOW")) h.dataflow = 0;
  if (getenv("METRO_DF_KEEP_ALTERNATE")) h.df_forward = 0;
  if (keep) h.dataflow = 0;      // debug handles keep grid-wide dependencies
  {
    void *q = nullptr;
    const size_t fb = ((h.gemms.size() + 1) * size_t(N) + h.gemms.size() + 1) * sizeof(unsigned int);   // + a word per layer
    if ((st = A.alloc(&q, fb)) != METRO_OK) return st;
    METRO_CUDA(cudaMemset(q, 0, fb));
    h.flags = static_cast<unsigned int *>(q);
  }
  METRO_CUDA(cudaDeviceSynchronize());
  return METRO_OK;
}

struct Timer {
  std::vector<cudaEvent_t> ev;
  std::vector<std::string> names;
  long long *role_prof = nullptr;   // device [launch][num_sms][8] role timers (METRO_ROLE_PROF=1)
  // device-side stamps instead of events: [launch][2] = earliest CTA start / latest CTA end of every launch, taken
  // inside the real pipeline (events between launches serialise it: no programmatic overlap, a launch latency each)
  unsigned long long *stamps = nullptr;
  unsigned long long *slot() const { return stamps ? stamps + 2 * names.size() : nullptr; }
};

// Dataflow wiring of convolution `li`: it waits on the counters of its producer (row li: the fused root for li == 0,
// else gemms[li - 1]) and reports into row li + 1.
// Only layers with long tiles report (ConvGemmLaunch::signals); a layer whose producer does not report keeps the
// grid-wide dependency.  `n` / `n_base`: this call's slice (the producer ran on the same slice).
void set_dataflow(metro_handle *h, int li, ConvGemmParams &prm) {
  if (!h->dataflow) return;
  const size_t N = size_t(h->max_batch);
  unsigned int *done = h->flags + (h->gemms.size() + 1) * N;     // one word per layer after the per-crop rows
  if (li > 0 && h->gemms[li - 1].signals) {
    prm.dep_flags = h->flags + size_t(li) * N;
    prm.dep_expected = h->gemms[li - 1].sig_expected;
    // "the producer layer is complete" = every CTA it has launched so far in this call has exited (a call may run a
    // layer in several slices; all of them precede this launch in the stream)
    prm.dep_done = done + li;
    prm.dep_ctas = h->df_launched[li];
  }
  if (h->gemms[li].signals) {
    prm.sig_flags = h->flags + size_t(li + 1) * N;
    prm.sig_done = done + li + 1;
    h->df_launched[li + 1] += unsigned(h->gemms[li].chain ? conv_chain_grid(prm, h->num_sms) : conv_gemm_grid(prm, h->num_sms));
  }
}

// zeroes the dataflow counters at the start of a call (every crop of a call is produced once per layer)
metro_status reset_dataflow(metro_handle *h, cudaStream_t s) {
  if (!h->dataflow) return METRO_OK;
  h->df_launched.assign(h->gemms.size() + 1, 0u);
  METRO_CUDA(cudaMemsetAsync(h->flags, 0, ((h->gemms.size() + 1) * size_t(h->max_batch) + h->gemms.size() + 1) * sizeof(unsigned int), s));
  return METRO_OK;
}

// Launches the stem (space-to-depth pack, conv1, pool1 and the first `stem_gemms` tensor-core
// convolutions) for crops [n_base, n_base + n) of the handle's buffers; `images` points at crop n_base.
metro_status run_stem(metro_handle *h, const void *images, bool u8, int n, int n_base, int stem_gemms, cudaStream_t s,
                      Timer *t) {
  const NetPlan &pl = h->plan;
  metro_status st;
  auto mark = [&](const char *name) {
    if (!t) return;
    if (!t->stamps) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); t->ev.push_back(e); }
    t->names.push_back(name);
  };
  (void)pl;
  if (h->root_v2) {
    if ((st = root_fused2_launch(images, u8, h->d_root_w2, h->d_root_bias, h->d_pool_scale, h->d_pool_shift,
                                 h->spec.keep_activations ? h->pool_raw : nullptr, h->pool_pre, h->buf_root, n, n_base, h->num_sms,
                                 s, t ? t->slot() : nullptr)) != METRO_OK) return st;
    mark("conv1+pool1");
  } else {
  if ((st = img_pack_launch(images, u8, h->buf_packed + size_t(n_base) * root_packed_image_elems(), n, s)) != METRO_OK) return st;
  mark("img_pack");
  if ((st = root_fused_launch(h->image_map, h->d_root_w, h->d_root_bias, h->d_pool_scale, h->d_pool_shift,
                              h->spec.keep_activations ? h->pool_raw : nullptr, h->pool_pre, h->buf_root, n, n_base, h->num_sms, s,
                              (t && t->role_prof) ? t->role_prof : nullptr, nullptr)) != METRO_OK) return st;
  mark("conv1+pool1");
  }
  for (int li = 0; li < stem_gemms; ++li) {
    // the handle's launch record is never written after metro_create: the batch slice, walk direction and
    // profiling pointer of THIS call travel in the by-value kernel parameters (so a call can be captured
    // into a CUDA graph, and two streams may use one handle's weights concurrently with separate arenas)
    const ConvGemmLaunch &L = h->gemms[li];
    ConvGemmParams prm = L.prm;
    conv_gemm_set_batch(prm, n, n_base);
    prm.reverse = h->alternate ? (li & 1) ^ 1 : 0;     // the root kernel walks forwards
    prm.prof = (t && t->role_prof) ? t->role_prof + size_t(li + 1) * h->num_sms * 16 : nullptr;
    set_dataflow(h, li, prm);
    if (prm.dep_flags && h->df_forward) prm.reverse = 0;
    prm.tstamp = t ? t->slot() : nullptr;
    if ((st = L.chain ? conv_chain_launch(prm, h->num_sms, s) : conv_gemm_launch(L, prm, h->num_sms, s)) != METRO_OK) return st;
    mark(L.name.c_str());
  }
  return METRO_OK;
}

// The remaining convolutions, the logits head and the soft-argmax for crops [n_base, n_base + n); `poses`
// points at crop 0 of the output.
metro_status run_tail(metro_handle *h, int n, int n_base, int first_gemm, float *poses, cudaStream_t s, Timer *t) {
  metro_status st;
  auto mark = [&](const char *name) {
    if (!t) return;
    if (!t->stamps) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); t->ev.push_back(e); }
    t->names.push_back(name);
  };
  for (size_t li = first_gemm; li < h->gemms.size(); ++li) {
    const ConvGemmLaunch &L = h->gemms[li];
    ConvGemmParams prm = L.prm;
    conv_gemm_set_batch(prm, n, n_base);
    prm.reverse = h->alternate ? (li & 1) ^ 1 : 0;
    prm.prof = (t && t->role_prof) ? t->role_prof + size_t(li + 1) * h->num_sms * 16 : nullptr;
    set_dataflow(h, int(li), prm);
    // a layer that follows its producer crop by crop walks the tile list in the producer's order, so that the CTAs
    // which replace the producer's early finishers start on crops that are already complete
    if (prm.dep_flags && h->df_forward) prm.reverse = 0;
    prm.tstamp = t ? t->slot() : nullptr;
    if ((st = L.chain ? conv_chain_launch(prm, h->num_sms, s) : conv_gemm_launch(L, prm, h->num_sms, s)) != METRO_OK) return st;
    mark(L.name.c_str());
  }
  SoftargmaxLaunch sl = h->sam;
  sl.tstamp = t ? t->slot() : nullptr;
  if (h->dataflow && h->gemms.back().signals) {
    sl.dep_flags = h->flags + h->gemms.size() * size_t(h->max_batch) + n_base;
    sl.dep_expected = h->gemms.back().sig_expected;
  }
  const size_t head_bytes = size_t(sl.H) * sl.W * sl.C * (sl.head_f16 ? 2 : 4);
  sl.n = n;
  sl.head = static_cast<const unsigned char *>(h->buf_head) + size_t(n_base) * head_bytes;
  sl.out = poses ? poses + size_t(n_base) * sl.n_out * 3 : nullptr;
  sl.coords01 = h->coords01_out ? h->coords01_out + size_t(n_base) * sl.J * 3 : nullptr;
  sl.counters = static_cast<unsigned int *>(h->sam_ws) + n_base;
  sl.partials = reinterpret_cast<double *>(static_cast<unsigned char *>(h->sam_ws) +
                                           ((size_t(h->max_batch) * 4 + 255) & ~size_t(255))) +
                size_t(n_base) * sl.splits * sl.J * 5;
  if ((st = softargmax_launch(sl, s)) != METRO_OK) return st;
  mark("softargmax");
  return METRO_OK;
}

// one call's launch sequence on stream `s` (capturable: nothing but stream operations, no handle state that a replay
// would need changed)
metro_status run_direct(metro_handle *h, const void *images, bool u8, int n, float *poses, cudaStream_t s, Timer *t) {
  {
    metro_status rst = reset_dataflow(h, s);
    if (rst != METRO_OK) return rst;
  }
  if (h->stem_chunk > 0 && h->stem_chunk < n && !t) {
    // experiment knob: the stem (maps of 32x32 and larger, HBM-bound) in slices small enough for a layer's
    // output to still sit in L2 when the next layer reads it; the deep blocks once on the whole batch
    const size_t img_bytes = size_t(h->plan.proc_side) * h->plan.proc_side * 3 * (u8 ? 1 : 4);
    for (int lo = 0; lo < n; lo += h->stem_chunk) {
      const int cnt = std::min(h->stem_chunk, n - lo);
      metro_status st = run_stem(h, static_cast<const unsigned char *>(images) + size_t(lo) * img_bytes, u8, cnt, lo,
                                 h->stem_gemms, s, nullptr);
      if (st != METRO_OK) return st;
    }
    return run_tail(h, n, 0, h->stem_gemms, poses, s, nullptr);
  }
  metro_status st = run_stem(h, images, u8, n, 0, 0, s, t);
  if (st != METRO_OK) return st;
  return run_tail(h, n, 0, 0, poses, s, t);
}

metro_status run(metro_handle *h, const void *images, bool u8, int n, float *poses, cudaStream_t s, Timer *t) {
  if (!h) return fail(METRO_ERR_VALUE, "handle is null");
  if (n < 0 || n > h->max_batch) return fail(METRO_ERR_VALUE, "batch %d outside [0, max_batch=%d]", n, h->max_batch);
  if (n == 0) return METRO_OK;
  if (!images || !poses) return fail(METRO_ERR_VALUE, "null image / pose buffer");
  METRO_CUDA(cudaSetDevice(h->device));
  if (t) {
    if (!t->stamps) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); t->ev.push_back(e); }
    t->names.push_back("start");
  }
  if (h->strict) return strict_run(h->strict, images, u8, n, poses, s);
  if (t || n > h->graph_max_batch) return run_direct(h, images, u8, n, poses, s, t);
  // ---- small batch: replay a captured graph of the same call (same batch, dtype and buffers) ----
  metro_handle::GraphEntry *e = nullptr;
  for (auto &g : h->graphs)
    if (g.n == n && g.u8 == u8 && g.images == images && g.poses == poses) e = &g;
  if (!e) {
    constexpr size_t kMaxGraphs = 8;
    if (h->graphs.size() >= kMaxGraphs) {          // evict the least recently used entry
      size_t lru = 0;
      for (size_t i = 1; i < h->graphs.size(); ++i)
        if (h->graphs[i].last_use < h->graphs[lru].last_use) lru = i;
      if (h->graphs[lru].exec) cudaGraphExecDestroy(h->graphs[lru].exec);
      if (h->graphs[lru].graph) cudaGraphDestroy(h->graphs[lru].graph);
      h->graphs.erase(h->graphs.begin() + lru);
    }
    h->graphs.emplace_back();
    e = &h->graphs.back();
    e->n = n; e->u8 = u8; e->images = images; e->poses = poses;
  }
  e->last_use = ++h->graph_clock;
  if (e->failed || e->seen++ == 0) return run_direct(h, images, u8, n, poses, s, nullptr);   // first sight: run as is
  if (!e->exec) {
    if (!h->cap_stream) METRO_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    bool ok = cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      const metro_status cst = run_direct(h, images, u8, n, poses, h->cap_stream, nullptr);
      ok = cudaStreamEndCapture(h->cap_stream, &e->graph) == cudaSuccess && cst == METRO_OK && e->graph;
    }
    if (ok) ok = cudaGraphInstantiate(&e->exec, e->graph, 0) == cudaSuccess;
    if (!ok) {
      cudaGetLastError();
      if (e->graph) { cudaGraphDestroy(e->graph); e->graph = nullptr; }
      e->exec = nullptr; e->failed = true;
      return run_direct(h, images, u8, n, poses, s, nullptr);
    }
  }
  METRO_CUDA(cudaGraphLaunch(e->exec, s));
  ++h->graph_replays;
  return METRO_OK;
}

}  // namespace

extern "C" {

const char *metro_last_error(void) { return g_last_error.c_str(); }
const char *metro_version(void) { return "metro-b200 0.1 (sm_100a)"; }

metro_status metro_blob_floats(const metro_spec *spec, uint64_t *n_floats) {
  if (!spec || !n_floats) return fail(METRO_ERR_VALUE, "null argument");
  NetPlan pl; std::string err;
  const metro_status st = build_plan(*spec, pl, err);
  if (st != METRO_OK) return fail(st, "%s", err.c_str());
  *n_floats = uint64_t(pl.blob_floats);
  return METRO_OK;
}

metro_status metro_plan_describe(const metro_spec *spec, char *buf, size_t buf_bytes, size_t *needed) {
  if (!spec) return fail(METRO_ERR_VALUE, "null argument");
  NetPlan pl; std::string err;
  const metro_status st = build_plan(*spec, pl, err);
  if (st != METRO_OK) return fail(st, "%s", err.c_str());
  const std::string js = plan_to_json(pl);
  if (needed) *needed = js.size() + 1;
  if (buf && buf_bytes > 0) {
    const size_t n = std::min(buf_bytes - 1, js.size());
    std::memcpy(buf, js.data(), n);
    buf[n] = 0;
  }
  return METRO_OK;
}

metro_status metro_create(const metro_spec *spec, const float *weights_blob, uint64_t n_floats, int32_t device,
                          metro_handle **out) {
  if (!out) return fail(METRO_ERR_VALUE, "out is null");
  *out = nullptr;
  metro_status st = check_spec(spec);
  if (st != METRO_OK) return st;
  if (spec->max_batch <= 0) return fail(METRO_ERR_VALUE, "max_batch must be positive");
  std::unique_ptr<metro_handle> h(new metro_handle());
  std::string err;
  st = build_plan(*spec, h->plan, err);
  if (st != METRO_OK) return fail(st, "%s", err.c_str());
  if (!weights_blob || n_floats != uint64_t(h->plan.blob_floats))
    return fail(METRO_ERR_VALUE, "weight blob has %llu floats, the model needs %lld", (unsigned long long)n_floats,
                (long long)h->plan.blob_floats);
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(METRO_ERR_NO_DEVICE, "no CUDA device: libmetro has no CPU fallback");
  if (device < 0 || device >= count) return fail(METRO_ERR_VALUE, "device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  METRO_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(METRO_ERR_NO_DEVICE, "device %d is sm_%d%d; libmetro is built for sm_100a only", device, prop.major, prop.minor);
  METRO_CUDA(cudaSetDevice(device));
  h->device = device; h->num_sms = prop.multiProcessorCount; h->max_batch = spec->max_batch;
  h->spec = *spec;
  h->perm.assign(spec->permutation, spec->permutation + spec->n_joints_out);
  h->spec.permutation = h->perm.data();
  if (spec->joint_names) h->joint_names = spec->joint_names;
  if (spec->n_joint_edges > 0) h->joint_edges.assign(spec->joint_edges, spec->joint_edges + 2 * spec->n_joint_edges);
  h->spec.joint_names = nullptr; h->spec.joint_edges = nullptr;     // the handle keeps its own copies
  if (spec->precision != METRO_PREC_F16) {
    st = strict_build(h->plan, weights_blob, h->max_batch, spec->precision == METRO_PREC_STRICT_F16 ? 1 : 0, spec->box_size_mm,
                      h->perm, spec->keep_activations != 0, &h->strict);
    if (st != METRO_OK) return st;
    METRO_CUDA(cudaDeviceSynchronize());
    *out = h.release();
    return METRO_OK;
  }
  st = build_handle(*h, weights_blob);
  if (st != METRO_OK) return st;
  *out = h.release();
  return METRO_OK;
}

metro_status metro_destroy(metro_handle *h) {
  if (!h) return METRO_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (auto e : h->ev_copied) if (e) cudaEventDestroy(e);
  if (h->stage_img) cudaFree(h->stage_img);
  if (h->stage_pose) cudaFree(h->stage_pose);
  if (h->strict) strict_destroy(h->strict);
  for (auto &g : h->graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
  }
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  delete h;
  return METRO_OK;
}

metro_status metro_workspace_bytes(const metro_handle *h, int32_t n, uint64_t *bytes) {
  if (!h || !bytes) return fail(METRO_ERR_VALUE, "null argument");
  const size_t total = h->strict ? strict_bytes(h->strict) : h->arena.total;
  if (n <= 0 || n >= h->max_batch || h->strict) { *bytes = total; return METRO_OK; }
  // weights and per-channel vectors do not scale with the batch; every activation buffer is `max_batch` crops long
  *bytes = h->weight_bytes + (total - h->weight_bytes) / uint64_t(h->max_batch) * uint64_t(n);
  return METRO_OK;
}

metro_status metro_get_joint_info(const metro_handle *h, char *names_buf, size_t names_bytes, size_t *names_needed,
                                  int32_t *edges_buf, int32_t edges_cap, int32_t *n_edges, int32_t *n_joints) {
  if (!h) return fail(METRO_ERR_VALUE, "handle is null");
  if (h->joint_names.empty() && h->joint_edges.empty())
    return fail(METRO_ERR_VALUE, "this handle was created without joint tables (spec.joint_names / spec.joint_edges)");
  if (names_needed) *names_needed = h->joint_names.size() + 1;
  if (names_buf && names_bytes > 0) {
    const size_t k = std::min(names_bytes - 1, h->joint_names.size());
    std::memcpy(names_buf, h->joint_names.data(), k);
    names_buf[k] = 0;
  }
  const int32_t ne = int32_t(h->joint_edges.size() / 2);
  if (n_edges) *n_edges = ne;
  if (n_joints) *n_joints = int32_t(h->perm.size());
  if (edges_buf && edges_cap > 0)
    std::memcpy(edges_buf, h->joint_edges.data(), size_t(std::min(ne, edges_cap)) * 2 * sizeof(int32_t));
  return METRO_OK;
}

metro_status metro_infer(metro_handle *h, const float *images_dev, int32_t n, float *poses_dev, void *stream) {
  return run(h, images_dev, false, n, poses_dev, static_cast<cudaStream_t>(stream), nullptr);
}

metro_status metro_infer_u8(metro_handle *h, const uint8_t *images_u8_dev, int32_t n, float *poses_dev, void *stream) {
  return run(h, images_u8_dev, true, n, poses_dev, static_cast<cudaStream_t>(stream), nullptr);
}

namespace {
metro_status infer_host(metro_handle *h, const void *images_host_v, bool u8, int32_t n, float *poses_host);
}

metro_status metro_infer_host(metro_handle *h, const float *images_host, int32_t n, float *poses_host) {
  return infer_host(h, images_host, false, n, poses_host);
}

metro_status metro_infer_host_u8(metro_handle *h, const uint8_t *images_u8_host, int32_t n, float *poses_host) {
  return infer_host(h, images_u8_host, true, n, poses_host);
}

namespace {
metro_status infer_host(metro_handle *h, const void *images_host_v, bool u8, int32_t n, float *poses_host) {
  const unsigned char *images_host = static_cast<const unsigned char *>(images_host_v);
  if (!h) return fail(METRO_ERR_VALUE, "handle is null");
  if (n < 0 || n > h->max_batch) return fail(METRO_ERR_VALUE, "batch %d outside [0, max_batch=%d]", n, h->max_batch);
  if (n == 0) return METRO_OK;
  if (!images_host || !poses_host) return fail(METRO_ERR_VALUE, "null image / pose buffer");
  METRO_CUDA(cudaSetDevice(h->device));
  const size_t img_elems = size_t(h->plan.proc_side) * h->plan.proc_side * 3;
  const size_t img_bytes = img_elems * (u8 ? sizeof(uint8_t) : sizeof(float));   // this call's element size
  const size_t pose_bytes = h->perm.size() * 3 * sizeof(float);
  if (!h->host_ready) {
    // a failed allocation returns early and leaves host_ready unset: the next call resumes where this one stopped
    if (!h->stream) METRO_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    if (!h->copy_stream) METRO_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->stage_img) METRO_CUDA(cudaMalloc(&h->stage_img, img_elems * sizeof(float) * h->max_batch));
    if (!h->stage_pose) METRO_CUDA(cudaMalloc(&h->stage_pose, pose_bytes * h->max_batch));
    for (int i = 0; i < 2; ++i)
      if (!h->ev_copied[i]) METRO_CUDA(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
    h->host_ready = true;
  }
  if (h->strict) {
    METRO_CUDA(cudaMemcpyAsync(h->stage_img, images_host, img_bytes * n, cudaMemcpyHostToDevice, h->stream));
    metro_status sst = strict_run(h->strict, h->stage_img, u8, n, h->stage_pose, h->stream);
    if (sst != METRO_OK) return sst;
    METRO_CUDA(cudaMemcpyAsync(poses_host, h->stage_pose, pose_bytes * n, cudaMemcpyDeviceToHost, h->stream));
    METRO_CUDA(cudaStreamSynchronize(h->stream));
    return METRO_OK;
  }
  // The crops arrive over PCIe in slices; the stem of the network (root, block1, block2: the layers whose
  // tile count per crop is large, so that a slice still fills the GPU) runs slice by slice underneath the
  // copies, the deep blocks (few, large tiles per crop) run once on the whole batch.  A crop's result does
  // not depend on the slicing (every tile holds whole rows of one crop).
  int chunk = u8 ? h->host_slice_u8 : h->host_chunk;
  if (chunk <= 0 || chunk >= n) chunk = n;
  const int stem_gemms = chunk < n ? h->stem_gemms : 0;
  // the deep blocks follow in slices of `tail` crops (a multiple of the stem slice; 64 by default, i.e. after
  // every stem slice).  Measured with METRO_HOST_TRACE at 256 crops: the four H2D slices land at 0.9 / 1.8 /
  // 2.7 / 3.7 ms, every stem slice costs 0.52 ms and every 64-crop tail 0.83 ms, so from the first slice on
  // the GPU never waits for the bus and the call ends ~2.6 ms after the last byte arrived; larger tail
  // slices are more efficient per crop (128 crops: 1.43 ms) but leave the GPU idle while they fill
  int tail = h->host_tail > 0 ? (h->host_tail + chunk - 1) / chunk * chunk : n;   // at least one stem slice
  if (chunk == n) tail = n;
  int i = 0, tail_lo = 0;
  // METRO_HOST_TRACE=1: timestamps of every slice's copy / stem / tail, printed after the call (debug only)
  static const bool trace = std::getenv("METRO_HOST_TRACE") != nullptr;
  std::vector<std::pair<const char *, cudaEvent_t>> marks;
  auto mark = [&](const char *what, cudaStream_t s) {
    if (!trace) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    marks.emplace_back(what, e);
  };
  if (trace) { cudaStreamSynchronize(h->stream); cudaStreamSynchronize(h->copy_stream); }
  {
    metro_status rst = reset_dataflow(h, h->stream);
    if (rst != METRO_OK) return rst;
  }
  mark("start", h->copy_stream);
  for (int lo = 0; lo < n; lo += chunk, ++i) {
    const int cnt = lo + chunk <= n ? chunk : n - lo;
    unsigned char *stage = reinterpret_cast<unsigned char *>(h->stage_img) + size_t(lo) * img_bytes;
    METRO_CUDA(cudaMemcpyAsync(stage, images_host + size_t(lo) * img_bytes, img_bytes * cnt, cudaMemcpyHostToDevice,
                               h->copy_stream));
    METRO_CUDA(cudaEventRecord(h->ev_copied[i & 1], h->copy_stream));
    mark("h2d", h->copy_stream);
    METRO_CUDA(cudaStreamWaitEvent(h->stream, h->ev_copied[i & 1], 0));
    metro_status st = run_stem(h, stage, u8, cnt, lo, stem_gemms, h->stream, nullptr);
    if (st != METRO_OK) return st;
    mark("stem", h->stream);
    const int done = lo + cnt;
    // the deep blocks follow a stem slice once `tail` crops are ready -- except after the second-to-last slice:
    // the GPU is the bottleneck of this call from the first slice on, nothing would fill the gap a split leaves,
    // and one tail over the last two slices costs less than two (fixed ramp of ~45 launches per tail)
    static const bool merge_last = std::getenv("METRO_HOST_NO_MERGE_LAST") == nullptr;
    const bool second_to_last = merge_last && lo + chunk < n && lo + 2 * chunk >= n;
    if (done == n || (done - tail_lo >= tail && !second_to_last)) {
      st = run_tail(h, done - tail_lo, tail_lo, stem_gemms, h->stage_pose, h->stream, nullptr);
      if (st != METRO_OK) return st;
      tail_lo = done;
      mark("tail", h->stream);
    }
  }
  METRO_CUDA(cudaMemcpyAsync(poses_host, h->stage_pose, pose_bytes * n, cudaMemcpyDeviceToHost, h->stream));
  mark("d2h", h->stream);
  METRO_CUDA(cudaStreamSynchronize(h->stream));
  if (trace) {
    std::fprintf(stderr, "[metro host trace]");
    for (auto &m : marks) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, marks[0].second, m.second);
      std::fprintf(stderr, " %s=%.3f", m.first, ms);
    }
    std::fprintf(stderr, "\n");
    for (auto &m : marks) cudaEventDestroy(m.second);
  }
  return METRO_OK;
}
}  // namespace

metro_status metro_extract_crops(const metro_crop_src *srcs, int32_t n, int32_t side, int32_t border_value, uint8_t *crops_u8_dev,
                                 void *stream) {
  if (n < 0 || side <= 0 || side > 4096) return fail(METRO_ERR_VALUE, "extract_crops: bad batch / side");
  if (border_value < 0 || border_value > 255) return fail(METRO_ERR_VALUE, "extract_crops: border value must be a byte");
  if (n == 0) return METRO_OK;
  if (!srcs || !crops_u8_dev) return fail(METRO_ERR_VALUE, "extract_crops: null argument");
  for (int i = 0; i < n; ++i) {
    const metro_crop_src &s = srcs[i];
    if (!s.frame_dev || s.height <= 0 || s.width <= 0 || s.row_stride_bytes < 3 * s.width)
      return fail(METRO_ERR_VALUE, "extract_crops: source %d is not a valid uint8 RGB frame", i);
    if (s.height > 32767 || s.width > 32767) return fail(METRO_ERR_VALUE, "extract_crops: frames above 32767 pixels a side are not supported (cv2.remap's int16 coordinates)");
  }
  return extract_crops_launch(srcs, n, side, border_value, crops_u8_dev, static_cast<cudaStream_t>(stream));
}

metro_status metro_to_orig_cam(const float *poses_dev, const float *rot_dev, const int32_t *mirror_mapping, int32_t n,
                               int32_t n_joints, float *out_dev, void *stream) {
  if (n < 0) return fail(METRO_ERR_VALUE, "to_orig_cam: negative batch");
  if (n_joints <= 0 || n_joints > kMaxJointsOut) return fail(METRO_ERR_VALUE, "to_orig_cam: n_joints must be in [1, %d]", kMaxJointsOut);
  if (!mirror_mapping) return fail(METRO_ERR_VALUE, "to_orig_cam: null mirror mapping");
  for (int i = 0; i < n_joints; ++i)
    if (mirror_mapping[i] < 0 || mirror_mapping[i] >= n_joints)
      return fail(METRO_ERR_VALUE, "to_orig_cam: mirror_mapping[%d]=%d out of range [0,%d)", i, mirror_mapping[i], n_joints);
  if (n == 0) return METRO_OK;
  if (!poses_dev || !rot_dev || !out_dev) return fail(METRO_ERR_VALUE, "to_orig_cam: null device buffer");
  if (poses_dev == out_dev) return fail(METRO_ERR_VALUE, "to_orig_cam: the output must not alias the input (joints are gathered)");
  return to_orig_cam_launch(poses_dev, rot_dev, mirror_mapping, n, n_joints, out_dev, static_cast<cudaStream_t>(stream));
}

metro_status metro_softargmax_workspace_bytes(const metro_softargmax_desc *d, int32_t n, uint64_t *bytes) {
  if (!d || !bytes) return fail(METRO_ERR_VALUE, "null argument");
  SoftargmaxLaunch L;
  const metro_status st = softargmax_plan(*d, n, L);
  if (st != METRO_OK) return st;
  *bytes = softargmax_workspace_bytes(L);
  return METRO_OK;
}

metro_status metro_softargmax(const metro_softargmax_desc *d, const void *head_dev, int32_t n, float *poses_dev,
                              void *workspace_dev, void *stream) {
  if (n > 0 && !poses_dev) return fail(METRO_ERR_VALUE, "null device buffer");
  return metro_softargmax_coords(d, head_dev, n, poses_dev, nullptr, workspace_dev, stream);
}

metro_status metro_heatmap_z(const metro_softargmax_desc *d, const void *head_dev, int32_t n, float *out_dev, void *stream) {
  if (!d) return fail(METRO_ERR_VALUE, "null argument");
  if (d->side <= 0 || d->n_joints_model <= 0 || d->depth <= 0 || n < 0) return fail(METRO_ERR_VALUE, "heatmap_z: bad shape");
  if (d->head_dtype != METRO_F16 && d->head_dtype != METRO_F32) return fail(METRO_ERR_VALUE, "heatmap_z: bad head_dtype");
  if (n == 0) return METRO_OK;
  if (!head_dev || !out_dev) return fail(METRO_ERR_VALUE, "null device buffer");
  return heatmap_z_launch(head_dev, d->head_dtype == METRO_F16, n, d->side, d->n_joints_model, d->depth, out_dev,
                          static_cast<cudaStream_t>(stream));
}

metro_status metro_back_project(const float *coords01_dev, const float *inv_intrinsics_dev, const float *z_offset_dev, int32_t n,
                                int32_t n_joints, int32_t stride, int32_t centered_stride, int32_t proc_side, float box_size_mm,
                                float *out_dev, void *stream) {
  if (n < 0 || n_joints <= 0 || stride <= 0 || proc_side <= 0) return fail(METRO_ERR_VALUE, "back_project: bad argument");
  if (n == 0) return METRO_OK;
  if (!coords01_dev || !inv_intrinsics_dev || !z_offset_dev || !out_dev) return fail(METRO_ERR_VALUE, "back_project: null device buffer");
  const int last = proc_side - 1;                                  // volumetric.py:288-291
  const double lrc = double(last - (last % stride) - 1);
  return back_project_launch(coords01_dev, inv_intrinsics_dev, z_offset_dev, n, n_joints, lrc,
                             centered_stride ? double(stride / 2) : 0.0, double(box_size_mm), out_dev,
                             static_cast<cudaStream_t>(stream));
}

metro_status metro_infer_coords(metro_handle *h, const float *images_dev, int32_t n, float *poses_dev, float *coords01_dev,
                                void *stream) {
  if (!h) return fail(METRO_ERR_VALUE, "handle is null");
  if (n < 0 || n > h->max_batch) return fail(METRO_ERR_VALUE, "batch %d outside [0, max_batch=%d]", n, h->max_batch);
  if (n == 0) return METRO_OK;
  if (!images_dev || !coords01_dev) return fail(METRO_ERR_VALUE, "null image / coordinate buffer");
  METRO_CUDA(cudaSetDevice(h->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (h->strict) return strict_run(h->strict, images_dev, false, n, poses_dev, s, coords01_dev);
  h->coords01_out = coords01_dev;                                  // picked up by run_tail for this call only
  const metro_status st = run_direct(h, images_dev, false, n, poses_dev, s, nullptr);
  h->coords01_out = nullptr;
  return st;
}

metro_status metro_softargmax_coords(const metro_softargmax_desc *d, const void *head_dev, int32_t n, float *poses_dev,
                                     float *coords01_dev, void *workspace_dev, void *stream) {
  if (!d) return fail(METRO_ERR_VALUE, "null argument");
  SoftargmaxLaunch L;
  metro_status st = softargmax_plan(*d, n, L);
  if (st != METRO_OK) return st;
  if (n == 0) return METRO_OK;
  if (!head_dev || (!poses_dev && !coords01_dev) || !workspace_dev) return fail(METRO_ERR_VALUE, "null device buffer");
  L.head = head_dev; L.out = poses_dev; L.coords01 = coords01_dev;
  L.counters = static_cast<unsigned int *>(workspace_dev);
  L.partials = reinterpret_cast<double *>(static_cast<unsigned char *>(workspace_dev) + ((size_t(n) * 4 + 255) & ~size_t(255)));
  return softargmax_launch(L, static_cast<cudaStream_t>(stream));
}

metro_status metro_conv2d(const metro_conv_desc *d, const void *x_dev, const float *w_host, const void *x2_dev,
                          const float *w2_host, const float *scale_host, const float *shift_host, const void *res_dev,
                          void *y_dev, const float *scale2_host, const float *shift2_host, void *y2_dev, int32_t device,
                          void *stream) {
  if (!d || !x_dev || !w_host || !scale_host || !shift_host) return fail(METRO_ERR_VALUE, "null argument");
  if (d->k != 1 && d->k != 3) return fail(METRO_ERR_VALUE, "metro_conv2d: k must be 1 or 3");
  if (d->n <= 0) return fail(METRO_ERR_VALUE, "metro_conv2d: n must be positive");
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(METRO_ERR_NO_DEVICE, "no CUDA device: libmetro has no CPU fallback");
  cudaDeviceProp prop;
  METRO_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(METRO_ERR_NO_DEVICE, "device is sm_%d%d; sm_100a required", prop.major, prop.minor);
  METRO_CUDA(cudaSetDevice(device));
  const int k_eff = d->k + (d->k - 1) * (d->rate - 1);
  const int pad_total = (d->stride == 1) ? (k_eff - 1) : -1;
  int out_side;
  if (d->stride == 1) out_side = d->in_side + pad_total - k_eff + 1;  // SAME
  else out_side = d->in_side / d->stride;
  DeviceArena arena;
  GemmSpec g; g.name = "metro_conv2d"; g.n_max = d->n;
  g.src = static_cast<const __half *>(x_dev); g.in_side = d->in_side; g.cin = d->cin;
  g.k = d->k; g.stride = d->stride; g.rate = d->rate; g.pad_lo = d->pad_lo; g.out_side = out_side; g.cout = d->cout;
  g.w = w_host;
  if (d->cin2 > 0) { g.src2 = static_cast<const __half *>(x2_dev); g.cin2 = d->cin2; g.w2 = w2_host; }
  g.scale.assign(scale_host, scale_host + d->cout);
  g.shift.assign(shift_host, shift_host + d->cout);
  g.relu = d->relu != 0;
  if (res_dev && d->res_stride > 0) {
    g.res = static_cast<const __half *>(res_dev); g.res_stride = d->res_stride; g.res_shift = d->res_shift;
    g.res_side = out_side * d->res_stride;
  }
  g.out1 = y_dev; g.out1_f32 = d->out_dtype == METRO_F32;
  if (y2_dev) {
    if (!scale2_host || !shift2_host) return fail(METRO_ERR_VALUE, "metro_conv2d: y2 needs scale2/shift2");
    g.out2 = static_cast<__half *>(y2_dev);
    g.scale2.assign(scale2_host, scale2_host + d->cout);
    g.shift2.assign(shift2_host, shift2_host + d->cout);
  }
  ConvGemmLaunch L;
  metro_status st = build_gemm(arena, g, L);
  if (st != METRO_OK) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  st = conv_gemm_launch(L, L.prm, prop.multiProcessorCount, s);
  if (st != METRO_OK) return st;
  METRO_CUDA(cudaStreamSynchronize(s));   // temporaries (packed weights) are freed on return
  return METRO_OK;
}

metro_status metro_debug_read(metro_handle *h, const char *name, void *host_buf, uint64_t buf_bytes, uint64_t *elems) {
  if (!h || !name) return fail(METRO_ERR_VALUE, "null argument");
  if (h->strict) {
    const double *ptr = nullptr; size_t per = 0;
    if (!strict_debug(h->strict, name, &ptr, &per)) return fail(METRO_ERR_VALUE, "no activation named '%s'", name);
    const uint64_t n_el = uint64_t(per) * h->max_batch;
    if (elems) *elems = n_el;
    if (host_buf) {
      if (buf_bytes < n_el * 8) return fail(METRO_ERR_VALUE, "buffer too small: need %llu bytes", (unsigned long long)(n_el * 8));
      METRO_CUDA(cudaSetDevice(h->device));
      METRO_CUDA(cudaDeviceSynchronize());
      METRO_CUDA(cudaMemcpy(host_buf, ptr, n_el * 8, cudaMemcpyDeviceToHost));
    }
    return METRO_OK;
  }
  auto it = h->debug.find(name);
  if (it == h->debug.end() || !it->second.first) return fail(METRO_ERR_VALUE, "no activation named '%s'", name);
  const bool is_head = std::string(name) == "head";
  const size_t esize = is_head ? (h->spec.head_dtype == METRO_F16 ? 2 : 4) : 2;
  const uint64_t n_el = uint64_t(it->second.second) * h->max_batch;
  if (elems) *elems = n_el;
  if (host_buf) {
    if (buf_bytes < n_el * esize) return fail(METRO_ERR_VALUE, "buffer too small: need %llu bytes", (unsigned long long)(n_el * esize));
    METRO_CUDA(cudaSetDevice(h->device));
    METRO_CUDA(cudaDeviceSynchronize());
    METRO_CUDA(cudaMemcpy(host_buf, it->second.first, n_el * esize, cudaMemcpyDeviceToHost));
  }
  return METRO_OK;
}

metro_status metro_graph_stats(const metro_handle *h, int32_t *graphs, int64_t *replays) {
  if (!h) return fail(METRO_ERR_VALUE, "handle is null");
  int32_t k = 0;
  for (const auto &g : h->graphs) k += g.exec != nullptr;
  if (graphs) *graphs = k;
  if (replays) *replays = int64_t(h->graph_replays);
  return METRO_OK;
}

metro_status metro_launch_count(const metro_handle *h, int32_t n, int32_t *launches) {
  if (!h || !launches) return fail(METRO_ERR_VALUE, "null argument");
  if (h->strict) {   // image cast + one kernel per conv / BN / pool step + decode + metric (strict.cu)
    *launches = n > 0 ? int32_t(h->plan.n_convs() + h->plan.units.size() + 2 + 3) : 0;
    return METRO_OK;
  }
  *launches = n > 0 ? int32_t(h->gemms.size()) + (h->root_v2 ? 2 : 3) : 0;   // + [image pack,] fused root, soft-argmax
  return METRO_OK;
}

metro_status metro_profile(metro_handle *h, const float *images_dev, int32_t n, float *poses_dev, float *ms_out,
                           char *names_buf, size_t names_bytes, int32_t *n_launches) {
  Timer t;
  if (!h) return fail(METRO_ERR_VALUE, "handle is null");
  if (h->strict) return fail(METRO_ERR_VALUE, "metro_profile: per-launch timing exists for the tensor-core path only");
  const bool roles = getenv("METRO_ROLE_PROF") != nullptr;
  const size_t role_elems = size_t(h->gemms.size() + 1) * h->num_sms * 16;
  if (roles) {
    METRO_CUDA(cudaMalloc(&t.role_prof, role_elems * sizeof(long long)));
    METRO_CUDA(cudaMemset(t.role_prof, 0, role_elems * sizeof(long long)));
  }
  // default: device-side time stamps per launch inside the real pipeline; METRO_PROFILE_EVENTS=1 (or the fallback root
  // kernel, which carries no stamps): CUDA events between launches, which serialise the launches
  const bool use_stamps = !getenv("METRO_PROFILE_EVENTS") && h->root_v2;
  const size_t max_launches = h->gemms.size() + 8;
  std::vector<unsigned long long> hstamps(2 * max_launches);
  if (use_stamps) {
    for (size_t i = 0; i < max_launches; ++i) { hstamps[2 * i] = ~0ull; hstamps[2 * i + 1] = 0ull; }
    METRO_CUDA(cudaMalloc(&t.stamps, hstamps.size() * sizeof(unsigned long long)));
    METRO_CUDA(cudaMemcpy(t.stamps, hstamps.data(), hstamps.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
  }
  METRO_CUDA(cudaDeviceSynchronize());
  metro_status st = run(h, images_dev, false, n, poses_dev, nullptr, &t);
  if (st != METRO_OK) return st;
  METRO_CUDA(cudaDeviceSynchronize());
  if (use_stamps) {
    METRO_CUDA(cudaMemcpy(hstamps.data(), t.stamps, hstamps.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(t.stamps);
  }
  if (roles) {
    // per launch, averaged over CTAs: total | producer waits for a free stage | MMA waits for operands |
    // MMA waits for a free accumulator | epilogue waits for an accumulator | epilogue busy | store wait | tiles
    std::vector<long long> host(role_elems);
    METRO_CUDA(cudaMemcpy(host.data(), t.role_prof, role_elems * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(t.role_prof);
    fprintf(stderr, "%-28s %9s %9s %9s %9s %9s %9s %9s %6s %8s %8s %8s %8s %8s %8s %8s %8s\n", "roles (kcycles, CTA mean)", "total", "prod_wait", "mma_full",
            "mma_acc", "epi_wait", "epi_busy", "st_wait", "tiles", "e_ld", "e_math", "e_sts", "e_issue", "e_par", "m_issue", "m_commit", "p_issue");
    for (size_t l = 0; l <= h->gemms.size(); ++l) {
      double acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      int ctas = 0;
      for (int c = 0; c < h->num_sms; ++c) {
        const long long *r = &host[(l * h->num_sms + c) * 16];
        if (r[0] == 0) continue;
        ++ctas;
        for (int k = 0; k < 16; ++k) acc[k] += double(r[k]);
      }
      if (!ctas) continue;
      acc[2] *= 2; acc[3] *= 2; acc[7] *= 2; acc[13] *= 2; acc[14] *= 2;        // MMA-thread columns exist in the pair leaders only
      if (l == 0) {   // fused root kernel
        double a[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        int ctas = 0;
        for (int c = 0; c < h->num_sms; ++c) {
          const long long *r = &host[size_t(c) * 16];
          if (r[2] == 0) continue;
          ++ctas;
          for (int k = 0; k < 16; ++k) a[k] += double(r[k]);
        }
        if (ctas)
          fprintf(stderr, "root_fused (kcycles): prod total %.1f wait_empty %.1f | mma total %.1f wait_acc %.1f wait_full %.1f rows %.1f | "
                          "epi total %.1f wait_acc %.1f tmem_ld %.1f barrier %.1f pool %.1f\n",
                  a[0] / ctas / 1e3, a[1] / ctas / 1e3, a[2] / ctas / 1e3, a[3] / ctas / 1e3, a[4] / ctas / 1e3, a[5] / ctas,
                  a[6] / ctas / 1e3, a[7] / ctas / 1e3, a[8] / ctas / 1e3, a[9] / ctas / 1e3, a[10] / ctas / 1e3);
        continue;
      }
      const std::string &nm = h->gemms[l - 1].name;
      fprintf(stderr, "%-28s %9.1f %9.1f %9.1f %9.1f %9.1f %9.1f %9.1f %6.1f %8.1f %8.1f %8.1f %8.1f %8.1f %8.1f %8.1f %8.1f\n", nm.c_str(), acc[0] / ctas / 1e3,
              acc[1] / ctas / 1e3, acc[2] / ctas / 1e3, acc[3] / ctas / 1e3, acc[4] / ctas / 1e3, acc[5] / ctas / 1e3,
              acc[6] / ctas / 1e3, acc[7] / ctas, acc[8] / ctas / 1e3, acc[9] / ctas / 1e3, acc[10] / ctas / 1e3,
              acc[11] / ctas / 1e3, acc[12] / ctas / 1e3, acc[13] / ctas / 1e3, acc[14] / ctas / 1e3, acc[15] / ctas / 1e3);
    }
  }
  std::string names;
  const int cnt = int(t.names.size()) - 1;
  for (int i = 0; i < cnt; ++i) {
    float ms = 0;
    if (use_stamps) {      // launch i + 1 of the name list stamped slot i + 1
      const unsigned long long b = hstamps[2 * (i + 1)], e = hstamps[2 * (i + 1) + 1];
      ms = (e > b && b != ~0ull) ? float(double(e - b) * 1e-6) : 0.f;
    } else {
      cudaEventElapsedTime(&ms, t.ev[i], t.ev[i + 1]);
    }
    if (ms_out) ms_out[i] = ms;
    names += t.names[i + 1];
    names += '\n';
  }
  for (auto e : t.ev) cudaEventDestroy(e);
  if (n_launches) *n_launches = cnt;
  if (names_buf && names_bytes) {
    const size_t k = std::min(names_bytes - 1, names.size());
    std::memcpy(names_buf, names.data(), k);
    names_buf[k] = 0;
  }
  return METRO_OK;
}

}  // extern "C"
