#include "plan.h"

#include <cmath>
#include <sstream>

namespace metro {

namespace {
const int kUnits50[4] = {3, 4, 6, 3};
const int kUnits101[4] = {3, 4, 23, 3};
const int kBaseDepth[4] = {64, 128, 256, 512};
const int kBlockStride[4] = {2, 2, 2, 1};

// TensorFlow 'SAME' padding.
void same_pad(int n, int k_eff, int s, int &out, int &lo, int &hi) {
  out = (n + s - 1) / s;
  int total = (out - 1) * s + k_eff - n;
  if (total < 0) total = 0;
  lo = total / 2;
  hi = total - lo;
}

struct Cursor {
  int64_t off = 0;
  int64_t take(int64_t n) { int64_t o = off; off += n; return o; }
};

void place(ConvGeom &c, Cursor &cur) {
  c.w_off = cur.take(int64_t(c.k) * c.k * c.cin * c.cout);
  if (c.has_bias) c.b_off = cur.take(c.cout);
  if (c.has_bn) c.bn_off = cur.take(4 * int64_t(c.cout));
}

ConvGeom make_conv(const std::string &name, int cin, int cout, int k, int stride, int rate, int lo,
                   int hi, int in_side, int out_side, bool bias, bool bn, bool relu) {
  ConvGeom c;
  c.name = name; c.cin = cin; c.cout = cout; c.k = k; c.stride = stride; c.rate = rate;
  c.pad_lo = lo; c.pad_hi = hi; c.in_side = in_side; c.out_side = out_side;
  c.has_bias = bias; c.has_bn = bn; c.relu = relu;
  return c;
}
}  // namespace

double NetPlan::flops_per_crop() const {
  double f = root.flops() + logits.flops();
  for (const auto &u : units) {
    if (u.proj) f += u.shortcut.flops();
    f += u.conv1.flops() + u.conv2.flops() + u.conv3.flops();
  }
  return f;
}

int NetPlan::n_convs() const {
  int n = 2;
  for (const auto &u : units) n += 3 + (u.proj ? 1 : 0);
  return n;
}

metro_status build_plan(const metro_spec &spec, NetPlan &p, std::string &err) {
  if (spec.arch != 50 && spec.arch != 101) { err = "unknown architecture (arch must be 50 or 101)"; return METRO_ERR_VALUE; }
  if (spec.stride <= 0 || spec.stride % 4 != 0) {  // resnet_v2.py:213-214
    err = "The output_stride needs to be a multiple of 4.";
    return METRO_ERR_VALUE;
  }
  if (spec.n_joints_model <= 0 || spec.depth <= 0 || spec.proc_side <= 0) {
    err = "n_joints_model, depth and proc_side must be positive";
    return METRO_ERR_VALUE;
  }
  p = NetPlan();
  p.arch = spec.arch; p.stride = spec.stride; p.n_joints = spec.n_joints_model;
  p.depth = spec.depth; p.centered = spec.centered_stride ? 1 : 0; p.proc_side = spec.proc_side;
  const int target = spec.stride / 4;   // resnet_v2.py:215
  const int *n_units = spec.arch == 50 ? kUnits50 : kUnits101;

  // centred-stride block selection: resnet_v2.py:277-281 (rn50, guarded) / :299-302 (rn101)
  bool c[3] = {false, false, false};
  if (p.centered) {
    const double l2 = std::log2(double(spec.stride));
    int i_last = (spec.arch == 50 ? int(std::nearbyint(l2)) : int(l2)) - 3;
    if (i_last > 2) { err = "The target output_stride cannot be reached."; return METRO_ERR_VALUE; }
    if (spec.arch == 50) { if (i_last >= 0) c[i_last] = true; }
    else c[(i_last + 3) % 3] = true;    // Python list index -1 wraps to block3
  }

  Cursor cur;
  int side = spec.proc_side;
  int out = (side + 6 - 7) / 2 + 1;     // conv2d_same(64, 7, stride=2): pad (3,3) + VALID
  p.root = make_conv("conv1", 3, 64, 7, 2, 1, 3, 3, side, out, true, false, false);
  place(p.root, cur);
  side = out;
  p.pool_in = side;
  side = (side + 2 - 3) / 2 + 1;        // max_pool2d_same(3, stride=2): zero pad (1,1) + VALID
  p.pool_out = side;

  int current_stride = 1, rate = 1, cin = 64;
  for (int b = 0; b < 4; ++b) {
    const int cb = kBaseDepth[b], depth = 4 * cb;
    for (int u = 0; u < n_units[b]; ++u) {
      const bool last = (u == n_units[b] - 1);
      const int unit_stride = last ? kBlockStride[b] : 1;
      const bool unit_centered = (last && b < 3) ? c[b] : false;
      int s, r;
      if (current_stride == target) {   // resnet_utils.py:325-327
        s = 1; r = rate; rate *= unit_stride;
      } else {                          // :329-333
        s = unit_stride; r = 1; current_stride *= unit_stride;
        if (current_stride > target) { err = "The target output_stride cannot be reached."; return METRO_ERR_VALUE; }
      }
      UnitPlan up;
      up.name = "block" + std::to_string(b + 1) + "/unit_" + std::to_string(u + 1);
      up.cin = cin; up.depth = depth; up.cb = cb; up.stride = s; up.rate = r;
      up.shift = (unit_centered && s == 2) ? 1 : 0;
      up.in_side = side;
      const int k_eff = 3 + 2 * (r - 1);
      int o, lo, hi;
      if (s == 1 || unit_centered) same_pad(side, k_eff, s, o, lo, hi);
      else { lo = (k_eff - 1) / 2; hi = (k_eff - 1) - lo; o = (side + lo + hi - k_eff) / s + 1; }
      up.out_side = o;
      up.proj = (depth != cin);
      up.preact_off = cur.take(4 * int64_t(cin));
      if (up.proj) {
        up.shortcut = make_conv(up.name + "/shortcut", cin, depth, 1, s, 1, 0, 0, side, o, true, false, false);
        place(up.shortcut, cur);
      }
      up.conv1 = make_conv(up.name + "/conv1", cin, cb, 1, 1, 1, 0, 0, side, side, false, true, true);
      place(up.conv1, cur);
      up.conv2 = make_conv(up.name + "/conv2", cb, cb, 3, s, r, lo, hi, side, o, false, true, true);
      place(up.conv2, cur);
      up.conv3 = make_conv(up.name + "/conv3", cb, depth, 1, 1, 1, 0, 0, o, o, true, false, false);
      place(up.conv3, cur);
      p.units.push_back(up);
      side = o; cin = depth;
    }
  }
  if (current_stride != target) { err = "The target output_stride cannot be reached."; return METRO_ERR_VALUE; }
  p.feat_side = side; p.feat_channels = cin;
  p.postnorm_off = cur.take(4 * int64_t(cin));
  p.logits = make_conv("logits", cin, spec.depth * spec.n_joints_model, 1, 1, 1, 0, 0, side, side, true, false, false);
  place(p.logits, cur);
  p.blob_floats = cur.off;
  return METRO_OK;
}

static void conv_json(std::ostringstream &o, const ConvGeom &c) {
  o << "{\"name\":\"" << c.name << "\",\"cin\":" << c.cin << ",\"cout\":" << c.cout << ",\"k\":" << c.k
    << ",\"stride\":" << c.stride << ",\"rate\":" << c.rate << ",\"pad_lo\":" << c.pad_lo
    << ",\"pad_hi\":" << c.pad_hi << ",\"in_side\":" << c.in_side << ",\"out_side\":" << c.out_side
    << ",\"has_bias\":" << (c.has_bias ? "true" : "false") << ",\"has_bn\":" << (c.has_bn ? "true" : "false")
    << ",\"relu\":" << (c.relu ? "true" : "false") << ",\"w_off\":" << c.w_off << "}";
}

std::string plan_to_json(const NetPlan &p) {
  std::ostringstream o;
  o << "{\"arch\":" << p.arch << ",\"stride\":" << p.stride << ",\"n_joints\":" << p.n_joints
    << ",\"feat_side\":" << p.feat_side << ",\"feat_channels\":" << p.feat_channels
    << ",\"pool_in\":" << p.pool_in << ",\"pool_out\":" << p.pool_out
    << ",\"blob_floats\":" << p.blob_floats << ",\"n_convs\":" << p.n_convs()
    << ",\"flops_per_crop\":" << std::fixed << p.flops_per_crop() << ",\"convs\":[";
  conv_json(o, p.root);
  for (const auto &u : p.units) {
    if (u.proj) { o << ","; conv_json(o, u.shortcut); }
    o << ","; conv_json(o, u.conv1);
    o << ","; conv_json(o, u.conv2);
    o << ","; conv_json(o, u.conv3);
  }
  o << ","; conv_json(o, p.logits);
  o << "],\"units\":[";
  for (size_t i = 0; i < p.units.size(); ++i) {
    const auto &u = p.units[i];
    if (i) o << ",";
    o << "{\"name\":\"" << u.name << "\",\"cin\":" << u.cin << ",\"depth\":" << u.depth << ",\"cb\":" << u.cb
      << ",\"stride\":" << u.stride << ",\"rate\":" << u.rate << ",\"shift\":" << u.shift
      << ",\"in_side\":" << u.in_side << ",\"out_side\":" << u.out_side
      << ",\"proj\":" << (u.proj ? "true" : "false") << "}";
  }
  o << "]}";
  return o.str();
}

}  // namespace metro
