// Host-side layer plan of the MeTRo inference graph (no CUDA here).
// Replays src/model/resnet_v2.py:272-312 (block table, centred-stride selection),
// src/model/resnet_v2.py:209-236 (root, pool, postnorm, logits) and
// src/model/resnet_utils.py:307-350 (stride / atrous bookkeeping) of the reference.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/metro.h"

namespace metro {

struct ConvGeom {
  std::string name;
  int cin = 0, cout = 0, k = 1, stride = 1, rate = 1, pad_lo = 0, pad_hi = 0;
  int in_side = 0, out_side = 0;
  bool has_bias = false, has_bn = false, relu = false;
  // offsets (in floats) into the weight blob; -1 = absent
  int64_t w_off = -1, b_off = -1, bn_off = -1;   // bn_off -> gamma, beta, mean, var (cout each)
  double flops() const { return 2.0 * out_side * out_side * cout * double(cin) * k * k; }
};

struct UnitPlan {
  std::string name;
  int cin = 0, depth = 0, cb = 0, stride = 1, rate = 1, shift = 0, in_side = 0, out_side = 0;
  bool proj = false;
  int64_t preact_off = -1;     // gamma, beta, mean, var (cin each)
  ConvGeom shortcut, conv1, conv2, conv3;
};

struct NetPlan {
  int arch = 50, stride = 16, n_joints = 17, depth = 8, centered = 1, proc_side = 256;
  ConvGeom root;
  int pool_in = 0, pool_out = 0;
  std::vector<UnitPlan> units;
  int feat_side = 0, feat_channels = 0;
  int64_t postnorm_off = -1;
  ConvGeom logits;
  int64_t blob_floats = 0;
  double flops_per_crop() const;
  int n_convs() const;
};

// Returns METRO_OK or METRO_ERR_VALUE with `err` set (same conditions as the reference's
// ValueErrors: stride % 4, unreachable stride).
metro_status build_plan(const metro_spec &spec, NetPlan &plan, std::string &err);
std::string plan_to_json(const NetPlan &plan);

}  // namespace metro
