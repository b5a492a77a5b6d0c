#pragma once
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>

#include <cuda_runtime.h>

#include "../../include/metro.h"

namespace metro {

void set_error(const char *fmt, ...);
metro_status fail(metro_status st, const char *fmt, ...);

#define METRO_CUDA(expr)                                                                              \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      return ::metro::fail(METRO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                           __FILE__, __LINE__);                                                       \
  } while (0)

constexpr int kMaxJointsOut = 64;
constexpr int kMaxDevices = 64;

// One-time work per CUDA device (cudaFuncSetAttribute is per device and context): runs `f` the first time the
// CURRENT device is seen; thread-safe.
struct PerDeviceOnce {
  std::mutex mu;
  bool done[kMaxDevices] = {};
  template <typename F>
  metro_status run(F &&f) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return fail(METRO_ERR_CUDA, "cudaGetDevice failed");
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev]) return METRO_OK;
    const metro_status st = f();
    if (st == METRO_OK) done[dev] = true;
    return st;
  }
};

// ---- soft-argmax ------------------------------------------------------------------------------------
struct SoftargmaxLaunch {
  const void *head;
  float *out;
  double *partials;        // [n][splits][J][5] : K (integer base-2 exponent), S, Sx, Sy, Sz
  unsigned int *counters;  // [n], zero between launches
  int n, H, W, J, D, C;
  int n_out, root;
  int perm[kMaxJointsOut];
  double mul_x, mul_y, mul_z;  // mm per unit of (Sx/S), (Sy/S), (Sz/S)
  int splits, ipx;         // work items per crop and pixels per item (shape-only rule)
  int lanes, slots, vec;   // CTA = slots channel words x lanes pixel lanes; vec channels per word
  int cluster;             // 1 = the splits of a crop form a thread-block cluster (DSMEM merge, no workspace traffic)
  int tpj;                 // threads per joint in the merge (power of two <= 32)
  long long *prof;         // debug (METRO_SAM_PROF): 8 clock64 stamps per CTA, or null
  int head_f16;
  float *coords01;         // optional second output [n][J][3]: heatmap coordinates in [0,1] of every MODEL joint
                           // (what net_output_to_heatmap_and_coords returns, volumetric.py:234); `out` may then be null
  double unmul_x, unmul_y, unmul_z;   // 1 / mm-per-unit: back from the scaled expectations to [0,1]
  int l2_prefetch;         // pull the item's bytes into L2 ahead of the dependency wait (stand-alone launches)
  int early;               // CTAs [0, early) also prefetch an equal share of the whole input (they start early)
  // dataflow (ptx.cuh): wait for the crop's counter of the logits layer instead of for the whole previous grid
  const unsigned int *dep_flags;   // indexed by crop of THIS launch (already offset), or null
  unsigned int dep_expected;
  unsigned long long *tstamp;      // optional [2]: earliest CTA start / latest CTA end (%globaltimer ns), metro_profile
};
metro_status softargmax_plan(const metro_softargmax_desc &d, int n, SoftargmaxLaunch &L);
size_t softargmax_workspace_bytes(const SoftargmaxLaunch &L);
metro_status softargmax_launch(const SoftargmaxLaunch &L, cudaStream_t stream);

// ---- post-path transforms ---------------------------------------------------------------------------
metro_status to_orig_cam_launch(const float *poses, const float *rot, const int32_t *mirror, int n, int j, float *out,
                                cudaStream_t stream);
metro_status extract_crops_launch(const metro_crop_src *srcs, int n, int side, int border, unsigned char *out, cudaStream_t stream);
metro_status back_project_launch(const float *coords01, const float *inv_k, const float *z_off, int n, int j, double lrc,
                                 double add_xy, double box, float *out, cudaStream_t stream);
metro_status heatmap_z_launch(const void *head, bool f16, int n, int side, int j, int depth, float *out, cudaStream_t stream);

}  // namespace metro
