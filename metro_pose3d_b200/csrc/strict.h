// Strict-precision evaluation of the same graph on CUDA cores (float64 arithmetic): host-side interface.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "common.h"
#include "plan.h"

namespace metro {

struct StrictNet;   // weights (double) + activation buffers + launch list

// quant = 0: every tensor float64 (METRO_PREC_STRICT): what the reference's graph computes in exact arithmetic.
// quant = 1: float64 arithmetic with operands rounded to float16 at the storage points of the reference's
//            default float16 graph (METRO_PREC_STRICT_F16): its ideal evaluation, free of summation-order noise.
metro_status strict_build(const NetPlan &plan, const float *blob, int max_batch, int quant, float box_size_mm,
                          const std::vector<int32_t> &perm, bool keep, StrictNet **out);
// poses_dev may be null when coords01_dev (float32 [n, J, 3]: heatmap coordinates in [0,1], model joint order) is given
metro_status strict_run(StrictNet *net, const void *images_dev, bool u8, int n, float *poses_dev, cudaStream_t stream,
                        float *coords01_dev = nullptr);
void strict_destroy(StrictNet *net);
size_t strict_bytes(const StrictNet *net);
// named activation of the last run (float64, NHWC); false if unknown
bool strict_debug(const StrictNet *net, const std::string &name, const double **ptr, size_t *elems_per_crop);

}  // namespace metro
