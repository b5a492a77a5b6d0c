/*
 * metro.h -- C-ABI of libmetro.so: the B200-native drop-in for the MeTRo inference hot path
 *            (RGB crops -> ResNet-v2-50/101 -> J x D volumetric heatmaps -> soft-argmax -> mm).
 *
 * Every entry point cites the reference interface it replaces (paths relative to the reference
 * repository isarandi/metro-pose3d).  The reference exposes this path as a TensorFlow frozen
 * graph: feed 'input:0', fetch 'output' / 'joint_edges' / 'joint_names' (inference.py:31-43,
 * src/main.py:106-161).  A C-ABI replacement therefore has: create (== import_graph_def + weight
 * constants), infer (== sess.run(poses_tensor)), joint tables (== the two Const nodes), destroy.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in signatures (streams travel as void*).
 *   - every function returns a metro_status; the message for the last failure on the calling
 *     thread is available from metro_last_error().  Nothing throws or exits across the ABI.
 *   - METRO_ERR_VALUE marks the conditions for which the reference raises ValueError
 *     (src/model/resnet_v2.py:213-214, src/model/resnet_utils.py:333,345,348).
 *   - device buffers are owned by the caller; the handle owns weights + a workspace arena sized at
 *     create time for `max_batch`.  No allocation, no synchronisation inside metro_infer.
 *   - a handle is bound to one device and is not thread-safe (the reference runs one tf.Session
 *     on the caller's thread, src/helpers.py:123-127).
 */
#ifndef METRO_H_
#define METRO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum metro_status {
  METRO_OK = 0,
  METRO_ERR_VALUE = 1,       /* invalid argument: the reference raises ValueError here          */
  METRO_ERR_CUDA = 2,        /* a CUDA runtime / driver call failed                              */
  METRO_ERR_NO_DEVICE = 3,   /* no sm_100 device: the product path refuses to run (no fallback)  */
  METRO_ERR_NOMEM = 4,
  METRO_ERR_INTERNAL = 5
} metro_status;

/* Element type of the head tensor (output of the 1x1 'logits' conv, input of the soft-argmax).
 * The reference computes the logits in FLAGS.dtype (float16 by default, src/options.py:73) and
 * casts to float32 (src/model/architectures.py:34). */
typedef enum metro_dtype { METRO_F32 = 0, METRO_F16 = 1 } metro_dtype;

/* Arithmetic of the backbone.  The reference exports its graph in FLAGS.dtype (src/options.py:73,
 * src/init.py:54-59, cast at src/model/architectures.py:29): float16 by default, float32 on request.
 *   METRO_PREC_F16        tensor-core path: float16 operands, float32 accumulation (the default export).
 *   METRO_PREC_STRICT     every tensor float64 on CUDA cores: within float64 summation order of the exact
 *                         graph, i.e. tighter than the float32 export (a float32 evaluation is itself
 *                         3-5e-3 mm from the exact graph).  ~50x slower; for
 *                         verification and accuracy-critical callers.
 *   METRO_PREC_STRICT_F16 float64 arithmetic with the float16 graph's roundings at its storage points:
 *                         the float16 export free of summation-order noise (verification). */
typedef enum metro_precision { METRO_PREC_F16 = 0, METRO_PREC_STRICT = 1, METRO_PREC_STRICT_F16 = 2 } metro_precision;

/* Static description of one exported model; the fields are the reference's FLAGS that shape the
 * frozen graph (src/options.py:41,96,113,118,119; src/main.py:119-127). */
typedef struct metro_spec {
  int32_t arch;               /* 50 | 101           FLAGS.architecture resnet_v2_50 / resnet_v2_101 */
  int32_t stride;             /* 4|8|16|32          FLAGS.stride_test                               */
  int32_t n_joints_model;     /* joints predicted by the head; the root is the LAST one            */
  int32_t depth;              /* 8                  FLAGS.depth                                     */
  int32_t centered_stride;    /* 1                  FLAGS.centered_stride                           */
  int32_t proc_side;          /* 256                FLAGS.proc_side                                 */
  float box_size_mm;          /* 2200               FLAGS.box_size_mm                               */
  int32_t n_joints_out;       /* length of `permutation`                                            */
  const int32_t *permutation; /* output joint j = model joint permutation[j]  (main.py:119-127)     */
  int32_t max_batch;          /* arena is sized for this many crops per metro_infer call            */
  int32_t head_dtype;         /* metro_dtype of the head tensor kept in HBM                         */
  int32_t keep_activations;   /* debug: give every layer its own buffer (metro_debug_read)          */
  int32_t precision;          /* metro_precision; 0 = the tensor-core float16 path                  */
  /* the two constant fetches of the frozen graph (src/main.py:128,140-141), returned by
   * metro_get_joint_info; both may be NULL / 0 when the caller keeps the tables itself */
  const char *joint_names;    /* n_joints_out names in OUTPUT order, separated by '\n'               */
  int32_t n_joint_edges;
  const int32_t *joint_edges; /* [n_joint_edges][2] indices into the output joints                  */
} metro_spec;

typedef struct metro_handle metro_handle;

/* ---- library ---------------------------------------------------------------------------------- */
const char *metro_last_error(void);
const char *metro_version(void);

/* ---- plan (host only, usable without a GPU) ---------------------------------------------------- */
/* Number of float32 values the weight blob must hold for `spec`
 * (== all variables under MainPart/resnet_v2_*, src/model/architectures.py:24). */
metro_status metro_blob_floats(const metro_spec *spec, uint64_t *n_floats);
/* Writes a JSON description of the layer plan (the replay of resnet_utils.py:307-350) into `buf`;
 * `*needed` receives the byte count including the terminator. */
metro_status metro_plan_describe(const metro_spec *spec, char *buf, size_t buf_bytes, size_t *needed);

/* ---- model: replaces estimate_pose()'s graph import (inference.py:31-38) ----------------------- */
/* `weights_blob`: host float32 array in the order documented in metro_pose3d_b200/weights.py
 * (TF variable order, conv filters HWIO, BatchNorm as gamma/beta/moving_mean/moving_variance). */
metro_status metro_create(const metro_spec *spec, const float *weights_blob, uint64_t n_floats,
                          int32_t device, metro_handle **out);
metro_status metro_destroy(metro_handle *h);
/* Device bytes a handle of this model holds for batches of up to `n` crops: weights plus the
 * activation arena, which is allocated once at create time for spec.max_batch (n <= 0 or n >
 * max_batch report the whole arena).  No allocation ever happens inside metro_infer. */
metro_status metro_workspace_bytes(const metro_handle *h, int32_t n, uint64_t *bytes);

/* ---- the graph's two constant fetches: 'joint_names' and 'joint_edges' (inference.py:36-38,
 *      src/main.py:128,140-141), as given in the spec at create time (dataset tables or the constants of
 *      an imported .pb).  names_buf receives the names separated by '\n' (NUL-terminated, truncated to
 *      names_bytes; *names_needed = bytes including the terminator); edges_buf receives up to edges_cap
 *      (a, b) pairs; *n_edges / *n_joints the counts.  Any output pointer may be NULL. ------------- */
metro_status metro_get_joint_info(const metro_handle *h, char *names_buf, size_t names_bytes,
                                  size_t *names_needed, int32_t *edges_buf, int32_t edges_cap,
                                  int32_t *n_edges, int32_t *n_joints);

/* ---- inference: replaces sess.run(poses_tensor) (inference.py:25-27) --------------------------- */
/* images_dev: device float32 NHWC [n,256,256,3] in [0,1] ('input:0', main.py:109-110)
 * poses_dev : device float32 [n, n_joints_out, 3], root-relative millimetres ('output', main.py:127)
 * stream    : cudaStream_t as void* (NULL = default stream).  Asynchronous. */
metro_status metro_infer(metro_handle *h, const float *images_dev, int32_t n, float *poses_dev,
                         void *stream);
/* Same through HOST buffers: pinned or pageable host memory in, host memory out; copies and a
 * final stream synchronisation happen inside the call (this is what sess.run does for numpy
 * feeds, inference.py:26-27). */
metro_status metro_infer_host(metro_handle *h, const float *images_host, int32_t n,
                              float *poses_host);
/* uint8 RGB crops (the /255 of src/improc.py:56-61 is fused into the first
 * kernel).  images_u8_dev: device uint8 NHWC [n,256,256,3]. */
metro_status metro_infer_u8(metro_handle *h, const uint8_t *images_u8_dev, int32_t n,
                            float *poses_dev, void *stream);
/* The same through HOST buffers: uint8 crops as the reference's loader holds them before
 * src/data/data_loading.py:102-103 / src/improc.py:56-61 turn them into floats -- a quarter of the
 * PCIe bytes of metro_infer_host. */
metro_status metro_infer_host_u8(metro_handle *h, const uint8_t *images_u8_host, int32_t n,
                                 float *poses_host);

/* ---- stand-alone soft-argmax: replaces net_output_to_heatmap_and_coords + heatmap_to_metric +
 *      root_relative + gather (volumetric.py:227-235,288-306; tfu.py:466-499; tfu3d.py:23-25;
 *      main.py:127) ------------------------------------------------------------------------------ */
typedef struct metro_softargmax_desc {
  int32_t side;               /* heatmap height == width (256 / stride)                            */
  int32_t n_joints_model;
  int32_t depth;
  int32_t stride;
  int32_t centered_stride;
  int32_t proc_side;
  float box_size_mm;
  int32_t n_joints_out;
  const int32_t *permutation;
  int32_t head_dtype;         /* metro_dtype                                                       */
  int32_t splits;             /* 0 = choose (shape-only rule); >0 = CTAs per crop (tuning / tests)  */
  int32_t lanes;              /* 0 = choose; >0 = pixel lanes per CTA, rounded down to a power of
                               * two <= 32 (tuning / tests)                                       */
  int32_t word_bytes;         /* 0 = choose (16 for fp32 heads, 8 for fp16); 8 | 16 = bytes of one
                               * pixel a thread owns (tuning)                                     */
} metro_softargmax_desc;

metro_status metro_softargmax_workspace_bytes(const metro_softargmax_desc *d, int32_t n,
                                              uint64_t *bytes);
/* head_dev: device NHWC [n, side, side, depth*n_joints_model], channel c = d*J + j
 * (volumetric.py:231-232).  workspace_dev must be zero-filled once after allocation; the kernel
 * leaves it zeroed. */
metro_status metro_softargmax(const metro_softargmax_desc *d, const void *head_dev, int32_t n,
                              float *poses_dev, void *workspace_dev, void *stream);

/* The same decode with the evaluation graph's other fetches.  coords01_dev [n, n_joints_model, 3]
 * receives the heatmap coordinates in [0,1] of every model joint (what net_output_to_heatmap_and_coords
 * returns, volumetric.py:234) -- the input of the absolute-scale variants below; poses_dev may then be NULL. */
metro_status metro_softargmax_coords(const metro_softargmax_desc *d, const void *head_dev, int32_t n,
                                     float *poses_dev, float *coords01_dev, void *workspace_dev, void *stream);
/* Whole network with that second fetch (tensor-core or strict handles; float32 images). */
metro_status metro_infer_coords(metro_handle *h, const float *images_dev, int32_t n, float *poses_dev,
                                float *coords01_dev, void *stream);
/* t.heatmap_pred_z = reduce_sum(softmaxed, axis=[2, 3]) (volumetric.py:165): the depth marginal of every model
 * joint's heatmap, out_dev float32 [n, n_joints_model, depth].  head_dev as for metro_softargmax. */
metro_status metro_heatmap_z(const metro_softargmax_desc *d, const void *head_dev, int32_t n, float *out_dev,
                             void *stream);
/* Absolute-scale variant 'true-root-depth' of build_inference_model (volumetric.py:190-198) and back_project
 * (:285): out[b,c,:] = (inv_intrinsics[b] @ [u, v, 1]) * ((z[b,c] - z[b,root]) * box_size_mm + z_offset[b]) with
 * (u, v) = heatmap_to_image(coords01[b,c,:2]) (:288-295) and root = the LAST joint.  coords01_dev / out_dev float32
 * [n, n_joints, 3] (model joint order), inv_intrinsics_dev float32 [n,3,3] row-major, z_offset_dev float32 [n]: the
 * true root depth ('true-root-depth') or the offset a bone-length fit produced ('bone-lengths*', whose optimiser is
 * host code in the reference too: a scipy py_func, src/model/bone_length_based_backproj.py). */
metro_status metro_back_project(const float *coords01_dev, const float *inv_intrinsics_dev, const float *z_offset_dev,
                                int32_t n, int32_t n_joints, int32_t stride, int32_t centered_stride, int32_t proc_side,
                                float box_size_mm, float *out_dev, void *stream);

/* ---- pre-path: crop extraction.  Replaces cameralib.reproject_image_fast (src/cameralib.py:406-429,
 *      called from src/data/data_loading.py:93): a homography warp of a full frame to a side x side crop with OpenCV's
 *      fixed-point bilinear cv2.remap(INTER_LINEAR, BORDER_CONSTANT) arithmetic, bit for bit.  Each source names a uint8 RGB
 *      frame in device memory and the float32 row-major homography that maps OUTPUT pixel (x, y, 1) to frame coordinates
 *      (np.linalg.solve(new_matrix.T, old_matrix.T).T, cameralib.py:411-413).  crops_u8_dev: uint8 [n, side, side, 3], the
 *      input layout of metro_infer_u8.  `srcs` is a HOST array; it travels in kernel parameters (no allocation, async). ---- */
typedef struct metro_crop_src {
  const uint8_t *frame_dev;   /* uint8 [height, width, 3], rows row_stride_bytes apart                       */
  int32_t height, width, row_stride_bytes;
  float homography[9];
} metro_crop_src;
metro_status metro_extract_crops(const metro_crop_src *srcs, int32_t n, int32_t side, int32_t border_value,
                                 uint8_t *crops_u8_dev, void *stream);

/* ---- post-path: back into the original camera frame.  Replaces to_orig_cam
 *      (volumetric.py:277-282): out[b,c,:] = R[b] @ poses[b,c',:], c' = c where det(R[b]) > 0, else
 *      mirror_mapping[c] (the crop was flipped, data_loading.py:80-83; JointInfo.mirror_mapping,
 *      datasets.py:76-79).  poses_dev / out_dev: device float32 [n, n_joints, 3] (must not alias);
 *      rot_dev: device float32 [n, 3, 3] row-major; mirror_mapping: HOST int32 [n_joints]. ----------- */
metro_status metro_to_orig_cam(const float *poses_dev, const float *rot_dev,
                               const int32_t *mirror_mapping, int32_t n, int32_t n_joints,
                               float *out_dev, void *stream);

/* ---- single fused convolution (operator-level entry point used by the parity tests) ----------- */
typedef struct metro_conv_desc {
  int32_t n, in_side, cin, cout, k, stride, rate, pad_lo;  /* conv2d_same geometry (resnet_utils.py:82-135) */
  int32_t relu;                 /* apply ReLU to output 1                                           */
  int32_t out_dtype;            /* metro_dtype of output 1                                          */
  int32_t res_stride, res_shift;/* residual = res[:, s::stride, s::stride, :] (resnet_v2.py:120-121); 0 = none */
  int32_t cin2;                 /* >0: second 1x1 source fused into the same accumulator (shortcut) */
} metro_conv_desc;
/* x: fp16 NHWC [n,in,in,cin]; w: host float32 HWIO [k,k,cin,cout]; x2/w2: optional second source
 * (fp16 NHWC [n,out,out,cin2], host float32 [1,1,cin2,cout]); scale/shift: host float32 [cout]
 * (folded BN or 1/bias); res: optional fp16 NHWC; y: output 1; scale2/shift2 + y2: optional second
 * output y2 = relu(scale2 * fp16(y) + shift2) as fp16 (the next unit's pre-activation,
 * resnet_v2.py:119). */
metro_status metro_conv2d(const metro_conv_desc *d, const void *x_dev, const float *w_host,
                          const void *x2_dev, const float *w2_host, const float *scale_host,
                          const float *shift_host, const void *res_dev, void *y_dev,
                          const float *scale2_host, const float *shift2_host, void *y2_dev,
                          int32_t device, void *stream);

/* ---- debug ------------------------------------------------------------------------------------- */
/* With spec.keep_activations: copy the named activation of the last metro_infer call to host.
 * Names: "conv1", "pool1", "<unit>/conv1", "<unit>/conv2", "<unit>/out", "<unit>/pre", "head".
 * `*elems` receives the element count; dtype is fp16 except "head" (spec.head_dtype); a strict-precision
 * handle returns float64 for every name. */
metro_status metro_debug_read(metro_handle *h, const char *name, void *host_buf, uint64_t buf_bytes,
                              uint64_t *elems);
/* Per-launch device times (ms) of the last metro_profile call, one entry per kernel launch. */
metro_status metro_profile(metro_handle *h, const float *images_dev, int32_t n, float *poses_dev,
                           float *ms_out, char *names_buf, size_t names_bytes, int32_t *n_launches);
/* Small batches are launch-bound (52 launches of a few microseconds each for ResNet-50), so metro_infer /
 * metro_infer_u8 capture a call's launch sequence into a CUDA graph the second time they see the same
 * (n, dtype, image buffer, pose buffer) with n <= 32 (environment METRO_GRAPH_MAX_BATCH; 0 disables) and
 * replay it afterwards -- the role the TensorFlow session's cached executor plays for the reference's
 * sess.run (inference.py:25-27).  Up to 8 graphs per handle, least recently used evicted.  Reports the number
 * of instantiated graphs and of replays so far. */
metro_status metro_graph_stats(const metro_handle *h, int32_t *graphs, int64_t *replays);
/* Number of kernel launches one metro_infer(n) enqueues. */
metro_status metro_launch_count(const metro_handle *h, int32_t n, int32_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* METRO_H_ */
