/* sleapnn_b200.h - C ABI of libsleapnn_b200.so, the sm_100a implementation of sleap-nn's
 * heat-map post-processing / target-synthesis hot path.
 *
 * The reference (talmolab/sleap-nn v0.3.3) has no FFI for this path: its boundary is a set of
 * module-level Python functions.  Each entry point below names the
 * reference function(s) it replaces as `file:line` relative to sleap_nn/.  The Python shim in
 * sleap_nn_b200/ keeps the reference signatures and calls these through ctypes (see
 * INTEGRATION.md for the reference-side binding).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with h_; the caller (torch) owns
 *     all memory, nothing is allocated or freed here and no entry point synchronises;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it;
 *   - strides are in ELEMENTS, not bytes;
 *   - return value: SNB_OK (0) or a negative SNB_ERR_* for argument / launch errors detected on
 *     the host; conditions only the device can see (capacity overflow, infeasible assignment,
 *     bad index) are OR-ed into the caller-provided `status` word (SNB_STATUS_* bits), which the
 *     caller reads whenever it next synchronises;
 *   - variable-length results use caller-provided fixed-capacity buffers plus a device count.
 */
#ifndef SLEAPNN_B200_H_
#define SLEAPNN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNB_ABI_VERSION 1

#define SNB_OK 0
#define SNB_ERR_BAD_ARG (-1)
#define SNB_ERR_UNSUPPORTED (-2)
#define SNB_ERR_CUDA_LAUNCH (-3)

#define SNB_STATUS_PEAK_OVERFLOW 1      /* a frame had more peaks than `cap`                     */
#define SNB_STATUS_CAND_OVERFLOW 2      /* a frame had more candidates than `cand_cap`           */
#define SNB_STATUS_LSAP_INFEASIBLE 4    /* scipy would raise "cost matrix is infeasible"         */
#define SNB_STATUS_LSAP_TOO_LARGE 8     /* an edge had more peaks per node than the solver cap   */
#define SNB_STATUS_INSTANCE_OVERFLOW 16 /* a frame produced more instances than `inst_cap`       */
#define SNB_STATUS_BAD_INDEX 32         /* an index argument pointed outside its table           */
#define SNB_STATUS_MATCH_OVERFLOW 64    /* a frame had more matches than `match_cap`             */

int snb_abi_version(void);

/* ---------------------------------------------------------------- peaks (inference/ops/peaks.py)
 *
 * snb_local_peaks: fused 3x3 NMS + threshold + ordered peak emission + integral refinement.
 *   Replaces find_local_peaks_rough (ops/peaks.py:184-218), find_local_peaks (:221-259),
 *   morphological_dilation (:26-63) and the refinement's make_centered_bboxes / crop_bboxes /
 *   integral_regression chain (data/instance_cropping.py:129-171, ops/crops.py:31-124,
 *   ops/peaks.py:66-86).
 *   cms: (B,C,H,W) fp32 with element strides (sb,sc,sh,sw).  threshold is already fp32 (the
 *   reference compares against float32(threshold)).  refine_size = 0 -> rough peaks, else the
 *   integral patch size.  xy_scale multiplies the final coordinates (1 = none;
 *   layers/bottomup.py:111 uses the confmap stride).
 *   Output is a padded per-frame table: frame b owns slots [b*cap, b*cap + min(count,cap)),
 *   sorted by (y, x, channel) = the reference's torch.where order.  frame_count[b] holds the TRUE
 *   count (may exceed cap -> SNB_STATUS_PEAK_OVERFLOW).  keys: scratch of B*cap u32.
 */
int snb_local_peaks(const float* cms, int B, int C, int H, int W, long long sb, long long sc, long long sh,
                    long long sw, float threshold, int refine_size, float xy_scale, int cap, int* frame_count,
                    uint32_t* keys, float* out_xy, float* out_val, int* out_chan, int* status, void* stream);

/* Padded table -> the reference's concatenated (points, vals, sample_inds, channel_inds). */
int snb_pack_peaks(const int* frame_count, int B, int cap, const float* xy, const float* val, const int* chan,
                   float* o_xy, float* o_val, int* o_sample, int* o_chan, void* stream);

/* snb_global_peaks: find_global_peaks_rough (ops/peaks.py:89-130) + find_global_peaks (:133-181).
 *   workspace: snb_global_peaks_workspace() bytes, zero-filled before FIRST use (self-resetting).
 *   out_xy (B*C*2), out_val (B*C). */
int snb_global_peaks_workspace(int B, int C, int H, int W, int* rows_per_chunk, int* n_chunks, long long* n_bytes);
int snb_global_peaks(const float* cms, int B, int C, int H, int W, long long sb, long long sc, long long sh,
                     long long sw, float threshold, int refine_size, void* workspace, float* out_xy, float* out_val,
                     void* stream);

/* crop_bboxes (ops/crops.py:31-124).  images (S,C,H,W) of elem_size bytes (1,2,4,8), bboxes
 * (n,4,2) fp32, sample_inds int64; crop_h/crop_w are read from bbox 0 by the caller, as the
 * reference does (ops/crops.py:66-67).  out (n,C,crop_h,crop_w) contiguous. */
int snb_crop_bboxes(const void* images, int elem_size, int S, int C, int H, int W, long long sb, long long sc,
                    long long sh, long long sw, const float* bboxes, const long long* sample_inds, long long n,
                    int crop_h, int crop_w, void* out, int* status, void* stream);

/* make_centered_bboxes (data/instance_cropping.py:129-171): centers (n,2) -> out (n,4,2),
 * half_h = box_height / 2, half_w = box_width / 2 (already fp32). */
int snb_centered_bboxes(const float* centers, long long n, float half_h, float half_w, float* out, void* stream);

/* integral_regression (ops/peaks.py:66-86) on contiguous (n_planes,h,w) patches. */
int snb_integral_regression(const float* patches, long long n_planes, int h, int w, const float* xv,
                            const float* yv, float* out_x, float* out_y, void* stream);

/* morphological_dilation (ops/peaks.py:26-63) on contiguous (n_planes,H,W). */
int snb_dilate8(const float* image, long long n_planes, int H, int W, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SLEAPNN_B200_H_ */
