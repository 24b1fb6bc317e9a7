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

#define SNB_ABI_VERSION 5

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
#define SNB_STATUS_LSAP_INVALID 128     /* scipy would raise "matrix contains invalid numeric entries" (NaN / -inf cost) */
#define SNB_STATUS_ASM_MISMATCH 256     /* make_predicted_instances' sanity assert would fail (ops/paf.py:873): a scored   */
                                        /* connection whose two peaks ended up in different instances                      */
#define SNB_STATUS_ASM_MISSING 512      /* ... or whose destination peak is in no (kept) instance: KeyError there          */

/* ABI v5: element type of the confidence-map / PAF tensors.  Under autocast the reference's backbone emits fp16 heads
 * and casts them back with .float() before any of these ops run (inference/layers/backends/torch_backend.py:125-146).
 * Every fp16 / bf16 value is exact in fp32, so the *_t entry points read the half tensors in place, up-cast in
 * registers and compute exactly what the fp32 ops compute on the .float() copy - without that copy's extra read +
 * write of the largest tensors.  The threshold stays an fp32 number (pass the dtype-rounded threshold to get the
 * semantics of calling the reference ops directly on a half tensor, where torch compares in the tensor's dtype). */
#define SNB_DTYPE_F32 0
#define SNB_DTYPE_F16 1
#define SNB_DTYPE_BF16 2

int snb_abi_version(void);

/* Device-visible alias of a PINNED host buffer (cudaHostGetDevicePointer).  Lets the bottom-up chain
 * sample the PAF tensor in place over PCIe when the maps arrive in host memory: only the sampled
 * sectors cross the link instead of the whole tensor.  SNB_ERR_UNSUPPORTED if the buffer is not
 * pinned / mapped. */
int snb_host_device_pointer(void* host_ptr, void** device_ptr);

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

/* The two halves of snb_local_peaks, exposed so that a pipeline can put them on different streams.
 * snb_local_peaks_detect zeroes frame_count and runs the streaming NMS kernel (the only kernel that
 * moves real bytes); ev_begin / ev_end are optional cudaEvent_t handles recorded right around it so
 * a benchmark can time the dominant kernel in situ.  snb_local_peaks_finalize sorts each frame's
 * keys, reads the values, refines and fills the padded table. */
int snb_local_peaks_detect(const float* cms, int B, int C, int H, int W, long long sb, long long sc, long long sh,
                           long long sw, float threshold, int cap, int* frame_count, uint32_t* keys, void* ev_begin,
                           void* ev_end, void* stream);
int snb_local_peaks_finalize(const float* cms, int B, int C, int H, int W, long long sb, long long sc, long long sh,
                             long long sw, int refine_size, float xy_scale, int cap, const int* frame_count,
                             uint32_t* keys, float* out_xy, float* out_val, int* out_chan, int* status, void* stream);

/* ABI v5: the same three entry points for fp32 / fp16 / bf16 maps (dtype = SNB_DTYPE_*; strides in elements of
 * that type).  The fp32-only names above are kept and forward here with SNB_DTYPE_F32. */
int snb_local_peaks_t(const void* cms, int dtype, int B, int C, int H, int W, long long sb, long long sc,
                      long long sh, long long sw, float threshold, int refine_size, float xy_scale, int cap,
                      int* frame_count, uint32_t* keys, float* out_xy, float* out_val, int* out_chan, int* status,
                      void* stream);
int snb_local_peaks_detect_t(const void* cms, int dtype, int B, int C, int H, int W, long long sb, long long sc,
                             long long sh, long long sw, float threshold, int cap, int* frame_count, uint32_t* keys,
                             void* ev_begin, void* ev_end, void* stream);
int snb_local_peaks_finalize_t(const void* cms, int dtype, int B, int C, int H, int W, long long sb, long long sc,
                               long long sh, long long sw, int refine_size, float xy_scale, int cap,
                               const int* frame_count, uint32_t* keys, float* out_xy, float* out_val, int* out_chan,
                               int* status, void* stream);

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

/* The coordinate ladder of the inference layers (inference/ops/coord.py:27-90), fusable into the peak kernels'
 * epilogues.  Applied in this order, each step one separately rounded fp32 op like the tensor op it replaces:
 *   xy * stride (undo_stride) -> / input_scale (undo_input_scale) -> / eff_scale[b] (undo_eff_scale)
 *   -> + crop_offset[b] (add_crop_offset) -> / eff_scale2[b] (TopDownLayer, layers/topdown.py:268-272).
 * NULL pointers skip a step; multiplying / dividing by exactly 1.0f is the identity.  scatter[b] (snb_global_peaks_ex
 * only) redirects sample b's output row (TopDownLayer's valid_idx scatter, layers/topdown.py:281-289); < 0 drops it. */
typedef struct snb_coord_ladder {
  float stride;
  float input_scale;
  const float* eff_scale;   /* (B,) */
  const float* crop_offset; /* (B, 2) x, y */
  const float* eff_scale2;  /* (B,) */
  const int* scatter;       /* (B,) */
} snb_coord_ladder;

/* snb_global_peaks + the ladder: CenteredInstanceLayer.postprocess (layers/centered_instance.py:199-230) and the
 * un-crop step of TopDownLayer._run_stage_2 (layers/topdown.py:259-289) in the same launch.  out_xy / out_val are
 * indexed by the scattered row; rows nobody writes keep whatever the caller put there (NaN-fill them first). */
int snb_global_peaks_ex(const float* cms, int B, int C, int H, int W, long long sb, long long sc, long long sh,
                        long long sw, float threshold, int refine_size, void* workspace,
                        const snb_coord_ladder* ladder, float* out_xy, float* out_val, void* stream);

/* ABI v5: snb_global_peaks_ex for fp32 / fp16 / bf16 maps (ladder may be NULL). */
int snb_global_peaks_t(const void* cms, int dtype, int B, int C, int H, int W, long long sb, long long sc,
                       long long sh, long long sw, float threshold, int refine_size, void* workspace,
                       const snb_coord_ladder* ladder, float* out_xy, float* out_val, void* stream);

/* CentroidLayer.postprocess after find_local_peaks (layers/centroid.py:196-258) on the padded table written by
 * snb_local_peaks (whose xy_scale already applied the stride): / input_scale, per-frame top max_instances by value
 * when there are more (torch.topk order), NaN padding, / eff_scale[b].  out_xy (B,max_instances,2), out_val (B,max_instances). */
int snb_peaks_topk(const int* frame_count, int B, int cap, const float* xy, const float* val, int max_instances,
                   float input_scale, const float* eff_scale, float* out_xy, float* out_val, void* stream);

/* undo_stride / undo_input_scale / undo_eff_scale / add_crop_offset (ops/coord.py:27-90) as one elementwise op on
 * contiguous (n_samples, pairs_per_sample, 2) coordinates. */
int snb_coord_ladder_apply(const float* xy, long long n_samples, long long pairs_per_sample,
                           const snb_coord_ladder* ladder, float* out, void* stream);

/* apply_input_scale (ops/coord.py:93-109): bilinear resize (align_corners=False, no antialias) of `planes` planes of
 * H x W (element strides sp, sh, sw) to contiguous (planes, oh, ow).  dtype: 0 fp32, 1 fp16, 2 bf16. */
int snb_bilinear_resize(const void* in, int dtype, long long planes, int H, int W, long long sp, long long sh,
                        long long sw, int oh, int ow, void* out, void* stream);

/* crop_bboxes (ops/crops.py:31-124).  images (S,C,H,W) of elem_size bytes (1,2,4,8), bboxes
 * (n,4,2) fp32, sample_inds int64; crop_h/crop_w are read from bbox 0 by the caller, as the
 * reference does (ops/crops.py:66-67).  out (n,C,crop_h,crop_w) contiguous. */
int snb_crop_bboxes(const void* images, int elem_size, int S, int C, int H, int W, long long sb, long long sc,
                    long long sh, long long sw, const float* bboxes, const long long* sample_inds, long long n,
                    int crop_h, int crop_w, void* out, int* status, void* stream);

/* make_centered_bboxes (data/instance_cropping.py:129-171): centers (n,2) -> out (n,4,2),
 * half_h = box_height / 2, half_w = box_width / 2 (already fp32). */
int snb_centered_bboxes(const float* centers, long long n, float half_h, float half_w, float* out, void* stream);

/* TopDownLayer stage B + the bookkeeping of stage 2 (layers/topdown.py:98-120, 186-236, 415-466), one launch:
 * valid = no NaN coordinate; optional greedy centroid NMS per frame (descending value, IoU of crop_h x crop_w boxes
 * centred on the centroids > nms_threshold drops; frames with <= 1 valid centroid untouched); then the valid (b, i)
 * pairs in torch.nonzero order as a crop list.  centroids (B, I, 2) are in IMAGE space; eff_scale (B) or NULL takes them
 * to sized space (x eff) for the boxes.  Outputs: n_valid (3 ints: the crop count, then the crop height and width
 * crop_bboxes would read off bbox 0, ops/crops.py:66-67), frame_off (B+1), and with capacity B*I rows:
 * sample_inds (int64), rows (= b*I + i), crop_bboxes (., 4, 2) = make_centered_bboxes in sized space, crop_topleft
 * (., 2), crop_eff (.); per slot: row_to_crop (B*I) (-1 = none), valid_mask (B, I) bytes, centroids_img (B, I, 2) =
 * (c x eff) / eff, full_bboxes (B, I, 4, 2) = box / eff or NaN.  Needs snb_topdown_select_smem_bytes(B, I) <= 200 KB. */
long long snb_topdown_select_smem_bytes(int B, int I);
int snb_topdown_select(const float* centroids, const float* centroid_vals, int B, int I, const float* eff_scale,
                       int crop_h, int crop_w, int centroid_nms, float nms_threshold, int* n_valid, int* frame_off,
                       long long* sample_inds, int* rows, int* row_to_crop, float* crop_bboxes, float* crop_topleft,
                       float* crop_eff, unsigned char* valid_mask, float* centroids_img, float* full_bboxes,
                       void* stream);

/* The lift of stage 2 (layers/topdown.py:259-291): stage-2 keypoints kpts (n, n_nodes, 2) / vals (n, n_nodes) of the
 * crop list -> full_kpts (n_slots, n_nodes, 2) = (kpts + crop_topleft) / crop_eff, full_crop_kpts (same shape, may be
 * NULL) = kpts, full_vals (n_slots, n_nodes); slots without a crop get NaN.  n_slots = B * max_instances. */
int snb_topdown_lift(const float* kpts, const float* vals, int n_nodes, long long n_slots, const int* row_to_crop,
                     const float* crop_topleft, const float* crop_eff, float* full_kpts, float* full_crop_kpts,
                     float* full_vals, void* stream);

/* integral_regression (ops/peaks.py:66-86) on contiguous (n_planes,h,w) patches. */
int snb_integral_regression(const float* patches, long long n_planes, int h, int w, const float* xv,
                            const float* yv, float* out_x, float* out_y, void* stream);

/* morphological_dilation (ops/peaks.py:26-63) on contiguous (n_planes,H,W). */
int snb_dilate8(const float* image, long long n_planes, int H, int W, float* out, void* stream);

/* ------------------------------------------------------- PAF grouping (inference/ops/paf.py)
 *
 * Tables.  Per-frame variable-length data lives either in a PADDED table (start == NULL: frame b
 * begins at b*stride, counts are clamped to stride) or a CSR table (start[b] given, count[b]
 * exact).  Peak tables: peak_xy (2 floats, image scale), peak_val, peak_chan (node index).
 *
 * snb_paf_prepare: per-frame grouping of peaks by node and per-edge candidate / match offsets.
 *   Replaces the argsort / meshgrid bookkeeping of get_connection_candidates (paf.py:84-130).
 *   edges: (n_edges,2) int32 node ids.  Outputs: node_start (B,N+1), node_peaks (same addressing
 *   as the peak table), edge_off (B,E+1) exclusive prefix of n_src*n_dst, match_off (B,E+1)
 *   exclusive prefix of min(n_src,n_dst).  Canonical candidate order: edge-major, source peak
 *   ascending, destination peak ascending (= a stable argsort; the reference's argsort is
 *   unstable for n >= 17). */
int snb_paf_prepare(const int* peak_chan, const int* frame_start, int frame_stride, const int* frame_count, int B,
                    const int* edges, int n_nodes, int n_edges, int* node_start, int* node_peaks, int* edge_off,
                    int* match_off, void* stream);

/* snb_paf_score: candidate enumeration + line sampling + gather + score, one thread per candidate.
 *   Replaces get_connection_candidates, make_line_subs (paf.py:133-234 with inference/utils.py:29-130),
 *   get_paf_lines (:237-287), compute_distance_penalty (:290-332), score_paf_lines (:335-410) and
 *   the per-sample loop of score_paf_lines_batch (:413-497).
 *   pafs: the (B,H,W,2E) VIEW with element strides (pb,py,px,pc), read in place; NULL = enumerate
 *   candidates only.  t_table: torch.linspace(0,1,n_points) computed on the HOST (bit-exactness).
 *   stride = float(pafs_stride); max_edge_length = ratio*max(H,W,2E)*stride (paf.py:457-461).
 *   Output candidate table (padded by cand_stride, or CSR via cand_start): cand_edge i32,
 *   cand_epi (2 x int64 peak indices local to the frame), cand_score f32.
 *   max_cand_per_frame only sizes the grid. */
int snb_paf_score(const float* pafs, long long pb, long long py, long long px, long long pc, int H, int W,
                  const float* t_table, int n_points, float stride, float max_edge_length, float penalty_weight,
                  const float* peak_xy, const int* frame_start, int frame_stride, int B, const int* edges,
                  int n_nodes, int n_edges, const int* node_start, const int* node_peaks, const int* edge_off,
                  const int* cand_start, int cand_stride, int max_cand_per_frame, int* cand_edge,
                  long long* cand_epi, float* cand_score, int* status, void* stream);

/* ABI v5: snb_paf_score on an fp32 / fp16 / bf16 PAF tensor (dtype = SNB_DTYPE_*, strides in elements). */
int snb_paf_score_t(const void* pafs, int dtype, long long pb, long long py, long long px, long long pc, int H, int W,
                    const float* t_table, int n_points, float stride, float max_edge_length, float penalty_weight,
                    const float* peak_xy, const int* frame_start, int frame_stride, int B, const int* edges,
                    int n_nodes, int n_edges, const int* node_start, const int* node_peaks, const int* edge_off,
                    const int* cand_start, int cand_stride, int max_cand_per_frame, int* cand_edge,
                    long long* cand_epi, float* cand_score, int* status, void* stream);

/* make_line_subs (paf.py:133-234): out (M,n_points,2,3) int32 [row,col,channel]. */
int snb_line_subs(const float* peaks, long long n_peaks, const long long* epi, const int* edge_inds, long long M,
                  const float* t_table, int n_points, float stride, int H, int W, int* out, int* status,
                  void* stream);
/* The gather of get_paf_lines (paf.py:282-287): subs (n_sub,3) int32 into an (H,W,Cn) strided view. */
int snb_paf_gather(const float* pafs, long long py, long long px, long long pc, int H, int W, int Cn,
                   const int* subs, long long n_sub, float* out, int* status, void* stream);
/* score_paf_lines on pre-gathered lines (paf.py:335-410): lines (M,n_points,2). */
int snb_score_lines(const float* lines, const float* peaks, long long n_peaks, const long long* epi, long long M,
                    int n_points, float max_edge_length, float weight, float* out, int* status, void* stream);
/* compute_distance_penalty (paf.py:290-332). */
int snb_distance_penalty(const float* lengths, long long n, float max_edge_length, float weight, float* out,
                         void* stream);

/* Matching = scipy.optimize.linear_sum_assignment semantics on cost = -score, NaN -> +inf
 * (match_candidates_sample, paf.py:500-619).  Matches of frame b / edge k are written at
 * match_off[b][k] inside the frame's slot, rows ascending; m_src / m_dst are RANKS within the
 * node's peaks (paf.py:596-599).
 *   snb_match_structured: candidates are the full cross product written by snb_paf_score.
 *     ws: B*n_edges*snb_lsap_workspace_bytes(ws_max_dim) bytes, used when an edge has more than
 *     32 peaks per node (NULL/0 -> such edges raise SNB_STATUS_LSAP_TOO_LARGE).
 *   snb_match_generic: arbitrary candidate lists (CSR).  phase 0 writes dims (B,E,2) = distinct
 *     (src,dst) counts; phase 1 needs cost_off (B*E) offsets into cost / cell_src scratch
 *     (n_src*n_dst each), match_start (B*E) output offsets and ws as above. */
long long snb_lsap_workspace_bytes(int max_dim);
int snb_match_structured(const float* cand_score, const int* cand_start, int cand_stride, const int* edges,
                         int n_nodes, int n_edges, const int* node_start, const int* edge_off, const int* match_off,
                         const int* match_start, int match_stride, void* ws, int ws_max_dim, int B, int* m_edge,
                         int* m_src, int* m_dst, float* m_score, int* m_count, int* status, void* stream);
int snb_match_generic(int phase, const int* cand_edge, const long long* cand_epi, const float* cand_score,
                      const int* cand_start, const int* cand_count, int B, int n_edges, int max_peak_id, int* dims,
                      const long long* cost_off, double* cost, int* cell_src, const int* match_start, void* ws,
                      int ws_max_dim, int* m_edge, int* m_src, int* m_dst, float* m_score, int* status,
                      void* stream);

/* snb_assemble: min_line_scores filter + assign_connections_to_instances (paf.py:705-820) +
 * make_predicted_instances (:823-887) + the per-sample loop of group_instances_batch (:1041-1149).
 *   sorted_edges: toposort_edges() order (host, paf.py:890-912).  min_instance_peaks is already
 *   an int (the caller applies int(frac*n_nodes), paf.py:802).  ws: B*4*ws_stride int32.
 *   Outputs (NaN-filled): inst_xy (B,inst_cap,N,2), inst_val (B,inst_cap,N), inst_score
 *   (B,inst_cap), n_inst (B). */
int snb_assemble(const float* peak_xy, const float* peak_val, const int* peak_chan, const int* frame_start,
                 int frame_stride, const int* frame_count, int B, const int* node_start, const int* node_peaks,
                 int n_nodes, const int* edges, const int* sorted_edges, int n_sorted, const int* m_edge,
                 const int* m_src, const int* m_dst, const float* m_score, const int* match_start, int match_stride,
                 const int* m_count, int min_instance_peaks, float min_line_scores, int* ws, int ws_stride,
                 int inst_cap, float* inst_xy, float* inst_val, float* inst_score, int* n_inst, int* status,
                 void* stream);

/* make_predicted_instances (paf.py:823-887), dict-API form: n_assign assignments in insertion order
 * (xy, val, compacted instance index, node), n_conn connections in visiting order (instance index of
 * the source peak or -1, score).  Outputs NaN-filled (n_inst,N,2), (n_inst,N) and (n_inst,). */
int snb_scatter_instances(const float* xy, const float* val, const int* inst, const int* node, int n_assign,
                          const int* conn_inst, const float* conn_score, int n_conn, int n_inst, int n_nodes,
                          float* o_xy, float* o_val, float* o_score, void* stream);

/* interp1d (inference/utils.py:29-130): x (x_rows,n), y (y_rows,n), xnew (xn_rows,p), *_rows in {1, rows};
 * out (rows,p). */
int snb_interp1d(const float* x, int x_rows, const float* y, int y_rows, const float* xnew, int xn_rows, int n, int p,
                 long long rows, float* out, void* stream);

/* ------------------------------------- training targets (data/confidence_maps.py, data/edge_maps.py)
 *
 * snb_confmaps: make_confmaps (confidence_maps.py:94-129, I = 1) and make_multi_confmaps (:132-166).
 *   points (G,I,N,2) fp32 (NaN = missing), xv (w), yv (h), den = fl32(2*sigma^2).
 *   out (G,N,h,w) fp32 or bf16 = max over I of nan_to_num(exp(-((xv-x)^2+(yv-y)^2)/den)). */
int snb_confmaps(const float* points, int G, int I, int N, const float* xv, const float* yv, int h, int w, float den,
                 int out_bf16, void* out, void* stream);

/* snb_pafs: make_pafs (edge_maps.py:120-164; accumulate = 0, I = 1, NaNs kept) and make_multi_pafs
 *   (:167-220; accumulate = 1: per-instance NaN -> 0, summed in instance order), for G frames in one launch
 *   (the reference API is per frame: G = 1).
 *   srcs/dsts (G,I,E,2); out (G,E,2,h,w) fp32 or bf16.  The weight is exp(-(d2*d2)/den) on the SQUARED
 *   point-segment distance d2, exactly as distance_to_edge + gaussian_pdf compose in the reference. */
int snb_pafs(const float* srcs, const float* dsts, int G, int I, int E, const float* xv, const float* yv, int h, int w,
             float den, int accumulate, int out_bf16, void* out, void* stream);

/* snb_confmaps with a strided point source, for batched dataset-side target generation (the call sites of
 *   generate_confmaps / generate_multiconfmaps / generate_class_maps in data/custom_datasets.py:1305-1327, 1489,
 *   1788, 2835, 2986): frame g / instance i / channel n reads its (x, y) pair at points + g*sg + i*si + n*sn
 *   (element strides), so (G,I,N,2), centroids (G,I,2) and the instance<->channel swapped view of
 *   generate_class_maps (data/identity.py:122-129) need no transposed copy.
 *   n_valid (G) or NULL: instances i >= n_valid[g] are missing (`instances[:, :num_instances]`,
 *   confidence_maps.py:79-84).  oob_w / oob_h > 0: filter_oob_points (data/providers.py:38-69) applied on the fly. */
int snb_confmaps_ex(const float* points, int G, int I, int N, long long sg, long long si, long long sn,
                    const int* n_valid, float oob_w, float oob_h, const float* xv, const float* yv, int h, int w,
                    float den, int out_bf16, void* out, void* stream);

/* generate_pafs (edge_maps.py:250-323) for G frames in one launch: get_edge_points (:223-247) gathers the edge
 *   endpoints from instances (G, I, N, 2) through edges (E, 2) int32 inside the kernel; in_xmax / in_ymax > 0 apply the
 *   in-image instance filter (:293-297; pass xv[-1], yv[-1]); then make_multi_pafs.  out (G, E, 2, h, w). */
int snb_pafs_from_instances(const float* instances, int G, int I, int N, const int* edges, int E, float in_xmax,
                            float in_ymax, const float* xv, const float* yv, int h, int w, float den, int out_bf16,
                            void* out, void* stream);

/* ABI v5: both targets of a bottom-up dataset item in one call - generate_multiconfmaps
 * (data/confidence_maps.py:46-91, with the datasets' `[:, :num_instances]` slice = n_valid and filter_oob_points =
 * oob_w / oob_h, as snb_confmaps_ex) and generate_pafs (data/edge_maps.py:250-323, with its in-image instance filter =
 * in_xmax / in_ymax, as snb_pafs_from_instances) for G frames, each head on its own grid (stride).  instances:
 * (G, I, N, 2); out_cms (G, N, h_cm, w_cm); out_pafs (G, E, 2, h_paf, w_paf); fp32 or bf16.
 * The two kernels are enqueued back to back as a programmatic-dependent-launch pair: the field kernel depends on
 * nothing the map kernel writes, starts while the map kernel is still running and shares the SMs with it, and ties its
 * own completion to the map kernel's - so a SINGLE frame (what Dataset.__getitem__ produces,
 * data/custom_datasets.py:1305-1327; 33.5 + 65 MB at cfg4 size, too little for either kernel alone to fill the GPU)
 * runs close to the store roofline.  Values are those of snb_confmaps_ex / snb_pafs_from_instances bit for bit. */
int snb_bottomup_targets(const float* instances, int G, int I, int N, const int* n_valid, float oob_w, float oob_h,
                         const int* edges, int E, float in_xmax, float in_ymax, const float* xv_cm, const float* yv_cm,
                         int h_cm, int w_cm, float den_cm, const float* xv_paf, const float* yv_paf, int h_paf, int w_paf,
                         float den_paf, int out_bf16, void* out_cms, void* out_pafs, void* stream);

/* Test hook: K7 divides by the per-launch constant 2*sigma^2 with a hoisted-reciprocal sequence instead of a full
 * div.rn per pixel; fast[i] = that sequence, exact[i] = __fdiv_rn(-a[i], den).  The parity tests require them equal
 * bit for bit. */
int snb_debug_neg_div(const float* a, long long n, float den, float* fast, float* exact, void* stream);

/* distance_to_edge (edge_maps.py:15-78; apply_pdf = 0) and make_edge_maps (:81-117; apply_pdf = 1).
 *   points (n_pts,2) or NULL for the (yv, xv) meshgrid with n_pts = h*w; out (n_pts,E). */
int snb_edge_distance(const float* points, const float* xv, const float* yv, int w, long long n_pts,
                      const float* src, const float* dst, int E, int apply_pdf, float den, float* out, void* stream);

/* gaussian_pdf (data/utils.py:114-125): out = exp(-(x*x)/den). */
int snb_gaussian_pdf(const float* x, long long n, float den, float* out, void* stream);

/* ------------------------------------ multi-class identity (inference/ops/identity.py, data/identity.py)
 *
 * snb_classify_peaks: group_class_peaks (ops/identity.py:13-71) and, with class_maps != NULL, the whole of
 *   classify_peaks_from_maps (:74-149) in one launch; one warp per (sample, channel) group.
 *   class_maps (n_samples, K, H, W) fp32 with element strides (ms, mk, mh, mw), or NULL when `probs` is given.
 *   Peaks are the concatenated lists find_local_peaks returns: peak_xy (P,2) in class-map pixels, peak_val (P),
 *   sample_inds / channel_inds (P) int32, in any order.
 *   probs (P, K): OUTPUT when class_maps != NULL (class_maps[sample, :, round(y), round(x)], half-to-even,
 *     clamped), INPUT otherwise.
 *   Per group the peaks (ascending index) are assigned to classes by scipy.optimize.linear_sum_assignment
 *   semantics on cost = -(double)prob; a match is kept only when prob == max over the peak's classes.
 *   g_peak / g_class (n_samples*n_channels, K) int64 + g_count (n_samples*n_channels): optional per-group kept
 *     matches in row order (feed snb_pack_class_matches).
 *   o_xy (n_samples, K, n_channels, 2), o_val, o_prob (n_samples, K, n_channels): optional fixed-size outputs,
 *     NaN-filled here.
 *   status bits: LSAP_INVALID (NaN or +inf probability: scipy raises ValueError), LSAP_INFEASIBLE,
 *     LSAP_TOO_LARGE (a group or K exceeds 128). */
int snb_classify_peaks(const float* class_maps, int n_samples, int K, int H, int W, long long ms, long long mk,
                       long long mh, long long mw, const float* peak_xy, const float* peak_val, const int* sample_inds,
                       const int* channel_inds, long long P, int n_channels, float* probs, long long* g_peak,
                       long long* g_class, int* g_count, float* o_xy, float* o_val, float* o_prob, int* status,
                       void* stream);
/* snb_classify_peaks on the PADDED peak table K1 writes (frame b owns slots [b*cap, b*cap + min(count, cap))), so
 * the chain find_local_peaks -> classify stays on the device without a pack step.  xy_div: the peaks are divided by
 * it before the class-map lookup and in o_xy (`peaks / class_maps_output_stride`, layers/bottomup_multiclass.py:86-87);
 * probs is a scratch of n_samples*cap*K floats. */
int snb_classify_peaks_padded(const float* class_maps, int n_samples, int K, int H, int W, long long ms, long long mk,
                              long long mh, long long mw, const int* frame_count, int cap, const float* peak_xy,
                              const float* peak_val, const int* peak_chan, float xy_div, int n_channels, float* probs,
                              float* o_xy, float* o_val, float* o_prob, int* status, void* stream);
/* BottomUpMultiClassLayer.postprocess after the classification (layers/bottomup_multiclass.py:99-146): xy (B,K,N,2) *
 * class_stride / input_scale / eff_scale[b]; o_scores = nanmean of val over nodes; o_tracking = nanmean of prob;
 * max_instances >= 0 applies _cap_instances_by_score (:148-190; np.argsort(scores)[::-1] order, NaN first). */
int snb_multiclass_outputs(const float* xy, const float* val, const float* prob, int B, int K, int N, float class_stride,
                           float input_scale, const float* eff_scale, int max_instances, float* o_kpts, float* o_vals,
                           float* o_scores, float* o_tracking, void* stream);
/* Per-group matches -> the concatenated (peak_inds, class_inds) of group_class_peaks in (sample, channel) order;
 * o_peak / o_class hold at most min(P, n_groups*K) entries, total[0] = how many were written. */
int snb_pack_class_matches(const long long* g_peak, const long long* g_class, const int* g_count, int n_groups, int K,
                           long long* o_peak, long long* o_class, int* total, void* stream);

/* get_class_inds_from_vectors (ops/identity.py:152-173): ONE assignment over probs (n, K); rows without a class
 * get -1 / NaN.  workspace: snb_class_inds_workspace_bytes(n, K) bytes. */
long long snb_class_inds_workspace_bytes(int n, int K);
int snb_class_inds_from_vectors(const float* probs, int n, int K, void* workspace, long long* o_inds, float* o_probs,
                                int* status, void* stream);

/* The per-frame re-assignment of TopDownLayer._run_stage_2 (layers/topdown.py:343-371): probs (n, K) are the class
 * vectors of the flattened crops, frame b owning crops [frame_off[b], frame_off[b+1]); rows_of_crop[r] = b*I + i.
 * One get_class_inds_from_vectors per frame, scattered into full_class_inds (B, I, n_nodes) int64 (-1 fill, the class
 * broadcast over the node axis), full_tracking (B, I) (NaN fill) and, when not NULL, full_vectors (B, I, K) (NaN fill).
 * workspace: snb_class_inds_grouped_workspace_bytes(B, I, K) bytes. */
long long snb_class_inds_grouped_workspace_bytes(int B, int I, int K);
int snb_class_inds_grouped(const float* probs, int K, const int* frame_off, const int* rows_of_crop, int B, int I,
                           int n_nodes, void* workspace, long long* full_class_inds, float* full_tracking,
                           float* full_vectors, int* status, void* stream);

/* make_class_vectors (data/identity.py:10-32): class_inds (n) fp32 (is_float) or int32 -> out (n, K) int32 one-hot,
 * index < 0 -> zero row; index >= K sets SNB_STATUS_BAD_INDEX (F.one_hot raises). */
int snb_class_vectors(const void* class_inds, int is_float, int n, int K, int* out, int* status, void* stream);

/* make_class_maps (data/identity.py:35-82) for G frames in one launch (the reference API is per frame: G = 1).
 *   confmaps (G, I, h, w) contiguous fp32 per-instance maps, class_inds (G, I) int32 (-1 = no class), n_valid (G) or
 *   NULL = per-frame instance count (the datasets' num_instances slice); threshold already fp32; out (G, K, h, w).
 *   The reference reshapes (not transposes) the (I, K) one-hot matrix to (K, I); that indexing is reproduced. */
int snb_class_maps(const float* confmaps, const int* class_inds, const int* n_valid, int G, int I, int K, int h, int w,
                   float threshold, float* out, void* stream);

/* ------------------------------------------------------ post-inference filters (inference/filters.py)
 *
 * snb_filter_instances: FilterPipeline.apply (inference/filters.py:100-163) on the padded outputs, one warp per
 *   frame, in the reference's fixed order: min_peak_value (:165-176) -> node count (:178-197) -> score filters
 *   (:199-243) -> greedy overlap NMS by bbox IoU or OKS (:245-344) -> centroid-distance NMS (:375-412); a dropped
 *   instance slot is NaN-filled in every field present (_nan_out_where, :346-373).  Thresholds that the reference
 *   compares against fp32 tensors are passed as fp32; the ones it compares as python floats (`.item()`) as double.
 *   A value <= 0 (overlapping == 0) disables its stage.  overlapping: 1 = "iou", 2 = "oks" (the caller applies the
 *   single-node OKS -> IoU fallback of :134-147).  oks_kappa_sq = fl32(0.1**2).
 *   kpts (B,I,N,2), vals (B,I,N), scores (B,I), centroids (B,I,2), centroid_vals (B,I): any may be NULL (the matching
 *   Outputs field is None), each o_* must be NULL exactly when its input is.  Inputs are not modified. */
typedef struct snb_filter_config {
  float min_peak_value;
  float min_visible_node_fraction;
  float min_instance_score;
  float min_mean_node_score;
  float oks_kappa_sq;
  int min_visible_nodes;
  int overlapping;
  double overlapping_threshold;
  double min_centroid_distance_sq;
} snb_filter_config;
/* FilterPipeline._bbox_iou / _oks (inference/filters.py:290-338) for one pair of keypoint sets a, b (N, 2), fp32
 * (is_f64 = 0) or float64: out[0] = IoU of the two NaN-aware boxes, out[1] = OKS with the scale from a's box area. */
int snb_pair_similarity(const void* a, const void* b, int N, int is_f64, double kappa, double* out, void* stream);
int snb_filter_instances(const snb_filter_config* cfg, int B, int I, int N, const float* kpts, const float* vals,
                         const float* scores, const float* centroids, const float* centroid_vals, float* o_kpts,
                         float* o_vals, float* o_scores, float* o_centroids, float* o_centroid_vals, void* stream);

/* The numeric cores of the Labels-level filters (inference/ops/filters.py), float64 like the numpy arrays they read,
 * for all frames of a Labels object in one launch.
 * snb_nms_greedy_f64: _nms_greedy_iou (:330-366) / _nms_greedy_oks (:369-404) with _compute_iou_one_to_many (:407-436),
 *   _instance_bbox (:300-316) and _compute_oks (:439-495).  pts (total, N, 2), scores (total); frame f owns instances
 *   [frame_start[f], frame_start[f+1]) (max_per_frame = the largest frame, sizes shared memory); method 0 = iou, 1 = oks.
 *   keep[frame_start[f] + k], k < keep_count[f]: kept LOCAL indices in keep order (np.argsort(scores)[::-1] visiting
 *   order); remaining slots -1.
 * snb_instance_stats_f64: _count_visible_nodes (:178-190) and _mean_node_score (:193-226; point_scores / mean_score
 *   may be NULL). */
int snb_nms_greedy_f64(const double* pts, const double* scores, const int* frame_start, int n_frames, int max_per_frame,
                       int N, int method, double threshold, double kappa, int* keep, int* keep_count, void* stream);
int snb_instance_stats_f64(const double* pts, const double* point_scores, long long total, int N, int* n_visible,
                           double* mean_score, void* stream);

/* ------------------------------------------------------------ fused bottom-up post-processing
 *
 * snb_bottomup_postproc enqueues K1 -> K4 -> K5 -> K6 on `stream` with padded tables: no host
 * sync, no allocation.  Replaces the sequence find_local_peaks -> "peaks * stride" -> per-sample
 * split -> PAFScorer.predict in BottomUpLayer (inference/layers/bottomup.py:95-236) and
 * group_scored_batch (inference/streaming.py:147-255).  n_nodes == C.  All pointers are caller
 * allocated device buffers:
 *   frame_count B | keys B*peak_cap | peak_xy B*peak_cap*2 | peak_val, peak_chan, node_peaks B*peak_cap
 *   node_start B*(C+1) | edge_off, match_off B*(n_edges+1)
 *   cand_edge, cand_score B*cand_cap | cand_epi B*cand_cap*2 (int64)
 *   m_edge, m_src, m_dst, m_score B*match_cap | m_count B
 *   lsap_ws B*n_edges*snb_lsap_workspace_bytes(lsap_max_dim) bytes (may be NULL when lsap_max_dim <= 32)
 *   asm_ws B*4*peak_cap int32 | inst_xy B*inst_cap*C*2 | inst_val B*inst_cap*C | inst_score B*inst_cap
 *   n_inst B | status 1 (zeroed by the caller). */
typedef struct snb_bottomup_args {
  const void* cms; /* (B, C, H, W), element type cms_dtype (ABI v5; fp32 when 0) */
  int B, C, H, W;
  long long cms_sb, cms_sc, cms_sh, cms_sw;
  const void* pafs; /* (B, paf_H, paf_W, 2*n_edges) view, element strides below, element type pafs_dtype */
  int paf_H, paf_W;
  long long paf_sb, paf_sy, paf_sx, paf_sc;
  const int* edges; /* (n_edges, 2) node ids */
  int n_edges;
  const int* sorted_edges; /* toposort_edges() order */
  int n_sorted;
  const float* t_table; /* torch.linspace(0, 1, n_points), host-computed */
  int n_points;
  float peak_threshold;
  int refine_size;   /* 0 = rough peaks, else integral patch size */
  float cms_stride;  /* peaks are multiplied by this (layers/bottomup.py:111) */
  float pafs_stride;
  float max_edge_length; /* ratio * max(paf_H, paf_W, 2*n_edges) * pafs_stride (paf.py:457-461) */
  float dist_penalty_weight;
  int min_instance_peaks;
  float min_line_scores;
  int peak_cap, cand_cap, match_cap, inst_cap, lsap_max_dim;
  int* frame_count;
  uint32_t* keys;
  float* peak_xy;
  float* peak_val;
  int* peak_chan;
  int* node_start;
  int* node_peaks;
  int* edge_off;
  int* match_off;
  int* cand_edge;
  long long* cand_epi;
  float* cand_score;
  int* m_edge;
  int* m_src;
  int* m_dst;
  float* m_score;
  int* m_count;
  void* lsap_ws;
  int* asm_ws;
  float* inst_xy;
  float* inst_val;
  float* inst_score;
  int* n_inst;
  int* status;
  void* ev_detect_begin; /* optional cudaEvent_t around the streaming detect kernel */
  void* ev_detect_end;
  /* Optional two-stream mode: detect runs on `stream`, the per-frame tail on tail_stream (ideally a
   * high-priority stream) after ev_handoff; ev_tail_done is recorded when the tail is enqueued and
   * waited on by the next call that reuses these buffers.  All three NULL = single stream. */
  void* tail_stream;
  void* ev_handoff;
  void* ev_tail_done;
  int flags; /* SNB_FLAG_* */
  /* ---- ABI v2: the BottomUpLayer / group_scored_batch epilogue (layers/bottomup.py:95-236,
   * inference/streaming.py:147-255), all optional.
   * max_peaks_per_node > 0 with skip_flag != NULL: the batch-wide guard of layers/bottomup.py:128-148 - if any
   *   node of any frame has more peaks than the limit, *skip_flag is set and the epilogue emits all-NaN outputs.
   * out_kpts != NULL: snb_bottomup_outputs is enqueued after the tail with the five fields below. */
  int max_peaks_per_node;
  int* skip_flag;       /* 1 int, zeroed by snb_bottomup_postproc at the start of every call */
  int max_instances;    /* rows per frame of the out_* tensors (>= 1) */
  float input_scale;    /* PreprocInfo.input_scale */
  const float* eff_scale; /* (B,) PreprocInfo.eff_scale or NULL */
  float* out_kpts;      /* (B, max_instances, C, 2) */
  float* out_vals;      /* (B, max_instances, C) */
  float* out_scores;    /* (B, max_instances) */
  /* ---- ABI v5: SNB_DTYPE_* of the two input tensors (0 = fp32).  Half-precision heads are read in place. */
  int cms_dtype;
  int pafs_dtype;
  /* SNB_FLAG_SELF_RESET_COUNTERS (fused tail only): frame_count must be all zero before the FIRST call; every call's
   * tail copies frame b's true peak count to n_peaks[b] and zeroes frame_count[b] again, so the chain is exactly two
   * kernel launches with no memset node in between. */
  int* n_peaks;
} snb_bottomup_args;

#define SNB_FLAG_UNFUSED_TAIL 1 /* chain the stand-alone kernels instead of the fused per-frame tail */
#define SNB_FLAG_SELF_RESET_COUNTERS 2 /* see snb_bottomup_args.n_peaks */
#define SNB_FLAG_NO_TAIL_CLUSTER 4     /* batches of <= 16 frames: one CTA per frame instead of a 4-CTA cluster (A/B, tests) */

/* The tail (everything after the streaming detect kernel) runs as ONE CTA per frame with all tables
 * in shared memory when snb_bottomup_tail_smem_bytes(...) <= 200 KB; the intermediate tables
 * (node_start, node_peaks, edge_off, match_off, cand_*, m_*) are then optional outputs (NULL = skip).
 * Otherwise, or with SNB_FLAG_UNFUSED_TAIL, the stand-alone kernels are chained and those tables
 * are required. */
int snb_bottomup_postproc(const snb_bottomup_args* args, void* stream);
int snb_bottomup_args_size(void); /* sizeof(snb_bottomup_args): lets a binding check its mirror of the layout */
long long snb_bottomup_tail_smem_bytes(int peak_cap, int n_nodes, int n_edges, int cand_cap, int match_cap,
                                       int n_sorted, int n_points);
int snb_bottomup_launches_per_call(const snb_bottomup_args* args);

/* snb_bottomup_outputs: the tail of group_scored_batch (inference/streaming.py:196-243) on device.
 *   Per frame: when n_inst > max_instances keep the top max_instances by instance score, in the order of
 *   np.argsort(scores)[::-1] (descending; NaN first; equal scores: higher index first), otherwise keep
 *   assembly order; coordinates are divided by fl32(input_scale) and then by eff_scale[b] (two separately
 *   rounded fp32 divisions, as `p / info.input_scale` and `p / eff[i]` are; dividing by exactly 1.0f is the
 *   identity, so the reference's `!= 1.0` short-circuits need no branch); rows >= n_inst are NaN.
 *   skip_flag (may be NULL): non-zero -> every output is NaN (the skip_paf short-circuit, :296-318).
 *   out_kpts (B, max_instances, N, 2), out_vals (B, max_instances, N), out_scores (B, max_instances). */
int snb_bottomup_outputs(const int* n_inst, const float* inst_xy, const float* inst_val, const float* inst_score,
                         int B, int inst_cap, int n_nodes, int max_instances, float input_scale,
                         const float* eff_scale, const int* skip_flag, float* out_kpts, float* out_vals,
                         float* out_scores, void* stream);

/* snb_pack_instances: append one batch's padded instance tables (the outputs of snb_bottomup_postproc /
 * snb_assemble) to a packed per-rank result table at a DEVICE-side running offset, so a rank can run its
 * whole frame shard with no host synchronisation and gather once at the end (the
 * reference concatenates per-sample python lists, inference/streaming.py:187-255).
 *   cursor: 24 bytes {u64 rows packed, u64 frames packed, u32 ticket}, zeroed before the first call.
 *   o_xy (out_cap,N,2), o_val (out_cap,N), o_score (out_cap), o_frame (out_cap) = frame_base + b,
 *   o_count[frames packed + b] = instances of that frame (may be NULL).  Overflow of out_cap sets
 *   SNB_STATUS_INSTANCE_OVERFLOW and drops the frame's rows (the cursor still advances). */
int snb_pack_instances(const int* n_inst, int B, int inst_cap, int n_nodes, const float* inst_xy,
                       const float* inst_val, const float* inst_score, int frame_base, void* cursor,
                       long long out_cap, float* o_xy, float* o_val, float* o_score, int* o_frame, int* o_count,
                       int* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SLEAPNN_B200_H_ */
