// Fused bottom-up post-processing: one C call enqueues the whole chain
//   K1 local peaks (+ integral refinement, x confmap stride)  ->  K4 candidates + PAF line scores
//   ->  K5 per-edge assignment  ->  K6 instance assembly
// on one stream with fixed-capacity tables, no host synchronisation and no allocation, so
// batches can be pipelined across streams (and captured into a CUDA graph).  It replaces the
// call sequence find_local_peaks -> peaks * stride -> per-sample split -> PAFScorer.predict of
// BottomUpLayer (layers/bottomup.py:95-236) + group_scored_batch (inference/streaming.py:147-255).
#include "common.cuh"

extern "C" int snb_local_peaks_ev(const float*, int, int, int, int, long long, long long, long long, long long, float,
                                  int, float, int, int*, uint32_t*, float*, float*, int*, int*, void*, void*, void*);

extern "C" int snb_bottomup_postproc(const snb_bottomup_args* a, void* stream) {
  if (!a) return SNB_ERR_BAD_ARG;
  const int n_nodes = a->C;
  int rc = snb_local_peaks_ev(a->cms, a->B, a->C, a->H, a->W, a->cms_sb, a->cms_sc, a->cms_sh, a->cms_sw,
                              a->peak_threshold, a->refine_size, a->cms_stride, a->peak_cap, a->frame_count, a->keys,
                              a->peak_xy, a->peak_val, a->peak_chan, a->status, a->ev_detect_begin, a->ev_detect_end,
                              stream);
  if (rc != SNB_OK) return rc;
  rc = snb_paf_prepare(a->peak_chan, nullptr, a->peak_cap, a->frame_count, a->B, a->edges, n_nodes, a->n_edges,
                       a->node_start, a->node_peaks, a->edge_off, a->match_off, stream);
  if (rc != SNB_OK) return rc;
  rc = snb_paf_score(a->pafs, a->paf_sb, a->paf_sy, a->paf_sx, a->paf_sc, a->paf_H, a->paf_W, a->t_table, a->n_points,
                     a->pafs_stride, a->max_edge_length, a->dist_penalty_weight, a->peak_xy, nullptr, a->peak_cap,
                     a->B, a->edges, n_nodes, a->n_edges, a->node_start, a->node_peaks, a->edge_off, nullptr,
                     a->cand_cap, a->cand_cap, a->cand_edge, a->cand_epi, a->cand_score, a->status, stream);
  if (rc != SNB_OK) return rc;
  rc = snb_match_structured(a->cand_score, nullptr, a->cand_cap, a->edges, n_nodes, a->n_edges, a->node_start,
                            a->edge_off, a->match_off, nullptr, a->match_cap, a->lsap_ws, a->lsap_max_dim, a->B,
                            a->m_edge, a->m_src, a->m_dst, a->m_score, a->m_count, a->status, stream);
  if (rc != SNB_OK) return rc;
  return snb_assemble(a->peak_xy, a->peak_val, a->peak_chan, nullptr, a->peak_cap, a->frame_count, a->B, a->node_start,
                      a->node_peaks, n_nodes, a->edges, a->sorted_edges, a->n_sorted, a->m_edge, a->m_src, a->m_dst,
                      a->m_score, nullptr, a->match_cap, a->m_count, a->min_instance_peaks, a->min_line_scores,
                      a->asm_ws, a->peak_cap, a->inst_cap, a->inst_xy, a->inst_val, a->inst_score, a->n_inst, a->status,
                      stream);
}

// Number of kernel launches snb_bottomup_postproc enqueues per call (detect, finalize, prepare,
// score, match, assemble); the frame_count memset is a memset node, not a kernel.
extern "C" int snb_bottomup_launches_per_call(void) { return 6; }
