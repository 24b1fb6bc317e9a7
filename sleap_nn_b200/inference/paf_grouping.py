"""`sleap_nn.inference.paf_grouping`'s import path (inference/paf_grouping.py:8-46): the names callers import from there,
bound to the CUDA-backed implementations of `sleap_nn_b200.inference.ops.paf`.

The table says which C-ABI entry point does the work behind each name (host-only helpers are marked as such).
"""

from sleap_nn_b200.inference.ops import paf as _impl

_BACKED_BY = {
    # value types of the dict API (host only)
    "PeakID": None, "EdgeType": None, "EdgeConnection": None,
    # K4: candidates, line subscripts, PAF taps, scores
    "get_connection_candidates": "snb_paf_prepare", "make_line_subs": "snb_line_subs", "get_paf_lines": "snb_paf_gather",
    "compute_distance_penalty": "snb_distance_penalty", "score_paf_lines": "snb_score_lines",
    "score_paf_lines_batch": "snb_paf_score",
    # K5: per-edge optimal assignment
    "match_candidates_sample": "snb_match_generic", "match_candidates_batch": "snb_match_generic",
    # K6: greedy assembly
    "assign_connections_to_instances": "snb_assemble", "make_predicted_instances": "snb_scatter_instances",
    "group_instances_sample": "snb_assemble", "group_instances_batch": "snb_assemble",
    "toposort_edges": None,  # once per scorer, on the host
    "PAFScorer": "snb_paf_score + snb_match_generic + snb_assemble",
}
globals().update({name: getattr(_impl, name) for name in _BACKED_BY})
__all__ = sorted(_BACKED_BY)
