"""`sleap_nn.inference.paf_grouping`'s import path (inference/paf_grouping.py:8-46): the names callers import from
there, re-exported from `sleap_nn_b200.inference.ops.paf` (the C-ABI entry point behind each is noted)."""

from sleap_nn_b200.inference.ops.paf import (
    EdgeConnection,                   # value types of the dict API (host only)
    EdgeType,
    PAFScorer,                        # snb_paf_score_t + snb_match_generic + snb_assemble
    PeakID,
    assign_connections_to_instances,  # K6: snb_assemble
    compute_distance_penalty,         # snb_distance_penalty
    get_connection_candidates,        # K4: snb_paf_prepare
    get_paf_lines,                    # snb_paf_gather
    group_instances_batch,            # snb_assemble
    group_instances_sample,
    make_line_subs,                   # snb_line_subs
    make_predicted_instances,         # snb_scatter_instances
    match_candidates_batch,           # K5: snb_match_generic
    match_candidates_sample,
    score_paf_lines,                  # snb_score_lines
    score_paf_lines_batch,            # snb_paf_score
    toposort_edges,                   # once per scorer, on the host
)

__all__ = [
    "EdgeConnection",
    "EdgeType",
    "PAFScorer",
    "PeakID",
    "assign_connections_to_instances",
    "compute_distance_penalty",
    "get_connection_candidates",
    "get_paf_lines",
    "group_instances_batch",
    "group_instances_sample",
    "make_line_subs",
    "make_predicted_instances",
    "match_candidates_batch",
    "match_candidates_sample",
    "score_paf_lines",
    "score_paf_lines_batch",
    "toposort_edges",
]
