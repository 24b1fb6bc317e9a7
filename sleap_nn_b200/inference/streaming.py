"""The GPU-stage -> grouping-stage seam of bottom-up inference - same names as sleap_nn/inference/streaming.py.

The reference splits bottom-up post-processing at this point because its grouping (scipy assignment + Python
dict assembly) runs on the CPU, optionally in a spawn pool (`PafGroupingPool`, streaming.py:328-441).  Here both
halves run on the device, so the production path is the fused `sleap_nn_b200.pipeline.BottomUpPostproc`; this module
keeps the seam for callers that already hold a `ScoredBatch`: `group_scored_batch` runs the per-edge assignment
(`snb_match_generic`), the assembly (`snb_assemble`) and the scale-undo / NaN-pad / top-N epilogue
(`snb_bottomup_outputs`) on the device and returns CPU tensors, like the reference.  No worker pool is needed.
"""

from __future__ import annotations

from typing import Any, List, Optional

import attrs
import torch

from sleap_nn_b200 import _native as N


@attrs.frozen(eq=False)
class ScoredBatch:
    """Peaks + scored PAF candidates of one batch (streaming.py:42-112); fields as in the reference."""

    cms_peaks: List[torch.Tensor]
    cms_peak_vals: List[torch.Tensor]
    cms_peak_channel_inds: List[torch.Tensor]
    edge_inds: List[torch.Tensor]
    edge_peak_inds: List[torch.Tensor]
    line_scores: List[torch.Tensor]
    info: Any
    n_samples: int
    n_nodes: int
    skip_paf: bool = False
    cms: Optional[torch.Tensor] = None
    pafs: Optional[torch.Tensor] = None

    def to_cpu(self) -> "ScoredBatch":
        """Detach + move every tensor field (and `info.eff_scale`) to the CPU (streaming.py:88-112)."""
        cpu = lambda ts: [t.detach().cpu() for t in ts]
        info = self.info
        if attrs.has(type(info)):
            info = attrs.evolve(info, eff_scale=info.eff_scale.detach().cpu())
        return attrs.evolve(self, cms_peaks=cpu(self.cms_peaks), cms_peak_vals=cpu(self.cms_peak_vals),
                            cms_peak_channel_inds=cpu(self.cms_peak_channel_inds), edge_inds=cpu(self.edge_inds),
                            edge_peak_inds=cpu(self.edge_peak_inds), line_scores=cpu(self.line_scores), info=info,
                            cms=None if self.cms is None else self.cms.detach().cpu(),
                            pafs=None if self.pafs is None else self.pafs.detach().cpu())


@attrs.frozen(eq=False)
class GroupingParams:
    """Layer-level grouping knobs (streaming.py:115-139)."""

    paf_scorer_kwargs: dict
    max_instances: Optional[int] = None
    return_confmaps: bool = False
    return_pafs: bool = False
    return_paf_graph: bool = False


@attrs.define
class GroupedOutputs:
    """The `Outputs` fields `group_scored_batch` fills (used when the reference's `Outputs` class is not importable)."""

    pred_keypoints: Optional[torch.Tensor] = None
    pred_peak_values: Optional[torch.Tensor] = None
    instance_scores: Optional[torch.Tensor] = None
    preprocess_info: Any = None
    pred_confmaps: Optional[torch.Tensor] = None
    pred_pafs: Optional[torch.Tensor] = None
    pred_paf_graph: Optional[tuple] = None


def _outputs_class():
    try:  # a box that has the reference installed gets the reference's own value type back
        from sleap_nn.inference.outputs import Outputs

        return Outputs
    except Exception:
        return GroupedOutputs


def group_scored_batch(scored: ScoredBatch, params: GroupingParams):
    """ScoredBatch -> per-batch outputs with NaN-padded (B, I, N, 2) keypoints (streaming.py:147-255)."""
    from sleap_nn_b200.inference.ops.paf import PAFScorer, _group_frames

    B, n_nodes, info = int(scored.n_samples), int(scored.n_nodes), scored.info
    input_scale = float(getattr(info, "input_scale", 1.0))
    eff = getattr(info, "eff_scale", None)
    nan = float("nan")
    if scored.skip_paf:  # the max_peaks_per_node guard tripped upstream: all-NaN outputs (streaming.py:296-318)
        mi = params.max_instances or 1
        kpts, vals, scores = (torch.full(s, nan) for s in ((B, mi, n_nodes, 2), (B, mi, n_nodes), (B, mi)))
    else:
        scorer = PAFScorer(**params.paf_scorer_kwargs)
        m_e, m_s, m_d, m_sc = scorer.match_candidates(scored.edge_inds, scored.edge_peak_inds, scored.line_scores)
        inst_xy, inst_val, inst_score, n_inst, status = _group_frames(
            list(scored.cms_peaks), list(scored.cms_peak_vals), list(scored.cms_peak_channel_inds), m_e, m_s, m_d, m_sc,
            scorer.n_nodes, scorer.sorted_edge_inds, scorer.edge_types, scorer.min_instance_peaks,
            scorer.min_line_scores, tables=True)
        dev = inst_xy.device
        mi = params.max_instances or (int(n_inst.max().item()) if B else 0)  # _infer_max_instances
        mi = max(int(mi), 1)
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        kpts, vals, scores = f32(B, mi, n_nodes, 2), f32(B, mi, n_nodes), f32(B, mi)
        eff_d = None
        if eff is not None:
            eff_d = eff.detach().to(device=dev, dtype=torch.float32).reshape(-1)
            eff_d = (eff_d.expand(B) if eff_d.numel() == 1 and B != 1 else eff_d).contiguous()
        with torch.cuda.device(dev):
            N.check(N.lib.snb_bottomup_outputs(N.ptr(n_inst), N.ptr(inst_xy), N.ptr(inst_val), N.ptr(inst_score), B,
                                               int(inst_xy.shape[1]), n_nodes, mi, input_scale, N.ptr(eff_d), None,
                                               N.ptr(kpts), N.ptr(vals), N.ptr(scores), N.stream_ptr(dev)),
                    "snb_bottomup_outputs")
        from sleap_nn_b200.inference.ops.paf import _status_check

        _status_check(status, "group_scored_batch")
        kpts, vals, scores = kpts.cpu(), vals.cpu(), scores.cpu()  # the reference builds them with torch.full on the host
    out = dict(pred_keypoints=kpts, pred_peak_values=vals, instance_scores=scores, preprocess_info=info)
    if params.return_confmaps and scored.cms is not None:
        out["pred_confmaps"] = scored.cms
    if params.return_pafs and scored.pafs is not None:
        out["pred_pafs"] = scored.pafs.permute(0, 3, 1, 2).contiguous()
    if params.return_paf_graph:  # streaming.py:256-285: the per-batch graph, concatenated over samples
        cat = lambda ts, empty: torch.cat(list(ts), dim=0) if ts else empty
        out["pred_paf_graph"] = (cat(scored.cms_peaks, torch.empty(0, 2)), cat(scored.edge_inds, torch.empty(0, dtype=torch.int32)),
                                 cat(scored.edge_peak_inds, torch.empty(0, 2, dtype=torch.int32)),
                                 cat(scored.line_scores, torch.empty(0)))
    return _outputs_class()(**out)


__all__ = ["ScoredBatch", "GroupingParams", "GroupedOutputs", "group_scored_batch"]
