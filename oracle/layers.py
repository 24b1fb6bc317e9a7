"""CPU oracle for the single-stage inference layers' post-model arithmetic.  TEST INFRASTRUCTURE ONLY.

Restates `CentroidLayer.postprocess` (sleap_nn/inference/layers/centroid.py:194-258), `CenteredInstanceLayer.postprocess`
(layers/centered_instance.py:199-230) and `SingleInstanceLayer.postprocess` (layers/single_instance.py:71-106) on top of
oracle.peaks, with explicit loops over frames.  Pinned against the unmodified layer classes by
tests/test_oracle_fuzz_vs_reference.py::test_single_stage_layers_fuzz.  Never imported by the product path.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from oracle import peaks as opeaks


def _ladder(xy: torch.Tensor, stride: int, input_scale: float) -> torch.Tensor:
    """undo_stride then undo_input_scale (ops/coord.py:27-55): separate fp32 ops, identities skipped."""
    if stride != 1:
        xy = xy * stride
    if input_scale != 1.0:
        xy = xy / input_scale
    return xy


def centroid_postprocess(cms: torch.Tensor, stride: int, input_scale: float, eff_scale: torch.Tensor,
                         max_instances: Optional[int], threshold: float = 0.2, refinement: Optional[str] = "integral",
                         patch: int = 5) -> Tuple[np.ndarray, np.ndarray]:
    """(B, max_instances, 2) centroids and (B, max_instances) values, NaN padded; per frame the top `max_instances` by
    value (torch.topk order) when there are more peaks than that, else (y, x) order.  centroid.py:194-258."""
    pts, vals, si, _ = opeaks.local_peaks(cms, threshold, refinement, patch)
    pts = _ladder(pts, stride, input_scale)
    B = cms.shape[0]
    counts = np.bincount(si.numpy(), minlength=B) if si.numel() else np.zeros(B, int)
    mi = max_instances or (int(counts.max()) if si.numel() else 0)   # _infer_max_instances
    mi = max(mi, 1)
    out = np.full((B, mi, 2), np.nan, np.float32)
    outv = np.full((B, mi), np.nan, np.float32)
    for b in range(B):
        sel = np.flatnonzero(si.numpy() == b)
        if sel.size == 0:
            continue
        p, v = pts[sel].numpy(), vals[sel].numpy()
        if sel.size > mi:
            order = sorted(range(sel.size), key=lambda i: (-v[i], i))[:mi]   # topk: descending, lower index first on ties
            p, v = p[order], v[order]
        out[b, : len(v)] = p
        outv[b, : len(v)] = v
    eff = eff_scale.numpy().astype(np.float32)
    if not (eff == 1.0).all():
        out = (out / eff[:, None, None]).astype(np.float32)            # undo_eff_scale comes last (centroid.py:248)
    return out, outv


def global_postprocess(cms: torch.Tensor, stride: int, input_scale: float, eff_scale: torch.Tensor, threshold: float = 0.2,
                       refinement: Optional[str] = "integral", patch: int = 5) -> Tuple[np.ndarray, np.ndarray]:
    """find_global_peaks + the ladder: (B, 1, N, 2), (B, 1, N).  centered_instance.py:199-230, single_instance.py:71-106."""
    pk, pv = opeaks.global_peaks(cms, threshold, refinement, patch)
    pk = _ladder(pk, stride, input_scale)
    if not bool((eff_scale == 1.0).all()):
        pk = pk / eff_scale.view(-1, 1, 1)
    return pk.unsqueeze(1).numpy(), pv.unsqueeze(1).numpy()
