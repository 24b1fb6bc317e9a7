"""Labels-level post-processing filters - same API as sleap_nn/inference/ops/filters.py; the numeric work (visible
node counts, mean node scores, greedy IoU / OKS non-maximum suppression) runs in CUDA kernels for ALL frames of a
`Labels` object in one launch (csrc/filters.cu), in float64 like the numpy arrays the reference reads.

`sleap_io` is not required: `labels` may be any object with `.labeled_frames`, each frame with a mutable
`.instances` list, each predicted instance with `.numpy()` -> (n_nodes, 2), `.score`, `.skeleton.nodes` and
(optionally) `.points["score"]`.  When `sleap_io` is importable, "predicted" means `isinstance(inst,
sio.PredictedInstance)` exactly as in the reference; otherwise the class name is used.

The tensor-level pipeline the inference layers use lives in `sleap_nn_b200.inference.filters`.
"""

from __future__ import annotations

from typing import List, Literal, Optional, Sequence, Tuple

import numpy as np
import torch

from sleap_nn_b200 import _native as N

try:  # pragma: no cover - sleap_io is absent from the build image
    import sleap_io as _sio
except Exception:  # noqa: BLE001
    _sio = None


def _is_predicted(inst) -> bool:
    if _sio is not None:
        return isinstance(inst, _sio.PredictedInstance)
    return any(c.__name__ == "PredictedInstance" for c in type(inst).__mro__)


def _device() -> torch.device:
    return N.compute_device()


def _stack_points(points_list: Sequence[np.ndarray]) -> Tuple[np.ndarray, int]:
    """(n, N, 2) float64, instances with fewer nodes padded with NaN rows (invisible either way)."""
    n_nodes = max((int(np.asarray(p).shape[0]) for p in points_list), default=0)
    out = np.full((len(points_list), n_nodes, 2), np.nan, dtype=np.float64)
    for i, p in enumerate(points_list):
        p = np.asarray(p, dtype=np.float64).reshape(-1, 2)
        out[i, : p.shape[0]] = p
    return out, n_nodes


def _nms_frames(points_list: Sequence[np.ndarray], scores: np.ndarray, frame_sizes: Sequence[int], threshold: float,
                method: int) -> List[List[int]]:
    """Greedy NMS of every frame in one launch; returns the kept local indices per frame, in keep order."""
    sizes = [int(s) for s in frame_sizes]
    total = sum(sizes)
    if total == 0:
        return [[] for _ in sizes]
    dev = _device()
    pts, n_nodes = _stack_points(points_list)
    start = np.zeros(len(sizes) + 1, dtype=np.int32)
    np.cumsum(sizes, out=start[1:])
    with torch.cuda.device(dev):
        d_pts = torch.from_numpy(pts).to(dev)
        d_sc = torch.from_numpy(np.asarray(scores, dtype=np.float64)).to(dev)
        d_start = torch.from_numpy(start).to(dev)
        keep = torch.empty((total,), dtype=torch.int32, device=dev)
        count = torch.empty((len(sizes),), dtype=torch.int32, device=dev)
        N.check(N.lib.snb_nms_greedy_f64(N.ptr(d_pts), N.ptr(d_sc), N.ptr(d_start), len(sizes), max(sizes), n_nodes, method,
                                         float(threshold), 0.1, N.ptr(keep), N.ptr(count), N.stream_ptr(dev)),
                "snb_nms_greedy_f64")
        keep_h, count_h = keep.cpu().numpy(), count.cpu().numpy()
    return [keep_h[start[f] : start[f] + count_h[f]].tolist() for f in range(len(sizes))]


def _nms_greedy_iou(bboxes: np.ndarray, scores: np.ndarray, threshold: float) -> List[int]:
    """Greedy NMS on (N, 4) [xmin, ymin, xmax, ymax] boxes; indices to keep, by decreasing score (ops/filters.py:330-366)."""
    bboxes = np.asarray(bboxes, dtype=np.float64).reshape(-1, 4)
    if len(bboxes) == 0:
        return []
    # a box is the bbox of its two corners: reuse the point kernel with two "keypoints" per instance
    corners = [np.array([[b[0], b[1]], [b[2], b[3]]]) for b in bboxes]
    return _nms_frames(corners, scores, [len(bboxes)], threshold, 0)[0]


def _nms_greedy_oks(points_list: List[np.ndarray], scores: np.ndarray, threshold: float) -> List[int]:
    """Greedy NMS by Object Keypoint Similarity; indices to keep, by decreasing score (ops/filters.py:369-404)."""
    if len(points_list) == 0:
        return []
    return _nms_frames(points_list, scores, [len(points_list)], threshold, 1)[0]


def _instance_score(instance) -> float:
    return getattr(instance, "score", 1.0)


def _instance_stats(instances: Sequence, want_scores: bool):
    """(n_visible (n,) int, mean_score (n,) float or None per instance) for a flat list of predicted instances."""
    if not instances:
        return np.zeros(0, np.int32), []
    dev = _device()
    pts, n_nodes = _stack_points([inst.numpy() for inst in instances])
    scores, has = None, [False] * len(instances)
    if want_scores:
        scores = np.full((len(instances), n_nodes), np.nan, dtype=np.float64)
        for i, inst in enumerate(instances):
            try:
                ps = inst.points["score"]
            except (KeyError, TypeError, IndexError, AttributeError):
                continue
            if ps is None or len(ps) == 0:
                continue
            ps = np.asarray(ps, dtype=np.float64).reshape(-1)
            scores[i, : ps.shape[0]] = ps
            has[i] = True
    with torch.cuda.device(dev):
        d_pts = torch.from_numpy(pts).to(dev)
        d_sc = torch.from_numpy(scores).to(dev) if scores is not None else None
        nv = torch.empty((len(instances),), dtype=torch.int32, device=dev)
        mean = torch.empty((len(instances),), dtype=torch.float64, device=dev) if scores is not None else None
        N.check(N.lib.snb_instance_stats_f64(N.ptr(d_pts), N.ptr(d_sc), len(instances), n_nodes, N.ptr(nv), N.ptr(mean),
                                             N.stream_ptr(dev)), "snb_instance_stats_f64")
        nv_h = nv.cpu().numpy()
        mean_h = mean.cpu().numpy() if mean is not None else None
    means = [float(mean_h[i]) if (mean_h is not None and has[i]) else None for i in range(len(instances))]
    return nv_h, means


def _predicted_of(labels):
    """[(frame, [predicted instances])] and the flat list of all predicted instances."""
    per_frame, flat = [], []
    for lf in labels.labeled_frames:
        pred = [inst for inst in lf.instances if _is_predicted(inst)]
        per_frame.append((lf, pred))
        flat.extend(pred)
    return per_frame, flat


def filter_by_node_count(labels, min_visible_nodes: int = 0, min_visible_node_fraction: float = 0.0):
    """Remove predicted instances with too few visible (non-NaN) keypoints, in place (ops/filters.py:13-88)."""
    if min_visible_nodes <= 0 and min_visible_node_fraction <= 0.0:
        return labels
    _, flat = _predicted_of(labels)
    nv, _ = _instance_stats(flat, want_scores=False)
    n_visible = {id(inst): int(v) for inst, v in zip(flat, nv)}
    for lf in labels.labeled_frames:
        if len(lf.instances) == 0:
            continue
        kept = []
        for inst in lf.instances:
            if not _is_predicted(inst):
                kept.append(inst)
                continue
            n_vis, n_total = n_visible[id(inst)], len(inst.skeleton.nodes)
            if min_visible_nodes > 0 and n_vis < min_visible_nodes:
                continue
            if min_visible_node_fraction > 0.0:
                fraction = n_vis / n_total if n_total > 0 else 0.0
                if fraction < min_visible_node_fraction:
                    continue
            kept.append(inst)
        lf.instances = kept
    return labels


def filter_by_node_confidence(labels, min_mean_node_score: float = 0.0, min_instance_score: float = 0.0):
    """Remove predicted instances by instance score and mean visible-node score, in place (ops/filters.py:91-175)."""
    if min_mean_node_score <= 0.0 and min_instance_score <= 0.0:
        return labels
    _, flat = _predicted_of(labels)
    means = {}
    if min_mean_node_score > 0.0:
        _, m = _instance_stats(flat, want_scores=True)
        means = {id(inst): v for inst, v in zip(flat, m)}
    for lf in labels.labeled_frames:
        if len(lf.instances) == 0:
            continue
        kept = []
        for inst in lf.instances:
            if not _is_predicted(inst):
                kept.append(inst)
                continue
            if min_instance_score > 0.0 and _instance_score(inst) < min_instance_score:
                continue
            if min_mean_node_score > 0.0:
                mean_score = means.get(id(inst))
                if mean_score is not None and mean_score < min_mean_node_score:
                    continue
            kept.append(inst)
        lf.instances = kept
    return labels


def filter_overlapping_instances(labels, threshold: float = 0.8, method: Literal["iou", "oks"] = "iou"):
    """Greedy NMS between the predicted instances of every frame, in place (ops/filters.py:229-297).

    "iou": bounding boxes of the non-NaN keypoints; "oks": keypoint similarity with the kept instance's bbox area as
    scale.  Kept predicted instances come first (by decreasing score), other instances after them.
    """
    if method not in ("iou", "oks"):
        raise ValueError(f"Unknown method: {method}. Use 'iou' or 'oks'.")
    work = []  # (frame, predicted, other)
    for lf in labels.labeled_frames:
        if len(lf.instances) <= 1:
            continue
        pred = [inst for inst in lf.instances if _is_predicted(inst)]
        other = [inst for inst in lf.instances if not _is_predicted(inst)]
        if len(pred) <= 1:
            continue
        work.append((lf, pred, other))
    if not work:
        return labels
    points = [inst.numpy() for _, pred, _ in work for inst in pred]
    scores = np.array([_instance_score(inst) for _, pred, _ in work for inst in pred], dtype=np.float64)
    keeps = _nms_frames(points, scores, [len(pred) for _, pred, _ in work], threshold, 1 if method == "oks" else 0)
    for (lf, pred, other), keep in zip(work, keeps):
        lf.instances = [pred[i] for i in keep] + other
    return labels
