"""CPU oracle for the post-inference filter pipeline.  TEST INFRASTRUCTURE ONLY.

Restates sleap_nn/inference/filters.py (`filters.py:NN`) with explicit per-frame / per-instance numpy loops in
the reference's arithmetic: fp32 for what the reference computes on fp32 tensors, python floats (double) after
its `.item()` calls.  Never imported by the product path.  Pinned against golden vectors produced by the
unmodified reference pipeline (tests/golden/ref_f4_filters.npz).
"""

from __future__ import annotations

from typing import Dict, Optional

import numpy as np

F = np.float32


def _rows_present(p: np.ndarray) -> np.ndarray:
    return ~np.isnan(p).any(axis=-1)


def _order_desc(scores: np.ndarray) -> list:
    """torch.argsort(descending=True): NaN first, then larger first, equal keys by ascending index."""
    idx = list(range(len(scores)))
    return sorted(idx, key=lambda i: (0 if np.isnan(scores[i]) else 1, -scores[i] if not np.isnan(scores[i]) else 0.0, i))


def bbox_iou(a: np.ndarray, b: np.ndarray) -> float:
    """filters.py:292-309."""
    a, b = a[_rows_present(a)], b[_rows_present(b)]
    if a.size == 0 or b.size == 0:
        return 0.0
    ax1, ay1, ax2, ay2 = a[:, 0].min(), a[:, 1].min(), a[:, 0].max(), a[:, 1].max()
    bx1, by1, bx2, by2 = b[:, 0].min(), b[:, 1].min(), b[:, 0].max(), b[:, 1].max()
    iw = max(F(min(ax2, bx2) - max(ax1, bx1)), F(0))
    ih = max(F(min(ay2, by2) - max(ay1, by1)), F(0))
    inter = float(F(iw * ih))
    area_a = float(F(F(ax2 - ax1) * F(ay2 - ay1)))
    area_b = float(F(F(bx2 - bx1) * F(by2 - by1)))
    union = area_a + area_b - inter
    return inter / union if union > 0 else 0.0


def oks(a: np.ndarray, b: np.ndarray, kappa: float = 0.1) -> float:
    """filters.py:311-344: scale is the bbox AREA of a's own valid keypoints."""
    va, vb = _rows_present(a), _rows_present(b)
    both = va & vb
    if both.sum() == 0:
        return 0.0
    own = a[va]
    if own.shape[0] < 2:
        return 0.0
    scale_sq = F(F(own[:, 0].max() - own[:, 0].min()) * F(own[:, 1].max() - own[:, 1].min()))
    if float(scale_sq) <= 0:
        return 0.0
    den = F(F(F(2) * scale_sq) * F(kappa**2))
    total, cnt = F(0), 0
    for n in np.flatnonzero(both):
        dx, dy = F(a[n, 0] - b[n, 0]), F(a[n, 1] - b[n, 1])
        d2 = F(F(dx * dx) + F(dy * dy))
        total = F(total + np.exp(F(-d2 / den), dtype=F))
        cnt += 1
    return float(F(total / F(cnt)))


def apply(cfg: Dict, kpts: Optional[np.ndarray] = None, vals: Optional[np.ndarray] = None,
          scores: Optional[np.ndarray] = None, cen: Optional[np.ndarray] = None, cenv: Optional[np.ndarray] = None):
    """FilterPipeline.apply (filters.py:100-163).  `cfg` holds FilterConfig's fields (missing = default).  Returns the
    five arrays (None where the input was None), float32 copies with dropped slots NaN-filled."""
    g = lambda k, d: cfg.get(k, d)
    f = {k: (None if v is None else np.array(v, dtype=F, copy=True)) for k, v in
         dict(kpts=kpts, vals=vals, scores=scores, cen=cen, cenv=cenv).items()}

    def nan_out(b, i):  # filters.py:346-373
        for v in f.values():
            if v is not None:
                v[b, i] = np.nan

    ref = f["kpts"] if f["kpts"] is not None else f["cen"]
    if ref is None:
        return tuple(f.values())
    B, I = ref.shape[:2]
    thr = F(g("min_peak_value", 0.0))
    if thr > 0 and f["kpts"] is not None and f["vals"] is not None:  # filters.py:165-176
        low = f["vals"] < thr
        f["kpts"][low] = np.nan
        f["vals"][low] = np.nan
    mv, mf = int(g("min_visible_nodes", 0)), F(g("min_visible_node_fraction", 0.0))
    if (mv > 0 or mf > 0) and f["kpts"] is not None:  # filters.py:178-197
        n_nodes = f["kpts"].shape[2]
        for b in range(B):
            for i in range(I):
                nv = int(_rows_present(f["kpts"][b, i]).sum())
                keep = True
                if mv > 0:
                    keep &= nv >= mv
                if mf > 0:
                    keep &= bool(F(F(nv) / F(max(n_nodes, 1))) >= mf)
                if not keep:
                    nan_out(b, i)
    mi, mm = F(g("min_instance_score", 0.0)), F(g("min_mean_node_score", 0.0))
    if mi > 0 or mm > 0:  # filters.py:199-243
        for b in range(B):
            for i in range(I):
                if f["kpts"] is None:
                    sc = f["scores"] if f["scores"] is not None else f["cenv"]
                    if f["cen"] is not None and mi > 0 and sc is not None and (sc[b, i] < mi or np.isnan(sc[b, i])):
                        nan_out(b, i)
                    continue
                keep = True
                if mi > 0 and f["scores"] is not None:
                    keep &= bool(f["scores"][b, i] >= mi)
                if mm > 0 and f["vals"] is not None:
                    row = f["vals"][b, i]
                    present = row[~np.isnan(row)]
                    total = F(0)
                    for v in present:
                        total = F(total + v)
                    mean = F(total / F(len(present))) if len(present) else F(0)
                    keep &= bool(mean >= mm)
                if not keep:
                    nan_out(b, i)
    if g("overlapping", False) and f["kpts"] is not None:  # filters.py:245-290
        method = g("overlapping_method", "iou")
        if method == "oks" and f["kpts"].shape[2] < 2:
            method = "iou"
        thr_o = float(g("overlapping_threshold", 0.8))
        drops = []
        for b in range(B):
            valid = [not np.isnan(f["kpts"][b, i]).all() for i in range(I)]
            if sum(valid) <= 1:
                continue
            sc = f["scores"][b] if f["scores"] is not None else np.zeros(I, F)
            kept = []
            for idx in _order_desc(sc):
                if not valid[idx]:
                    continue
                sim = (lambda k: oks(f["kpts"][b, idx], f["kpts"][b, k])) if method == "oks" else \
                      (lambda k: bbox_iou(f["kpts"][b, idx], f["kpts"][b, k]))
                if any(sim(k) > thr_o for k in kept):
                    drops.append((b, idx))
                else:
                    kept.append(idx)
        for b, i in drops:
            nan_out(b, i)
    md = float(g("min_centroid_distance", 0.0))
    if md > 0 and f["cen"] is not None:  # filters.py:375-412
        sc_all = f["cenv"] if f["cenv"] is not None else (f["scores"] if f["scores"] is not None else np.zeros((B, I), F))
        drops = []
        for b in range(B):
            valid = _rows_present(f["cen"][b])
            if valid.sum() <= 1:
                continue
            sc = np.where(np.isnan(sc_all[b]), -np.inf, sc_all[b]).astype(F)
            kept = []
            for idx in _order_desc(sc):
                if not valid[idx]:
                    continue
                p = f["cen"][b, idx]
                close_ = False
                for k in kept:
                    dx, dy = F(p[0] - f["cen"][b, k, 0]), F(p[1] - f["cen"][b, k, 1])
                    if float(F(F(dx * dx) + F(dy * dy))) < md**2:
                        close_ = True
                        break
                if close_:
                    drops.append((b, idx))
                else:
                    kept.append(idx)
        for b, i in drops:
            nan_out(b, i)
    return tuple(f.values())


# --------------------------------------------------------------------------------------------------------------
# The float64 numpy cores of the Labels-level filters, sleap_nn/inference/ops/filters.py (`ops/filters.py:NN`).
# --------------------------------------------------------------------------------------------------------------
def instance_bbox64(pts: np.ndarray) -> np.ndarray:
    """ops/filters.py:300-316: [xmin, ymin, xmax, ymax] of the rows without NaN, zeros when there is none."""
    rows = pts[_rows_present(pts)]
    if rows.shape[0] == 0:
        return np.zeros(4)
    return np.array([rows[:, 0].min(), rows[:, 1].min(), rows[:, 0].max(), rows[:, 1].max()])


def iou64(a: np.ndarray, b: np.ndarray) -> float:
    """ops/filters.py:407-436 for one pair of boxes."""
    iw = max(0.0, min(a[2], b[2]) - max(a[0], b[0]))
    ih = max(0.0, min(a[3], b[3]) - max(a[1], b[1]))
    inter = iw * ih
    union = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return inter / union if union > 0 else 0.0


def oks64(a: np.ndarray, b: np.ndarray, kappa: float = 0.1) -> float:
    """ops/filters.py:439-495: a's own bbox area is the scale; mean over the keypoints present in both."""
    va, vb = _rows_present(a), _rows_present(b)
    both = va & vb
    if not both.any() or va.sum() < 2:
        return 0.0
    own = a[va]
    scale_sq = (own[:, 0].max() - own[:, 0].min()) * (own[:, 1].max() - own[:, 1].min())
    if scale_sq <= 0:
        return 0.0
    total, cnt = 0.0, 0
    for n in np.flatnonzero(both):
        d2 = (a[n, 0] - b[n, 0]) ** 2 + (a[n, 1] - b[n, 1]) ** 2
        total += float(np.exp(-d2 / (2 * scale_sq * kappa**2)))
        cnt += 1
    return total / cnt


def nms_greedy64(points_list, scores, threshold: float, method: str):
    """ops/filters.py:330-404: visit in np.argsort(scores)[::-1] order (NaN first, ties by descending index), keep an
    instance unless an already kept one is more similar than `threshold` (similarity(kept, candidate))."""
    n = len(points_list)
    asc = sorted(range(n), key=lambda i: (1 if np.isnan(scores[i]) else 0, scores[i] if not np.isnan(scores[i]) else 0.0, i))
    boxes = [instance_bbox64(np.asarray(p, float)) for p in points_list]
    kept = []
    for idx in asc[::-1]:
        if method == "iou":
            hit = any(iou64(boxes[k], boxes[idx]) > threshold for k in kept)
        else:
            hit = any(oks64(np.asarray(points_list[k], float), np.asarray(points_list[idx], float)) > threshold for k in kept)
        if not hit:
            kept.append(idx)
    return kept


def count_visible64(pts: np.ndarray) -> int:
    """ops/filters.py:178-190."""
    return int(_rows_present(np.asarray(pts, float)).sum())


def mean_node_score64(pts: np.ndarray, point_scores) -> Optional[float]:
    """ops/filters.py:193-226: mean of the finite scores of the visible nodes; 0.0 when there is none."""
    if point_scores is None or len(point_scores) == 0:
        return None
    vis = _rows_present(np.asarray(pts, float))
    sc = np.asarray(point_scores, float)[vis]
    sc = sc[~np.isnan(sc)]
    if len(sc) == 0:
        return 0.0
    total = 0.0
    for v in sc:
        total += float(v)
    return total / len(sc)
