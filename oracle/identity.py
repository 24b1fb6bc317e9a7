"""CPU oracle for multi-class identity grouping and class-map targets.  TEST INFRASTRUCTURE ONLY.

Restates sleap_nn/inference/ops/identity.py (`ops/identity.py:NN`) and sleap_nn/data/identity.py
(`data/identity.py:NN`) with explicit loops over (sample, channel) groups and pixels' instances.  The
assignment solver is `oracle.paf.lsap_jv` (the restatement of scipy.optimize.linear_sum_assignment that
tests/test_oracle_golden.py pins against scipy).  Never imported by the product path.  Pinned against golden
vectors generated from the unmodified reference (tests/golden/ref_f3_identity.npz).
"""

from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from oracle.paf import lsap_jv
from oracle.targets import grid_vectors, multi_confmaps


def class_probs_at_peaks(class_maps: torch.Tensor, peak_points: torch.Tensor, sample_inds: torch.Tensor) -> torch.Tensor:
    """(n_peaks, n_classes): the class-map column under each peak.  ops/identity.py:103-116.

    Subscripts are round-half-even of (y, x), converted to int32 and clamped to the map.
    """
    n_samples, n_classes, h, w = class_maps.shape
    out = torch.empty((peak_points.shape[0], n_classes), dtype=class_maps.dtype)

    def sub(v: float, hi: int) -> int:
        # torch.round (half to even) in fp32, then .to(int32): x86 turns NaN / inf / |r| >= 2^31 into INT_MIN,
        # which the clamp sends to 0
        r = np.rint(np.float32(v))
        if not np.isfinite(r) or abs(float(r)) >= 2147483648.0:
            return 0
        return int(min(max(int(r), 0), hi))

    for p in range(peak_points.shape[0]):
        ry, rx = sub(float(peak_points[p, 1]), h - 1), sub(float(peak_points[p, 0]), w - 1)
        out[p] = class_maps[int(sample_inds[p]), :, ry, rx]
    return out


def group_class_peaks(peak_class_probs: torch.Tensor, peak_sample_inds: torch.Tensor, peak_channel_inds: torch.Tensor,
                      n_samples: int, n_channels: int, solver=lsap_jv) -> Tuple[torch.Tensor, torch.Tensor]:
    """Optimal peak-to-class assignment per (sample, channel) group, kept only where the assigned class is
    also the peak's most probable class.  ops/identity.py:13-71.

    Groups are visited sample-major, channel-minor; inside a group peaks keep ascending index order; the
    assignment maximises the summed probability (cost = -prob in float64, as scipy converts it).
    """
    probs = peak_class_probs.detach().cpu().numpy()
    s_inds = peak_sample_inds.detach().cpu().numpy()
    c_inds = peak_channel_inds.detach().cpu().numpy()
    keep_peaks, keep_classes = [], []
    for s in range(n_samples):
        for c in range(n_channels):
            members = np.flatnonzero((s_inds == s) & (c_inds == c))
            if members.size == 0 or probs.shape[1] == 0:
                continue
            rows, cols = solver(-probs[members])
            for r, k in zip(rows, cols):
                p = int(members[r])
                row = probs[p]
                best = np.nan if np.isnan(row).any() else row.max()  # torch.max propagates NaN
                if row[k] == best:
                    keep_peaks.append(p)
                    keep_classes.append(int(k))
    return torch.tensor(keep_peaks, dtype=torch.int64), torch.tensor(keep_classes, dtype=torch.int64)


def classify_peaks_from_maps(class_maps, peak_points, peak_vals, peak_sample_inds, peak_channel_inds, n_channels: int,
                             solver=lsap_jv):
    """(points (S,K,C,2), point_vals (S,K,C), class_probs (S,K,C)), NaN where nothing was assigned.
    ops/identity.py:74-149."""
    n_samples, n_classes, _, _ = class_maps.shape
    s32 = peak_sample_inds.to(torch.int32)
    c32 = peak_channel_inds.to(torch.int32)
    probs = class_probs_at_peaks(class_maps, peak_points, s32)
    peak_inds, class_inds = group_class_peaks(probs, s32, c32, n_samples, n_channels, solver=solver)
    points = torch.full((n_samples, n_classes, n_channels, 2), float("nan"))
    vals = torch.full((n_samples, n_classes, n_channels), float("nan"))
    cprobs = torch.full((n_samples, n_classes, n_channels), float("nan"))
    for p, k in zip(peak_inds.tolist(), class_inds.tolist()):
        s, c = int(s32[p]), int(c32[p])
        points[s, k, c] = peak_points[p]
        vals[s, k, c] = peak_vals[p]
        cprobs[s, k, c] = probs[p, k]
    return points, vals, cprobs


def class_inds_from_vectors(peak_class_probs: torch.Tensor, solver=lsap_jv):
    """One assignment over the whole (n_samples, n_classes) matrix; unassigned rows get -1 / NaN.
    ops/identity.py:152-173."""
    n = peak_class_probs.shape[0]
    rows, cols = solver(-peak_class_probs.detach().cpu().numpy())
    inds = torch.full((n,), -1, dtype=torch.int64)
    vals = torch.full((n,), float("nan"))
    for r, k in zip(rows, cols):
        inds[int(r)] = int(k)
        vals[int(r)] = peak_class_probs[int(r), int(k)]
    return inds, vals


# ------------------------------------------------------------------ training targets
def class_vectors(class_inds: torch.Tensor, n_classes: int) -> torch.Tensor:
    """(n_instances, n_classes) int32 one-hot rows; index < 0 gives an all-zero row.  data/identity.py:10-32."""
    out = torch.zeros((class_inds.shape[0], n_classes), dtype=torch.int32)
    for i, k in enumerate(class_inds.tolist()):
        if k >= 0:
            out[i, int(k)] = 1
    return out


def class_maps(confmaps: torch.Tensor, class_inds: torch.Tensor, n_classes: int, threshold: float = 0.2) -> torch.Tensor:
    """(1, n_classes, H, W) from per-instance confidence maps (1, I, H, W).  data/identity.py:35-82.

    Quirk kept from the reference: the (I, n_classes) one-hot matrix is RESHAPED (not transposed) to
    (n_classes, I, 1, 1), so class c / instance i reads flat element c*I + i of the row-major one-hot matrix.
    """
    n_inst = confmaps.shape[-3]
    weights = class_vectors(class_inds, n_classes).to(torch.float32).reshape(-1)  # flat (I * n_classes)
    cm = confmaps.reshape(n_inst, *confmaps.shape[-2:])
    total = cm[0].clone()
    for i in range(1, n_inst):  # torch.sum over a strided dim accumulates in order for the few instances used here
        total = total + cm[i]
    out = torch.empty((n_classes,) + tuple(cm.shape[-2:]), dtype=torch.float32)
    thr = torch.tensor(threshold, dtype=torch.float32)
    for c in range(n_classes):
        acc = None
        for i in range(n_inst):
            share = torch.where(cm[i] > thr, cm[i] / total, torch.zeros(()))
            term = share * weights[c * n_inst + i]
            if acc is None:
                acc = term
            else:
                both = torch.maximum(acc, term)  # torch.maximum propagates NaN like torch.max(dim)
                acc = both
        out[c] = acc
    return out.unsqueeze(0)


def generate_class_maps(instances, img_hw, num_instances, class_inds, num_tracks, class_map_threshold=0.2, sigma=1.5,
                        output_stride=2, is_centroids=False):
    """data/identity.py:85-137: per-INSTANCE confidence maps (max over that instance's nodes) -> class maps."""
    height, width = img_hw
    xv, yv = grid_vectors(height, width, output_stride)
    if is_centroids:
        points = instances[:, :num_instances, :].unsqueeze(dim=-3)
    else:
        points = instances[:, :num_instances, :, :].permute(0, 2, 1, 3)
    cms = multi_confmaps(points, xv, yv, sigma * output_stride)
    return class_maps(cms, class_inds, num_tracks, class_map_threshold)


# ------------------------------------------------------------------ layer glue (layers/bottomup_multiclass.py)
def cap_instances_by_score(instances, peak_scores, instance_scores, tracking_scores, max_instances: int):
    """NaN-out all but the first `max_instances` entries of np.argsort(scores)[::-1] in frames with more present
    classes than that.  layers/bottomup_multiclass.py:148-190 (NaN scores sort FIRST in that order)."""
    inst, pv, sc, tr = (t.clone() for t in (instances, peak_scores, instance_scores, tracking_scores))
    for b in range(sc.shape[0]):
        row = sc[b].numpy()
        present = int((~np.isnan(row)).sum())
        if present <= max_instances:
            continue
        # ascending with NaN last and equal keys in index order, then reversed
        asc = sorted(range(len(row)), key=lambda i: (1 if np.isnan(row[i]) else 0, row[i] if not np.isnan(row[i]) else 0.0, i))
        keep = set(asc[::-1][:max_instances])
        for k in range(len(row)):
            if k not in keep:
                inst[b, k] = float("nan"); pv[b, k] = float("nan"); sc[b, k] = float("nan"); tr[b, k] = float("nan")
    return inst, pv, sc, tr


def bottomup_multiclass_postprocess(cms, class_maps_t, cms_stride, class_stride, input_scale, eff_scale, max_instances=None,
                                    threshold=0.2, refinement="integral", patch=5):
    """BottomUpMultiClassLayer.postprocess, layers/bottomup_multiclass.py:75-146."""
    from oracle import peaks as opeaks

    pts, vals, si, ci = opeaks.local_peaks(cms, threshold, refinement, patch)
    pts = pts * torch.tensor(float(cms_stride))
    for_map = pts / torch.tensor(float(class_stride))
    inst, pv, cp = classify_peaks_from_maps(class_maps_t, for_map, vals, si, ci, cms.shape[1])
    inst = inst * torch.tensor(float(class_stride))
    if input_scale != 1.0:
        inst = inst / torch.tensor(float(input_scale), dtype=torch.float32)
    if not bool((eff_scale == 1.0).all()):
        inst = inst / eff_scale.view(-1, 1, 1, 1)

    def nanmean_rows(t):
        out = torch.empty(t.shape[:-1])
        for idx in np.ndindex(*t.shape[:-1]):
            row = t[idx]
            keep = row[~torch.isnan(row)]
            acc = torch.tensor(0.0)
            for v in keep:
                acc = acc + v
            out[idx] = acc / len(keep) if len(keep) else float("nan")
        return out

    sc, tr = nanmean_rows(pv), nanmean_rows(cp)
    if max_instances is not None:
        inst, pv, sc, tr = cap_instances_by_score(inst, pv, sc, tr, max_instances)
    return inst, pv, sc, tr
