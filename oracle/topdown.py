"""CPU oracle for the top-down composition.  TEST INFRASTRUCTURE ONLY.

Restates the post-centroid half of sleap_nn/inference/layers/topdown.py (`topdown.py:NN`): the NaN-centroid mask,
the greedy centroid NMS, the crop list and the lift of the centred-instance peaks into (B, max_inst, ...) tensors,
plus the per-frame class assignment of the multi-class variant.  Written as explicit loops over frames and slots
with numpy float32 scalars (every operation a separate fp32 rounding, as the reference's tensor ops are).
Never imported by the product path.  Pinned against tests/golden/ref_f2_topdown.npz, which
tests/golden/make_golden.py produced by calling the unmodified reference methods.
"""

from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch

from oracle import peaks as opeaks
from oracle.identity import class_inds_from_vectors

f32 = np.float32


def centred_iou(c1, c2, h: int, w: int) -> np.float32:
    """IoU of two h x w boxes centred on c1, c2 (x, y).  topdown.py:448-460."""
    hh, hw = f32(h / 2.0), f32(w / 2.0)
    ay1, ax1, ay2, ax2 = f32(c1[1] - hh), f32(c1[0] - hw), f32(c1[1] + hh), f32(c1[0] + hw)
    by1, bx1, by2, bx2 = f32(c2[1] - hh), f32(c2[0] - hw), f32(c2[1] + hh), f32(c2[0] + hw)
    ih = max(f32(min(ay2, by2) - max(ay1, by1)), f32(0))
    iw = max(f32(min(ax2, bx2) - max(ax1, bx1)), f32(0))
    inter = f32(ih * iw)
    return f32(inter / f32(f32(2.0 * (h * w)) - inter))


def valid_slots(centroids: np.ndarray, vals: np.ndarray, crop_hw: Tuple[int, int], nms: bool, thr: float) -> np.ndarray:
    """(B, I) bool: slots that get a crop.  topdown.py:98-104 (NaN mask) and :415-446 (greedy NMS, descending value)."""
    B, I = centroids.shape[:2]
    valid = ~np.isnan(centroids).any(-1)
    if not nms:
        return valid
    out = valid.copy()
    for b in range(B):
        slots = [i for i in range(I) if valid[b, i]]
        if len(slots) <= 1:
            continue
        # argsort(descending=True): NaN first, larger first (tie order is implementation-defined; inputs are tie-free)
        order = sorted(slots, key=lambda i: (0 if np.isnan(vals[b, i]) else 1, -vals[b, i] if not np.isnan(vals[b, i]) else 0, i))
        kept = []
        for i in order:
            if any(centred_iou(centroids[b, i], centroids[b, k], *crop_hw) > f32(thr) for k in kept):
                out[b, i] = False
            else:
                kept.append(i)
    return out


def stage_2(image: torch.Tensor, centroids: torch.Tensor, centroid_vals: torch.Tensor, eff_scale: Optional[torch.Tensor],
            crop_hw: Tuple[int, int], model: Callable, nms: bool = False, nms_threshold: float = 0.5,
            output_stride: int = 1, input_scale: float = 1.0, threshold: float = 0.2,
            refinement: Optional[str] = "integral", patch: int = 5) -> Dict[str, np.ndarray]:
    """predict() after stage 1 + _run_stage_2 (topdown.py:98-150, 186-371).  `model(crops)` returns confmaps or
    (confmaps, class_vectors); centroids are in image space, `image` is the sized image."""
    cen = centroids.numpy().astype(np.float32)
    val = centroid_vals.numpy().astype(np.float32)
    B, I = cen.shape[:2]
    eff = np.ones(B, np.float32) if eff_scale is None else eff_scale.numpy().astype(np.float32)
    valid = valid_slots(cen, val, crop_hw, nms, nms_threshold)
    sized = (cen * eff[:, None, None]).astype(np.float32)
    res = dict(valid=valid, centroids=(sized / eff[:, None, None]).astype(np.float32), scores=val)
    pairs = [(b, i) for b in range(B) for i in range(I) if valid[b, i]]  # nonzero order
    if not pairs:
        return res
    centers = torch.from_numpy(np.stack([sized[b, i] for b, i in pairs]))
    bboxes = opeaks.centered_bboxes(centers, crop_hw[0], crop_hw[1])
    crops = opeaks.crop_patches(image, bboxes, torch.tensor([b for b, _ in pairs]))
    raw = model(crops)
    cms, vecs = raw if isinstance(raw, (tuple, list)) else (raw, None)
    pk, pv = opeaks.global_peaks(cms, threshold, refinement, patch)
    pk = (pk * output_stride) if output_stride != 1 else pk
    pk = (pk / input_scale) if input_scale != 1.0 else pk
    Nn = pk.shape[1]
    kpts = np.full((B, I, Nn, 2), np.nan, np.float32)
    ckpts, vals_o = kpts.copy(), np.full((B, I, Nn), np.nan, np.float32)
    boxes = np.full((B, I, 4, 2), np.nan, np.float32)
    full_crops = np.zeros((B, I) + tuple(crops.shape[1:]), dtype=crops.numpy().dtype)
    bb = bboxes.numpy()
    for r, (b, i) in enumerate(pairs):
        k = pk[r].numpy()
        ckpts[b, i] = k
        kpts[b, i] = ((k + bb[r, 0][None, :]).astype(np.float32) / eff[b]).astype(np.float32)
        vals_o[b, i] = pv[r].numpy()
        boxes[b, i] = (bb[r] / eff[b]).astype(np.float32)
        full_crops[b, i] = crops[r].numpy()
    res.update(kpts=kpts, crop_kpts=ckpts, vals=vals_o, bboxes=boxes, crops=full_crops)
    if vecs is not None:
        K = vecs.shape[1]
        cls = np.full((B, I, Nn), -1, np.int64)
        trk = np.full((B, I), np.nan, np.float32)
        cv = np.full((B, I, K), np.nan, np.float32)
        for b in range(B):  # one assignment PER FRAME (topdown.py:343-371)
            rows = [r for r, (bb_, _) in enumerate(pairs) if bb_ == b]
            if not rows:
                continue
            inds, probs = class_inds_from_vectors(vecs[rows])
            for j, r in enumerate(rows):
                i = pairs[r][1]
                cls[b, i, :] = int(inds[j])
                trk[b, i] = float(probs[j])
                cv[b, i] = vecs[r].numpy()
        res.update(class_inds=cls, tracking=trk, class_vectors=cv)
    return res
