"""Seeded, tie-free synthetic inputs for parity tests and the CPU baseline.  TEST INFRASTRUCTURE ONLY.

Poses follow SURVEY.md section 8(d): per frame `n_inst` roots uniform inside the image
(margin away from the border), node chains by cumulative steps U(-step, step), clamped
to [8, size-8]; heat-maps are rendered with the ORACLE's multi_confmaps / multi_pafs
(i.e. the reference's arithmetic) plus U(0, 1e-3) noise on the confidence maps to
break plateau ties.  `certify` rejects frames whose integer decisions sit closer to a
rounding / threshold / assignment boundary than the fp32 noise floor.
"""

from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import paf as opaf
from . import peaks as opeaks
from . import targets as otargets


def chain_edges(n_nodes: int) -> List[Tuple[int, int]]:
    """A simple path skeleton 0-1-2-...; n_nodes-1 edges."""
    return [(k, k + 1) for k in range(n_nodes - 1)]


def star_chain_edges(n_nodes: int, fan: int = 3) -> List[Tuple[int, int]]:
    """A tree: node k's parent is (k-1)//fan.  Exercises BFS order != edge order."""
    return [((k - 1) // fan, k) for k in range(1, n_nodes)]


def make_poses(seed: int, n_frames: int, n_inst: int, n_nodes: int, img_hw: Tuple[int, int],
               margin: float = 150.0, step: float = 30.0, edges: Sequence[Tuple[int, int]] = None,
               min_sep: float = 12.0) -> torch.Tensor:
    """(n_frames, n_inst, n_nodes, 2) float32 (x, y) image coordinates.

    Same-node points of different instances are kept at least `min_sep` px apart
    (rejection sampling) so confidence-map blobs of one channel never merge.
    """
    g = np.random.default_rng(seed)
    h, w = img_hw
    if edges is None:
        edges = chain_edges(n_nodes)
    parent = {b: a for a, b in edges}
    out = np.zeros((n_frames, n_inst, n_nodes, 2), np.float32)
    mx, my = min(margin, w / 4), min(margin, h / 4)
    for f in range(n_frames):
        for i in range(n_inst):
            for _try in range(1000):
                p = np.zeros((n_nodes, 2), np.float64)
                for k in range(n_nodes):
                    if k in parent and parent[k] < k:
                        p[k] = p[parent[k]] + g.uniform(-step, step, 2)
                        # keep a minimum limb length so src != dst and lines have direction
                        d = p[k] - p[parent[k]]
                        n = np.hypot(*d)
                        if n < 10.0:
                            p[k] = p[parent[k]] + (d / max(n, 1e-6)) * 10.0 if n > 1e-6 else p[parent[k]] + [10.0, 0.0]
                    else:
                        p[k] = [g.uniform(mx, w - mx), g.uniform(my, h - my)]
                p[:, 0] = np.clip(p[:, 0], 8, w - 8)
                p[:, 1] = np.clip(p[:, 1], 8, h - 8)
                ok = True
                for j in range(i):
                    if (np.hypot(*(out[f, j] - p).T) < min_sep).any():
                        ok = False
                        break
                if ok:
                    # distinct nodes of the SAME instance may overlap freely (different channels)
                    break
            out[f, i] = p.astype(np.float32)
    return torch.from_numpy(out)


def render(poses: torch.Tensor, img_hw, stride: int, edges, sigma_cm: float = 2.5, sigma_paf: float = 2.5,
           noise: float = 1e-3, seed: int = 0):
    """Confidence maps (B,N,h,w) and PAFs (B,2E,h,w), fp32, via the oracle target code."""
    h, w = img_hw
    xv, yv = otargets.grid_vectors(h, w, stride)
    e = torch.as_tensor(list(edges), dtype=torch.int64).reshape(-1, 2)
    cms, pafs = [], []
    gen = torch.Generator().manual_seed(seed)
    for f in range(poses.shape[0]):
        cm = otargets.multi_confmaps(poses[f][None], xv, yv, sigma_cm * stride)[0]
        if noise:
            cm = cm + torch.rand(cm.shape, generator=gen) * noise
        cms.append(cm)
        if e.shape[0]:
            s, d = otargets.edge_points(poses[f], e)
            pafs.append(otargets.multi_pafs(xv, yv, s, d, sigma_paf).reshape(-1, yv.shape[0], xv.shape[0]))
        else:
            pafs.append(torch.zeros((0, yv.shape[0], xv.shape[0])))
    return torch.stack(cms), torch.stack(pafs)


def certify(cms, pafs_bchw, edges, n_nodes, stride, threshold=0.2, n_points=10, ratio=0.25, weight=1.0,
            min_line_scores=0.25, coord_margin=2e-5, score_margin=1e-5, gap_margin=1e-5) -> np.ndarray:
    """Per-frame bool: True when every integer decision has a safe margin.

    Checks: pre-round line coordinate not within `coord_margin` of k+0.5;
    |score - min_line_scores| >= score_margin; optimal assignment beats the best
    assignment that avoids any one of its pairs by >= gap_margin; no coincident peaks.
    """
    B = cms.shape[0]
    ok = np.ones(B, bool)
    pts, vals, si, ci = opeaks.local_peaks(cms, threshold, "integral")
    pts = pts * stride
    pafs_hwc = pafs_bchw.permute(0, 2, 3, 1)
    max_len = opaf.max_edge_length_for(pafs_hwc.shape, stride, ratio)
    t = torch.linspace(0, 1, n_points)
    for b in range(B):
        m = si == b
        p, c = pts[m], ci[m]
        if torch.isnan(p).any():
            ok[b] = False
            continue
        e_i, ep = opaf.connection_candidates(c, edges, n_nodes)
        if e_i.numel() == 0:
            continue
        src, dst = p[ep[:, 0]], p[ep[:, 1]]
        if ((dst - src).abs().sum(1) == 0).any():
            ok[b] = False
            continue
        val = (src[:, :, None] + ((dst - src) / (1 + opaf.F32_EPS))[:, :, None] * t) / stride
        frac = (val - torch.floor(val) - 0.5).abs()
        if (frac < coord_margin).any():
            ok[b] = False
            continue
        sc = opaf.score_lines(opaf.paf_lines(pafs_hwc[b], p, ep, e_i, n_points, stride), p, ep, max_len, weight)
        if torch.isnan(sc).any() or ((sc - min_line_scores).abs() < score_margin).any():
            ok[b] = False
            continue
        for k in range(len(edges)):
            sel = e_i == k
            if not sel.any():
                continue
            s_ids, s_r = torch.unique(ep[sel, 0], return_inverse=True)
            d_ids, d_r = torch.unique(ep[sel, 1], return_inverse=True)
            cost = np.full((len(s_ids), len(d_ids)), np.inf)
            cost[s_r.numpy(), d_r.numpy()] = -sc[sel].double().numpy()
            r, cidx = opaf.lsap_jv(cost)
            best = cost[r, cidx].sum()
            for rr, cc in zip(r, cidx):
                alt = cost.copy()
                alt[rr, cc] = 1e6
                r2, c2 = opaf.lsap_jv(alt)
                if alt[r2, c2].sum() - best < gap_margin:
                    ok[b] = False
    return ok


def split_by_sample(values: torch.Tensor, sample_inds: torch.Tensor, n_samples: int):
    """Concatenated per-peak tensor -> list of per-sample tensors (the bottom-up layer's split)."""
    return [values[sample_inds == b] for b in range(n_samples)]
