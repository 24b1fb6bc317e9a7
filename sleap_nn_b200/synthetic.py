"""On-device synthetic bottom-up heat-maps (for benchmarks, smoke tests and full-size property tests).

Poses are drawn on the host from a seeded generator (per frame `n_inst`
roots inside the image, node chains by bounded random steps); the confidence maps and PAFs are
rendered ON THE DEVICE by the target-synthesis kernels (`snb_confmaps`, `snb_pafs`) - i.e. with
the reference's target arithmetic - plus U(0, noise) on the confidence maps to break plateau ties.
"""

from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from sleap_nn_b200 import _native as N
from sleap_nn_b200.data.confidence_maps import _confmaps
from sleap_nn_b200.data.edge_maps import _pafs
from sleap_nn_b200.data.utils import make_grid_vectors


def chain_edges(n_nodes: int) -> List[Tuple[int, int]]:
    return [(k, k + 1) for k in range(n_nodes - 1)]


def random_poses(seed: int, n_frames: int, n_inst: int, n_nodes: int, img_hw: Tuple[int, int],
                 edges: Sequence[Tuple[int, int]], margin: float = 150.0, step: float = 30.0,
                 min_limb: float = 10.0, min_sep: float = 12.0) -> torch.Tensor:
    """(n_frames, n_inst, n_nodes, 2) fp32 (x, y): rejection-sampled so same-node blobs never merge."""
    g = np.random.default_rng(seed)
    h, w = img_hw
    parent = {b: a for a, b in edges}
    mx, my = min(margin, w / 4), min(margin, h / 4)
    out = np.zeros((n_frames, n_inst, n_nodes, 2), np.float32)
    for f in range(n_frames):
        for i in range(n_inst):
            for _ in range(1000):
                p = np.zeros((n_nodes, 2))
                for k in range(n_nodes):
                    if k in parent and parent[k] < k:
                        d = g.uniform(-step, step, 2)
                        n = float(np.hypot(*d))
                        if n < min_limb:
                            d = np.array([min_limb, 0.0]) if n < 1e-6 else d / n * min_limb
                        p[k] = p[parent[k]] + d
                    else:
                        p[k] = [g.uniform(mx, w - mx), g.uniform(my, h - my)]
                p[:, 0] = np.clip(p[:, 0], 8, w - 8)
                p[:, 1] = np.clip(p[:, 1], 8, h - 8)
                if all((np.hypot(*(out[f, j] - p).T) >= min_sep).all() for j in range(i)):
                    break
            out[f, i] = p
    return torch.from_numpy(out)


def render_batch(poses: torch.Tensor, img_hw: Tuple[int, int], stride: int, edges: Sequence[Tuple[int, int]],
                 device: torch.device, sigma_cm: float = 2.5, sigma_paf: float = 2.5, noise: float = 1e-3,
                 seed: int = 0):
    """poses (B, I, N, 2) -> (cms (B, N, h, w), pafs (B, 2E, h, w)) fp32 on `device`."""
    B, I, Nn, _ = poses.shape
    xv, yv = make_grid_vectors(img_hw[0], img_hw[1], stride)
    h, w = int(yv.shape[0]), int(xv.shape[0])
    with torch.cuda.device(device):
        cms = _confmaps(poses, xv, yv, sigma_cm * stride, torch.float32, device)  # G = B frames at once
        if noise:
            g = torch.Generator(device=device).manual_seed(seed)
            cms += torch.rand(cms.shape, generator=g, device=device) * noise
        e = torch.tensor(list(edges), dtype=torch.int64).reshape(-1, 2)
        E = int(e.shape[0])
        if E:
            pd = poses.to(device)
            pafs = _pafs(xv, yv, pd[:, :, e[:, 0]], pd[:, :, e[:, 1]], sigma_paf, True, torch.float32, device,
                         batched=True).reshape(B, 2 * E, h, w)
        else:
            pafs = torch.empty((B, 0, h, w), dtype=torch.float32, device=device)
    return cms, pafs
