"""Confidence-map targets - same API as sleap_nn/data/confidence_maps.py, computed by CUDA kernels.

Results live on the device of the input points (the reference's datasets pass CPU tensors and
get CPU tensors back).  `out_dtype=torch.bfloat16` and `device=` are extensions for on-device
training pipelines; defaults reproduce the reference exactly (fp32).
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch

from sleap_nn_b200 import _native as N
from sleap_nn_b200.data.utils import make_grid_vectors


def _confmaps(points_gin2: torch.Tensor, xv: torch.Tensor, yv: torch.Tensor, sigma: float,
              out_dtype: torch.dtype, dev: torch.device) -> torch.Tensor:
    """points (G, I, N, 2) -> (G, N, h, w) on `dev` (max over I)."""
    G, I, Nn, _ = points_gin2.shape
    f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    pts, xd, yd = f(points_gin2), f(xv), f(yv)
    h, w = int(yd.shape[0]), int(xd.shape[0])
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("confidence maps are produced in float32 or bfloat16")
    out = torch.empty((G, Nn, h, w), dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        N.check(
            N.lib.snb_confmaps(N.ptr(pts), G, I, Nn, N.ptr(xd), N.ptr(yd), h, w, float(2 * sigma**2),
                               int(out_dtype == torch.bfloat16), N.ptr(out), N.stream_ptr(dev)),
            "snb_confmaps",
        )
    return out


def make_confmaps(points_batch: torch.Tensor, xv: torch.Tensor, yv: torch.Tensor, sigma: float,
                  out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """`exp(-((xv-x)^2 + (yv-y)^2) / (2 sigma^2))` per point, NaN points give all-zero maps.

    points_batch (n_samples, n_nodes, 2) -> (n_samples, n_nodes, grid_h, grid_w);
    sleap_nn/data/confidence_maps.py:94-129.
    """
    dev = N.compute_device(points_batch)
    out = _confmaps(points_batch.unsqueeze(1), xv, yv, sigma, out_dtype, dev)
    return out.to(points_batch.device)


def make_multi_confmaps(points_batch: torch.Tensor, xv: torch.Tensor, yv: torch.Tensor, sigma: float,
                        out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Per-node maximum over instances; sleap_nn/data/confidence_maps.py:132-166.

    points_batch (n_samples, n_instances, n_nodes, 2) -> (n_samples, n_nodes, grid_h, grid_w).
    Faithful to the reference loop: with n_samples > 1 the maximum runs over the instances of
    ALL samples and every output sample holds that same reduction (callers pass one sample).
    """
    dev = N.compute_device(points_batch)
    S, I, Nn, _ = points_batch.shape
    out = _confmaps(points_batch.reshape(1, S * I, Nn, 2), xv, yv, sigma, out_dtype, dev)
    if S != 1:
        out = out.expand(S, -1, -1, -1).contiguous()
    return out.to(points_batch.device)


def generate_confmaps(instance: torch.Tensor, img_hw: Tuple[int], sigma: float = 1.5,
                      output_stride: int = 2) -> torch.Tensor:
    """Single-instance confidence maps; sigma is scaled by the stride (confidence_maps.py:8-43)."""
    if instance.ndim != 3:
        instance = instance.view(instance.shape[0], -1, 2)
    height, width = img_hw
    xv, yv = make_grid_vectors(height, width, output_stride)
    return make_confmaps(instance, xv, yv, sigma * output_stride)


def generate_multiconfmaps(instances: torch.Tensor, img_hw: Tuple[int], num_instances: int, sigma: float = 1.5,
                           output_stride: int = 2, is_centroids: bool = False) -> torch.Tensor:
    """Multi-instance (or centroid) confidence maps (confidence_maps.py:46-91)."""
    if is_centroids:
        points = instances[:, :num_instances, :].unsqueeze(dim=-2)
    else:
        points = instances[:, :num_instances, :, :]
    height, width = img_hw
    xv, yv = make_grid_vectors(height, width, output_stride)
    return make_multi_confmaps(points, xv, yv, sigma * output_stride)
