"""Class-vector / class-map targets for identity models - same API as sleap_nn/data/identity.py, CUDA-computed."""

from __future__ import annotations

from typing import Tuple

import torch

from sleap_nn_b200 import _native as N
from sleap_nn_b200.data.confidence_maps import make_grid_vectors, make_multi_confmaps


def _class_vectors_dev(class_inds: torch.Tensor, n_classes: int, dev: torch.device) -> torch.Tensor:
    n = int(class_inds.shape[0])
    is_float = class_inds.dtype.is_floating_point
    src = class_inds.detach().to(device=dev, dtype=torch.float32 if is_float else torch.int32).contiguous()
    out = torch.empty((n, int(n_classes)), dtype=torch.int32, device=dev)
    if n * int(n_classes):
        status = torch.zeros((1,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            N.check(N.lib.snb_class_vectors(N.ptr(src), int(is_float), n, int(n_classes), N.ptr(out), N.ptr(status),
                                            N.stream_ptr(dev)), "snb_class_vectors")
        if int(status.item()) & N.STATUS_BAD_INDEX:
            raise RuntimeError("Class values must be smaller than num_classes.")  # F.one_hot's error
    return out


def _class_inds_i32(class_inds: torch.Tensor, dev: torch.device) -> torch.Tensor:
    """Class indices as int32 on `dev`; float indices truncate toward zero like the reference's `.long()`, and every
    negative value means "no class" (data/identity.py:24-31)."""
    ci = class_inds.detach().to(dev)
    if ci.dtype.is_floating_point:
        ci = torch.where(ci >= 0, ci, torch.full_like(ci, -1.0))
    return ci.to(torch.int32).contiguous()


def make_class_vectors(class_inds: torch.Tensor, n_classes: int) -> torch.Tensor:
    """One-hot int32 rows `(n_instances, n_classes)`; an index of -1 gives an all-zero row (data/identity.py:10-32)."""
    dev = N.compute_device(class_inds)
    return _class_vectors_dev(class_inds, n_classes, dev).to(class_inds.device)


def make_class_maps(confmaps: torch.Tensor, class_inds: torch.Tensor, n_classes: int, threshold: float = 0.2) -> torch.Tensor:
    """Identity class maps from per-instance confidence maps; data/identity.py:35-82.

    confmaps (1, n_instances, h, w) -> (1, n_classes, h, w): each instance's share of the summed confidence,
    kept where its own confidence exceeds `threshold`, weighted by the class vectors and max-reduced over
    instances.  The reference reshapes (not transposes) the (n_instances, n_classes) one-hot matrix to
    (n_classes, n_instances, 1, 1); that indexing is reproduced as is.
    """
    dev = N.compute_device(confmaps, class_inds)
    n_inst, h, w = (int(v) for v in confmaps.shape[-3:])
    cms = confmaps.detach().to(device=dev, dtype=torch.float32).reshape(-1, h, w).contiguous()
    if cms.shape[0] != n_inst:
        raise ValueError("make_class_maps expects confmaps of shape (1, n_instances, height, width)")
    if int(class_inds.shape[0]) != n_inst:  # the reference's reshape to [n_classes, n_instances, 1, 1] fails too
        raise RuntimeError(f"shape '[{n_classes}, {n_inst}, 1, 1]' is invalid for input of size "
                           f"{int(class_inds.shape[0]) * int(n_classes)}")
    ci = _class_inds_i32(class_inds, dev).reshape(1, n_inst)
    if n_inst and int(ci.max()) >= int(n_classes):
        raise RuntimeError("Class values must be smaller than num_classes.")  # F.one_hot's error
    out = torch.empty((1, int(n_classes), h, w), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib.snb_class_maps(N.ptr(cms), N.ptr(ci), None, 1, n_inst, int(n_classes), h, w, float(threshold),
                                     N.ptr(out), N.stream_ptr(dev)), "snb_class_maps")
    return out.to(confmaps.device)


def generate_class_maps(instances: torch.Tensor, img_hw: Tuple[int], num_instances: int, class_inds: torch.Tensor,
                        num_tracks: int, class_map_threshold: float = 0.2, sigma: float = 1.5, output_stride: int = 2,
                        is_centroids: bool = False):
    """Class maps from track indices; data/identity.py:85-137.

    instances (1, n_instances, n_nodes, 2), or (1, n_instances, 2) for centroids.  Per-INSTANCE confidence maps
    (max over the instance's nodes, sigma scaled by the stride) are turned into class maps.
    """
    height, width = img_hw
    xv, yv = make_grid_vectors(height, width, output_stride)
    if is_centroids:
        points = instances[:, :num_instances, :].unsqueeze(dim=-3)
    else:
        points = instances[:, :num_instances, :, :].permute(0, 2, 1, 3)
    cms = make_multi_confmaps(points, xv, yv, sigma * output_stride)
    return make_class_maps(cms, class_inds=class_inds, n_classes=num_tracks, threshold=class_map_threshold)
