"""CPU oracle for training-target synthesis (confidence maps, PAFs).  TEST INFRASTRUCTURE ONLY.

Restates sleap_nn/data/confidence_maps.py (`confidence_maps.py:NN`),
sleap_nn/data/edge_maps.py (`edge_maps.py:NN`) and sleap_nn/data/utils.py:55-125
(`utils.py:NN`) with torch CPU ops in fp32, every product / sum separately rounded as
the reference's ATen chain does.  Never imported by the product path.  Pinned against
golden vectors generated from the unmodified reference (tests/golden).
"""

from __future__ import annotations

from typing import Tuple

import torch


def grid_vectors(height: int, width: int, stride: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    """xv = 0, s, 2s, ... < width ; yv likewise.  utils.py:55-85."""
    return (
        torch.arange(0, width, step=stride, dtype=torch.float32),
        torch.arange(0, height, step=stride, dtype=torch.float32),
    )


def confmaps(points: torch.Tensor, xv: torch.Tensor, yv: torch.Tensor, sigma: float) -> torch.Tensor:
    """exp(-((xv-x)^2 + (yv-y)^2) / (2 sigma^2)), NaN -> 0.  confidence_maps.py:94-129.

    points (S,N,2) -> (S,N,h,w).
    """
    x = points[..., 0][..., None, None]
    y = points[..., 1][..., None, None]
    dx = xv.view(1, 1, 1, -1) - x
    dy = yv.view(1, 1, -1, 1) - y
    arg = -(dx * dx + dy * dy) / (2 * sigma**2)
    return torch.nan_to_num(torch.exp(arg))


def multi_confmaps(points: torch.Tensor, xv, yv, sigma: float) -> torch.Tensor:
    """Per-node max over instances (and, faithfully, over ALL samples).  confidence_maps.py:132-166.

    points (S,I,N,2) -> (S,N,h,w); every output sample holds the same reduction when
    S > 1, exactly as the reference loop does (real callers pass S = 1).
    """
    s, i, n, _ = points.shape
    out = torch.zeros((s, n, yv.shape[0], xv.shape[0]), dtype=torch.float32)
    for inst in points.reshape(s * i, n, 2):
        out = torch.maximum(out, confmaps(inst[None], xv, yv, sigma))
    return out


def generate_confmaps(instance, img_hw, sigma: float = 1.5, output_stride: int = 2):
    """confidence_maps.py:8-43: sigma is scaled by the stride."""
    if instance.ndim != 3:
        instance = instance.view(instance.shape[0], -1, 2)
    xv, yv = grid_vectors(img_hw[0], img_hw[1], output_stride)
    return confmaps(instance, xv, yv, sigma * output_stride)


def generate_multiconfmaps(instances, img_hw, num_instances, sigma=1.5, output_stride=2, is_centroids=False):
    """confidence_maps.py:46-91."""
    pts = instances[:, :num_instances]
    if is_centroids:
        pts = pts.unsqueeze(-2)
    xv, yv = grid_vectors(img_hw[0], img_hw[1], output_stride)
    return multi_confmaps(pts, xv, yv, sigma * output_stride)


def distance_to_edge(points: torch.Tensor, src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """SQUARED distance from each point to each segment.  edge_maps.py:15-78.

    points (..., 2), src/dst (E,2) -> (..., E).  edge_length is max(|v|^2, 1).
    """
    points = points.reshape((1,) * max(2 - points.dim(), 0) + tuple(points.shape))
    src = src.reshape((1,) * max(2 - src.dim(), 0) + tuple(src.shape))
    dst = dst.reshape((1,) * max(2 - dst.dim(), 0) + tuple(dst.shape))
    v = dst - src  # (E,2)
    length = torch.maximum((v * v).sum(dim=1), torch.tensor(1.0))
    rel = points.unsqueeze(-2) - src  # (...,E,2)
    proj = ((rel * v).sum(dim=-1) / length).clamp(min=0, max=1)
    err = proj.unsqueeze(-1) * v - rel
    return (err * err).sum(dim=-1)


def gaussian_pdf(x: torch.Tensor, sigma: float) -> torch.Tensor:
    """exp(-x^2 / (2 sigma^2)).  utils.py:114-125."""
    return torch.exp(-(x * x) / (2 * sigma**2))


def edge_maps(xv, yv, src, dst, sigma: float) -> torch.Tensor:
    """(h,w,E).  edge_maps.py:81-117.  The squared distance is squared AGAIN by gaussian_pdf."""
    yy, xx = torch.meshgrid(yv, xv, indexing="ij")
    return gaussian_pdf(distance_to_edge(torch.stack((xx, yy), dim=-1), src, dst), sigma)


def pafs(xv, yv, src, dst, sigma: float) -> torch.Tensor:
    """(E,2,h,w) = edge map * unit(dst - src).  edge_maps.py:120-164.  NaNs are kept."""
    v = dst - src
    unit = v / torch.linalg.vector_norm(v, dim=-1, keepdim=True)
    em = edge_maps(xv, yv, src, dst, sigma)  # (h,w,E)
    return (em.unsqueeze(-1) * unit.view(1, 1, -1, 2)).permute(2, 3, 0, 1)


def multi_pafs(xv, yv, srcs, dsts, sigma: float) -> torch.Tensor:
    """Sum over instances, in order, NaN -> 0 per instance.  edge_maps.py:167-220."""
    out = torch.zeros((srcs.shape[1], 2, yv.shape[0], xv.shape[0]), dtype=torch.float32)
    for i in range(srcs.shape[0]):
        one = pafs(xv, yv, srcs[i], dsts[i], sigma)
        out += torch.where(torch.isnan(one), torch.zeros(()), one)
    return out


def edge_points(instances: torch.Tensor, edge_inds: torch.Tensor):
    """edge_maps.py:223-247."""
    e = torch.as_tensor(edge_inds).to(torch.int64)
    return instances[:, e[:, 0]], instances[:, e[:, 1]]


def generate_pafs(instances, img_hw, sigma=1.5, output_stride=2, edge_inds=None, flatten_channels=False):
    """edge_maps.py:250-323: sample 0 only; keep instances with any node strictly inside
    (0, xv[-1]) x (0, yv[-1]); PAF sigma is NOT scaled by the stride."""
    xv, yv = grid_vectors(img_hw[0], img_hw[1], output_stride)
    inst = instances[0]
    bound = torch.stack([xv[-1], yv[-1]]).view(1, 1, 2)
    keep = ((inst > 0) & (inst < bound)).all(dim=-1).any(dim=1)
    s, d = edge_points(inst[keep], torch.as_tensor(edge_inds))
    out = multi_pafs(xv, yv, s, d, sigma)
    if flatten_channels:
        out = out.reshape(-1, yv.shape[0], xv.shape[0])
    return out


def filter_oob_points(points: torch.Tensor, img_height: int, img_width: int) -> torch.Tensor:
    """Keypoints with a negative coordinate, x >= width or y >= height become NaN (both coordinates).
    sleap_nn/data/providers.py:38-69."""
    out = points.clone()
    flat = out.reshape(-1, 2)
    for k in range(flat.shape[0]):
        x, y = float(flat[k, 0]), float(flat[k, 1])
        if x < 0 or x >= img_width or y < 0 or y >= img_height:
            flat[k, 0] = float("nan")
            flat[k, 1] = float("nan")
    return flat.reshape(points.shape)


def batched_dataset_targets(instances, num_instances, edges, img_hw, tracks=None):
    """The per-frame target calls of the datasets' __getitem__ (data/custom_datasets.py:1305-1327, 1489-1511, 1788,
    2835, 2986) over a collated batch, stacked the way the default collate stacks samples.  Knobs are the ones
    tests/golden/make_golden.py:f4_batched_targets uses."""
    from oracle import identity as oid

    H, W = img_hw
    out = {k: [] for k in ("confidence_maps", "part_affinity_fields", "centroid_maps", "single_maps", "class_maps",
                           "class_maps_centroids")}
    for b in range(instances.shape[0]):
        n = int(num_instances[b])
        fr = instances[b]
        out["confidence_maps"].append(generate_multiconfmaps(fr, (H, W), n, sigma=1.5, output_stride=2))
        out["part_affinity_fields"].append(generate_pafs(fr, (H, W), sigma=4.0, output_stride=4,
                                                         edge_inds=torch.as_tensor(edges), flatten_channels=True))
        out["centroid_maps"].append(generate_multiconfmaps(fr[:, :, 0, :], (H, W), n, sigma=2.0, output_stride=2,
                                                           is_centroids=True))
        out["single_maps"].append(generate_confmaps(filter_oob_points(fr[:, 0], H, W), (H, W), sigma=1.5, output_stride=2))
        if tracks is not None and n > 0:
            out["class_maps"].append(oid.generate_class_maps(fr, (H, W), n, tracks[b, :n], 4, 0.2, 3.0, 2))
            out["class_maps_centroids"].append(oid.generate_class_maps(fr[:, :, 0, :], (H, W), n, tracks[b, :n], 4, 0.1,
                                                                       3.0, 4, is_centroids=True))
    return {k: torch.stack(v) for k, v in out.items() if v}
