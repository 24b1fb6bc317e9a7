"""Edge maps and part-affinity-field targets - same API as sleap_nn/data/edge_maps.py, CUDA-computed."""

from __future__ import annotations

from typing import Optional, Tuple

import attrs
import torch

from sleap_nn_b200 import _native as N
from sleap_nn_b200.data.utils import ensure_list, expand_to_rank, gaussian_pdf, make_grid_vectors  # noqa: F401


def _f32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def distance_to_edge(points: torch.Tensor, edge_source: torch.Tensor, edge_destination: torch.Tensor) -> torch.Tensor:
    """SQUARED distance between points (..., 2) and segments (n_edges, 2) -> (..., n_edges).

    sleap_nn/data/edge_maps.py:15-78 (edge length is max(|v|^2, 1); projections clamped to [0, 1]).
    """
    dev = N.compute_device(points)
    pts = expand_to_rank(points, 2)
    src = _f32(expand_to_rank(edge_source, 2), dev)
    dst = _f32(expand_to_rank(edge_destination, 2), dev)
    lead = tuple(pts.shape[:-1])
    p = _f32(pts, dev).reshape(-1, 2)
    E = int(src.shape[0])
    out = torch.empty((p.shape[0], E), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib.snb_edge_distance(N.ptr(p), None, None, 1, int(p.shape[0]), N.ptr(src), N.ptr(dst), E, 0, 0.0,
                                        N.ptr(out), N.stream_ptr(dev)), "snb_edge_distance")
    return out.reshape(lead + (E,)).to(points.device)


def make_edge_maps(xv: torch.Tensor, yv: torch.Tensor, edge_source: torch.Tensor, edge_destination: torch.Tensor,
                   sigma: float) -> torch.Tensor:
    """Edge confidence maps (grid_h, grid_w, n_edges); sleap_nn/data/edge_maps.py:81-117."""
    dev = N.compute_device(edge_source)
    xd, yd, src, dst = _f32(xv, dev), _f32(yv, dev), _f32(edge_source, dev), _f32(edge_destination, dev)
    h, w, E = int(yd.shape[0]), int(xd.shape[0]), int(src.shape[0])
    out = torch.empty((h, w, E), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib.snb_edge_distance(None, N.ptr(xd), N.ptr(yd), w, h * w, N.ptr(src), N.ptr(dst), E, 1,
                                        float(2 * sigma**2), N.ptr(out), N.stream_ptr(dev)), "snb_edge_distance")
    return out.to(edge_source.device)


def _pafs(xv, yv, srcs, dsts, sigma, accumulate: bool, out_dtype, dev, batched: bool = False) -> torch.Tensor:
    """srcs/dsts (I, E, 2) -> (E, 2, h, w); with `batched` (G, I, E, 2) -> (G, E, 2, h, w) in one launch."""
    xd, yd, s, d = _f32(xv, dev), _f32(yv, dev), _f32(srcs, dev), _f32(dsts, dev)
    G = int(s.shape[0]) if batched else 1
    I, E = int(s.shape[-3]), int(s.shape[-2])
    h, w = int(yd.shape[0]), int(xd.shape[0])
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("part affinity fields are produced in float32 or bfloat16")
    out = torch.empty((G, E, 2, h, w) if batched else (E, 2, h, w), dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        N.check(
            N.lib.snb_pafs(N.ptr(s), N.ptr(d), G, I, E, N.ptr(xd), N.ptr(yd), h, w, float(2 * sigma**2), int(accumulate),
                           int(out_dtype == torch.bfloat16), N.ptr(out), N.stream_ptr(dev)),
            "snb_pafs",
        )
    return out


def make_multi_pafs_batch(xv: torch.Tensor, yv: torch.Tensor, edge_sources: torch.Tensor, edge_destinations: torch.Tensor,
                          sigma: float, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """`make_multi_pafs` for G frames in one launch (an extension; the reference API is per frame).

    edge_sources / edge_destinations (G, n_instances, n_edges, 2) -> (G, n_edges, 2, grid_h, grid_w).
    """
    dev = N.compute_device(edge_sources)
    out = _pafs(xv, yv, edge_sources, edge_destinations, sigma, True, out_dtype, dev, batched=True)
    return out.to(edge_sources.device)


def make_pafs(xv: torch.Tensor, yv: torch.Tensor, edge_source: torch.Tensor, edge_destination: torch.Tensor,
              sigma: float) -> torch.Tensor:
    """PAFs of one instance, (n_edges, 2, grid_h, grid_w); NaNs are kept (edge_maps.py:120-164)."""
    dev = N.compute_device(edge_source)
    out = _pafs(xv, yv, edge_source.unsqueeze(0), edge_destination.unsqueeze(0), sigma, False, torch.float32, dev)
    return out.to(edge_source.device)


def make_multi_pafs(xv: torch.Tensor, yv: torch.Tensor, edge_sources: torch.Tensor, edge_destinations: torch.Tensor,
                    sigma: float, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """PAFs summed over instances (n_instances, n_edges, 2) -> (n_edges, 2, grid_h, grid_w).

    sleap_nn/data/edge_maps.py:167-220: NaNs of an instance become 0 before the sum.
    """
    dev = N.compute_device(edge_sources)
    out = _pafs(xv, yv, edge_sources, edge_destinations, sigma, True, out_dtype, dev)
    return out.to(edge_sources.device)


def get_edge_points(instances: torch.Tensor, edge_inds: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Source / destination points of each edge, two (n_instances, n_edges, 2) tensors (edge_maps.py:223-247).

    A pure index selection (no arithmetic), kept as tensor indexing on the instances' device.
    """
    source_inds = edge_inds[:, 0].to(torch.int32)
    destination_inds = edge_inds[:, 1].to(torch.int32)
    return instances[:, source_inds], instances[:, destination_inds]


def generate_pafs(
    instances: torch.Tensor,
    img_hw: Tuple[int],
    sigma: float = 1.5,
    output_stride=2,
    edge_inds: Optional[torch.Tensor] = attrs.field(default=None, converter=attrs.converters.optional(ensure_list)),
    flatten_channels: bool = False,
) -> torch.Tensor:
    """PAF targets of a frame (edge_maps.py:250-323).

    instances (1, n_instances, n_nodes, 2); instances without any node strictly inside
    (0, xv[-1]) x (0, yv[-1]) are dropped; PAF sigma is NOT scaled by the stride.  Returns
    (n_edges, 2, grid_h, grid_w) or (2 * n_edges, grid_h, grid_w) when `flatten_channels`.
    """
    image_height, image_width = img_hw
    xv, yv = make_grid_vectors(image_height=image_height, image_width=image_width, output_stride=output_stride)
    grid_height, grid_width = len(yv), len(xv)
    n_edges = len(edge_inds)
    instances = instances[0]
    bound = torch.stack([xv[-1], yv[-1]]).view(1, 1, 2).to(instances.device)
    in_img = ((instances > 0) & (instances < bound)).all(dim=-1).any(dim=1)
    assert len(in_img.shape) == 1
    instances = instances[in_img]
    edge_sources, edge_destinations = get_edge_points(instances, edge_inds)
    assert len(edge_sources.shape) == 3
    assert edge_sources.shape[1:] == (n_edges, 2)
    assert len(edge_destinations.shape) == 3
    assert edge_destinations.shape[1:] == (n_edges, 2)
    pafs = make_multi_pafs(xv=xv, yv=yv, edge_sources=edge_sources, edge_destinations=edge_destinations, sigma=sigma)
    assert pafs.shape == (n_edges, 2, grid_height, grid_width)
    if flatten_channels:
        pafs = pafs.reshape(n_edges * 2, grid_height, grid_width)
    return pafs
