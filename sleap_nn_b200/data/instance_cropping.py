"""`make_centered_bboxes` - same API as sleap_nn/data/instance_cropping.py:129-171, CUDA-computed.

Only the hot-path function of that module is provided (the crop-size / dataset helpers are
out of scope).
"""

from __future__ import annotations

import torch

from sleap_nn_b200 import _native as N


def make_centered_bboxes(centroids: torch.Tensor, box_height: int, box_width: int) -> torch.Tensor:
    """Corner boxes (top-left, top-right, bottom-right, bottom-left) centred on `centroids`.

    centroids (..., 2) in (x, y) -> (..., 4, 2); corners are centre -/+ size/2 inset by 0.5 px.
    """
    dev = N.compute_device(centroids)
    out_dev, out_dtype = centroids.device, centroids.dtype
    c = centroids.to(device=dev, dtype=torch.float32).contiguous()
    lead = tuple(c.shape[:-1])
    n = c.numel() // 2
    out = torch.empty(lead + (4, 2), dtype=torch.float32, device=dev)
    if n:
        with torch.cuda.device(dev):
            N.check(N.lib.snb_centered_bboxes(N.ptr(c), n, float(box_height / 2), float(box_width / 2), N.ptr(out),
                                              N.stream_ptr(dev)), "snb_centered_bboxes")
    if out_dtype.is_floating_point and out_dtype != torch.float32:
        out = out.to(out_dtype)
    return out.to(out_dev)
