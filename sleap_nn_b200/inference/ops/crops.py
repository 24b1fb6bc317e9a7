"""Bbox creation + cropping - same API as sleap_nn/inference/ops/crops.py, CUDA-computed."""

from __future__ import annotations

import torch

from sleap_nn_b200 import _native as N
from sleap_nn_b200.data.instance_cropping import make_centered_bboxes  # noqa: F401  (re-export, as the reference)

_ELEM_OK = (1, 2, 4, 8)


def crop_bboxes(images: torch.Tensor, bboxes: torch.Tensor, sample_inds: torch.Tensor) -> torch.Tensor:
    """Crop integer-aligned, zero-padded boxes (sleap_nn/inference/ops/crops.py:31-124).

    images (samples, channels, H, W) of any 1/2/4/8-byte dtype; bboxes (n, 4, 2) float32 corners
    (TL, TR, BR, BL); sample_inds (n,).  Returns (n, channels, crop_h, crop_w) in the images'
    dtype and on the images' device; crop size is read from bbox 0; the top-left is
    trunc(tl + size // 2) - size // 2; out-of-image taps are 0.
    """
    n = bboxes.shape[0]
    if n == 0:
        return torch.empty(0, images.shape[1], 0, 0, device=images.device, dtype=images.dtype)
    # Crop size from the first bbox (host read, exactly like the reference's .item()).
    first = bboxes[0].detach().to("cpu", torch.float32)
    crop_h = int(abs(first[3, 1] - first[0, 1]).item()) + 1
    crop_w = int(abs(first[1, 0] - first[0, 0]).item()) + 1
    dev = N.compute_device(images, bboxes)
    out_dev = images.device
    if images.element_size() not in _ELEM_OK:
        raise TypeError(f"crop_bboxes: unsupported dtype {images.dtype}")
    img = images.to(dev)
    bb = bboxes.to(device=dev, dtype=torch.float32).contiguous()
    if not isinstance(sample_inds, torch.Tensor):
        sample_inds = torch.tensor(sample_inds)
    si = sample_inds.to(device=dev, dtype=torch.int64).contiguous()
    S, Cn, H, W = img.shape
    out = torch.empty((n, Cn, crop_h, crop_w), dtype=img.dtype, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    sb, sc, sh, sw = img.stride()
    with torch.cuda.device(dev):
        N.check(
            N.lib.snb_crop_bboxes(N.ptr(img), img.element_size(), S, Cn, H, W, sb, sc, sh, sw, N.ptr(bb), N.ptr(si), n,
                                  crop_h, crop_w, N.ptr(out), N.ptr(status), N.stream_ptr(dev)),
            "snb_crop_bboxes",
        )
    if not images.is_cuda:  # CPU caller: we synchronise anyway, so surface index errors like torch would
        if int(status.item()) & N.STATUS_BAD_INDEX:
            raise IndexError("crop_bboxes: sample_inds out of range")
    return out.to(out_dev)
