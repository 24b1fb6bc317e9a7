"""Coordinate-space ops - same API as sleap_nn/inference/ops/coord.py, computed on the device.

The reference's four `undo_*` / `add_crop_offset` functions are single elementwise tensor ops; here they
share one kernel (`snb_coord_ladder_apply`) whose steps are separately rounded fp32 ops, so chaining them
reproduces the reference bit for bit.  The same ladder is fused into the peak kernels' epilogues
(`sleap_nn_b200.inference.layers`), which is where production traffic goes.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from sleap_nn_b200 import _native as N


def make_ladder(dev: torch.device, n_samples: int, stride: float = 1.0, input_scale: float = 1.0,
                eff_scale: Optional[torch.Tensor] = None, crop_offset: Optional[torch.Tensor] = None,
                eff_scale2: Optional[torch.Tensor] = None, scatter: Optional[torch.Tensor] = None):
    """Build a `CoordLadder` struct; returns (struct, keep-alive tensors)."""
    keep = []

    def per_sample(t, width, dtype=torch.float32):
        if t is None:
            return None
        t = t.detach().to(device=dev, dtype=dtype).reshape(-1, *([width] if width > 1 else []))
        if t.shape[0] == 1 and n_samples != 1:
            t = t.expand(n_samples, *t.shape[1:])
        if t.shape[0] != n_samples:
            raise ValueError(f"expected one entry per sample ({n_samples}), got {t.shape[0]}")
        t = t.contiguous()
        keep.append(t)
        return N.ptr(t)

    lad = N.CoordLadder(float(stride), float(input_scale), per_sample(eff_scale, 1), per_sample(crop_offset, 2),
                        per_sample(eff_scale2, 1), per_sample(scatter, 1, torch.int32))
    return lad, keep


def _apply(coords: torch.Tensor, **kw) -> torch.Tensor:
    if coords.shape[-1] != 2:
        raise ValueError("coords must end in an xy axis of size 2")
    dev = N.compute_device(coords)
    x = coords.detach().to(device=dev, dtype=torch.float32).contiguous()
    n_samples = int(x.shape[0]) if x.dim() > 1 else 1
    pairs = x.numel() // 2 // max(n_samples, 1) if x.numel() else 0
    out = torch.empty_like(x)
    if x.numel():
        with torch.cuda.device(dev):
            lad, keep = make_ladder(dev, n_samples, **kw)
            N.check(N.lib.snb_coord_ladder_apply(N.ptr(x), n_samples, pairs, C.byref(lad), N.ptr(out), N.stream_ptr(dev)),
                    "snb_coord_ladder_apply")
            del keep
    return out.to(device=coords.device, dtype=coords.dtype if coords.dtype.is_floating_point else torch.float32)


def undo_stride(coords: torch.Tensor, output_stride: int) -> torch.Tensor:
    """Confmap pixels -> input pixels: `coords * output_stride` (ops/coord.py:27-39); stride 1 is the identity."""
    if output_stride == 1:
        return coords
    return _apply(coords, stride=float(output_stride))


def undo_input_scale(coords: torch.Tensor, input_scale: float) -> torch.Tensor:
    """`coords / input_scale` (ops/coord.py:42-55); 1.0 is the identity."""
    if input_scale == 1.0:
        return coords
    return _apply(coords, input_scale=float(input_scale))


def undo_eff_scale(coords: torch.Tensor, eff_scale: torch.Tensor) -> torch.Tensor:
    """Per-sample `coords[b] / eff_scale[b]` (ops/coord.py:58-76); all-ones is the identity."""
    if torch.all(eff_scale == 1.0):
        return coords
    return _apply(coords, eff_scale=eff_scale)


def add_crop_offset(peaks: torch.Tensor, crop_topleft: torch.Tensor) -> torch.Tensor:
    """Crop-local peaks -> full-image coordinates (ops/coord.py:79-90).

    peaks (B*I, N, 2) or (B, I, N, 2); crop_topleft (B*I, 2) in (x, y) order.
    """
    shape = peaks.shape
    flat = peaks.reshape(crop_topleft.reshape(-1, 2).shape[0], -1, 2)
    return _apply(flat, crop_offset=crop_topleft.reshape(-1, 2)).reshape(shape)


_RESIZE_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def apply_input_scale(image: torch.Tensor, input_scale: float) -> torch.Tensor:
    """Bilinear resize of an image batch by `input_scale` (ops/coord.py:93-109); 1.0 returns the input itself.

    image (B, C, H, W) float -> (B, C, int(H * s), int(W * s)), same dtype and device: what
    `F.interpolate(mode="bilinear", align_corners=False)` computes, in one kernel that reads the input through its
    strides.
    """
    if input_scale == 1.0:
        return image
    if image.dim() != 4 or image.dtype not in _RESIZE_DTYPES:
        raise TypeError("apply_input_scale expects a (B, C, H, W) float32 / float16 / bfloat16 tensor")
    B, Cn, H, W = (int(v) for v in image.shape)
    oh, ow = int(H * input_scale), int(W * input_scale)
    if oh <= 0 or ow <= 0:
        raise RuntimeError(f"Input and output sizes should be greater than 0, but got input (H: {H}, W: {W}) output (H: {oh}, W: {ow})")
    dev = N.compute_device(image)
    x = image.detach().to(dev)
    if x.stride(0) != Cn * x.stride(1):  # (B, C) must collapse into one plane axis
        x = x.contiguous()
    out = torch.empty((B, Cn, oh, ow), dtype=x.dtype, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib.snb_bilinear_resize(N.ptr(x), _RESIZE_DTYPES[x.dtype], B * Cn, H, W, x.stride(1), x.stride(2),
                                          x.stride(3), oh, ow, N.ptr(out), N.stream_ptr(dev)), "snb_bilinear_resize")
    return out.to(image.device)


__all__: Tuple[str, ...] = ("undo_stride", "undo_input_scale", "undo_eff_scale", "add_crop_offset", "apply_input_scale")
