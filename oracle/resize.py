"""CPU oracle for apply_input_scale (sleap_nn/inference/ops/coord.py:93-109).  TEST INFRASTRUCTURE ONLY.

The reference calls `F.interpolate(image, size=(int(H*s), int(W*s)), mode="bilinear", align_corners=False)`; this
restates that resampling rule with numpy so the kernel is checked against a formula, and the formula against ATen
on the CPU (tests/test_oracle_golden.py::test_bilinear_resize_restatement_matches_aten).  Never imported by the
product path.
"""

from __future__ import annotations

import numpy as np


def _axis(n_in: int, n_out: int):
    scale = np.float32(n_in) / np.float32(n_out)
    d = np.arange(n_out, dtype=np.float32)
    src = np.maximum(scale * (d + np.float32(0.5)) - np.float32(0.5), np.float32(0)).astype(np.float32)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = i0 + (i0 < n_in - 1)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    return i0, i1, (np.float32(1) - l1).astype(np.float32), l1


def apply_input_scale(image: np.ndarray, input_scale: float) -> np.ndarray:
    """image (B, C, H, W) -> (B, C, int(H*s), int(W*s)); weights in fp32, accumulation in fp64 (rounded once)."""
    if input_scale == 1.0:
        return image
    H, W = image.shape[-2:]
    oh, ow = int(H * input_scale), int(W * input_scale)
    y0, y1, ly0, ly1 = _axis(H, oh)
    x0, x1, lx0, lx1 = _axis(W, ow)
    img = image.astype(np.float64)
    top = img[..., y0, :][..., :, x0] * lx0 + img[..., y0, :][..., :, x1] * lx1
    bot = img[..., y1, :][..., :, x0] * lx0 + img[..., y1, :][..., :, x1] * lx1
    return (top * ly0[:, None] + bot * ly1[:, None]).astype(image.dtype)
