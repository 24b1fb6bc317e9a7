"""Grid / Gaussian helpers - same API as the hot-path part of sleap_nn/data/utils.py:55-125.

The rest of that module (sleap-io / psutil helpers) is out of scope.
"""

from __future__ import annotations

from typing import Tuple

import torch

from sleap_nn_b200 import _native as N


def ensure_list(x):
    """Wrap non-lists in a list (used as an attrs converter by generate_pafs' signature)."""
    return x if isinstance(x, list) else [x]


def make_grid_vectors(image_height: int, image_width: int, output_stride: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sampling grid vectors `(xv, yv)`: 0, stride, 2*stride, ... below the image size (fp32, CPU).

    sleap_nn/data/utils.py:55-85.  Two tiny host aranges, exact in fp32; they are inputs of the
    target kernels, not compute.
    """
    xv = torch.arange(0, image_width, step=output_stride, dtype=torch.float32)
    yv = torch.arange(0, image_height, step=output_stride, dtype=torch.float32)
    return xv, yv


def expand_to_rank(x: torch.Tensor, target_rank: int, prepend: bool = True) -> torch.Tensor:
    """Add singleton dims up to `target_rank` (a view; sleap_nn/data/utils.py:88-111)."""
    n = max(target_rank - x.dim(), 0)
    shape = [1] * n + list(x.shape) if prepend else list(x.shape) + [1] * n
    return x.reshape(shape)


def gaussian_pdf(x: torch.Tensor, sigma: float) -> torch.Tensor:
    """Unnormalised zero-centred Gaussian `exp(-x^2 / (2 sigma^2))` (sleap_nn/data/utils.py:114-125)."""
    dev = N.compute_device(x)
    xd = x.detach().to(device=dev, dtype=torch.float32).contiguous()
    out = torch.empty_like(xd)
    if xd.numel():
        with torch.cuda.device(dev):
            N.check(N.lib.snb_gaussian_pdf(N.ptr(xd), xd.numel(), float(2 * sigma**2), N.ptr(out), N.stream_ptr(dev)),
                    "snb_gaussian_pdf")
    return out.to(x.device)
