"""`interp1d` - same API as sleap_nn/inference/utils.py:29-130, computed by a CUDA kernel.

Only the hot-path function of that module is provided (`get_skeleton_from_config` needs
sleap-io and is out of scope).
"""

from __future__ import annotations

import torch

from sleap_nn_b200 import _native as N


def interp1d(x: torch.Tensor, y: torch.Tensor, xnew: torch.Tensor) -> torch.Tensor:
    """Linear 1-D interpolation with the reference's broadcasting and rounding.

    x: (N,) or (D, N) sorted knots; y: (N,) or (D, N) float; xnew: (P,) or (D, P).
    Returns (P,) when `y` is 1-D, else (D, P) - or (1, D*P) when x and y have one row and xnew
    several, exactly like the reference (it never reshapes that case back).
    """
    v = {}
    for name, vec in {"x": x, "y": y, "xnew": xnew}.items():
        assert len(vec.shape) <= 2, "interp1d: all inputs must be at most 2-D."
        v[name] = vec[None, :] if len(vec.shape) == 1 else vec
    assert v["x"].shape[1] == v["y"].shape[1] and (
        v["x"].shape[0] == v["y"].shape[0] or v["x"].shape[0] == 1 or v["y"].shape[0] == 1
    ), (
        "x and y must have the same number of columns, and either "
        "the same number of row or one of them having only one "
        "row."
    )
    if (v["x"].shape[0] == 1) and (v["y"].shape[0] == 1) and (v["xnew"].shape[0] > 1):
        v["xnew"] = v["xnew"].contiguous().view(1, -1)
    rows = max(v["x"].shape[0], v["xnew"].shape[0])
    dev = N.compute_device(y, x, xnew)
    out_dev = y.device
    f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    xd, yd, qd = f(v["x"]), f(v["y"]), f(v["xnew"])
    if yd.shape[0] not in (1, rows):
        raise ValueError("interp1d: y has an incompatible number of rows")
    n, p = int(xd.shape[1]), int(qd.shape[1])
    out = torch.empty((rows, p), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        N.check(
            N.lib.snb_interp1d(N.ptr(xd), int(xd.shape[0]), N.ptr(yd), int(yd.shape[0]), N.ptr(qd), int(qd.shape[0]), n,
                               p, rows, N.ptr(out), N.stream_ptr(dev)),
            "snb_interp1d",
        )
    out = out.to(out_dev)
    if len(y.shape) == 1:
        out = out.view(-1)
    return out
