"""CPU oracle for interp1d (sleap_nn/inference/utils.py:29-130).  TEST INFRASTRUCTURE ONLY.

Explicit loops, numpy float32 scalars (one rounding per operation, as the reference's tensor ops).  For every query
point: the left knot is searchsorted(x, q) - 1 clamped to [0, n - 2] (so queries outside the knots extrapolate along the
first / last segment), the segment slope is (y[k+1] - y[k]) / (eps + (x[k+1] - x[k])) with eps = float32 machine epsilon,
and the value is y[k] + slope * (q - x[k]).  One quirk is kept: when x has a single row but y several, the reference
indexes its slope table flat, so every row uses the slopes of y's FIRST row (utils.py:112-124).  Pinned against the live reference by
tests/test_oracle_fuzz_vs_reference.py::test_interp1d_fuzz.  Never imported by the product path.
"""

from __future__ import annotations

import numpy as np
import torch

F = np.float32


def interp1d(x: torch.Tensor, y: torch.Tensor, xnew: torch.Tensor) -> torch.Tensor:
    as2d = lambda t: t[None, :] if t.dim() == 1 else t
    xv, yv, qv = as2d(x), as2d(y), as2d(xnew)
    if xv.shape[0] == 1 and yv.shape[0] == 1 and qv.shape[0] > 1:
        qv = qv.contiguous().view(1, -1)
    rows = max(xv.shape[0], qv.shape[0])
    xn, yn, qn = (t.detach().cpu().numpy().astype(np.float32) for t in (xv, yv, qv))
    n = xn.shape[1]
    eps = F(np.finfo(np.float32).eps)
    out = np.zeros((rows, qn.shape[1]), np.float32)
    for r in range(rows):
        xr = xn[r if xn.shape[0] > 1 else 0]
        yr = yn[r if yn.shape[0] > 1 else 0]
        qr = qn[r if qn.shape[0] > 1 else 0]
        for j, q in enumerate(qr):
            k = int(np.searchsorted(xr, q, side="left")) - 1   # torch.searchsorted default: right=False
            k = min(max(k, 0), n - 2)
            ys = yn[0] if xn.shape[0] == 1 else yr   # reference quirk: one row of knots -> the slope table is indexed
            slope = F(F(ys[k + 1] - ys[k]) / F(eps + F(xr[k + 1] - xr[k])))  # flat, so every row uses row 0's slopes
            out[r, j] = F(yr[k] + F(slope * F(q - xr[k])))
    res = torch.from_numpy(out)
    return res.view(-1) if y.dim() == 1 else res
