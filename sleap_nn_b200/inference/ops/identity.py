"""Multi-class identity grouping - same API as sleap_nn/inference/ops/identity.py, computed by CUDA kernels.

`group_class_peaks`, `classify_peaks_from_maps` and `get_class_inds_from_vectors` keep the reference's
signatures, return dtypes and device convention.  The per-(sample, channel) assignment problems are solved on
the device by the same LSAP kernel the PAF matcher uses (scipy.optimize.linear_sum_assignment semantics, one
warp per group); nothing is copied to the host except, for `group_class_peaks`, the one count word that sizes
its variable-length result.  Kernels: sleap_nn_b200/csrc/identity.cu.
"""

from __future__ import annotations

from typing import Tuple

import torch

from sleap_nn_b200 import _native as N


def _raise_for_status(status: torch.Tensor) -> None:
    st = int(status.item())
    if st & N.STATUS_LSAP_INVALID:
        raise ValueError("matrix contains invalid numeric entries")  # scipy's message for NaN / -inf costs
    if st & N.STATUS_LSAP_INFEASIBLE:
        raise ValueError("cost matrix is infeasible")
    if st & N.STATUS_LSAP_TOO_LARGE:
        raise RuntimeError("a (sample, channel) group or the class count exceeds the device solver's limit of 128")


def _i32(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.int32).contiguous()


def _f32(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def group_class_peaks(peak_class_probs: torch.Tensor, peak_sample_inds: torch.Tensor, peak_channel_inds: torch.Tensor,
                      n_samples: int, n_channels: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Match peaks to classes per (sample, channel) group; sleap_nn/inference/ops/identity.py:13-71.

    Returns int64 `(peak_inds, class_inds)` on the device of `peak_sample_inds`, groups in (sample, channel)
    order, only where the assigned class is also the peak's most probable class.
    """
    out_dev = peak_sample_inds.device
    dev = N.compute_device(peak_class_probs, peak_sample_inds)
    P = int(peak_class_probs.shape[0])
    K = int(peak_class_probs.shape[1]) if peak_class_probs.dim() > 1 else 0
    groups = int(n_samples) * int(n_channels)
    empty = (torch.empty(0, dtype=torch.int64, device=out_dev), torch.empty(0, dtype=torch.int64, device=out_dev))
    if P == 0 or K == 0 or groups == 0:
        return empty
    probs, si, ci = _f32(peak_class_probs, dev), _i32(peak_sample_inds, dev), _i32(peak_channel_inds, dev)
    with torch.cuda.device(dev):
        st = N.stream_ptr(dev)
        g_peak = torch.empty((groups, K), dtype=torch.int64, device=dev)
        g_class = torch.empty((groups, K), dtype=torch.int64, device=dev)
        g_count = torch.empty((groups,), dtype=torch.int32, device=dev)
        status = torch.zeros((2,), dtype=torch.int32, device=dev)  # [status, total]
        N.check(N.lib.snb_classify_peaks(None, int(n_samples), K, 0, 0, 0, 0, 0, 0, None, None, N.ptr(si), N.ptr(ci), P,
                                         int(n_channels), N.ptr(probs), N.ptr(g_peak), N.ptr(g_class), N.ptr(g_count),
                                         None, None, None, N.ptr(status), st), "snb_classify_peaks")
        cap = min(P, groups * K)
        o_peak = torch.empty((cap,), dtype=torch.int64, device=dev)
        o_class = torch.empty((cap,), dtype=torch.int64, device=dev)
        N.check(N.lib.snb_pack_class_matches(N.ptr(g_peak), N.ptr(g_class), N.ptr(g_count), groups, K, N.ptr(o_peak),
                                             N.ptr(o_class), N.ptr(status[1:]), st), "snb_pack_class_matches")
        _raise_for_status(status[:1])
        total = int(status[1].item())
    return o_peak[:total].to(out_dev), o_class[:total].to(out_dev)


def classify_peaks_from_maps(class_maps: torch.Tensor, peak_points: torch.Tensor, peak_vals: torch.Tensor,
                             peak_sample_inds: torch.Tensor, peak_channel_inds: torch.Tensor, n_channels: int
                             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Classify and group local peaks by their class-map probability; ops/identity.py:74-149.

    class_maps (n_samples, n_classes, H, W); peaks as `find_local_peaks` returns them, in class-map pixels.
    Returns `(points (S, K, C, 2), point_vals (S, K, C), class_probs (S, K, C))`, NaN where no peak was assigned,
    on the device of `class_maps`.  One launch: gather under the rounded peak positions, per-group assignment,
    arg-max filter and scatter.
    """
    out_dev = class_maps.device
    dev = N.compute_device(class_maps, peak_points)
    S, K, H, W = (int(v) for v in class_maps.shape)
    Cn, P = int(n_channels), int(peak_points.shape[0])
    maps = class_maps.detach().to(device=dev, dtype=torch.float32)  # any strides: read in place
    with torch.cuda.device(dev):
        points = torch.empty((S, K, Cn, 2), dtype=torch.float32, device=dev)
        vals = torch.empty((S, K, Cn), dtype=torch.float32, device=dev)
        cprobs = torch.empty((S, K, Cn), dtype=torch.float32, device=dev)
        if S * Cn == 0 or K == 0:
            for t in (points, vals, cprobs):
                t.fill_(float("nan"))
            return points.to(out_dev), vals.to(out_dev), cprobs.to(out_dev)
        xy, pv = _f32(peak_points.reshape(-1, 2), dev), _f32(peak_vals, dev)
        si, ci = _i32(peak_sample_inds, dev), _i32(peak_channel_inds, dev)
        probs = torch.empty((max(P, 1), K), dtype=torch.float32, device=dev)
        status = torch.zeros((1,), dtype=torch.int32, device=dev)
        ms, mk, mh, mw = maps.stride()
        N.check(N.lib.snb_classify_peaks(N.ptr(maps), S, K, H, W, ms, mk, mh, mw, N.ptr(xy), N.ptr(pv), N.ptr(si),
                                         N.ptr(ci), P, Cn, N.ptr(probs), None, None, None, N.ptr(points), N.ptr(vals),
                                         N.ptr(cprobs), N.ptr(status), N.stream_ptr(dev)), "snb_classify_peaks")
        _raise_for_status(status)
    return points.to(out_dev), vals.to(out_dev), cprobs.to(out_dev)


def get_class_inds_from_vectors(peak_class_probs: torch.Tensor):
    """One optimal assignment of samples (crops) to classes; ops/identity.py:152-173.

    peak_class_probs (n_samples, n_classes) -> `(class_inds (n_samples,) int64, class_probs (n_samples,) f32)`
    as CPU tensors (the reference builds them with `torch.full` on the host); unassigned rows are -1 / NaN.
    """
    dev = N.compute_device(peak_class_probs)
    n, K = (int(v) for v in peak_class_probs.shape)
    if n == 0:
        return torch.full((0,), -1, dtype=torch.int64), torch.full((0,), float("nan"))
    probs = _f32(peak_class_probs, dev)
    with torch.cuda.device(dev):
        ws = torch.empty((int(N.lib.snb_class_inds_workspace_bytes(n, K)) + 15,), dtype=torch.uint8, device=dev)
        inds = torch.empty((n,), dtype=torch.int64, device=dev)
        vals = torch.empty((n,), dtype=torch.float32, device=dev)
        status = torch.zeros((1,), dtype=torch.int32, device=dev)
        ws_ptr = (ws.data_ptr() + 15) & ~15
        N.check(N.lib.snb_class_inds_from_vectors(N.ptr(probs), n, K, ws_ptr, N.ptr(inds), N.ptr(vals), N.ptr(status),
                                                  N.stream_ptr(dev)), "snb_class_inds_from_vectors")
        _raise_for_status(status)
    return inds.cpu(), vals.cpu()
