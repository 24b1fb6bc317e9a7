"""Peak finding ops - same API as sleap_nn/inference/ops/peaks.py, computed by CUDA kernels.

Each function keeps the reference's signature, defaults, return dtypes and device
convention (results live on the input's device; CPU inputs are staged to the current
CUDA device, never computed on the host).  Kernels: sleap_nn_b200/csrc/peaks.cu.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from sleap_nn_b200 import _native as N
from sleap_nn_b200.inference.ops.crops import crop_bboxes, make_centered_bboxes  # noqa: F401  (API parity)

# Initial per-frame peak capacity of the padded table; grown automatically on overflow.
DEFAULT_PEAK_CAP = int(os.environ.get("SLEAPNN_B200_PEAK_CAP", "1024"))


def _as_f32_cuda(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    if t.device != dev or t.dtype != torch.float32:
        t = t.to(device=dev, dtype=torch.float32)
    return t


_NATIVE_DTYPES = (torch.float32, torch.float16, torch.bfloat16)


def _as_map_cuda(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    """Confidence maps on the compute device IN THEIR OWN dtype when the kernels read it natively (fp32 / fp16 / bf16:
    no up-cast copy of the largest tensor on the hot path); anything else (float64, ints) is converted to fp32."""
    if t.dtype not in _NATIVE_DTYPES:
        return t.to(device=dev, dtype=torch.float32)
    return t if t.device == dev else t.to(device=dev)


def _threshold_for(dtype: torch.dtype, threshold: float) -> float:
    """`cms > threshold` compares in the TENSOR's dtype (the python scalar is cast to it): for fp16 / bf16 maps the
    threshold is first rounded to that dtype.  The up-cast values are exact, so comparing them with the rounded threshold
    in fp32 reproduces the reference's decisions (an element equal to a threshold that rounds UP must not pass)."""
    if dtype in (torch.float16, torch.bfloat16):
        return float(torch.tensor(float(threshold), dtype=dtype))
    return float(threshold)


def _next_pow2(n: int) -> int:
    p = 1
    while p < n:
        p <<= 1
    return p


def local_peaks_padded(cms: torch.Tensor, threshold: float, refine_size: int, xy_scale: float = 1.0,
                       cap: Optional[int] = None):
    """Run K1 and return the padded per-frame peak table (all on device, no host sync).

    Returns (frame_count (B,) i32, xy (B,cap,2) f32, val (B,cap) f32, chan (B,cap) i32,
    status (1,) i32, cap).  `cms` must be a CUDA fp32 / fp16 / bf16 tensor (any strides); half-precision maps are
    read in place and compared / refined on their exact fp32 values.
    """
    B, Cn, H, W = cms.shape
    cap = int(cap or DEFAULT_PEAK_CAP)
    dev = cms.device
    frame_count = torch.empty((max(B, 1),), dtype=torch.int32, device=dev)
    keys = torch.empty((max(B, 1) * cap,), dtype=torch.int32, device=dev)
    xy = torch.empty((B, cap, 2), dtype=torch.float32, device=dev)
    val = torch.empty((B, cap), dtype=torch.float32, device=dev)
    chan = torch.empty((B, cap), dtype=torch.int32, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    sb, sc, sh, sw = cms.stride()
    N.check(
        N.lib.snb_local_peaks_t(N.ptr(cms), N.dtype_code(cms.dtype), B, Cn, H, W, sb, sc, sh, sw, float(threshold),
                                int(refine_size), float(xy_scale), cap, N.ptr(frame_count), N.ptr(keys), N.ptr(xy),
                                N.ptr(val), N.ptr(chan), N.ptr(status), N.stream_ptr(dev)),
        "snb_local_peaks_t",
    )
    return frame_count[:B], xy, val, chan, status, cap


def _local_peaks(cms: torch.Tensor, threshold: float, refine_size: int):
    if cms.dim() != 4:
        raise ValueError(f"cms must be (samples, channels, height, width), got {tuple(cms.shape)}")
    dev = N.compute_device(cms)
    out_dev, out_dtype = cms.device, cms.dtype
    threshold = _threshold_for(cms.dtype, threshold)
    x = _as_map_cuda(cms, dev)
    B = x.shape[0]
    if x.numel() == 0:
        z = torch.zeros
        return (z((0, 2), dtype=torch.float32, device=out_dev), z((0,), dtype=out_dtype, device=out_dev),
                z((0,), dtype=torch.int32, device=out_dev), z((0,), dtype=torch.int32, device=out_dev))
    with torch.cuda.device(dev):
        cap = DEFAULT_PEAK_CAP
        while True:
            frame_count, xy, val, chan, status, cap = local_peaks_padded(x, threshold, refine_size, 1.0, cap)
            counts = frame_count.cpu()  # the one host sync: the result length is data dependent
            worst = int(counts.max()) if B else 0
            if worst <= cap:
                break
            cap = _next_pow2(worst)  # overflow: re-run with room for the busiest frame
        n = int(counts.sum())
        o_xy = torch.empty((n, 2), dtype=torch.float32, device=dev)
        o_val = torch.empty((n,), dtype=torch.float32, device=dev)
        o_s = torch.empty((n,), dtype=torch.int32, device=dev)
        o_c = torch.empty((n,), dtype=torch.int32, device=dev)
        if n:
            N.check(
                N.lib.snb_pack_peaks(N.ptr(frame_count), B, cap, N.ptr(xy), N.ptr(val), N.ptr(chan), N.ptr(o_xy),
                                     N.ptr(o_val), N.ptr(o_s), N.ptr(o_c), N.stream_ptr(dev)),
                "snb_pack_peaks",
            )
    if out_dtype != torch.float32:
        o_val = o_val.to(out_dtype)
    return o_xy.to(out_dev), o_val.to(out_dev), o_s.to(out_dev), o_c.to(out_dev)


def find_local_peaks_rough(
    cms: torch.Tensor, threshold: float = 0.2
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Strict 3x3 local maxima above `threshold`, ordered (sample, y, x, channel).

    Same contract as sleap_nn/inference/ops/peaks.py:184-218: returns
    (peak_points (n,2) f32 x/y, peak_vals (n,), peak_sample_inds (n,) i32, peak_channel_inds (n,) i32).
    """
    return _local_peaks(cms, threshold, 0)


def find_local_peaks(
    cms: torch.Tensor,
    threshold: float = 0.2,
    refinement: Optional[str] = None,
    integral_patch_size: int = 5,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Local peaks with optional integral refinement (sleap_nn/inference/ops/peaks.py:221-259).

    Any `refinement` other than "integral" (including unknown strings) returns rough peaks.
    """
    size = int(integral_patch_size) if refinement == "integral" else 0
    return _local_peaks(cms, threshold, size)


def _global_peaks(cms: torch.Tensor, threshold: float, refine_size: int):
    if cms.dim() != 4:
        raise ValueError(f"cms must be (samples, channels, height, width), got {tuple(cms.shape)}")
    dev = N.compute_device(cms)
    out_dev, out_dtype = cms.device, cms.dtype
    threshold = _threshold_for(cms.dtype, threshold)
    x = _as_map_cuda(cms, dev)
    B, Cn, H, W = x.shape
    if H == 0 or W == 0:
        raise ValueError("find_global_peaks: empty spatial dimensions")
    pts = torch.empty((B, Cn, 2), dtype=torch.float32, device=dev)
    vals = torch.empty((B, Cn), dtype=torch.float32, device=dev)
    if B * Cn:
        with torch.cuda.device(dev):
            rpc, nch, nbytes = C.c_int(), C.c_int(), C.c_longlong()
            N.check(N.lib.snb_global_peaks_workspace(B, Cn, H, W, C.byref(rpc), C.byref(nch), C.byref(nbytes)),
                    "snb_global_peaks_workspace")
            ws = torch.zeros(((nbytes.value + 3) // 4,), dtype=torch.int32, device=dev)
            sb, sc, sh, sw = x.stride()
            N.check(
                N.lib.snb_global_peaks_t(N.ptr(x), N.dtype_code(x.dtype), B, Cn, H, W, sb, sc, sh, sw, float(threshold),
                                         int(refine_size), N.ptr(ws), None, N.ptr(pts), N.ptr(vals), N.stream_ptr(dev)),
                "snb_global_peaks_t",
            )
    if out_dtype != torch.float32:
        vals = vals.to(out_dtype)
    return pts.to(out_dev), vals.to(out_dev)


def find_global_peaks_rough(cms: torch.Tensor, threshold: float = 0.1) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global maximum per (sample, channel); below threshold -> NaN coords, 0 value.

    sleap_nn/inference/ops/peaks.py:89-130.  Returns (points (S,C,2) x/y f32, vals (S,C)).
    """
    return _global_peaks(cms, threshold, 0)


def find_global_peaks(
    cms: torch.Tensor,
    threshold: float = 0.2,
    refinement: Optional[str] = None,
    integral_patch_size: int = 5,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global peaks with optional integral refinement (sleap_nn/inference/ops/peaks.py:133-181)."""
    size = int(integral_patch_size) if refinement == "integral" else 0
    return _global_peaks(cms, threshold, size)


def integral_regression(cms: torch.Tensor, xv: torch.Tensor, yv: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Expected (x, y) under each patch's mass (sleap_nn/inference/ops/peaks.py:66-86).

    cms (samples, channels, h, w); xv (w,), yv (h,).  Returns two (samples, channels) tensors.
    """
    dev = N.compute_device(cms)
    out_dev = cms.device
    p = _as_f32_cuda(cms, dev).contiguous()
    n, c, h, w = p.shape
    xv_d = _as_f32_cuda(xv, dev).contiguous()
    yv_d = _as_f32_cuda(yv, dev).contiguous()
    ox = torch.empty((n, c), dtype=torch.float32, device=dev)
    oy = torch.empty((n, c), dtype=torch.float32, device=dev)
    if h * w == 0:  # empty patches: sum over nothing = 0, 0/0 = NaN (as torch)
        ox.fill_(float("nan")); oy.fill_(float("nan"))
    elif n * c:
        with torch.cuda.device(dev):
            N.check(N.lib.snb_integral_regression(N.ptr(p), n * c, h, w, N.ptr(xv_d), N.ptr(yv_d), N.ptr(ox), N.ptr(oy),
                                                  N.stream_ptr(dev)), "snb_integral_regression")
    return ox.to(out_dev), oy.to(out_dev)


def morphological_dilation(image: torch.Tensor, kernel: torch.Tensor) -> torch.Tensor:
    """Max over the 8-neighbourhood, -inf outside (sleap_nn/inference/ops/peaks.py:26-63).

    `kernel` is accepted and ignored, as in the reference.  image: (B, 1, H, W).
    """
    del kernel
    dev = N.compute_device(image)
    out_dev = image.device
    x = _as_f32_cuda(image, dev).contiguous()
    out = torch.empty_like(x)
    if x.numel():
        H, W = x.shape[-2:]
        with torch.cuda.device(dev):
            N.check(N.lib.snb_dilate8(N.ptr(x), x.numel() // (H * W), H, W, N.ptr(out), N.stream_ptr(dev)), "snb_dilate8")
    return out.to(device=out_dev, dtype=image.dtype)
