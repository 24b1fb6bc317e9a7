"""Device-resident `postprocess()` of the reference's inference layers.

Each class replaces the post-model half of one reference layer with a short launch chain and no host
synchronisation when the output shape is fixed:

    CentroidPostproc          CentroidLayer.postprocess        (inference/layers/centroid.py:196-258)
    CenteredInstancePostproc  CenteredInstanceLayer.postprocess (inference/layers/centered_instance.py:199-230)
                              + TopDownLayer._run_stage_2's un-crop / scatter (inference/layers/topdown.py:259-289)
    BottomUpPostproc          BottomUpLayer.postprocess         (sleap_nn_b200/pipeline.py)
    BottomUpMultiClassPostproc BottomUpMultiClassLayer.postprocess (inference/layers/bottomup_multiclass.py:75-190)

Knob names follow `PostprocessConfig` / `PreprocInfo`; returned tensors live on the device.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from sleap_nn_b200 import _native as N
from sleap_nn_b200.inference.ops.coord import make_ladder
from sleap_nn_b200.inference.ops.peaks import DEFAULT_PEAK_CAP, local_peaks_padded
from sleap_nn_b200.pipeline import BottomUpPostproc  # noqa: F401  (re-export: the third layer)


def _refine_size(refinement: Optional[str], integral_patch_size: int) -> int:
    return int(integral_patch_size) if refinement == "integral" else 0


class CentroidPostproc:
    """confmaps (B, 1, H, W) -> pred_centroids (B, max_instances, 2), pred_centroid_values (B, max_instances).

    find_local_peaks -> x stride -> / input_scale -> per-frame top-k by value when there are more peaks than
    `max_instances` -> NaN padding -> / eff_scale.  Two launches (K1 + the top-k epilogue reuse the padded peak
    table) plus the key sort; with `max_instances=None` the busiest frame's count is read back first
    (`_infer_max_instances`, centroid.py:264-270), which is the only host sync.
    """

    def __init__(self, peak_threshold: float = 0.2, refinement: Optional[str] = "integral", integral_patch_size: int = 5,
                 max_instances: Optional[int] = None, peak_cap: int = DEFAULT_PEAK_CAP):
        self.peak_threshold, self.refine_size = float(peak_threshold), _refine_size(refinement, integral_patch_size)
        self.max_instances, self.peak_cap = max_instances, int(peak_cap)

    def __call__(self, confmaps: torch.Tensor, output_stride: int = 1, input_scale: float = 1.0,
                 eff_scale: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if not confmaps.is_cuda or confmaps.dtype != torch.float32:
            raise TypeError("CentroidPostproc expects an fp32 CUDA tensor")
        dev, B = confmaps.device, int(confmaps.shape[0])
        with torch.cuda.device(dev):
            count, xy, val, _chan, status, cap = local_peaks_padded(confmaps.detach(), self.peak_threshold,
                                                                    self.refine_size, float(output_stride), self.peak_cap)
            max_inst = self.max_instances
            if not max_inst:
                max_inst = int(count.clamp(max=cap).max().item()) if B else 0
            max_inst = max(int(max_inst), 1)  # always at least one slot (centroid.py:222-223)
            o_xy = torch.empty((B, max_inst, 2), dtype=torch.float32, device=dev)
            o_val = torch.empty((B, max_inst), dtype=torch.float32, device=dev)
            eff = None if eff_scale is None else eff_scale.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
            if eff is not None and eff.numel() == 1 and B != 1:
                eff = eff.expand(B).contiguous()
            N.check(N.lib.snb_peaks_topk(N.ptr(count), B, cap, N.ptr(xy), N.ptr(val), max_inst, float(input_scale),
                                         N.ptr(eff), N.ptr(o_xy), N.ptr(o_val), N.stream_ptr(dev)), "snb_peaks_topk")
        self.last_status = status  # SNB_STATUS_PEAK_OVERFLOW if a frame had more than peak_cap peaks
        return o_xy, o_val


class CenteredInstancePostproc:
    """confmaps (n, N, h, w) -> keypoints + values in ONE launch, ladder and un-crop scatter included.

    Stand-alone (CenteredInstanceLayer): returns (n, 1, N, 2), (n, 1, N) with stride / input_scale / eff_scale
    undone.  Inside a top-down pipeline pass `crop_topleft` (n, 2), `per_crop_eff_scale` (n,) and
    `scatter_rows` (n,) = b * max_instances + i of each crop together with `out_shape=(B, max_instances)`:
    the kernel adds the crop offset, divides by the per-crop scale and writes straight into the NaN-filled
    (B, max_instances, N, 2) / (B, max_instances, N) tensors (topdown.py:259-289).
    """

    def __init__(self, peak_threshold: float = 0.2, refinement: Optional[str] = "integral", integral_patch_size: int = 5):
        self.peak_threshold, self.refine_size = float(peak_threshold), _refine_size(refinement, integral_patch_size)
        self._ws = None

    def __call__(self, confmaps: torch.Tensor, output_stride: int = 1, input_scale: float = 1.0,
                 eff_scale: Optional[torch.Tensor] = None, crop_topleft: Optional[torch.Tensor] = None,
                 per_crop_eff_scale: Optional[torch.Tensor] = None, scatter_rows: Optional[torch.Tensor] = None,
                 out_shape: Optional[Tuple[int, int]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if not confmaps.is_cuda or confmaps.dtype != torch.float32:
            raise TypeError("CenteredInstancePostproc expects an fp32 CUDA tensor")
        dev = confmaps.device
        n, Cn, H, W = (int(v) for v in confmaps.shape)
        with torch.cuda.device(dev):
            if scatter_rows is not None:
                if out_shape is None:
                    raise ValueError("scatter_rows needs out_shape=(B, max_instances)")
                rows = int(out_shape[0]) * int(out_shape[1])
                kpts = torch.full((rows, Cn, 2), float("nan"), dtype=torch.float32, device=dev)
                vals = torch.full((rows, Cn), float("nan"), dtype=torch.float32, device=dev)
            else:
                kpts = torch.empty((n, Cn, 2), dtype=torch.float32, device=dev)
                vals = torch.empty((n, Cn), dtype=torch.float32, device=dev)
            if n * Cn:
                rpc, nch, nbytes = C.c_int(), C.c_int(), C.c_longlong()
                N.check(N.lib.snb_global_peaks_workspace(n, Cn, H, W, C.byref(rpc), C.byref(nch), C.byref(nbytes)), "workspace")
                need = (nbytes.value + 3) // 4
                if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                    self._ws = torch.zeros((need,), dtype=torch.int32, device=dev)  # tickets self-reset after each use
                lad, keep = make_ladder(dev, n, stride=float(output_stride), input_scale=float(input_scale),
                                        eff_scale=eff_scale, crop_offset=crop_topleft, eff_scale2=per_crop_eff_scale,
                                        scatter=scatter_rows)
                x = confmaps.detach()
                N.check(N.lib.snb_global_peaks_ex(N.ptr(x), n, Cn, H, W, *x.stride(), self.peak_threshold, self.refine_size,
                                                  N.ptr(self._ws), C.byref(lad), N.ptr(kpts), N.ptr(vals), N.stream_ptr(dev)),
                        "snb_global_peaks_ex")
                del keep
        if scatter_rows is not None:
            return kpts.view(out_shape[0], out_shape[1], Cn, 2), vals.view(out_shape[0], out_shape[1], Cn)
        return kpts.unsqueeze(1), vals.unsqueeze(1)


class BottomUpMultiClassPostproc:
    """confmaps (B, N, H, W) + class maps (B, K, Hc, Wc) -> per-class instances, all on the device.

    `BottomUpMultiClassLayer.postprocess` (inference/layers/bottomup_multiclass.py:75-146): find_local_peaks ->
    x cms_output_stride -> / class_maps_output_stride -> classify_peaks_from_maps (gather under the rounded peak,
    per-(frame, node) optimal assignment, arg-max filter) -> x class_maps_output_stride -> / input_scale ->
    / eff_scale -> nanmean scores -> `_cap_instances_by_score`.  Three launches (K1 + key sort/refine, the
    classification kernel, the epilogue), no host synchronisation; slot k of the outputs IS class k.

    Returns `(pred_keypoints (B, K, N, 2), pred_peak_values (B, K, N), instance_scores (B, K),
    instance_tracking_scores (B, K))`.  Status bits (a NaN class probability, a frame with more than `peak_cap`
    peaks, more than 128 peaks of one node in one frame) are raised by `.check()`, the only call that synchronises.
    """

    def __init__(self, peak_threshold: float = 0.2, refinement: Optional[str] = "integral", integral_patch_size: int = 5,
                 cms_output_stride: int = 2, class_maps_output_stride: int = 2, max_instances: Optional[int] = None,
                 peak_cap: int = DEFAULT_PEAK_CAP):
        self.peak_threshold, self.refine_size = float(peak_threshold), _refine_size(refinement, integral_patch_size)
        self.cms_output_stride, self.class_maps_output_stride = cms_output_stride, class_maps_output_stride
        self.max_instances, self.peak_cap = max_instances, int(peak_cap)
        self._status = []

    def __call__(self, confmaps: torch.Tensor, class_maps: torch.Tensor, input_scale: float = 1.0,
                 eff_scale: Optional[torch.Tensor] = None):
        if not (confmaps.is_cuda and class_maps.is_cuda) or confmaps.dtype != torch.float32 or class_maps.dtype != torch.float32:
            raise TypeError("BottomUpMultiClassPostproc expects fp32 CUDA tensors")
        dev = confmaps.device
        B, Nn = int(confmaps.shape[0]), int(confmaps.shape[1])
        if int(class_maps.shape[0]) != B:
            raise ValueError("confmaps and class maps must hold the same frames")
        K, Hc, Wc = (int(v) for v in class_maps.shape[1:])
        with torch.cuda.device(dev):
            count, xy, val, chan, status, cap = local_peaks_padded(confmaps.detach(), self.peak_threshold, self.refine_size,
                                                                   float(self.cms_output_stride), self.peak_cap)
            f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
            probs = f32(max(B * cap, 1), max(K, 1))
            c_xy, c_val, c_prob = f32(B, K, Nn, 2), f32(B, K, Nn), f32(B, K, Nn)
            cm = class_maps.detach()
            st = N.stream_ptr(dev)
            N.check(N.lib.snb_classify_peaks_padded(N.ptr(cm), B, K, Hc, Wc, *cm.stride(), N.ptr(count), cap, N.ptr(xy),
                                                    N.ptr(val), N.ptr(chan), float(self.class_maps_output_stride), Nn,
                                                    N.ptr(probs), N.ptr(c_xy), N.ptr(c_val), N.ptr(c_prob), N.ptr(status),
                                                    st), "snb_classify_peaks_padded")
            eff = None if eff_scale is None else eff_scale.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
            if eff is not None and eff.numel() == 1 and B != 1:
                eff = eff.expand(B).contiguous()
            if eff is not None and eff.numel() != B:
                raise ValueError("eff_scale must hold one factor per frame")
            kpts, vals, scores, tracking = f32(B, K, Nn, 2), f32(B, K, Nn), f32(B, K), f32(B, K)
            N.check(N.lib.snb_multiclass_outputs(N.ptr(c_xy), N.ptr(c_val), N.ptr(c_prob), B, K, Nn,
                                                 float(self.class_maps_output_stride), float(input_scale), N.ptr(eff),
                                                 -1 if self.max_instances is None else int(self.max_instances),
                                                 N.ptr(kpts), N.ptr(vals), N.ptr(scores), N.ptr(tracking), st),
                    "snb_multiclass_outputs")
        self._status.append(status)
        del self._status[:-8]
        return kpts, vals, scores, tracking

    def check(self) -> None:
        """Synchronise and raise if any recent call overflowed a table or met an invalid class probability."""
        bits = 0
        for s in self._status:
            bits |= int(s.item())
        self._status.clear()
        if bits & N.STATUS_LSAP_INVALID:
            raise ValueError("matrix contains invalid numeric entries")
        if bits & N.STATUS_LSAP_INFEASIBLE:
            raise ValueError("cost matrix is infeasible")
        if bits:
            raise RuntimeError(f"multi-class post-processing overflowed a fixed-capacity table (status 0x{bits:x})")
