"""Post-inference filtering - `FilterConfig` + `FilterPipeline` with the API of sleap_nn/inference/filters.py,
computed by one CUDA kernel (csrc/filters.cu) on the padded outputs of the grouping kernels.

The reference applies up to five filters as separate passes of tensor ops and per-frame python loops with
`.item()` synchronisations (filters.py:100-163); here the whole pipeline is ONE launch with one warp per frame and
no host round trip.  The pipeline accepts any object carrying the `Outputs` fields it reads (`pred_keypoints`,
`pred_peak_values`, `instance_scores`, `pred_centroids`, `pred_centroid_values`): the reference's attrs `Outputs`,
or a plain namespace such as `FilterableOutputs` below.
"""

from __future__ import annotations

import copy
import ctypes as C
import warnings
from typing import Literal, Optional

import attrs
import torch

from sleap_nn_b200 import _native as N


@attrs.frozen
class FilterConfig:
    """Post-inference filter configuration (value type, picklable); defaults are the no-op identity.

    Same fields, defaults and meaning as sleap_nn/inference/filters.py:41-86.
    """

    min_peak_value: float = 0.0
    min_instance_score: float = 0.0
    min_mean_node_score: float = 0.0
    min_visible_nodes: int = 0
    min_visible_node_fraction: float = 0.0
    overlapping: bool = False
    overlapping_threshold: float = 0.8
    overlapping_method: Literal["iou", "oks"] = "iou"
    min_centroid_distance: float = 0.0


@attrs.define
class FilterableOutputs:
    """The five `Outputs` fields the filters read, for callers that do not carry the reference's `Outputs`."""

    pred_keypoints: Optional[torch.Tensor] = None       # (B, I, N, 2)
    pred_peak_values: Optional[torch.Tensor] = None     # (B, I, N)
    instance_scores: Optional[torch.Tensor] = None      # (B, I)
    pred_centroids: Optional[torch.Tensor] = None       # (B, I, 2)
    pred_centroid_values: Optional[torch.Tensor] = None  # (B, I)


_FIELDS = ("pred_keypoints", "pred_peak_values", "instance_scores", "pred_centroids", "pred_centroid_values")


def _evolve(outputs, **changes):
    if attrs.has(type(outputs)):
        return attrs.evolve(outputs, **changes)
    new = copy.copy(outputs)
    for k, v in changes.items():
        setattr(new, k, v)
    return new


@attrs.define
class FilterPipeline:
    """Apply a `FilterConfig` to an `Outputs`-like object (filters.py:88-163); `__call__` is sugar for `apply`."""

    config: FilterConfig

    def __call__(self, outputs):
        return self.apply(outputs)

    @classmethod
    def run(cls, outputs, config: FilterConfig):
        return cls(config=config)(outputs)

    @staticmethod
    def _pair_similarity(a: torch.Tensor, b: torch.Tensor, kappa: float):
        """(IoU, OKS) of two keypoint sets (N, 2), computed on the device in the tensors' dtype (fp32 or float64)."""
        dev = N.compute_device(a, b)
        f64 = a.dtype == torch.float64
        dt = torch.float64 if f64 else torch.float32
        a_d, b_d = (t.detach().to(device=dev, dtype=dt).reshape(-1, 2).contiguous() for t in (a, b))
        out = torch.empty((2,), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            N.check(N.lib.snb_pair_similarity(N.ptr(a_d), N.ptr(b_d), int(a_d.shape[0]), int(f64), float(kappa), N.ptr(out),
                                              N.stream_ptr(dev)), "snb_pair_similarity")
        iou, oks = out.tolist()
        return iou, oks

    @staticmethod
    def _bbox_iou(a: torch.Tensor, b: torch.Tensor) -> float:
        """IoU of the boxes implied by two keypoint sets, NaN-aware (filters.py:290-307)."""
        return FilterPipeline._pair_similarity(a, b, 0.1)[0]

    @staticmethod
    def _oks(a: torch.Tensor, b: torch.Tensor, kappa: float = 0.1) -> float:
        """Object-keypoint similarity with the scale taken from a's own box area (filters.py:309-338)."""
        return FilterPipeline._pair_similarity(a, b, kappa)[1]

    def apply(self, outputs):
        """Run all configured filters in the reference's cheap -> expensive order, in one kernel launch."""
        cfg = self.config
        kpts = getattr(outputs, "pred_keypoints", None)
        cen = getattr(outputs, "pred_centroids", None)
        overlapping = 0
        if cfg.overlapping:
            if kpts is None and cen is not None:
                warnings.warn(
                    "overlapping NMS (iou/oks) is not meaningful for centroid-only outputs (single points have "
                    "degenerate bbox-IoU / OKS); use FilterConfig.min_centroid_distance for centroid de-duplication. "
                    "Skipping overlap NMS.", stacklevel=2)
            else:
                method = cfg.overlapping_method
                if method == "oks" and kpts is not None and kpts.shape[-2] < 2:
                    warnings.warn("OKS overlap NMS is degenerate for single-node (centroid) keypoints; falling back "
                                  "to IoU.", stacklevel=2)
                    method = "iou"
                if method not in ("iou", "oks"):
                    raise ValueError(f"Unknown method: {method}. Use 'iou' or 'oks'.")
                overlapping = 2 if method == "oks" else 1
        active = (cfg.min_peak_value > 0.0 or cfg.min_visible_nodes > 0 or cfg.min_visible_node_fraction > 0.0
                  or cfg.min_instance_score > 0.0 or cfg.min_mean_node_score > 0.0 or overlapping
                  or cfg.min_centroid_distance > 0.0)
        if not active or (kpts is None and cen is None):
            return outputs
        ref = kpts if kpts is not None else cen
        out_dev = ref.device
        dev = N.compute_device(ref)
        B, I = int(ref.shape[0]), int(ref.shape[1])
        Nn = int(kpts.shape[2]) if kpts is not None else 0
        ins, outs = {}, {}
        for name in _FIELDS:
            t = getattr(outputs, name, None)
            if t is None:
                ins[name] = outs[name] = None
                continue
            ins[name] = t.detach().to(device=dev, dtype=torch.float32).contiguous()
            outs[name] = torch.empty_like(ins[name])
        # Without keypoints the keypoint-only stages are skipped by the kernel; peak values alone are passed through.
        st = N.FilterConfigStruct(float(cfg.min_peak_value), float(cfg.min_visible_node_fraction),
                                  float(cfg.min_instance_score), float(cfg.min_mean_node_score), float(0.1**2),
                                  int(cfg.min_visible_nodes), int(overlapping), float(cfg.overlapping_threshold),
                                  float(cfg.min_centroid_distance) ** 2)
        if B * I:
            with torch.cuda.device(dev):
                N.check(N.lib.snb_filter_instances(C.byref(st), B, I, Nn, *(N.ptr(ins[k]) for k in _FIELDS),
                                                   *(N.ptr(outs[k]) for k in _FIELDS), N.stream_ptr(dev)),
                        "snb_filter_instances")
        changes = {}
        for name in _FIELDS:
            if outs[name] is None:
                continue
            src = getattr(outputs, name)
            if src.dtype == torch.float32:
                changes[name] = outs[name].to(out_dev)
            else:
                # the kernel decides in fp32 (the dtype every inference layer emits); values of another dtype are
                # never rounded through it: the NaN pattern is applied to a copy of the original tensor
                kept = src.detach().clone().to(torch.float64 if not src.dtype.is_floating_point else src.dtype)
                kept[torch.isnan(outs[name]).to(out_dev)] = float("nan")
                changes[name] = kept
        return _evolve(outputs, **changes)
