"""`sleap_nn.inference.peak_finding`'s import path (inference/peak_finding.py:9-27): the names callers import from
there, re-exported from the CUDA-backed implementations."""

from sleap_nn_b200.inference.ops.crops import crop_bboxes  # snb_crop_bboxes
from sleap_nn_b200.inference.ops.peaks import (
    find_global_peaks,        # K2 + K3: snb_global_peaks_t
    find_global_peaks_rough,
    find_local_peaks,         # K1 + K3: snb_local_peaks_t + snb_pack_peaks
    find_local_peaks_rough,
    integral_regression,      # snb_integral_regression
    morphological_dilation,   # snb_dilate8
)

__all__ = [
    "crop_bboxes",
    "find_global_peaks",
    "find_global_peaks_rough",
    "find_local_peaks",
    "find_local_peaks_rough",
    "integral_regression",
    "morphological_dilation",
]
