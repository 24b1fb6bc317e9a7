"""Re-export shim, same names as sleap_nn/inference/peak_finding.py:9-27."""

from sleap_nn_b200.inference.ops.crops import crop_bboxes
from sleap_nn_b200.inference.ops.peaks import (
    find_global_peaks,
    find_global_peaks_rough,
    find_local_peaks,
    find_local_peaks_rough,
    integral_regression,
    morphological_dilation,
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
