"""`sleap_nn.inference.peak_finding`'s import path (inference/peak_finding.py:9-27): the names callers import from there,
bound to the CUDA-backed implementations; the table names the C-ABI entry point behind each one."""

from sleap_nn_b200.inference.ops import crops as _crops
from sleap_nn_b200.inference.ops import peaks as _peaks

_BACKED_BY = {
    "find_local_peaks": (_peaks, "snb_local_peaks + snb_pack_peaks"),        # K1 + K3
    "find_local_peaks_rough": (_peaks, "snb_local_peaks + snb_pack_peaks"),
    "find_global_peaks": (_peaks, "snb_global_peaks"),                        # K2 + K3
    "find_global_peaks_rough": (_peaks, "snb_global_peaks"),
    "integral_regression": (_peaks, "snb_integral_regression"),
    "morphological_dilation": (_peaks, "snb_dilate8"),
    "crop_bboxes": (_crops, "snb_crop_bboxes"),
}
globals().update({name: getattr(mod, name) for name, (mod, _entry) in _BACKED_BY.items()})
__all__ = sorted(_BACKED_BY)
