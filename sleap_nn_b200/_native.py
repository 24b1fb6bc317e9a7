"""ctypes binding of libsleapnn_b200.so (the C ABI declared in include/sleapnn_b200.h).

There is NO fallback: if the shared library is missing or a symbol is absent, importing
this module raises.  torch is used only for device memory, streams and dtype plumbing.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SLEAPNN_B200_LIB", os.path.join(_HERE, "lib", "libsleapnn_b200.so"))

EXPECTED_ABI = 5  # include/sleapnn_b200.h SNB_ABI_VERSION this binding was written against
DTYPE_F32, DTYPE_F16, DTYPE_BF16 = 0, 1, 2

OK = 0
ERRORS = {-1: "bad argument", -2: "unsupported configuration", -3: "CUDA launch failed"}

STATUS_PEAK_OVERFLOW = 1
STATUS_CAND_OVERFLOW = 2
STATUS_LSAP_INFEASIBLE = 4
STATUS_LSAP_TOO_LARGE = 8
STATUS_INSTANCE_OVERFLOW = 16
STATUS_BAD_INDEX = 32
STATUS_MATCH_OVERFLOW = 64
STATUS_LSAP_INVALID = 128
STATUS_ASM_MISMATCH = 256
STATUS_ASM_MISSING = 512

_i, _ll, _f, _p = C.c_int, C.c_longlong, C.c_float, C.c_void_p
_ip, _llp = C.POINTER(C.c_int), C.POINTER(C.c_longlong)

# name -> argtypes; every function returns int.  Keep in the order of include/sleapnn_b200.h.
SIGNATURES = {
    "snb_abi_version": [],
    "snb_host_device_pointer": [_p, C.POINTER(_p)],
    "snb_local_peaks": [_p, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _f, _i, _f, _i, _p, _p, _p, _p, _p, _p, _p],
    "snb_local_peaks_detect": [_p, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _f, _i, _p, _p, _p, _p, _p],
    "snb_local_peaks_finalize": [_p, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _i, _f, _i, _p, _p, _p, _p, _p, _p, _p],
    "snb_local_peaks_t": [_p, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _f, _i, _f, _i, _p, _p, _p, _p, _p, _p, _p],
    "snb_local_peaks_detect_t": [_p, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _f, _i, _p, _p, _p, _p, _p],
    "snb_local_peaks_finalize_t": [_p, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _i, _f, _i, _p, _p, _p, _p, _p, _p, _p],
    "snb_pack_peaks": [_p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    "snb_global_peaks_workspace": [_i, _i, _i, _i, _ip, _ip, _llp],
    "snb_global_peaks": [_p, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _f, _i, _p, _p, _p, _p],
    "snb_global_peaks_ex": [_p, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _f, _i, _p, _p, _p, _p, _p],
    "snb_global_peaks_t": [_p, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _f, _i, _p, _p, _p, _p, _p],
    "snb_peaks_topk": [_p, _i, _i, _p, _p, _i, _f, _p, _p, _p, _p],
    "snb_coord_ladder_apply": [_p, _ll, _ll, _p, _p, _p],
    "snb_bilinear_resize": [_p, _i, _ll, _i, _i, _ll, _ll, _ll, _i, _i, _p, _p],
    "snb_crop_bboxes": [_p, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _p, _p, _ll, _i, _i, _p, _p, _p],
    "snb_centered_bboxes": [_p, _ll, _f, _f, _p, _p],
    "snb_topdown_select": [_p, _p, _i, _i, _p, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "snb_topdown_lift": [_p, _p, _i, _ll, _p, _p, _p, _p, _p, _p, _p],
    "snb_integral_regression": [_p, _ll, _i, _i, _p, _p, _p, _p, _p],
    "snb_dilate8": [_p, _ll, _i, _i, _p, _p],
    "snb_paf_prepare": [_p, _p, _i, _p, _i, _p, _i, _i, _p, _p, _p, _p, _p],
    "snb_paf_score": [_p, _ll, _ll, _ll, _ll, _i, _i, _p, _i, _f, _f, _f, _p, _p, _i, _i, _p, _i, _i, _p, _p, _p, _p,
                      _i, _i, _p, _p, _p, _p, _p],
    "snb_paf_score_t": [_p, _i, _ll, _ll, _ll, _ll, _i, _i, _p, _i, _f, _f, _f, _p, _p, _i, _i, _p, _i, _i, _p, _p, _p, _p,
                        _i, _i, _p, _p, _p, _p, _p],
    "snb_line_subs": [_p, _ll, _p, _p, _ll, _p, _i, _f, _i, _i, _p, _p, _p],
    "snb_paf_gather": [_p, _ll, _ll, _ll, _i, _i, _i, _p, _ll, _p, _p, _p],
    "snb_score_lines": [_p, _p, _ll, _p, _ll, _i, _f, _f, _p, _p, _p],
    "snb_distance_penalty": [_p, _ll, _f, _f, _p, _p],
    "snb_match_structured": [_p, _p, _i, _p, _i, _i, _p, _p, _p, _p, _i, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p],
    "snb_match_generic": [_i, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p],
    "snb_assemble": [_p, _p, _p, _p, _i, _p, _i, _p, _p, _i, _p, _p, _i, _p, _p, _p, _p, _p, _i, _p, _i, _f, _p, _i,
                     _i, _p, _p, _p, _p, _p, _p],
    "snb_scatter_instances": [_p, _p, _p, _p, _i, _p, _p, _i, _i, _i, _p, _p, _p, _p],
    "snb_interp1d": [_p, _i, _p, _i, _p, _i, _i, _i, _ll, _p, _p],
    "snb_confmaps": [_p, _i, _i, _i, _p, _p, _i, _i, _f, _i, _p, _p],
    "snb_pafs": [_p, _p, _i, _i, _i, _p, _p, _i, _i, _f, _i, _i, _p, _p],
    "snb_confmaps_ex": [_p, _i, _i, _i, _ll, _ll, _ll, _p, _f, _f, _p, _p, _i, _i, _f, _i, _p, _p],
    "snb_pafs_from_instances": [_p, _i, _i, _i, _p, _i, _f, _f, _p, _p, _i, _i, _f, _i, _p, _p],
    "snb_bottomup_targets": [_p, _i, _i, _i, _p, _f, _f, _p, _i, _f, _f, _p, _p, _i, _i, _f, _p, _p, _i, _i, _f, _i, _p, _p,
                             _p],
    "snb_debug_neg_div": [_p, _ll, _f, _p, _p, _p],
    "snb_edge_distance": [_p, _p, _p, _i, _ll, _p, _p, _i, _i, _f, _p, _p],
    "snb_gaussian_pdf": [_p, _ll, _f, _p, _p],
    "snb_classify_peaks": [_p, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _p, _p, _p, _p, _ll, _i, _p, _p, _p, _p, _p, _p, _p,
                           _p, _p],
    "snb_classify_peaks_padded": [_p, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _p, _i, _p, _p, _p, _f, _i, _p, _p, _p, _p, _p, _p],
    "snb_multiclass_outputs": [_p, _p, _p, _i, _i, _i, _f, _f, _p, _i, _p, _p, _p, _p, _p],
    "snb_pack_class_matches": [_p, _p, _p, _i, _i, _p, _p, _p, _p],
    "snb_class_inds_from_vectors": [_p, _i, _i, _p, _p, _p, _p, _p],
    "snb_class_inds_grouped": [_p, _i, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p],
    "snb_class_vectors": [_p, _i, _i, _i, _p, _p, _p],
    "snb_class_maps": [_p, _p, _p, _i, _i, _i, _i, _i, _f, _p, _p],
    "snb_pair_similarity": [_p, _p, _i, _i, C.c_double, _p, _p],
    "snb_filter_instances": [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "snb_nms_greedy_f64": [_p, _p, _p, _i, _i, _i, _i, C.c_double, C.c_double, _p, _p, _p],
    "snb_instance_stats_f64": [_p, _p, _ll, _i, _p, _p, _p],
    "snb_bottomup_postproc": [_p, _p],
    "snb_bottomup_launches_per_call": [_p],
    "snb_bottomup_args_size": [],
    "snb_bottomup_outputs": [_p, _p, _p, _p, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p],
    "snb_pack_instances": [_p, _i, _i, _i, _p, _p, _p, _i, _p, _ll, _p, _p, _p, _p, _p, _p, _p],
}
RETURNS_LONGLONG = {"snb_lsap_workspace_bytes": [_i], "snb_class_inds_workspace_bytes": [_i, _i],
                   "snb_class_inds_grouped_workspace_bytes": [_i, _i, _i], "snb_topdown_select_smem_bytes": [_i, _i], "snb_bottomup_tail_smem_bytes": [_i, _i, _i, _i, _i, _i, _i]}
FLAG_UNFUSED_TAIL = 1
FLAG_SELF_RESET_COUNTERS = 2
FLAG_NO_TAIL_CLUSTER = 4


class BottomUpArgs(C.Structure):
    """Mirror of `snb_bottomup_args` (include/sleapnn_b200.h); field order must match."""

    _fields_ = [
        ("cms", _p), ("B", _i), ("C", _i), ("H", _i), ("W", _i),
        ("cms_sb", _ll), ("cms_sc", _ll), ("cms_sh", _ll), ("cms_sw", _ll),
        ("pafs", _p), ("paf_H", _i), ("paf_W", _i),
        ("paf_sb", _ll), ("paf_sy", _ll), ("paf_sx", _ll), ("paf_sc", _ll),
        ("edges", _p), ("n_edges", _i), ("sorted_edges", _p), ("n_sorted", _i),
        ("t_table", _p), ("n_points", _i),
        ("peak_threshold", _f), ("refine_size", _i), ("cms_stride", _f), ("pafs_stride", _f),
        ("max_edge_length", _f), ("dist_penalty_weight", _f), ("min_instance_peaks", _i), ("min_line_scores", _f),
        ("peak_cap", _i), ("cand_cap", _i), ("match_cap", _i), ("inst_cap", _i), ("lsap_max_dim", _i),
        ("frame_count", _p), ("keys", _p), ("peak_xy", _p), ("peak_val", _p), ("peak_chan", _p),
        ("node_start", _p), ("node_peaks", _p), ("edge_off", _p), ("match_off", _p),
        ("cand_edge", _p), ("cand_epi", _p), ("cand_score", _p),
        ("m_edge", _p), ("m_src", _p), ("m_dst", _p), ("m_score", _p), ("m_count", _p),
        ("lsap_ws", _p), ("asm_ws", _p),
        ("inst_xy", _p), ("inst_val", _p), ("inst_score", _p), ("n_inst", _p), ("status", _p),
        ("ev_detect_begin", _p), ("ev_detect_end", _p),
        ("tail_stream", _p), ("ev_handoff", _p), ("ev_tail_done", _p), ("flags", _i),
        ("max_peaks_per_node", _i), ("skip_flag", _p), ("max_instances", _i), ("input_scale", _f),
        ("eff_scale", _p), ("out_kpts", _p), ("out_vals", _p), ("out_scores", _p),
        ("cms_dtype", _i), ("pafs_dtype", _i), ("n_peaks", _p),
    ]


class CoordLadder(C.Structure):
    """Mirror of `snb_coord_ladder` (include/sleapnn_b200.h)."""

    _fields_ = [("stride", _f), ("input_scale", _f), ("eff_scale", _p), ("crop_offset", _p), ("eff_scale2", _p),
                ("scatter", _p)]


class FilterConfigStruct(C.Structure):
    """Mirror of `snb_filter_config` (include/sleapnn_b200.h)."""

    _fields_ = [("min_peak_value", _f), ("min_visible_node_fraction", _f), ("min_instance_score", _f),
                ("min_mean_node_score", _f), ("oks_kappa_sq", _f), ("min_visible_nodes", _i), ("overlapping", _i),
                ("overlapping_threshold", C.c_double), ("min_centroid_distance_sq", C.c_double)]


class NativeLibraryError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not os.path.isfile(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or sleap_nn_b200/csrc/build.sh. There is no CPU / PyTorch fallback for this path."
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.argtypes = argtypes
        fn.restype = C.c_int
    for name, argtypes in RETURNS_LONGLONG.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.argtypes = argtypes
        fn.restype = C.c_longlong
    return lib


lib = _load()
ABI_VERSION = lib.snb_abi_version()
if ABI_VERSION != EXPECTED_ABI:  # a stale / foreign .so called with these argtypes would corrupt memory silently
    raise NativeLibraryError(f"{LIB_PATH} reports ABI v{ABI_VERSION}, this binding needs v{EXPECTED_ABI}: rebuild it "
                             "(sleap_nn_b200/csrc/build.sh)")
if lib.snb_bottomup_args_size() != C.sizeof(BottomUpArgs):
    raise NativeLibraryError(
        f"snb_bottomup_args layout mismatch: library {lib.snb_bottomup_args_size()} bytes, binding {C.sizeof(BottomUpArgs)}")


def check(rc: int, what: str) -> None:
    if rc != OK:
        raise RuntimeError(f"{what}: {ERRORS.get(rc, f'error {rc}')}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


_DTYPES = {torch.float32: DTYPE_F32, torch.float16: DTYPE_F16, torch.bfloat16: DTYPE_BF16}


def dtype_code(dtype: torch.dtype) -> int:
    """SNB_DTYPE_* of a map tensor the kernels read natively (fp32 / fp16 / bf16)."""
    try:
        return _DTYPES[dtype]
    except KeyError:
        raise TypeError(f"confidence maps / PAFs must be float32, float16 or bfloat16, got {dtype}") from None


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def compute_device(*tensors) -> torch.device:
    """The CUDA device kernels run on: the first CUDA input's device, else the current device.

    CPU tensors are accepted by the public API (the reference's own tests pass them) and are
    staged to this device; compute never happens on the host.
    """
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise NativeLibraryError(
            "sleap_nn_b200 needs a CUDA device: every op runs in hand-written sm_100a kernels and there is no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())
