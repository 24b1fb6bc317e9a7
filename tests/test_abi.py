"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/sleapnn_b200.h declares with the declared arity, and the Python API mirrors the
reference's module layout and signatures.  No compute calls (there is no GPU here)."""

import ctypes
import inspect
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sleapnn_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|long long)\s+(snb_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        out[m.group(2)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_library_exports_every_declared_symbol():
    from sleap_nn_b200 import _native as N

    decl = _declared()
    assert len(decl) >= 25
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    table = dict(N.SIGNATURES)
    table.update(N.RETURNS_LONGLONG)
    assert set(table) == set(decl), set(table) ^ set(decl)
    for name, argtypes in table.items():
        assert len(argtypes) == decl[name], f"{name}: ctypes arity {len(argtypes)} != header {decl[name]}"
    assert N.ABI_VERSION == N.EXPECTED_ABI == 5


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sleap_nn_b200 import _native as N
    from sleap_nn_b200.inference import peak_finding

    with pytest.raises(N.NativeLibraryError):
        peak_finding.find_local_peaks(torch.zeros(1, 1, 8, 8))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sleap_nn_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle"


REFERENCE_SIGNATURES = {
    "sleap_nn_b200.inference.peak_finding": {
        "find_local_peaks": ["cms", "threshold", "refinement", "integral_patch_size"],
        "find_local_peaks_rough": ["cms", "threshold"],
        "find_global_peaks": ["cms", "threshold", "refinement", "integral_patch_size"],
        "find_global_peaks_rough": ["cms", "threshold"],
        "integral_regression": ["cms", "xv", "yv"],
        "morphological_dilation": ["image", "kernel"],
        "crop_bboxes": ["images", "bboxes", "sample_inds"],
    },
    "sleap_nn_b200.inference.paf_grouping": {
        "get_connection_candidates": ["peak_channel_inds_sample", "skeleton_edges", "n_nodes"],
        "make_line_subs": ["peaks_sample", "edge_peak_inds", "edge_inds", "n_line_points", "pafs_stride", "pafs_hw"],
        "get_paf_lines": ["pafs_sample", "peaks_sample", "edge_peak_inds", "edge_inds", "n_line_points", "pafs_stride"],
        "compute_distance_penalty": ["spatial_vec_lengths", "max_edge_length", "dist_penalty_weight"],
        "score_paf_lines": ["paf_lines_sample", "peaks_sample", "edge_peak_inds_sample", "max_edge_length",
                            "dist_penalty_weight"],
        "score_paf_lines_batch": ["pafs", "peaks", "peak_channel_inds", "skeleton_edges", "n_line_points",
                                  "pafs_stride", "max_edge_length_ratio", "dist_penalty_weight", "n_nodes"],
        "match_candidates_sample": ["edge_inds_sample", "edge_peak_inds_sample", "line_scores_sample", "n_edges"],
        "match_candidates_batch": ["edge_inds", "edge_peak_inds", "line_scores", "n_edges"],
        "assign_connections_to_instances": ["connections", "min_instance_peaks", "n_nodes"],
        "make_predicted_instances": ["peaks", "peak_scores", "connections", "instance_assignments"],
        "toposort_edges": ["edge_types"],
        "group_instances_sample": ["peaks_sample", "peak_scores_sample", "peak_channel_inds_sample",
                                   "match_edge_inds_sample", "match_src_peak_inds_sample",
                                   "match_dst_peak_inds_sample", "match_line_scores_sample", "n_nodes",
                                   "sorted_edge_inds", "edge_types", "min_instance_peaks", "min_line_scores"],
        "group_instances_batch": ["peaks", "peak_vals", "peak_channel_inds", "match_edge_inds", "match_src_peak_inds",
                                  "match_dst_peak_inds", "match_line_scores", "n_nodes", "sorted_edge_inds",
                                  "edge_types", "min_instance_peaks", "min_line_scores"],
    },
    "sleap_nn_b200.data.confidence_maps": {
        "make_confmaps": ["points_batch", "xv", "yv", "sigma"],
        "make_multi_confmaps": ["points_batch", "xv", "yv", "sigma"],
        "generate_confmaps": ["instance", "img_hw", "sigma", "output_stride"],
        "generate_multiconfmaps": ["instances", "img_hw", "num_instances", "sigma", "output_stride", "is_centroids"],
    },
    "sleap_nn_b200.data.edge_maps": {
        "distance_to_edge": ["points", "edge_source", "edge_destination"],
        "make_edge_maps": ["xv", "yv", "edge_source", "edge_destination", "sigma"],
        "make_pafs": ["xv", "yv", "edge_source", "edge_destination", "sigma"],
        "make_multi_pafs": ["xv", "yv", "edge_sources", "edge_destinations", "sigma"],
        "get_edge_points": ["instances", "edge_inds"],
        "generate_pafs": ["instances", "img_hw", "sigma", "output_stride", "edge_inds", "flatten_channels"],
    },
    "sleap_nn_b200.data.utils": {
        "make_grid_vectors": ["image_height", "image_width", "output_stride"],
        "expand_to_rank": ["x", "target_rank", "prepend"],
        "gaussian_pdf": ["x", "sigma"],
    },
    "sleap_nn_b200.inference.utils": {"interp1d": ["x", "y", "xnew"]},
    "sleap_nn_b200.inference.ops.identity": {
        "group_class_peaks": ["peak_class_probs", "peak_sample_inds", "peak_channel_inds", "n_samples", "n_channels"],
        "classify_peaks_from_maps": ["class_maps", "peak_points", "peak_vals", "peak_sample_inds", "peak_channel_inds",
                                     "n_channels"],
        "get_class_inds_from_vectors": ["peak_class_probs"],
    },
    "sleap_nn_b200.data.identity": {
        "make_class_vectors": ["class_inds", "n_classes"],
        "make_class_maps": ["confmaps", "class_inds", "n_classes", "threshold"],
        "generate_class_maps": ["instances", "img_hw", "num_instances", "class_inds", "num_tracks", "class_map_threshold",
                                "sigma", "output_stride", "is_centroids"],
    },
    "sleap_nn_b200.data.instance_cropping": {"make_centered_bboxes": ["centroids", "box_height", "box_width"]},
    "sleap_nn_b200.inference.ops.coord": {
        "undo_stride": ["coords", "output_stride"],
        "undo_input_scale": ["coords", "input_scale"],
        "undo_eff_scale": ["coords", "eff_scale"],
        "add_crop_offset": ["peaks", "crop_topleft"],
        "apply_input_scale": ["image", "input_scale"],
    },
    "sleap_nn_b200.inference.streaming": {"group_scored_batch": ["scored", "params"]},
}


def test_api_mirrors_reference_signatures():
    import importlib

    for modname, fns in REFERENCE_SIGNATURES.items():
        mod = importlib.import_module(modname)
        for fn, params in fns.items():
            got = list(inspect.signature(getattr(mod, fn)).parameters)
            assert got[: len(params)] == params, f"{modname}.{fn}: {got} vs reference {params}"
    from sleap_nn_b200.inference import paf_grouping as pg

    s = pg.PAFScorer(part_names=["a", "b", "c"], edges=[("a", "b"), ("b", "c")], pafs_stride=2)
    assert (s.max_edge_length_ratio, s.dist_penalty_weight, s.n_points, s.min_instance_peaks, s.min_line_scores) == \
        (0.25, 1.0, 10, 0, 0.25)
    assert s.edge_inds == [(0, 1), (1, 2)] and s.n_nodes == 3 and s.n_edges == 2 and s.sorted_edge_inds == (0, 1)
    assert s.edge_types == [pg.EdgeType(0, 1), pg.EdgeType(1, 2)]
    for m in ("from_config", "score_paf_lines", "match_candidates", "group_instances", "predict"):
        assert callable(getattr(pg.PAFScorer, m))


def test_filter_config_mirrors_the_reference_when_present():
    import attrs

    from oracle import ref_loader
    from sleap_nn_b200.inference import filters as ours

    assert [a.name for a in attrs.fields(ours.FilterConfig)] == [
        "min_peak_value", "min_instance_score", "min_mean_node_score", "min_visible_nodes", "min_visible_node_fraction",
        "overlapping", "overlapping_threshold", "overlapping_method", "min_centroid_distance"]
    if not ref_loader.available():
        pytest.skip("reference tree not present on this box")
    theirs = ref_loader.ref().filters.FilterConfig
    assert [(a.name, a.default) for a in attrs.fields(ours.FilterConfig)] == [(a.name, a.default) for a in attrs.fields(theirs)]
    for m in ("apply", "run", "__call__"):
        assert callable(getattr(ours.FilterPipeline, m))


def test_reference_signatures_table_matches_the_reference_when_present():
    """In the build container, check REFERENCE_SIGNATURES against the real reference modules."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference tree not present on this box")
    R = ref_loader.ref()
    mods = {
        "sleap_nn_b200.inference.peak_finding": [R.peaks, R.crops],
        "sleap_nn_b200.inference.paf_grouping": [R.paf],
        "sleap_nn_b200.data.confidence_maps": [R.confidence_maps],
        "sleap_nn_b200.data.edge_maps": [R.edge_maps],
        "sleap_nn_b200.data.utils": [R.data_utils],
        "sleap_nn_b200.inference.utils": [R.interp],
        "sleap_nn_b200.inference.ops.identity": [R.identity],
        "sleap_nn_b200.data.identity": [R.data_identity],
        "sleap_nn_b200.data.instance_cropping": [R.instance_cropping],
        "sleap_nn_b200.inference.ops.coord": [R.coord],
        "sleap_nn_b200.inference.streaming": [R.streaming],
    }
    for modname, fns in REFERENCE_SIGNATURES.items():
        for fn, params in fns.items():
            ref_fn = next(getattr(m, fn) for m in mods[modname] if hasattr(m, fn))
            assert list(inspect.signature(ref_fn).parameters) == params, fn


def test_streaming_value_types_mirror_the_reference_when_present():
    import attrs

    from oracle import ref_loader
    from sleap_nn_b200.inference import streaming as ours

    assert [a.name for a in attrs.fields(ours.ScoredBatch)] == [
        "cms_peaks", "cms_peak_vals", "cms_peak_channel_inds", "edge_inds", "edge_peak_inds", "line_scores", "info",
        "n_samples", "n_nodes", "skip_paf", "cms", "pafs"]
    assert [a.name for a in attrs.fields(ours.GroupingParams)] == [
        "paf_scorer_kwargs", "max_instances", "return_confmaps", "return_pafs", "return_paf_graph"]
    if not ref_loader.available():
        pytest.skip("reference tree not present on this box")
    theirs = ref_loader.ref().streaming
    for name in ("ScoredBatch", "GroupingParams"):
        assert [a.name for a in attrs.fields(getattr(ours, name))] == [a.name for a in attrs.fields(getattr(theirs, name))]


def _build_c_demo(tmp_path):
    import subprocess

    exe = str(tmp_path / "c_abi_demo")
    lib_dir = os.path.join(ROOT, "sleap_nn_b200", "lib")
    cmd = ["gcc", "-O2", "-std=c99", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", exe, "-L", lib_dir, "-lsleapnn_b200",
           "-L", "/usr/local/cuda/lib64", "-lcudart", "-lm", f"-Wl,-rpath,{lib_dir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_links_without_python(tmp_path):
    """include/sleapnn_b200.h compiles as C99 and a program that uses only it and libcudart links against the library."""
    _build_c_demo(tmp_path)


@pytest.mark.gpu
def test_c_program_drives_the_library_without_torch(tmp_path):
    """examples/c_abi_demo.c: make_multi_confmaps -> find_local_peaks / find_global_peaks from plain C; the planted points
    come back in the reference's order."""
    import subprocess

    r = subprocess.run([_build_c_demo(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "C ABI demo: OK" in r.stdout, r.stdout + r.stderr


def test_shims_are_plain_reexports_of_the_reference_names():
    """The two backward-compatibility import paths (inference/peak_finding.py:9-27, inference/paf_grouping.py:8-46) are
    explicit re-exports (visible to static analysis), and every name is the ops-module object itself."""
    from sleap_nn_b200.inference import paf_grouping, peak_finding
    from sleap_nn_b200.inference.ops import crops, paf, peaks

    for name in peak_finding.__all__:
        assert getattr(peak_finding, name) is getattr(crops if name == "crop_bboxes" else peaks, name)
    for name in paf_grouping.__all__:
        assert getattr(paf_grouping, name) is getattr(paf, name)
    for mod in (paf_grouping, peak_finding):
        src = open(mod.__file__).read()
        assert "globals()" not in src
        for name in mod.__all__:
            assert callable(getattr(mod, name)) or isinstance(getattr(mod, name), type)
