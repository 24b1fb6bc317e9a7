"""compat.install() routes the reference's import paths to this package (no GPU needed: imports only)."""

import importlib
import sys

import pytest


def test_install_aliases_reference_import_paths_and_uninstall_restores():
    import sleap_nn_b200.compat as compat

    had = {k: sys.modules.get(k) for k in list(sys.modules) if k == "sleap_nn" or k.startswith("sleap_nn.")}
    compat.install()
    try:
        from sleap_nn.inference.peak_finding import find_global_peaks, find_local_peaks  # noqa: F401
        from sleap_nn.inference.paf_grouping import PAFScorer, toposort_edges  # noqa: F401
        from sleap_nn.data.confidence_maps import make_multi_confmaps  # noqa: F401
        from sleap_nn.data.edge_maps import make_pafs  # noqa: F401
        import sleap_nn.inference.ops.paf as ref_paf
        import sleap_nn_b200.inference.ops.paf as our_paf

        assert ref_paf is our_paf
        assert find_local_peaks.__module__.startswith("sleap_nn_b200.")
        assert importlib.import_module("sleap_nn.inference.ops.peaks").find_local_peaks is find_local_peaks
        compat.install()  # idempotent
    finally:
        compat.uninstall()
    now = {k: sys.modules.get(k) for k in list(sys.modules) if k == "sleap_nn" or k.startswith("sleap_nn.")}
    assert now == had


def test_every_alias_target_exists():
    import sleap_nn_b200.compat as compat

    for ref_name, ours in compat.ALIASES.items():
        assert importlib.import_module(ours) is not None, ref_name


def test_install_patches_only_the_hot_path_functions_of_shared_modules(tmp_path, monkeypatch):
    """When the real sleap_nn is importable, modules that hold more than the hot path (data.utils, data.instance_cropping,
    inference.utils) keep their identity and their other functions; only the hot-path functions are re-pointed, and
    uninstall() puts the originals back."""
    import sleap_nn_b200.compat as compat

    pkg = tmp_path / "sleap_nn"
    for sub in ("", "data", "inference", "inference/ops"):
        (pkg / sub).mkdir(parents=True, exist_ok=True)
        (pkg / sub / "__init__.py").write_text("")
    (pkg / "data" / "utils.py").write_text(
        "def make_grid_vectors(image_height, image_width, output_stride=1):\n    return 'theirs'\n"
        "def gaussian_pdf(x, sigma):\n    return 'theirs'\n"
        "def check_memory(*a):\n    return 'not on the hot path'\n")
    (pkg / "data" / "instance_cropping.py").write_text("def make_centered_bboxes(c, h, w):\n    return 'theirs'\n")
    (pkg / "inference" / "utils.py").write_text(
        "def interp1d(x, y, xnew):\n    return 'theirs'\ndef get_skeleton_from_config(c):\n    return 'not on the hot path'\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    for k in [k for k in sys.modules if k == "sleap_nn" or k.startswith("sleap_nn.")]:
        monkeypatch.delitem(sys.modules, k)
    import sleap_nn.data.utils as real_utils
    import sleap_nn.inference.utils as real_iutils

    compat.install()
    try:
        import sleap_nn.data.utils as now
        import sleap_nn_b200.data.utils as ours

        assert now is real_utils                                    # the module object is untouched ...
        assert now.make_grid_vectors is ours.make_grid_vectors      # ... its hot-path functions are ours ...
        assert now.gaussian_pdf is ours.gaussian_pdf
        assert now.check_memory() == "not on the hot path"          # ... and the rest is still theirs
        assert real_iutils.interp1d.__module__.startswith("sleap_nn_b200.")
        assert real_iutils.get_skeleton_from_config(None) == "not on the hot path"
        from sleap_nn.inference.ops.peaks import find_local_peaks   # whole-module aliases still apply

        assert find_local_peaks.__module__.startswith("sleap_nn_b200.")
    finally:
        compat.uninstall()
    assert real_utils.make_grid_vectors(1, 1) == "theirs" and real_iutils.interp1d(0, 0, 0) == "theirs"
    for k in [k for k in sys.modules if k == "sleap_nn" or k.startswith("sleap_nn.")]:
        sys.modules.pop(k, None)
