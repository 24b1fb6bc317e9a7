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
