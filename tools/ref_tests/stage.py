#!/usr/bin/env python
"""Stage UNMODIFIED reference files under baseline/_ref/ (build container only; `__graft_entry__.build()` calls this):

  * the reference's OWN hot-path unit tests, for a run against this package (tests/test_reference_suite_gpu.py), and
  * the hot-path source files themselves (the list `oracle/ref_loader.py` exec's in place), so that `bench.py`'s CPU arm
    (`--impl reference`, `cpu_baseline`) times the reference's own implementation on the GPU box's host cores, where
    /root/reference does not exist.

    python tools/ref_tests/stage.py            # copies into baseline/_ref/ (git-ignored, travels with gpurun)
    gpurun -- 'python tools/ref_tests/run.py'  # runs them on the GPU box through sleap_nn_b200.compat.install()
    python tools/ref_tests/stage.py --clean    # removes the staged copy again

Nothing staged here is ever committed: baseline/_ref/ is listed in .gitignore (it is the place the task reserves for
an unmodified copy of the reference).  Staged: the test modules that exercise the hot path with plain tensors, their
two .pt assets, and two attrs-only reference modules the filter tests import (`Outputs`, `PreprocInfo`).  The
conftest that wires them to this package is tools/ref_tests/conftest_staged.py (ours).
"""
import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("SLEAPNN_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = [
    "tests/inference/test_peak_finding.py",
    "tests/inference/test_paf_grouping.py",
    "tests/inference/test_filters.py",
    "tests/inference/ops/test_crops.py",
    "tests/inference/ops/test_coord.py",
    "tests/data/test_edge_maps.py",
    "tests/data/test_identity.py",
    "tests/data/test_utils.py",
    "tests/inference/test_utils.py",
    "tests/inference/test_cuda.py",
    "tests/assets/inference/minimal_cms.pt",
    "tests/assets/inference/minimal_bboxes.pt",
    "sleap_nn/inference/outputs.py",
    "sleap_nn/inference/preprocess_info.py",
]


def hot_path_sources():
    """Relative paths of the reference source files oracle/ref_loader.py loads (kept in one place: its _HOT_FILES)."""
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ref_loader

    return [rel for _name, rel in ref_loader._HOT_FILES]


def stage(verbose: bool = True) -> int:
    """Copy the files (byte for byte) from REF into baseline/_ref/; returns how many.  No-op when REF is absent."""
    if not os.path.isdir(REF):
        return 0
    files = list(dict.fromkeys(FILES + hot_path_sources()))
    for rel in files:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
        os.chmod(dst, 0o644)
    shutil.copyfile(os.path.join(ROOT, "tools", "ref_tests", "conftest_staged.py"), os.path.join(DST, "tests", "conftest.py"))
    if verbose:
        print(f"staged {len(files)} unmodified reference files under {DST}")
    return len(files)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clean", action="store_true")
    args = ap.parse_args()
    if args.clean:
        shutil.rmtree(DST, ignore_errors=True)
        print(f"removed {DST}")
        return
    if not os.path.isdir(REF):
        sys.exit(f"{REF} not found: staging only works in the build container")
    stage()


if __name__ == "__main__":
    main()
