#!/usr/bin/env python
"""Stage the reference's OWN hot-path unit tests for a run against this package (build container only).

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
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
        os.chmod(dst, 0o644)
    shutil.copyfile(os.path.join(ROOT, "tools", "ref_tests", "conftest_staged.py"), os.path.join(DST, "tests", "conftest.py"))
    print(f"staged {len(FILES)} files under {DST}")


if __name__ == "__main__":
    main()
