"""The reference's OWN hot-path unit tests, run unmodified against this package (drop-in check).

They are staged into baseline/_ref/ (git-ignored) by `python tools/ref_tests/stage.py` in the build container and
travel to the GPU box with the snapshot; where nothing is staged this test skips.  A recorded run is kept in
profiles/r1_d_reference_own_tests.log.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_reference_own_unit_tests_pass_against_the_drop_in():
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "tests")):
        pytest.skip("reference tests not staged (tools/ref_tests/stage.py, build container only)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_tests", "run.py")], capture_output=True, text=True)
    tail = "\n".join(r.stdout.splitlines()[-40:])
    assert r.returncode == 0, tail
    assert " passed" in tail and " failed" not in tail, tail
