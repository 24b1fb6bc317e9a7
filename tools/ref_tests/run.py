#!/usr/bin/env python
"""Run the staged reference tests (see stage.py) against sleap_nn_b200 and print a summary."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tests = os.path.join(ROOT, "baseline", "_ref", "tests")
if not os.path.isdir(tests):
    sys.exit("nothing staged: run tools/ref_tests/stage.py in the build container first")
# deselected by name: tests of functions outside the hot path that live in the same files (dataset memory estimates,
# skeleton configs, the TorchBackend / layer classes) - they need sleap-io, a dataset or the backbone wrappers
DESELECT = ("not check_memory and not cache_memory and not get_skeleton_from_config and not torch_backend "
            "and not single_instance_layer_cross_device")
cmd = [sys.executable, "-m", "pytest", tests, "-q", "-rA", "--rootdir", os.path.join(ROOT, "baseline", "_ref"), "-p",
       "no:cacheprovider", "-c", os.devnull, "--import-mode=importlib", "-k", DESELECT] + sys.argv[1:]
sys.exit(subprocess.call(cmd, cwd=os.path.join(ROOT, "baseline", "_ref")))
