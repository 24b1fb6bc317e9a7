"""conftest for the STAGED reference tests (copied to baseline/_ref/tests/conftest.py by stage.py).

Routes the reference's import paths to sleap_nn_b200 (compat.install()), stubs the third-party modules the test
files import at module top but the hot path never touches (omegaconf, sleap_io), registers the two attrs-only
reference modules the filter tests need, and provides the two asset fixtures of tests/fixtures/inference.py.
"""
import importlib.util
import sys
import types
from pathlib import Path

import pytest

HERE = Path(__file__).resolve().parent                      # baseline/_ref/tests
REPO = HERE.parent.parent.parent
sys.path.insert(0, str(REPO))


class _AttrDict(dict):
    """OmegaConf.create() stand-in: nested attribute access is all PAFScorer.from_config uses."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return _AttrDict(v) if isinstance(v, dict) else v


def _stub(name, **attrs):
    if name in sys.modules:
        return
    try:
        importlib.import_module(name)
        return
    except Exception:
        pass
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


_stub("omegaconf", OmegaConf=types.SimpleNamespace(create=lambda d: _AttrDict(d)), DictConfig=_AttrDict)
_stub("sleap_io")

import sleap_nn_b200.compat as compat  # noqa: E402

compat.install()


def _load_ref_module(name, rel):
    spec = importlib.util.spec_from_file_location(name, str(HERE.parent / rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    parent, _, leaf = name.rpartition(".")
    setattr(sys.modules[parent], leaf, mod)
    return mod


# attrs-only value types of the reference (no arithmetic): Outputs / PreprocInfo, used by test_filters.py
_load_ref_module("sleap_nn.inference.preprocess_info", "sleap_nn/inference/preprocess_info.py")
_load_ref_module("sleap_nn.inference.outputs", "sleap_nn/inference/outputs.py")
import sleap_nn_b200.inference.filters as _our_filters  # noqa: E402

sys.modules["sleap_nn.inference.filters"] = _our_filters
sys.modules["sleap_nn.inference"].filters = _our_filters
# tests/data/test_edge_maps.py imports process_lf at module top; only its sleap-io test (deselected) calls it
_prov = types.ModuleType("sleap_nn.data.providers")
_prov.process_lf = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("sleap-io is not available here"))
sys.modules.setdefault("sleap_nn.data.providers", _prov)


# names the staged test modules import at module top but whose tests are deselected (they need sleap-io / a dataset)
def _not_here(*a, **k):
    raise RuntimeError("outside the hot path: not provided by sleap_nn_b200")


for _mod, _names in (("sleap_nn.data.utils", ("check_memory", "check_cache_memory", "estimate_cache_memory")),
                     ("sleap_nn.inference.utils", ("get_skeleton_from_config",))):
    for _n in _names:
        if not hasattr(sys.modules[_mod], _n):
            setattr(sys.modules[_mod], _n, _not_here)


@pytest.fixture
def minimal_cms():
    return HERE / "assets" / "inference" / "minimal_cms.pt"


@pytest.fixture
def minimal_bboxes():
    return HERE / "assets" / "inference" / "minimal_bboxes.pt"


@pytest.fixture
def minimal_instance():
    pytest.skip("needs sleap-io and a .slp asset (not available offline)")
