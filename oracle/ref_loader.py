"""Load the REAL reference hot-path modules from /root/reference, in place, unmodified.

TEST INFRASTRUCTURE ONLY.  Works only in the build container (the GPU box has no
/root/reference).  Used by `tests/golden/make_golden.py` to generate the committed
golden vectors and by `tests/test_oracle_vs_reference.py` (skipped when the
reference tree is absent) to pin `oracle/` against the reference itself.

`import sleap_nn` fails here (sleap_io / omegaconf / lightning are absent), so the
hot-path files (and the five caller-side files of the section-8f rows) are exec'd individually under empty namespace packages and
stub modules for the names they import but never touch on this path
(SURVEY.md section 8c).  No reference source is copied: files are read where they
lie.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("SLEAPNN_REFERENCE_ROOT", "/root/reference")

# (module name, path relative to REF_ROOT) in dependency order.
_HOT_FILES = [
    ("sleap_nn.data.providers", "sleap_nn/data/providers.py"),  # only filter_oob_points is used (sleap_io stays a stub)
    ("sleap_nn.data.instance_cropping", "sleap_nn/data/instance_cropping.py"),
    ("sleap_nn.data.utils", "sleap_nn/data/utils.py"),
    ("sleap_nn.data.confidence_maps", "sleap_nn/data/confidence_maps.py"),
    ("sleap_nn.data.edge_maps", "sleap_nn/data/edge_maps.py"),
    ("sleap_nn.inference.utils", "sleap_nn/inference/utils.py"),
    ("sleap_nn.inference.ops.crops", "sleap_nn/inference/ops/crops.py"),
    ("sleap_nn.inference.ops.peaks", "sleap_nn/inference/ops/peaks.py"),
    ("sleap_nn.inference.ops.paf", "sleap_nn/inference/ops/paf.py"),
    # the callers either side of the path (SURVEY.md section 8f "next" rows): pure attrs / torch / numpy / scipy
    ("sleap_nn.inference.preprocess_info", "sleap_nn/inference/preprocess_info.py"),
    ("sleap_nn.inference.outputs", "sleap_nn/inference/outputs.py"),
    ("sleap_nn.inference.streaming", "sleap_nn/inference/streaming.py"),
    ("sleap_nn.inference.ops.coord", "sleap_nn/inference/ops/coord.py"),
    ("sleap_nn.inference.ops.identity", "sleap_nn/inference/ops/identity.py"),
    ("sleap_nn.data.identity", "sleap_nn/data/identity.py"),
    ("sleap_nn.inference.filters", "sleap_nn/inference/filters.py"),
    ("sleap_nn.inference.ops.filters", "sleap_nn/inference/ops/filters.py"),  # numpy cores + Labels-level wrappers (sio stubbed)
    # layer glue of the f3 row: only BottomUpMultiClassLayer.postprocess / _cap_instances_by_score are called, with a
    # stand-in `self` (the base classes resolve to stubs)
    ("sleap_nn.inference.layers.bottomup_multiclass", "sleap_nn/inference/layers/bottomup_multiclass.py"),
    # layer glue of the f2 row: postprocess / _run_stage_2 / _centroid_nms_mask are called with stand-in `self` objects
    ("sleap_nn.inference.layers.bottomup", "sleap_nn/inference/layers/bottomup.py"),
    ("sleap_nn.inference.layers.centered_instance", "sleap_nn/inference/layers/centered_instance.py"),
    ("sleap_nn.inference.layers.single_instance", "sleap_nn/inference/layers/single_instance.py"),
    ("sleap_nn.inference.layers.centroid", "sleap_nn/inference/layers/centroid.py"),
    ("sleap_nn.inference.layers.topdown", "sleap_nn/inference/layers/topdown.py"),
    ("sleap_nn.inference.layers.topdown_multiclass", "sleap_nn/inference/layers/topdown_multiclass.py"),
]

_NAMESPACE_PKGS = [
    "sleap_nn",
    "sleap_nn.inference",
    "sleap_nn.inference.ops",
    "sleap_nn.data",
    "sleap_nn.config",
    "sleap_nn.inference.layers",
]


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "sleap_nn/inference/ops/peaks.py"))


class _Anything:
    """Attribute sink: any attribute access / call returns another sink."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)

    def _module_getattr(attr):
        if attr.startswith("__"):  # keep inspect / importlib machinery honest
            raise AttributeError(attr)
        return _Anything

    mod.__getattr__ = _module_getattr  # type: ignore[attr-defined]
    return mod


def load(prefix: str = "_sleapnn_ref") -> dict:
    """Return {short name: module} for the eight reference hot-path files.

    The modules are registered under their real dotted names while they are being
    exec'd (they import one another), then the `sleap_nn*` entries are moved out of
    `sys.modules` so the reference never shadows `sleap_nn_b200.compat.install()`.
    """
    if not available():
        raise FileNotFoundError(f"reference tree not found under {REF_ROOT}")
    import attrs, networkx, numpy, scipy.optimize, torch  # noqa: F401  real deps, import before stubbing

    saved = {k: v for k, v in sys.modules.items() if k == "sleap_nn" or k.startswith("sleap_nn.")}
    for k in saved:
        del sys.modules[k]
    stub_names = []
    try:
        for pkg in _NAMESPACE_PKGS:
            m = types.ModuleType(pkg)
            m.__path__ = []  # namespace package marker
            sys.modules[pkg] = m
        for name, attrs in [
            ("omegaconf", dict(OmegaConf=_Anything, DictConfig=_Anything)),
            ("sleap_io", {}),
            ("sleap_io.io", {}),
            ("sleap_io.io.skeleton", dict(SkeletonYAMLDecoder=_Anything)),
            ("sleap_nn.data.skia_augmentation", {}),
            ("sleap_nn.config.utils", {}),
        ]:
            if name not in sys.modules:
                sys.modules[name] = _stub(name, **attrs)
                stub_names.append(name)
        # psutil / loguru etc. are importable here; anything else missing gets a stub lazily.
        out = {}
        for modname, rel in _HOT_FILES:
            path = os.path.join(REF_ROOT, rel)
            spec = importlib.util.spec_from_file_location(modname, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[modname] = mod
            while True:
                try:
                    spec.loader.exec_module(mod)
                    break
                except ModuleNotFoundError as e:  # stub whatever else is absent
                    missing = e.name
                    if missing in sys.modules:
                        raise
                    sys.modules[missing] = _stub(missing)
                    stub_names.append(missing)
            out[modname.rsplit(".", 1)[-1]] = mod
        return out
    finally:
        for k in [k for k in sys.modules if k == "sleap_nn" or k.startswith("sleap_nn.")]:
            mod = sys.modules.pop(k)
            sys.modules[f"{prefix}.{k}"] = mod
        for k in stub_names:
            sys.modules.pop(k, None)
        sys.modules.update(saved)


import contextlib


@contextlib.contextmanager
def reference_imports(prefix: str = "_sleapnn_ref"):
    """Temporarily register the loaded reference modules under their real `sleap_nn.*` names.

    A few reference functions import lazily at CALL time (`group_scored_batch` does
    `from sleap_nn.inference.ops.paf import PAFScorer`, streaming.py:163); wrap such calls in this.
    """
    ref()
    saved = {k: v for k, v in sys.modules.items() if k == "sleap_nn" or k.startswith("sleap_nn.")}
    for k in saved:
        del sys.modules[k]
    added = []
    try:
        for k, m in list(sys.modules.items()):
            if k.startswith(prefix + "."):
                real = k[len(prefix) + 1:]
                sys.modules[real] = m
                added.append(real)
        yield
    finally:
        for k in added:
            sys.modules.pop(k, None)
        sys.modules.update(saved)


_CACHE = None


def ref() -> types.SimpleNamespace:
    """Cached namespace: ref().peaks, .crops, .paf, .utils (inference), .confidence_maps,
    .edge_maps, .data_utils, .instance_cropping."""
    global _CACHE
    if _CACHE is None:
        mods = load()
        # two files are called utils.py; disambiguate
        full = {name: sys.modules[f"_sleapnn_ref.{name}"] for name, _ in _HOT_FILES}
        _CACHE = types.SimpleNamespace(
            peaks=mods["peaks"],
            crops=mods["crops"],
            paf=mods["paf"],
            interp=full["sleap_nn.inference.utils"],
            confidence_maps=mods["confidence_maps"],
            edge_maps=mods["edge_maps"],
            data_utils=full["sleap_nn.data.utils"],
            instance_cropping=mods["instance_cropping"],
            preprocess_info=mods["preprocess_info"],
            outputs=mods["outputs"],
            streaming=mods["streaming"],
            coord=mods["coord"],
            identity=full["sleap_nn.inference.ops.identity"],
            data_identity=full["sleap_nn.data.identity"],
            providers=full["sleap_nn.data.providers"],
            filters=full["sleap_nn.inference.filters"],
            ops_filters=full["sleap_nn.inference.ops.filters"],
            bottomup_multiclass=full["sleap_nn.inference.layers.bottomup_multiclass"],
            bottomup=full["sleap_nn.inference.layers.bottomup"],
            centroid=full["sleap_nn.inference.layers.centroid"],
            centered_instance=full["sleap_nn.inference.layers.centered_instance"],
            single_instance=full["sleap_nn.inference.layers.single_instance"],
            topdown=full["sleap_nn.inference.layers.topdown"],
            topdown_multiclass=full["sleap_nn.inference.layers.topdown_multiclass"],
        )
    return _CACHE
