"""Alias this package's modules over the reference's import paths.

    import sleap_nn_b200.compat as compat
    compat.install()            # sleap_nn.inference.peak_finding etc. now resolve to the CUDA path

The reference reaches the hot path through two import paths per function
(sleap_nn/inference/peak_finding.py:9-27 and sleap_nn/inference/paf_grouping.py:8-46 re-export
sleap_nn/inference/ops/*; sleap_nn/data/{confidence_maps,edge_maps,utils,instance_cropping}.py are
imported directly by the datasets).  `install()` registers this package's modules under those names
in `sys.modules`.  When the real `sleap_nn` package is importable it is left in place and only the
hot-path submodules are replaced (and their already-bound names re-pointed on the parent packages);
when it is not, empty namespace packages are created so `from sleap_nn.inference.peak_finding import
find_local_peaks` works on a box that has only this repo.  `uninstall()` restores the previous state.
"""

from __future__ import annotations

import importlib
import sys
import types
from typing import Dict, Optional

# reference module name -> module of this package that replaces it
ALIASES = {
    "sleap_nn.inference.peak_finding": "sleap_nn_b200.inference.peak_finding",
    "sleap_nn.inference.paf_grouping": "sleap_nn_b200.inference.paf_grouping",
    "sleap_nn.inference.ops.peaks": "sleap_nn_b200.inference.ops.peaks",
    "sleap_nn.inference.ops.crops": "sleap_nn_b200.inference.ops.crops",
    "sleap_nn.inference.ops.paf": "sleap_nn_b200.inference.ops.paf",
    "sleap_nn.data.confidence_maps": "sleap_nn_b200.data.confidence_maps",
    "sleap_nn.data.edge_maps": "sleap_nn_b200.data.edge_maps",
    "sleap_nn.inference.ops.identity": "sleap_nn_b200.inference.ops.identity",
    "sleap_nn.data.identity": "sleap_nn_b200.data.identity",
    "sleap_nn.inference.ops.coord": "sleap_nn_b200.inference.ops.coord",
}
# Reference modules that hold more than the hot path: when the real module is importable only these functions are
# re-pointed on it; when it is not (a box with only this repo) the whole name resolves to this package's module.
PATCHES = {
    "sleap_nn.data.utils": ("sleap_nn_b200.data.utils", ("make_grid_vectors", "gaussian_pdf")),
    "sleap_nn.data.instance_cropping": ("sleap_nn_b200.data.instance_cropping", ("make_centered_bboxes",)),
    "sleap_nn.inference.utils": ("sleap_nn_b200.inference.utils", ("interp1d",)),
}
# sleap_nn.inference.streaming is likewise left alone (it also holds the spawn pool); `ScoredBatch`, `GroupingParams` and
# `group_scored_batch` of sleap_nn_b200.inference.streaming are drop-ins for the three names of that seam.
# sleap_nn.inference.filters is NOT aliased wholesale: the reference module also re-exports `Outputs`; a maintainer
# swaps `FilterPipeline` / `FilterConfig` for sleap_nn_b200.inference.filters' (see INTEGRATION.md).

_saved: Optional[Dict[str, Optional[types.ModuleType]]] = None
_patched: Dict[str, Dict[str, object]] = {}


def _ensure_parent(name: str) -> types.ModuleType:
    """Import (or create as an empty namespace package) the parent package `name`."""
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        mod = types.ModuleType(name)
        mod.__path__ = []  # namespace-like: lets `import a.b.c` resolve through sys.modules
        mod.__sleapnn_b200_stub__ = True
        sys.modules[name] = mod
        if "." in name:
            parent, _, leaf = name.rpartition(".")
            setattr(_ensure_parent(parent), leaf, mod)
        return mod


def install() -> None:
    """Route the reference's hot-path import paths to the CUDA implementation (idempotent)."""
    global _saved
    if _saved is not None:
        return
    _saved = {}
    for ref_name, our_name in ALIASES.items():
        ours = importlib.import_module(our_name)
        parent_name, _, leaf = ref_name.rpartition(".")
        parent = _ensure_parent(parent_name)
        _saved[ref_name] = sys.modules.get(ref_name)
        sys.modules[ref_name] = ours
        setattr(parent, leaf, ours)
    for ref_name, (our_name, names) in PATCHES.items():
        ours = importlib.import_module(our_name)
        parent_name, _, leaf = ref_name.rpartition(".")
        parent = _ensure_parent(parent_name)
        try:
            real = sys.modules.get(ref_name) or importlib.import_module(ref_name)
        except Exception:
            real = None
        if real is None or real is ours:
            _saved[ref_name] = None
            sys.modules[ref_name] = ours
            setattr(parent, leaf, ours)
        else:
            _patched[ref_name] = {n: getattr(real, n, None) for n in names if hasattr(ours, n)}
            for n in _patched[ref_name]:
                setattr(real, n, getattr(ours, n))


def uninstall() -> None:
    """Undo `install()`."""
    global _saved
    if _saved is None:
        return
    for ref_name, prev in _saved.items():
        parent_name, _, leaf = ref_name.rpartition(".")
        if prev is None:
            sys.modules.pop(ref_name, None)
            parent = sys.modules.get(parent_name)
            if parent is not None and hasattr(parent, leaf):
                delattr(parent, leaf)
        else:
            sys.modules[ref_name] = prev
            parent = sys.modules.get(parent_name)
            if parent is not None:
                setattr(parent, leaf, prev)
    for ref_name, names in _patched.items():
        real = sys.modules.get(ref_name)
        for n, prev in names.items():
            if real is not None:
                setattr(real, n, prev) if prev is not None else delattr(real, n)
    _patched.clear()
    for name in [n for n, m in sys.modules.items() if getattr(m, "__sleapnn_b200_stub__", False)]:
        sys.modules.pop(name, None)
    _saved = None
