"""Shared helpers for the test-suite: golden loading and keyed comparisons."""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLT_MIN = 1.1754944e-38


def golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name))


def T(a):
    return torch.from_numpy(np.asarray(a))


def ragged(d, prefix):
    return [T(d[f"{prefix}_{i}"]) for i in range(int(d[f"{prefix}_n"]))]


def eq(a, b):
    """Bit-exact equality, NaN == NaN."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    np.testing.assert_array_equal(a, b)


def close(a, b, rtol=0.0, atol=0.0):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def candidate_table(edge_inds, edge_peak_inds, scores):
    """{(edge, src peak, dst peak): score} - candidate order inside an edge is
    implementation-defined in the reference (unstable argsort)."""
    e, p, s = npy(edge_inds), npy(edge_peak_inds), npy(scores)
    out = {}
    for i in range(len(e)):
        out[(int(e[i]), int(p[i, 0]), int(p[i, 1]))] = float(s[i])
    assert len(out) == len(e), "duplicate candidates"
    return out


def fake_labels(d):
    """Rebuild the fake Labels object of tests/golden/make_golden.py:f4_labels_filters from ref_f4_labels_filters.npz.
    Predictions are instances of a class NAMED PredictedInstance (what the duck-typed filters look for)."""
    import types

    Pred = type("PredictedInstance", (), {})
    sizes = [int(d[f"lab_n_{i}"]) for i in range(int(d["lab_n_n"]))]
    lfs, k = [], 0
    for n in sizes:
        objs = []
        for uid in range(n):
            o = Pred() if int(d["lab_kind"][k]) else types.SimpleNamespace()
            pts = np.array(d["lab_pts"][k])
            o.uid, o.score = uid, float(d["lab_score"][k])
            o.numpy = (lambda p: (lambda: p))(pts)
            o.skeleton = types.SimpleNamespace(nodes=list(range(pts.shape[0])))
            o.points = {"score": np.array(d["lab_ps"][k])} if bool(d["lab_has_ps"][k]) else {}
            objs.append(o)
            k += 1
        lfs.append(types.SimpleNamespace(instances=objs))
    return types.SimpleNamespace(labeled_frames=lfs)


def topdown_model(crops, gain, pattern, cgain=None):
    """The stand-in centred-instance "network" of the top-down goldens (tests/golden/make_golden.py:f2_topdown):
    crops (n, 1, h, w) uint8 / float -> confmaps (n, N, h, w) [, class vectors (n, K)].  Every op is an exactly
    rounded elementwise fp32 op or a max, so CPU and CUDA runs see bit-identical maps."""
    x = crops.to(torch.float32) * 0.00390625  # 1 / 256: exact
    h, w = x.shape[-2:]  # crop_bboxes reads the size off bbox 0: it can be one less than configured (ops/crops.py:66-67)
    cms = x * gain[:, :h, :w].unsqueeze(0) + pattern[:, :h, :w].unsqueeze(0)
    if cgain is None:
        return cms
    return cms, (x * cgain[:, :h, :w].unsqueeze(0)).amax(dim=(2, 3))
