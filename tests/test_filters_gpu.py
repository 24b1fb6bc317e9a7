"""GPU parity of the post-inference filter pipeline: sleap_nn_b200.inference.filters against golden
vectors produced by the unmodified reference FilterPipeline and against the CPU oracle on random batches.  Which
slots survive is an integer decision: bit-exact (inputs certified to keep every similarity / score away from its
threshold); surviving values are copies of the inputs: bit-exact too."""

import numpy as np
import pytest
import torch

from tests.helpers import T, eq, golden, npy
from tests.test_oracle_golden import FILTER_CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fl():
    from sleap_nn_b200.inference import filters

    return filters


def test_filter_pipeline_golden(fl):
    d = golden("ref_f4_filters.npz")
    for dev in ("cuda", "cpu"):
        kpts, vals, scores = (T(d[k]).to(dev) for k in ("kpts", "vals", "scores"))
        for tag, kw in FILTER_CASES.items():
            out = fl.FilterPipeline.run(fl.FilterableOutputs(kpts, vals, scores), fl.FilterConfig(**kw))
            assert out.pred_keypoints.device.type == dev
            eq(npy(out.pred_keypoints), d[f"{tag}_kpts"]); eq(npy(out.pred_peak_values), d[f"{tag}_vals"])
            eq(npy(out.instance_scores), d[f"{tag}_scores"])
        eq(npy(kpts), d["kpts"])  # inputs are never modified
    out = fl.FilterPipeline(fl.FilterConfig(overlapping=True, overlapping_threshold=0.5))(
        fl.FilterableOutputs(T(d["kpts"]).cuda(), T(d["vals"]).cuda()))
    eq(npy(out.pred_keypoints), d["noscore_kpts"]); eq(npy(out.pred_peak_values), d["noscore_vals"])
    assert out.instance_scores is None
    cen, cenv = T(d["cen"]).cuda(), T(d["cenv"]).cuda()
    o = fl.FilterPipeline.run(fl.FilterableOutputs(pred_centroids=cen, pred_centroid_values=cenv),
                              fl.FilterConfig(min_instance_score=0.3, min_centroid_distance=12.0))
    eq(npy(o.pred_centroids), d["cen_a"]); eq(npy(o.pred_centroid_values), d["cenv_a"])
    o = fl.FilterPipeline.run(fl.FilterableOutputs(pred_centroids=cen, pred_centroid_values=cenv, instance_scores=1 - cenv),
                              fl.FilterConfig(min_centroid_distance=20.0))
    eq(npy(o.pred_centroids), d["cen_b"]); eq(npy(o.pred_centroid_values), d["cenv_b"]); eq(npy(o.instance_scores), d["cens_b"])
    o = fl.FilterPipeline.run(fl.FilterableOutputs(pred_centroids=cen, instance_scores=cenv),
                              fl.FilterConfig(min_instance_score=0.5, min_centroid_distance=6.0))
    eq(npy(o.pred_centroids), d["cen_c"]); eq(npy(o.instance_scores), d["cens_c"])
    # defaults are the identity and return the very same object; centroid-only overlap NMS warns and is skipped
    same = fl.FilterableOutputs(T(d["kpts"]).cuda())
    assert fl.FilterPipeline.run(same, fl.FilterConfig()) is same
    with pytest.warns(UserWarning):
        o = fl.FilterPipeline.run(fl.FilterableOutputs(pred_centroids=cen), fl.FilterConfig(overlapping=True))
    eq(npy(o.pred_centroids), d["cen"])


def _certified(kpts, thr_iou, thr_oks):
    from oracle import filters as ofil

    k = kpts.numpy()
    for b in range(k.shape[0]):
        for i in range(k.shape[1]):
            for j in range(k.shape[1]):
                if i != j and (abs(ofil.bbox_iou(k[b, i], k[b, j]) - thr_iou) < 1e-4 or abs(ofil.oks(k[b, i], k[b, j]) - thr_oks) < 1e-4):
                    return False
    return True


@pytest.mark.parametrize("B,I,Nn", [(3, 5, 1), (4, 40, 13), (2, 70, 4)])
def test_filter_pipeline_vs_oracle_random(fl, B, I, Nn):
    """More instance slots than lanes (I > 32), single-node skeletons (OKS falls back to IoU), heavy duplication."""
    from oracle import filters as ofil

    seed = 10 * I + Nn
    while True:
        g = torch.Generator().manual_seed(seed)
        base = torch.rand((B, 6, 1, 2), generator=g) * 400 + torch.cumsum(torch.rand((B, 6, Nn, 2), generator=g) * 40 - 15, dim=2)
        kpts = base[:, torch.randint(0, 6, (I,), generator=g)] + (torch.rand((B, I, Nn, 2), generator=g) - 0.5) * \
            torch.tensor([0.0, 2.0, 8.0, 30.0])[torch.randint(0, 4, (B, I, 1, 1), generator=g)]
        kpts[torch.rand((B, I, Nn), generator=g) < 0.25] = float("nan")
        kpts[torch.rand((B, I), generator=g) < 0.15] = float("nan")
        if _certified(kpts, 0.45, 0.35):
            break
        seed += 1000
    vals = torch.rand((B, I, Nn), generator=g)
    scores = torch.rand((B, I), generator=g)
    scores[torch.rand((B, I), generator=g) < 0.1] = float("nan")
    cfgs = [dict(overlapping=True, overlapping_threshold=0.45, overlapping_method="iou"),
            dict(overlapping=True, overlapping_threshold=0.35, overlapping_method="oks"),
            dict(min_peak_value=0.2, min_visible_nodes=1, min_visible_node_fraction=0.4, min_instance_score=0.15,
                 min_mean_node_score=0.35, overlapping=True, overlapping_threshold=0.35, overlapping_method="oks")]
    import warnings

    for cfg in cfgs:
        want = ofil.apply(cfg, kpts.numpy(), vals.numpy(), scores.numpy())
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = fl.FilterPipeline.run(fl.FilterableOutputs(kpts.cuda(), vals.cuda(), scores.cuda()), fl.FilterConfig(**cfg))
        eq(npy(got.pred_keypoints), want[0]); eq(npy(got.pred_peak_values), want[1]); eq(npy(got.instance_scores), want[2])
    cen = torch.rand((B, I, 2), generator=g) * 120
    cen[torch.rand((B, I), generator=g) < 0.1] = float("nan")
    cfg = dict(min_instance_score=0.2, min_centroid_distance=9.0)
    want = ofil.apply(cfg, scores=None, cen=cen.numpy(), cenv=scores.numpy())
    got = fl.FilterPipeline.run(fl.FilterableOutputs(pred_centroids=cen.cuda(), pred_centroid_values=scores.cuda()), fl.FilterConfig(**cfg))
    eq(npy(got.pred_centroids), want[3]); eq(npy(got.pred_centroid_values), want[4])


def test_filters_on_the_bottomup_outputs(fl):
    """End of the device chain: grouping outputs (B, I, N, 2) stay in HBM and are filtered there."""
    from sleap_nn_b200 import synthetic
    from sleap_nn_b200.pipeline import BottomUpPostproc

    dev = torch.device("cuda", 0)
    edges = synthetic.chain_edges(5)
    poses = synthetic.random_poses(5, 4, 3, 5, (256, 256), edges, margin=60.0, step=24.0)
    cms, pafs = synthetic.render_batch(poses, (256, 256), 2, edges, dev)
    pipe = BottomUpPostproc(5, edges, 4, tuple(cms.shape[-2:]), cms_stride=2, pafs_stride=2, device=dev)
    k, v, s = pipe(cms, pafs).outputs(max_instances=6)
    out = fl.FilterPipeline.run(fl.FilterableOutputs(k, v, s), fl.FilterConfig(min_visible_nodes=5, overlapping=True,
                                                                                overlapping_threshold=0.8))
    assert out.pred_keypoints.is_cuda and tuple(out.pred_keypoints.shape) == tuple(k.shape)
    n_in = int((~torch.isnan(k).all(-1).all(-1)).sum())
    n_out = int((~torch.isnan(out.pred_keypoints).all(-1).all(-1)).sum())
    assert n_in == 12 and n_out == 12  # three well-separated full skeletons per frame survive


def test_labels_level_filters_golden():
    """sleap_nn_b200.inference.ops.filters on fake Labels == the reference's ops/filters.py (kept instances per frame,
    in order), and the numeric cores on plain arrays."""
    from sleap_nn_b200.inference.ops import filters as opf
    from tests.helpers import fake_labels

    d = golden("ref_f4_labels_filters.npz")
    cases = {
        "count": lambda L: opf.filter_by_node_count(L, min_visible_nodes=4, min_visible_node_fraction=0.6),
        "conf": lambda L: opf.filter_by_node_confidence(L, min_mean_node_score=0.5, min_instance_score=0.3),
        "iou": lambda L: opf.filter_overlapping_instances(L, threshold=0.45, method="iou"),
        "oks": lambda L: opf.filter_overlapping_instances(L, threshold=0.3, method="oks"),
    }
    for tag, fn in cases.items():
        L = fake_labels(d)
        assert fn(L) is L
        for f, lf in enumerate(L.labeled_frames):
            assert [o.uid for o in lf.instances] == d[f"lab_{tag}_{f}"].tolist(), (tag, f)
    L = fake_labels(d)
    assert opf.filter_by_node_count(L) is L and opf.filter_by_node_confidence(L) is L  # defaults: no-ops
    with pytest.raises(ValueError):
        opf.filter_overlapping_instances(L, method="giou")
    pred = [o for o in fake_labels(d).labeled_frames[4].instances if type(o).__name__ == "PredictedInstance"]
    pts, sc = [o.numpy() for o in pred], np.array([o.score for o in pred])
    assert opf._nms_greedy_oks(pts, sc, 0.3) == d["core_oks_keep"].tolist()
    from oracle import filters as ofil

    boxes = np.array([ofil.instance_bbox64(p) for p in pts])
    assert opf._nms_greedy_iou(boxes, sc, 0.45) == d["core_iou_keep"].tolist()
    assert opf._nms_greedy_iou(np.zeros((0, 4)), np.zeros(0), 0.5) == [] and opf._nms_greedy_oks([], np.zeros(0), 0.5) == []


def _ref_outputs(fl):
    """`_make_outputs` of the reference's tests/inference/test_filters.py:30-57."""
    kpts = torch.tensor([[[[1.0, 1.0], [2.0, 2.0], [3.0, 3.0], [4.0, 4.0]],
                          [[10.0, 10.0], [11.0, 11.0], [12.0, 12.0], [13.0, 13.0]],
                          [[20.0, 20.0], [21.0, 21.0], [22.0, 22.0], [23.0, 23.0]]]])
    vals = torch.tensor([[[0.9, 0.8, 0.7, 0.6], [0.4, 0.3, 0.2, 0.1], [0.95, 0.05, 0.95, 0.05]]])
    return fl.FilterableOutputs(kpts, vals, torch.tensor([[0.85, 0.30, 0.55]]))


def test_reference_known_answers_replayed(fl):
    """The known-answer tests of the reference's tests/inference/test_filters.py, through the CUDA pipeline."""
    import pickle

    FP, FC = fl.FilterPipeline, fl.FilterConfig
    nan_all = lambda t: bool(torch.isnan(t).all())
    nan_any = lambda t: bool(torch.isnan(t).any())
    o = _ref_outputs(fl)
    out = FP(FC())(o)                                                           # :65-70
    assert torch.equal(out.pred_keypoints, o.pred_keypoints) and torch.equal(out.pred_peak_values, o.pred_peak_values)
    out = FP(FC(min_peak_value=0.5))(o)                                         # :78-85
    assert nan_all(out.pred_keypoints[0, 1]) and nan_all(out.pred_keypoints[0, 2, 1]) and not nan_any(out.pred_keypoints[0, 2, 0])
    k = o.pred_keypoints.clone(); k[0, 2, 2:] = float("nan")                    # :93-108
    out = FP(FC(min_visible_nodes=3))(fl.FilterableOutputs(k, o.pred_peak_values, o.instance_scores))
    assert not nan_any(out.pred_keypoints[0, 0]) and not nan_any(out.pred_keypoints[0, 1]) and nan_all(out.pred_keypoints[0, 2])
    k = o.pred_keypoints.clone(); k[0, 0, 2:] = float("nan"); k[0, 1, 1:] = float("nan")   # :116-129
    out = FP(FC(min_visible_node_fraction=0.4))(fl.FilterableOutputs(k, o.pred_peak_values, o.instance_scores))
    assert not nan_all(out.pred_keypoints[0, 0]) and nan_all(out.pred_keypoints[0, 1])
    out = FP(FC(min_instance_score=0.5))(o)                                     # :137-143
    assert not nan_any(out.pred_keypoints[0, 0]) and nan_all(out.pred_keypoints[0, 1]) and not nan_any(out.pred_keypoints[0, 2])
    out = FP(FC(min_mean_node_score=0.5))(o)                                    # :151-157 (instance 2: mean exactly 0.5 is kept)
    assert not nan_any(out.pred_keypoints[0, 0]) and nan_all(out.pred_keypoints[0, 1]) and not nan_any(out.pred_keypoints[0, 2])
    two = fl.FilterableOutputs(torch.tensor([[[[10.0, 10.0], [12.0, 12.0]], [[10.1, 10.1], [12.1, 12.1]]]]),
                               torch.tensor([[[0.9, 0.9], [0.5, 0.5]]]), torch.tensor([[0.9, 0.5]]))
    out = FP(FC(overlapping=True, overlapping_threshold=0.5, overlapping_method="iou"))(two)  # :165-184
    assert not nan_any(out.pred_keypoints[0, 0]) and nan_all(out.pred_keypoints[0, 1])
    one = fl.FilterableOutputs(torch.tensor([[[[1.0, 1.0]], [[1.05, 1.05]]]]), torch.tensor([[[1.0], [1.0]]]), torch.tensor([[0.9, 0.5]]))
    with pytest.warns(UserWarning):                                             # :192-208 (+ the single-node fallback, #586)
        out = FP(FC(overlapping=True, overlapping_threshold=0.5, overlapping_method="oks"))(one)
    assert int((~torch.isnan(out.pred_keypoints).all(dim=-1).all(dim=-1)).sum()) >= 1
    a = [(0.0, 0.0), (10.0, 0.0), (0.0, 10.0), (10.0, 10.0)]                    # :303-333, float64 Outputs
    b = [(0.1, 0.1), (10.1, 0.0), (0.0, 10.1), (10.1, 10.1)]
    o64 = fl.FilterableOutputs(torch.tensor([[a, b]], dtype=torch.float64), torch.ones(1, 2, 4, dtype=torch.float64),
                               torch.tensor([[0.9, 0.5]], dtype=torch.float64))
    out = FP(FC(overlapping=True, overlapping_threshold=0.5, overlapping_method="oks"))(o64)
    assert out.pred_keypoints.dtype == torch.float64 and torch.equal(out.pred_keypoints[0, 0], o64.pred_keypoints[0, 0])
    assert nan_all(out.pred_keypoints[0, 1])
    out = FP(FC(min_peak_value=0.1, min_visible_nodes=1, min_instance_score=0.5))(o)   # :342-349
    assert nan_all(out.pred_keypoints[0, 1])
    cfg = FC(min_peak_value=0.2, min_instance_score=0.3, overlapping=True, overlapping_threshold=0.5, overlapping_method="oks")
    back = pickle.loads(pickle.dumps(cfg))                                      # :357-368
    assert back == cfg and back.overlapping_method == "oks"
    via_run, via_inst = FP.run(o, FC(min_instance_score=0.5)), FP(FC(min_instance_score=0.5))(o)   # :376-383
    assert torch.equal(torch.nan_to_num(via_run.pred_keypoints, nan=-1.0), torch.nan_to_num(via_inst.pred_keypoints, nan=-1.0))
