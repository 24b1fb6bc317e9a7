"""Pin oracle/ against the LIVE reference (unmodified files loaded in place from /root/reference).

Runs only where the reference tree is mounted (the build container); skipped elsewhere - there the
committed golden vectors (tests/test_oracle_golden.py) carry the pin.  Fresh random inputs here, so this
is independent evidence from the goldens: the oracle restates the reference, it does not memorise it.
"""

import numpy as np
import pytest
import torch

from oracle import paf as opaf
from oracle import peaks as opeaks
from oracle import ref_loader, synth
from oracle import targets as otgt
from tests.helpers import FLT_MIN, candidate_table, close, eq, npy

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def R():
    return ref_loader.ref()


@pytest.mark.parametrize("seed,shape,thr", [(0, (2, 3, 40, 36), 0.2), (1, (1, 5, 33, 65), 0.6), (2, (3, 2, 16, 16), 0.95)])
def test_local_and_global_peaks(R, seed, shape, thr):
    g = torch.Generator().manual_seed(seed)
    cms = torch.rand(shape, generator=g)
    for a, b in zip(opeaks.local_peaks_rough(cms, thr), R.peaks.find_local_peaks_rough(cms, threshold=thr)):
        eq(npy(a), npy(b))
    for size in (3, 5, 6):
        a = opeaks.local_peaks(cms, thr, "integral", size)
        b = R.peaks.find_local_peaks(cms, threshold=thr, refinement="integral", integral_patch_size=size)
        close(npy(a[0]), npy(b[0]), atol=1e-6)
        for x, y in zip(a[1:], b[1:]):
            eq(npy(x), npy(y))
    for a, b in zip(opeaks.global_peaks_rough(cms, 0.1), R.peaks.find_global_peaks_rough(cms, threshold=0.1)):
        eq(npy(a), npy(b))
    a = opeaks.global_peaks(cms, thr, "integral", 5)
    b = R.peaks.find_global_peaks(cms, threshold=thr, refinement="integral", integral_patch_size=5)
    close(npy(a[0]), npy(b[0]), atol=1e-6)
    eq(npy(a[1]), npy(b[1]))


def test_loop_restatement_is_the_reference(R):
    """The 4-nested-loop statement of the algorithm (what the CUDA kernel implements) == the reference."""
    g = torch.Generator().manual_seed(5)
    cms = torch.rand((2, 3, 12, 14), generator=g)
    cms[0, 1, 4, 4] = float("nan")
    for a, b in zip(opeaks.local_peaks_rough_loops(cms, 0.3), R.peaks.find_local_peaks_rough(cms, threshold=0.3)):
        eq(npy(a), npy(b))


@pytest.mark.parametrize("seed", [11, 12])
def test_bottomup_chain_on_synthetic_frames(R, seed):
    n_nodes, n_inst, hw, stride = 4, 2, (160, 192), 2
    edges = synth.chain_edges(n_nodes)
    poses = synth.make_poses(seed, 2, n_inst, n_nodes, hw, edges=edges, margin=40.0, step=20.0)
    cms, pafs = synth.render(poses, hw, stride, edges, seed=seed)
    pts, vals, si, ci = R.peaks.find_local_peaks(cms, threshold=0.2, refinement="integral")
    o = opeaks.local_peaks(cms, 0.2, "integral")
    close(npy(o[0]), npy(pts), atol=1e-6)
    eq(npy(o[3]), npy(ci))
    B = cms.shape[0]
    peaks, pvs, pcs = (synth.split_by_sample(x, si, B) for x in (pts * stride, vals, ci))
    pafs_v = pafs.permute(0, 2, 3, 1)
    scorer = R.paf.PAFScorer(part_names=[str(i) for i in range(n_nodes)], edges=[(str(a), str(b)) for a, b in edges],
                             pafs_stride=stride)
    want = scorer.predict(pafs_v, peaks, pvs, pcs)
    got = opaf.predict(pafs_v, peaks, pvs, pcs, edges, n_nodes, stride)
    for b in range(B):
        assert got[0][b].shape == want[0][b].shape
        close(npy(got[0][b]), npy(want[0][b]), atol=1e-6)
        eq(npy(got[1][b]), npy(want[1][b]))
        close(npy(got[2][b]), npy(want[2][b]), rtol=1e-6, atol=1e-6)
    # the PAF graph, keyed by (edge, src, dst): candidate order inside an edge is implementation-defined
    ei, epi, sc = scorer.score_paf_lines(pafs_v, peaks, pcs)
    oi, oepi, osc = opaf.score_lines_batch(pafs_v, peaks, pcs, edges, 10, stride, 0.25, 1.0, n_nodes)
    for b in range(B):
        w, g_ = candidate_table(ei[b], epi[b], sc[b]), candidate_table(oi[b], oepi[b], osc[b])
        assert w.keys() == g_.keys()
        for k in w:
            assert abs(w[k] - g_[k]) <= 1e-6 + 1e-6 * abs(w[k]) or (np.isnan(w[k]) and np.isnan(g_[k]))


def test_targets(R):
    g = torch.Generator().manual_seed(3)
    pts = torch.rand((1, 3, 4, 2), generator=g) * torch.tensor([90.0, 60.0])
    pts[0, 1, 2] = float("nan")
    xv, yv = R.data_utils.make_grid_vectors(60, 90, 2)
    oxv, oyv = otgt.grid_vectors(60, 90, 2)
    eq(npy(xv), npy(oxv)); eq(npy(yv), npy(oyv))
    close(npy(otgt.multi_confmaps(pts, xv, yv, 3.0)), npy(R.confidence_maps.make_multi_confmaps(pts, xv, yv, 3.0)),
          rtol=1e-6, atol=FLT_MIN)
    close(npy(otgt.confmaps(pts[0], xv, yv, 2.0)), npy(R.confidence_maps.make_confmaps(pts[0], xv, yv, 2.0)),
          rtol=1e-6, atol=FLT_MIN)
    e = torch.tensor([[0, 1], [1, 2], [2, 3]])
    srcs, dsts = pts[0][:, e[:, 0]], pts[0][:, e[:, 1]]
    close(npy(otgt.multi_pafs(xv, yv, srcs, dsts, 2.5)), npy(R.edge_maps.make_multi_pafs(xv, yv, srcs, dsts, 2.5)),
          rtol=1e-6, atol=1e-7)
    close(npy(otgt.pafs(xv, yv, srcs[0], dsts[0], 1.5)), npy(R.edge_maps.make_pafs(xv, yv, srcs[0], dsts[0], 1.5)),
          rtol=1e-6, atol=1e-7)
    close(npy(otgt.generate_pafs(pts, (60, 90), 1.5, 2, e, True)),
          npy(R.edge_maps.generate_pafs(pts, (60, 90), sigma=1.5, output_stride=2, edge_inds=e, flatten_channels=True)),
          rtol=1e-6, atol=1e-7)


def test_lsap_and_toposort_against_the_third_party_libraries(R):
    from scipy.optimize import linear_sum_assignment

    rng = np.random.default_rng(0)
    for _ in range(200):
        nr, nc = rng.integers(1, 9, 2)
        cost = rng.standard_normal((nr, nc))
        if rng.random() < 0.3:
            cost = np.round(cost, 1)  # ties
        r, c = opaf.lsap_jv(cost)
        wr, wc = linear_sum_assignment(cost)
        eq(r, wr); eq(c, wc)
    for edges in ([(0, 1), (1, 2), (1, 3)], [(2, 0), (0, 1), (0, 3), (3, 4)], [(0, 1), (2, 3)], [(0, 1), (0, 2), (1, 2)]):
        types = [R.paf.EdgeType(a, b) for a, b in edges]
        assert tuple(R.paf.toposort_edges(types)) == tuple(opaf.toposort_edge_order(edges))


@pytest.mark.parametrize("seed,thr", [(0, 0.5), (1, 0.2), (2, 0.05)])
def test_topdown_stage_b_and_stage_2_on_random_frames(R, seed, thr):
    """oracle.topdown vs TopDownLayer._centroid_nms_mask / _run_stage_2 (unmodified reference methods, stand-in self)
    on fresh random centroids: the NMS order / IoU arithmetic, the crop list, the lift and the per-frame classes."""
    import types

    from oracle import topdown as otd
    from tests.helpers import topdown_model

    TL = R.topdown.TopDownLayer
    TM = R.topdown_multiclass.CenteredInstanceMultiClassLayer
    P = R.preprocess_info.PreprocInfo
    g = torch.Generator().manual_seed(300 + seed)
    B, I, H, W, crop_hw, Nn, K = 5, 9, 96, 120, (20, 28), 3, 4
    cen = torch.rand((B, I, 2), generator=g) * torch.tensor([W - 1.0, H - 1.0])
    cen[torch.rand((B, I), generator=g) < 0.25] = float("nan")
    cen[1] = float("nan")
    val = torch.rand((B, I), generator=g)
    eff = torch.rand((B,), generator=g) * 0.6 + 0.7
    img = (torch.rand((B, 1, H, W), generator=g) * 255).to(torch.uint8)
    gain = torch.rand((Nn, *crop_hw), generator=g) * 0.5 + 0.5
    pattern = torch.rand((Nn, *crop_hw), generator=g) * 1e-3
    cgain = torch.rand((K, *crop_hw), generator=g) * 0.5 + 0.5
    cfg = types.SimpleNamespace(peak_threshold=0.2, effective_refinement="integral", integral_patch_size=5,
                                return_confmaps=False, return_class_vectors=True)

    def predict(crops):
        cms, vec = topdown_model(crops, gain, pattern, cgain)
        me2 = types.SimpleNamespace(postprocess_config=cfg, _extract_confmaps=lambda raw: raw["CenteredInstanceConfmapsHead"])
        info = P(eff_scale=torch.ones(crops.shape[0]), input_scale=0.5, output_stride=2)
        return TM.postprocess(me2, {"CenteredInstanceConfmapsHead": cms, "ClassVectorsHead": vec}, info)

    inner = types.SimpleNamespace(predict=predict, postprocess_config=cfg)
    me = types.SimpleNamespace(crop_size=crop_hw, centered_instance_layer=inner, return_crops=False, centroid_nms=True,
                               centroid_nms_threshold=thr, _bbox_iou=TL._bbox_iou, _infer_n_nodes=lambda: Nn)
    with ref_loader.reference_imports():
        valid = ~torch.isnan(cen).any(dim=-1)
        valid = valid & TL._centroid_nms_mask(me, cen, val, valid)
        o = TL._run_stage_2(me, img, cen * eff.view(-1, 1, 1), val, valid, eff_scale=eff)
    got = otd.stage_2(img, cen, val, eff, crop_hw, lambda c: topdown_model(c, gain, pattern, cgain), nms=True,
                      nms_threshold=thr, output_stride=2, input_scale=0.5)
    eq(got["valid"], npy(valid))
    assert npy(valid).sum() < (~np.isnan(npy(cen)).any(-1)).sum() or thr >= 0.5
    eq(got["kpts"], npy(o.pred_keypoints)); eq(got["crop_kpts"], npy(o.pred_crop_keypoints))
    eq(got["vals"], npy(o.pred_peak_values)); eq(got["centroids"], npy(o.pred_centroids))
    eq(got["bboxes"], npy(o.instance_bboxes))  # (return_crops is off: the reference itself raises when bbox 0 rounds short)
    eq(got["class_inds"], npy(o.pred_class_inds)); eq(got["tracking"], npy(o.instance_tracking_scores))
    eq(got["class_vectors"], npy(o.pred_class_vectors))


def test_layer_goldens_equal_the_layers_own_postprocess(R):
    """The f1 / f2 goldens were composed from the reference's ops in the order its layers call them; here the layers' OWN
    `postprocess` methods (unmodified classes, stand-in `self`) are run on the same inputs and must reproduce them."""
    import types

    from tests.helpers import T, golden

    P = R.preprocess_info.PreprocInfo
    cfg = lambda **kw: types.SimpleNamespace(peak_threshold=0.2, effective_refinement="integral", integral_patch_size=5,
                                             return_confmaps=False, return_pafs=False, return_paf_graph=False, **kw)
    # ---- f2: CentroidLayer.postprocess / CenteredInstanceLayer.postprocess vs ref_f2_layers.npz
    d = golden("ref_f2_layers.npz")
    CL, CI = R.centroid.CentroidLayer, R.centered_instance.CenteredInstanceLayer
    for tag in ("dyn", "top3", "pad7"):
        mi = int(d[f"cen_{tag}_max"])
        me = types.SimpleNamespace(postprocess_config=cfg(max_instances=None), max_instances=None if mi < 0 else mi,
                                   _extract_confmaps=lambda raw: raw["x"], _infer_max_instances=CL._infer_max_instances)
        info = P(eff_scale=T(d["cen_eff"]), input_scale=float(d[f"cen_{tag}_scale"]), output_stride=int(d["cen_stride"]))
        o = CL.postprocess(me, {"x": T(d["cen_cms"])}, info)
        eq(npy(o.pred_centroids), d[f"cen_{tag}_xy"])
        eq(npy(o.pred_centroid_values), d[f"cen_{tag}_val"])
    me = types.SimpleNamespace(postprocess_config=cfg(), _extract_confmaps=lambda raw: raw["x"])
    o = CI.postprocess(me, {"x": T(d["ci_cms"])}, P(eff_scale=T(d["ci_eff"]), input_scale=float(d["ci_scale"]), output_stride=2))
    eq(npy(o.pred_keypoints), d["ci_xy"])
    eq(npy(o.pred_peak_values), d["ci_val"])
    # ---- f1: BottomUpLayer.postprocess (= _score_pafs_on_gpu + group_scored_batch) vs ref_f1_outputs.npz
    p, f1 = golden("ref_pipeline_tree.npz"), golden("ref_f1_outputs.npz")
    BU = R.bottomup.BottomUpLayer
    edges, n_nodes, stride = p["edges"].tolist(), int(p["n_nodes"]), int(p["stride"])
    scorer = R.paf.PAFScorer(part_names=[str(i) for i in range(n_nodes)], edges=[(str(a), str(b)) for a, b in edges],
                             pafs_stride=stride, min_instance_peaks=int(p["min_instance_peaks"]))
    raw = {"MultiInstanceConfmapsHead": T(p["cms"]), "PartAffinityFieldsHead": T(p["pafs"])}
    for tag in f1["cases"].tolist():
        mi = int(f1[f"{tag}_max_instances"])
        skip = bool(f1[f"{tag}_skip"])
        me = types.SimpleNamespace(postprocess_config=cfg(max_instances=None), cms_output_stride=stride, paf_scorer=scorer,
                                   max_instances=None if mi < 0 else mi,
                                   max_peaks_per_node=int(f1["max_node_peaks"]) - 1 if skip else None)
        me._score_pafs_on_gpu = lambda r, i, me=me: BU._score_pafs_on_gpu(me, r, i)
        me.grouping_params = lambda me=me: BU.grouping_params(me)
        info = P(eff_scale=T(f1[f"{tag}_eff"]), input_scale=float(f1[f"{tag}_input_scale"]))
        with ref_loader.reference_imports():
            o = BU.postprocess(me, raw, info)
        eq(npy(o.pred_keypoints), f1[f"{tag}_kpts"])
        eq(npy(o.pred_peak_values), f1[f"{tag}_vals"])
        eq(npy(o.instance_scores), f1[f"{tag}_scores"])
