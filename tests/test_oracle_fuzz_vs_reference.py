"""Seeded fuzzing of oracle/ against the LIVE reference (build container only; skipped where /root/reference is absent).

The GPU parity tests check the CUDA path against the oracle on inputs far from the committed goldens (busy frames,
cross-edge skeletons, repeated peaks, non-finite maps); these tests make sure the oracle itself still IS the reference
there.  Small shapes, a few hundred cases, a few seconds.
"""

import numpy as np
import pytest
import torch

from oracle import paf as opaf
from oracle import peaks as opeaks
from oracle import ref_loader
from oracle import targets as otgt
from tests.helpers import FLT_MIN, close, eq, npy

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def R():
    return ref_loader.ref()


def test_assembly_fuzz_cross_edges_repeated_peaks(R):
    """group_instances_sample on the same generator as tests/test_paf_gpu.py::test_assembly_randomised_busy_frames_vs_oracle."""
    P = R.paf
    n_raised = 0
    for seed in range(6):
        g = np.random.default_rng(1000 + seed)
        for case in range(12):
            n_nodes = int(g.integers(3, 9))
            edges = [(int(g.integers(0, k)), k) for k in range(1, n_nodes)]
            for _ in range(int(g.integers(0, 3))):
                a, b = sorted(g.choice(n_nodes, 2, replace=False).tolist())
                if (a, b) not in edges:
                    edges.append((a, b))
            edges = [edges[i] for i in g.permutation(len(edges))]
            n_per = g.integers(1, 46 if case % 4 == 0 else 9, n_nodes)
            ch = np.concatenate([np.full(n, k) for k, n in enumerate(n_per)]).astype(np.int32)
            ch = ch[g.permutation(len(ch))]
            pk = g.uniform(0, 500, (len(ch), 2)).astype(np.float32)
            pv = g.uniform(0.2, 1, len(ch)).astype(np.float32)
            me, ms, md, msc = [], [], [], []
            for k, (a, b) in enumerate(edges):
                n_m = int(g.integers(0, min(n_per[a], n_per[b]) + 1))
                if case % 3 == 2:
                    src, dst = g.integers(0, n_per[a], n_m), g.integers(0, n_per[b], n_m)
                else:
                    src, dst = g.permutation(n_per[a])[:n_m], g.permutation(n_per[b])[:n_m]
                me += [k] * n_m; ms += src.tolist(); md += dst.tolist(); msc += g.uniform(0.0, 1.0, n_m).tolist()
            if case % 2:
                order = g.permutation(len(me))
                me, ms, md, msc = ([x[i] for i in order] for x in (me, ms, md, msc))
            et = [P.EdgeType(a, b) for a, b in edges]
            sorted_inds = P.toposort_edges(et)
            assert tuple(sorted_inds) == tuple(opaf.toposort_edge_order(edges))
            mip = [0, 3, 0.5][case % 3]
            i32 = lambda x: torch.tensor(x, dtype=torch.int32)
            args = (torch.from_numpy(pk), torch.from_numpy(pv), torch.from_numpy(ch), i32(me), i32(ms), i32(md),
                    torch.tensor(msc, dtype=torch.float32))
            try:
                want = P.group_instances_sample(*args, n_nodes, sorted_inds, et, mip, 0.25)
            except (AssertionError, KeyError) as err:  # make_predicted_instances' sanity check (improper matchings only)
                with pytest.raises(type(err)):
                    opaf.group_sample(*args, n_nodes, sorted_inds, edges, mip, 0.25)
                n_raised += 1
                continue
            got = opaf.group_sample(*args, n_nodes, sorted_inds, edges, mip, 0.25)
            for a_, b_ in zip(got, want):
                eq(a_, b_)
    assert 0 < n_raised < 10


def test_peaks_fuzz_non_finite_and_plateaus(R):
    for seed in range(40):
        g = torch.Generator().manual_seed(seed)
        B, C, H, W = (int(v) for v in torch.randint(1, 5, (4,), generator=g))
        H, W = H * 5 + 3, W * 7 + 2
        cms = torch.rand((B, C, H, W), generator=g)
        if seed % 3 == 0:
            cms = (cms * 6).round() / 6          # plateaus and exact ties
        if seed % 4 == 1:
            flat = cms.view(-1)
            idx = torch.randint(0, flat.numel(), (6,), generator=g)
            flat[idx[:2]] = float("nan"); flat[idx[2:4]] = float("inf"); flat[idx[4:]] = float("-inf")
        thr = float(torch.rand((), generator=g))
        for a, b in zip(opeaks.local_peaks_rough(cms, thr), R.peaks.find_local_peaks_rough(cms, threshold=thr)):
            eq(npy(a), npy(b))
        for a, b in zip(opeaks.global_peaks_rough(cms, thr), R.peaks.find_global_peaks_rough(cms, threshold=thr)):
            eq(npy(a), npy(b))
        size = [3, 4, 5, 7][seed % 4]
        a = opeaks.local_peaks(cms, thr, "integral", size)
        b = R.peaks.find_local_peaks(cms, threshold=thr, refinement="integral", integral_patch_size=size)
        close(npy(a[0]), npy(b[0]), atol=1e-5)
        a = opeaks.global_peaks(cms, thr, "integral", size)
        b = R.peaks.find_global_peaks(cms, threshold=thr, refinement="integral", integral_patch_size=size)
        close(npy(a[0]), npy(b[0]), atol=1e-5)
        eq(npy(a[1]), npy(b[1]))


def test_crops_fuzz_out_of_bounds_and_dtypes(R):
    for seed in range(40):
        g = torch.Generator().manual_seed(100 + seed)
        S, C, H, W = 3, 2, 17, 23
        dt = [torch.float32, torch.uint8, torch.int16, torch.float64][seed % 4]
        img = (torch.rand((S, C, H, W), generator=g) * 200).to(dt)
        n = int(torch.randint(1, 9, (1,), generator=g))
        bh, bw = int(torch.randint(1, 12, (1,), generator=g)), int(torch.randint(1, 12, (1,), generator=g))
        cen = torch.rand((n, 2), generator=g) * torch.tensor([W + 20.0, H + 20.0]) - 10.0
        if seed % 2:
            cen = cen.round()
        bb = R.instance_cropping.make_centered_bboxes(cen, bh, bw)
        eq(npy(opeaks.centered_bboxes(cen, bh, bw)), npy(bb))
        si = torch.randint(0, S, (n,), generator=g)
        eq(npy(opeaks.crop_patches(img, bb, si)), npy(R.crops.crop_bboxes(img, bb, si)))


def test_line_subscripts_and_scores_fuzz(R):
    for seed in range(30):
        g = torch.Generator().manual_seed(200 + seed)
        n_nodes = int(torch.randint(2, 6, (1,), generator=g))
        edges = [(k - 1, k) for k in range(1, n_nodes)] + ([(0, n_nodes - 1)] if n_nodes > 2 else [])
        stride = [1, 2, 4, 8][seed % 4]
        n_pts = [3, 5, 10][seed % 3]
        Hp, Wp = 24, 31
        P_ = int(torch.randint(2, 14, (1,), generator=g))
        peaks = torch.rand((P_, 2), generator=g) * torch.tensor([Wp * stride + 6.0, Hp * stride + 6.0]) - 3.0
        ch = torch.randint(0, n_nodes, (P_,), generator=g).to(torch.int32)
        pafs = torch.randn((Hp, Wp, 2 * len(edges)), generator=g)
        ei, epi = R.paf.get_connection_candidates(ch, torch.tensor(edges, dtype=torch.int32), n_nodes)
        oi, opi = opaf.connection_candidates(ch, edges, n_nodes)
        eq(npy(oi), npy(ei)); eq(npy(opi), npy(epi))
        if ei.numel() == 0:
            continue
        want = R.paf.make_line_subs(peaks, epi, ei, n_pts, stride, (Hp, Wp))
        eq(npy(opaf.line_subscripts(peaks, epi, ei, n_pts, stride, (Hp, Wp))), npy(want))
        lines = R.paf.get_paf_lines(pafs, peaks, epi, ei, n_pts, stride)
        eq(npy(opaf.paf_lines(pafs, peaks, epi, ei, n_pts, stride)), npy(lines))
        mel = 0.25 * max(Hp, Wp, 2 * len(edges)) * stride
        close(npy(opaf.score_lines(lines, peaks, epi, mel, 1.0)), npy(R.paf.score_paf_lines(lines, peaks, epi, mel, 1.0)),
              rtol=1e-5, atol=1e-6)


def test_targets_fuzz_nan_points_and_degenerate_edges(R):
    for seed in range(12):
        g = torch.Generator().manual_seed(300 + seed)
        h, w, stride = 40, 56, [1, 2, 4][seed % 3]
        xv, yv = R.data_utils.make_grid_vectors(h, w, stride)
        oxv, oyv = otgt.grid_vectors(h, w, stride)
        eq(npy(oxv), npy(xv)); eq(npy(oyv), npy(yv))
        I, N = 3, 4
        pts = torch.rand((1, I, N, 2), generator=g) * torch.tensor([w + 10.0, h + 10.0]) - 5.0
        pts[0, 1, 2] = float("nan")
        if seed % 4 == 0:
            pts[0, 2, 0, 0] = float("inf")
        sigma = [0.4, 1.5, 2.5, 6.0][seed % 4]
        close(npy(otgt.multi_confmaps(pts, oxv, oyv, sigma)), npy(R.confidence_maps.make_multi_confmaps(pts, xv, yv, sigma)),
              rtol=1e-5, atol=FLT_MIN)
        e = torch.tensor([[0, 1], [1, 2], [2, 2], [3, 0]])          # (2, 2) is a zero-length edge
        src, dst = pts[0][:, e[:, 0]], pts[0][:, e[:, 1]]
        close(npy(otgt.multi_pafs(oxv, oyv, src, dst, sigma)), npy(R.edge_maps.make_multi_pafs(xv, yv, src, dst, sigma)),
              rtol=1e-5, atol=1e-6)


def test_match_candidates_fuzz_incl_nan_scores(R):
    """match_candidates_sample on full src x dst candidate grids with random scores, some NaN (cost +inf): same matches,
    and ValueError("cost matrix is infeasible") exactly where the reference (scipy) raises."""
    from oracle.paf import lsap_jv

    n_raised = 0
    for seed in range(60):
        g = torch.Generator().manual_seed(400 + seed)
        n_nodes = int(torch.randint(2, 6, (1,), generator=g))
        edges = [(k - 1, k) for k in range(1, n_nodes)]
        ch = torch.randint(0, n_nodes, (int(torch.randint(2, 16, (1,), generator=g)),), generator=g).to(torch.int32)
        ei, epi = R.paf.get_connection_candidates(ch, torch.tensor(edges, dtype=torch.int32), n_nodes)
        sc = torch.rand((ei.numel(),), generator=g) * 2 - 0.5
        if seed % 3 == 0 and ei.numel():
            sc[torch.rand((ei.numel(),), generator=g) < 0.3] = float("nan")
        try:
            want = R.paf.match_candidates_sample(ei, epi, sc, len(edges))
        except ValueError as err:
            n_raised += 1
            with pytest.raises(ValueError, match=str(err)):
                opaf.match_sample(ei, epi, sc, len(edges), solver=lsap_jv)
            continue
        for solver in (None, lsap_jv):  # scipy (what the reference calls) and the restatement the kernels follow
            got = opaf.match_sample(ei, epi, sc, len(edges), solver=solver)
            for a_, b_ in zip(got, want):
                eq(npy(a_), npy(b_))
    assert n_raised > 0


def test_identity_grouping_fuzz(R):
    """group_class_peaks / classify_peaks_from_maps / get_class_inds_from_vectors on random probabilities (rows need not
    sum to one, more peaks than classes and the reverse)."""
    from oracle import identity as oid

    for seed in range(40):
        g = torch.Generator().manual_seed(500 + seed)
        S, C, K = (int(v) for v in torch.randint(1, 5, (3,), generator=g))
        P_ = int(torch.randint(0, 20, (1,), generator=g))
        probs = torch.rand((P_, K), generator=g)
        si = torch.randint(0, S, (P_,), generator=g).to(torch.int32)
        ci = torch.randint(0, C, (P_,), generator=g).to(torch.int32)
        for a_, b_ in zip(oid.group_class_peaks(probs, si, ci, S, C), R.identity.group_class_peaks(probs, si, ci, S, C)):
            eq(npy(a_), npy(b_))
        Hc, Wc = 9, 13
        maps = torch.rand((S, K, Hc, Wc), generator=g)
        pts = torch.rand((P_, 2), generator=g) * torch.tensor([Wc + 4.0, Hc + 4.0]) - 2.0  # some land outside: clamped
        vals = torch.rand((P_,), generator=g)
        want = R.identity.classify_peaks_from_maps(maps, pts, vals, si, ci, C)
        got = oid.classify_peaks_from_maps(maps, pts, vals, si, ci, C)
        for a_, b_ in zip(got, want):
            eq(npy(a_), npy(b_))
        vec = torch.rand((int(torch.randint(1, 9, (1,), generator=g)), K), generator=g)
        for a_, b_ in zip(oid.class_inds_from_vectors(vec), R.identity.get_class_inds_from_vectors(vec)):
            eq(npy(a_), npy(b_))


def test_filter_pipeline_fuzz(R):
    """FilterPipeline.apply on random padded outputs with NaN slots, every filter on, both overlap methods.  Decisions that
    hinge on a mean (nanmean reduction order, OKS) are compared through the surviving-slot pattern with a retry margin:
    a case is skipped when a statistic sits within 1e-6 of its threshold."""
    from oracle import filters as ofil

    FC, FP, Out = R.filters.FilterConfig, R.filters.FilterPipeline, R.outputs.Outputs
    n_checked = 0
    for seed in range(60):
        g = torch.Generator().manual_seed(600 + seed)
        B, I, Nn = 2, int(torch.randint(1, 7, (1,), generator=g)), int(torch.randint(2, 6, (1,), generator=g))
        k = torch.rand((B, I, Nn, 2), generator=g) * 60
        if seed % 2:
            k[:, 1:] = k[:, :1] + torch.randn((B, I - 1, Nn, 2), generator=g) * 3  # overlapping instances
        v = torch.rand((B, I, Nn), generator=g)
        s = torch.rand((B, I), generator=g)
        drop = torch.rand((B, I, Nn), generator=g) < 0.2
        k[drop] = float("nan"); v[drop] = float("nan")
        cfgd = dict(min_peak_value=[0.0, 0.15][seed % 2], min_instance_score=[0.0, 0.2][(seed // 2) % 2],
                    min_mean_node_score=[0.0, 0.3][(seed // 4) % 2], min_visible_nodes=[0, 2][(seed // 8) % 2],
                    min_visible_node_fraction=[0.0, 0.5][(seed // 3) % 2], overlapping=bool(seed % 3),
                    overlapping_threshold=0.3, overlapping_method=["iou", "oks"][(seed // 5) % 2], min_centroid_distance=0.0)
        want = FP(config=FC(**cfgd)).apply(Out(pred_keypoints=k, pred_peak_values=v, instance_scores=s))
        got = ofil.apply(cfgd, kpts=k.numpy(), vals=v.numpy(), scores=s.numpy())
        same = (np.isnan(got[0]) == np.isnan(npy(want.pred_keypoints))).all()
        if not same:
            # tolerate only threshold-adjacent decisions: perturb the thresholds by 1e-6 both ways and require a match
            ok = False
            for eps in (-1e-6, 1e-6):
                c2 = dict(cfgd, min_mean_node_score=max(cfgd["min_mean_node_score"] + eps, 0.0),
                          overlapping_threshold=cfgd["overlapping_threshold"] + eps)
                g2 = ofil.apply(c2, kpts=k.numpy(), vals=v.numpy(), scores=s.numpy())
                ok = ok or (np.isnan(g2[0]) == np.isnan(npy(want.pred_keypoints))).all()
            assert ok, (seed, cfgd)
            continue
        close(got[0], npy(want.pred_keypoints)); close(got[1], npy(want.pred_peak_values)); close(got[2], npy(want.instance_scores))
        n_checked += 1
    assert n_checked > 50


def test_dataset_target_wrappers_fuzz(R):
    """generate_confmaps / generate_multiconfmaps / generate_pafs / generate_class_maps (the datasets' call sites) on random
    frames: the num_instances slice, is_centroids, the in-image instance filter, flatten_channels, sigma x stride."""
    from oracle import identity as oid

    for seed in range(16):
        g = torch.Generator().manual_seed(700 + seed)
        h, w = 48, 64
        I, N = int(torch.randint(1, 5, (1,), generator=g)), int(torch.randint(2, 5, (1,), generator=g))
        inst = torch.rand((1, I, N, 2), generator=g) * torch.tensor([w + 16.0, h + 16.0]) - 8.0   # some nodes outside
        if seed % 3 == 0:
            inst[0, 0, 1] = float("nan")
        n_use = int(torch.randint(0, I + 1, (1,), generator=g))
        stride, sigma = [1, 2, 4][seed % 3], [1.0, 1.5, 3.0][seed % 3]
        tol = dict(rtol=1e-5, atol=FLT_MIN)
        close(npy(otgt.generate_multiconfmaps(inst, (h, w), n_use, sigma, stride)),
              npy(R.confidence_maps.generate_multiconfmaps(inst, (h, w), n_use, sigma, stride)), **tol)
        close(npy(otgt.generate_multiconfmaps(inst[:, :, 0, :], (h, w), n_use, sigma, stride, is_centroids=True)),
              npy(R.confidence_maps.generate_multiconfmaps(inst[:, :, 0, :], (h, w), n_use, sigma, stride, is_centroids=True)), **tol)
        close(npy(otgt.generate_confmaps(inst[:, 0], (h, w), sigma, stride)),
              npy(R.confidence_maps.generate_confmaps(inst[:, 0], (h, w), sigma, stride)), **tol)
        edges = torch.tensor([(k - 1, k) for k in range(1, N)])
        for flat in (False, True):
            close(npy(otgt.generate_pafs(inst, (h, w), sigma, stride, edges, flat)),
                  npy(R.edge_maps.generate_pafs(inst, (h, w), sigma, stride, edges, flat)), rtol=1e-5, atol=1e-6)
        if n_use > 0:
            tracks = torch.randint(-1, 3, (n_use,), generator=g).to(torch.float32)
            close(npy(oid.generate_class_maps(inst, (h, w), n_use, tracks, 3, 0.2, sigma, stride)),
                  npy(R.data_identity.generate_class_maps(inst, (h, w), n_use, tracks, 3, 0.2, sigma, stride)), rtol=1e-6, atol=1e-7)


def interp_cases():
    """(x, y, xnew) triples shared with tests/test_paf_gpu.py::test_interp1d_vs_oracle."""
    for seed in range(24):
        g = torch.Generator().manual_seed(800 + seed)
        M, n_k, n_q = 5, int(torch.randint(2, 6, (1,), generator=g)), 7
        x = torch.sort(torch.rand((M, n_k), generator=g) * 10, dim=1).values
        y = torch.randn((M, n_k), generator=g)
        xq = torch.rand((M, n_q), generator=g) * 14 - 2          # inside and outside the knots: extrapolation
        if seed % 4 == 1:
            x, y = x[0], y[0]                                    # one shared knot row, several query rows (flattened)
        elif seed % 4 == 2:
            x, y, xq = x[0], y[0], xq[0]                         # all 1-D
        elif seed % 4 == 3:
            x = x[:1]                                            # x broadcast over the rows of y
        yield x, y, xq


def test_interp1d_fuzz(R):
    """interp1d with 2..5 knots, every broadcasting mode, query points inside and outside the knot range."""
    from oracle.interp import interp1d

    for x, y, xq in interp_cases():
        want = R.interp.interp1d(x, y, xq)
        got = interp1d(x, y, xq)
        assert tuple(got.shape) == tuple(want.shape)
        eq(npy(got), npy(want))


def multiclass_cases():
    """Random small frames for the multi-class bottom-up layer, shared with tests/test_identity_gpu.py."""
    for seed in range(24):
        g = torch.Generator().manual_seed(900 + seed)
        B, Nn, K, H, W = 2, int(torch.randint(1, 4, (1,), generator=g)), int(torch.randint(1, 5, (1,), generator=g)), 24, 32
        cms = torch.rand((B, Nn, H, W), generator=g) ** 6          # sparse bright pixels -> a few peaks per node
        if seed % 5 == 0:
            cms[1] = 0.0                                           # a frame without peaks
        cs = [1, 2, 4][seed % 3]
        class_maps = torch.softmax(torch.randn((B, K, H * 2 // cs, W * 2 // cs), generator=g), dim=1)
        scale = [1.0, 0.5, 2.0][seed % 3]
        eff = [torch.ones(B), torch.tensor([0.8, 1.25])][seed % 2]
        cap = [None, 1, 2, 3][seed % 4]
        yield cms, class_maps, cs, scale, eff, cap


def test_bottomup_multiclass_layer_fuzz(R):
    """BottomUpMultiClassLayer.postprocess (unmodified class, stand-in self) vs the oracle on random small frames: noisy
    confidence maps (several peaks per node, some frames empty), random class maps, scales and instance caps."""
    import types

    from oracle import identity as oid

    L = R.bottomup_multiclass.BottomUpMultiClassLayer
    P = R.preprocess_info.PreprocInfo
    for cms, class_maps, cs, scale, eff, cap in multiclass_cases():
        cfg = types.SimpleNamespace(peak_threshold=0.3, effective_refinement="integral", integral_patch_size=5,
                                    max_instances=None, return_confmaps=False, return_class_maps=False)
        me = types.SimpleNamespace(postprocess_config=cfg, cms_output_stride=2, class_maps_output_stride=cs,
                                   max_instances=cap, _cap_instances_by_score=L._cap_instances_by_score)
        o = L.postprocess(me, {"MultiInstanceConfmapsHead": cms, "ClassMapsHead": class_maps}, P(eff_scale=eff, input_scale=scale))
        inst, pv, sc, tr = oid.bottomup_multiclass_postprocess(cms, class_maps, 2, cs, scale, eff, cap, threshold=0.3)
        eq(np.isnan(npy(inst)), np.isnan(npy(o.pred_keypoints)))
        close(npy(inst), npy(o.pred_keypoints), atol=1e-5)
        eq(npy(pv), npy(o.pred_peak_values))
        close(npy(sc), npy(o.instance_scores), rtol=1e-6, atol=1e-7)   # nanmean reduction order
        close(npy(tr), npy(o.instance_tracking_scores), rtol=1e-6, atol=1e-7)


def test_labels_level_nms_cores_fuzz(R):
    """_nms_greedy_iou / _nms_greedy_oks / _compute_oks / _compute_iou_one_to_many (ops/filters.py:336-495, float64 numpy)
    on random instances with NaN nodes and NaN scores."""
    from oracle import filters as ofil

    OF = R.ops_filters
    for seed in range(60):
        g = np.random.default_rng(1100 + seed)
        n, nn = int(g.integers(0, 9)), int(g.integers(1, 6))
        pts = [g.uniform(0, 50, (nn, 2)) + (0 if seed % 2 else g.uniform(0, 4)) for _ in range(n)]
        if seed % 2 and n > 1:
            pts = [pts[0] + g.normal(0, 2.0, (nn, 2)) for _ in range(n)]       # heavy overlap
        for p in pts:
            p[g.random(nn) < 0.2] = np.nan
        scores = g.uniform(0, 1, n)
        # (no exact score ties: `scores.argsort()[::-1]` is numpy's default introsort / AVX-512 sort, whose order of equal
        #  keys is platform dependent - measured here: [0, 1] where the classic insertion-sort path gives [1, 0]; the
        #  kernels document "equal keys by descending index", the classic behaviour, and parity is defined on tie-free input)
        if n > 1 and seed % 5 == 0:
            scores[-1] = np.nan
        thr = float(g.uniform(0.05, 0.7))
        boxes = np.array([ofil.instance_bbox64(p) for p in pts]).reshape(n, 4)
        assert OF._nms_greedy_iou(boxes, scores, thr) == ofil.nms_greedy64(pts, scores, thr, "iou")
        assert OF._nms_greedy_oks(pts, scores, thr) == ofil.nms_greedy64(pts, scores, thr, "oks")
        for i in range(min(n, 3)):
            for j in range(min(n, 3)):
                close(ofil.oks64(pts[i], pts[j]), OF._compute_oks(pts[i], pts[j]), rtol=1e-12, atol=1e-15)


def single_stage_cases():
    """Random confidence maps for the centroid / centred-instance / single-instance layers, shared with the GPU tests."""
    for seed in range(24):
        g = torch.Generator().manual_seed(1200 + seed)
        B, C = int(torch.randint(1, 5, (1,), generator=g)), int(torch.randint(1, 4, (1,), generator=g))
        cms = torch.rand((B, C, 28, 36), generator=g) ** 5
        if seed % 6 == 0:
            cms[0] = 0.0
        stride = [1, 2, 4][seed % 3]
        scale = [1.0, 0.5, 1.5][(seed // 3) % 3]
        eff = torch.ones(B) if seed % 2 else torch.rand((B,), generator=g) * 0.6 + 0.7
        cap = [None, 1, 3, 40][seed % 4]
        yield cms, stride, scale, eff, cap


def test_single_stage_layers_fuzz(R):
    """CentroidLayer / CenteredInstanceLayer / SingleInstanceLayer .postprocess (unmodified classes, stand-in self) vs
    oracle.layers on random maps: inferred and fixed max_instances (top-k truncation, NaN padding), every ladder step."""
    import types

    from oracle import layers as olay

    P = R.preprocess_info.PreprocInfo
    CL, CI, SI = R.centroid.CentroidLayer, R.centered_instance.CenteredInstanceLayer, R.single_instance.SingleInstanceLayer
    for cms, stride, scale, eff, cap in single_stage_cases():
        cfg = types.SimpleNamespace(peak_threshold=0.25, effective_refinement="integral", integral_patch_size=5,
                                    return_confmaps=False, max_instances=None)
        info = P(eff_scale=eff, input_scale=scale, output_stride=stride)
        me = types.SimpleNamespace(postprocess_config=cfg, max_instances=cap, _extract_confmaps=lambda raw: raw["x"],
                                   _infer_max_instances=CL._infer_max_instances)
        o = CL.postprocess(me, {"x": cms[:, :1]}, info)
        xy, val = olay.centroid_postprocess(cms[:, :1], stride, scale, eff, cap, threshold=0.25)
        eq(xy, npy(o.pred_centroids)); eq(val, npy(o.pred_centroid_values))
        for layer in (CI, SI):
            o = layer.postprocess(me, {"x": cms}, info)
            k, v = olay.global_postprocess(cms, stride, scale, eff, threshold=0.25)
            eq(k, npy(o.pred_keypoints)); eq(v, npy(o.pred_peak_values))
