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
