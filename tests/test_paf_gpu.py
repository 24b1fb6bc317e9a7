"""GPU parity: PAF scoring / matching / assembly kernels vs reference goldens and the oracle.

The first block restates the reference's own unit tests (tests/inference/test_paf_grouping.py)
against the CUDA-backed API, so it reads like the reference's suite.  Bars: candidate sets,
line subscripts, matches and instance grouping bit-exact; line scores within 1e-5 relative
with a 1e-6 absolute floor (north_star + SURVEY section 7 "Score numerics").
"""

import numpy as np
import pytest
import torch
from numpy.testing import assert_array_equal
from torch.testing import assert_close

from tests.helpers import T, candidate_table, close, eq, golden, npy, ragged

pytestmark = pytest.mark.gpu

SCORE_RTOL, SCORE_ATOL = 1e-5, 1e-6


@pytest.fixture(scope="module")
def pg():
    from sleap_nn_b200.inference import paf_grouping

    return paf_grouping


# ----------------------------------------------------- the reference's unit tests, restated
def test_get_connection_candidates(pg):
    edge_inds, edge_peak_inds = pg.get_connection_candidates(
        torch.tensor([0, 0, 0, 1, 1, 2]), torch.tensor([[0, 1], [1, 2], [2, 3]]), 4
    )
    assert edge_inds.numpy().tolist() == [0, 0, 0, 0, 0, 0, 1, 1]
    assert edge_peak_inds.numpy().tolist() == [[0, 3], [0, 4], [1, 3], [1, 4], [2, 3], [2, 4], [3, 5], [4, 5]]
    assert edge_inds.dtype == torch.int32 and edge_peak_inds.dtype == torch.int64


def test_make_line_subs(pg):
    line_subs = pg.make_line_subs(
        torch.tensor([[0, 0], [4, 8]], dtype=torch.float32), torch.tensor([[0, 1]], dtype=torch.int32),
        torch.tensor([0], dtype=torch.int32), n_line_points=3, pafs_stride=2, pafs_hw=(9, 9),
    )
    assert line_subs.numpy().tolist() == [[[[0, 0, 0], [0, 0, 1]], [[2, 1, 0], [2, 1, 1]], [[4, 2, 0], [4, 2, 1]]]]


def test_get_paf_lines_and_score(pg):
    pafs_sample = torch.arange(6 * 4 * 2).view(6, 4, 2).float()
    peaks_sample = torch.tensor([[0, 0], [4, 8]], dtype=torch.float32)
    epi = torch.tensor([[0, 1]], dtype=torch.int32)
    ei = torch.tensor([0], dtype=torch.int32)
    paf_lines = pg.get_paf_lines(pafs_sample, peaks_sample, epi, ei, n_line_points=3, pafs_stride=2)
    assert paf_lines.numpy().tolist() == [[[0, 1], [18, 19], [36, 37]]]
    scores = pg.score_paf_lines(paf_lines, peaks_sample, epi, max_edge_length=2)
    assert_close(scores, torch.tensor([24.27]), atol=1e-2, rtol=1e-2)
    d = golden("ref_paf_units.npz")
    close(npy(scores), d["ut_score"], rtol=SCORE_RTOL, atol=SCORE_ATOL)


def test_compute_distance_penalty(pg):
    p1 = pg.compute_distance_penalty(torch.tensor([1, 2, 3, 4], dtype=torch.float32), max_edge_length=2)
    assert_close(p1, torch.tensor([0, 0, 2 / 3 - 1, 2 / 4 - 1]), atol=1e-6, rtol=1e-6)
    p2 = pg.compute_distance_penalty(torch.tensor([1, 2, 3, 4], dtype=torch.float32), max_edge_length=2,
                                     dist_penalty_weight=2)
    assert_close(p2, torch.tensor([0, 0, -0.6666666, -1]), atol=1e-6, rtol=1e-6)
    eq(npy(p2), golden("ref_paf_units.npz")["ut_pen"])


def test_score_paf_lines_batch(pg):
    pafs = torch.arange(6 * 4 * 2, dtype=torch.float32).reshape(1, 6, 4, 2)
    peaks = [torch.tensor([[0, 0], [4, 8]], dtype=torch.float32)]
    ch = [torch.tensor([0, 1], dtype=torch.int32)]
    edges = torch.tensor([[0, 1], [1, 2], [2, 3]], dtype=torch.int32)
    edge_inds, edge_peak_inds, line_scores = pg.score_paf_lines_batch(pafs, peaks, ch, edges, 3, 2, 2 / 12, 1.0, 4)
    assert len(edge_inds) == 1 and edge_inds[0].numpy().tolist() == [0]
    assert edge_peak_inds[0].numpy().tolist() == [[0, 1]]
    assert_close(line_scores[0], torch.tensor([24.27]), rtol=8e-2, atol=8e-2)


def test_match_candidates_sample_and_batch(pg):
    e = torch.tensor([0, 0], dtype=torch.int32)
    p = torch.tensor([[0, 1], [2, 1]], dtype=torch.int32)
    s = torch.tensor([-0.5, 1.0], dtype=torch.float32)
    m = pg.match_candidates_sample(e, p, s, 1)
    assert_array_equal(m[0], [0]); assert_array_equal(m[1], [1]); assert_array_equal(m[2], [0]); assert_array_equal(m[3], [1.0])
    assert all(t.device.type == "cpu" for t in m) and m[0].dtype == torch.int32 and m[3].dtype == torch.float32
    mb = pg.match_candidates_batch([e], [p], [s], 1)
    assert [x[0].numpy().tolist() for x in mb] == [[0], [1], [0], [1.0]]


def test_toposort_edges(pg):
    d = golden("ref_assembly.npz")
    for i in range(int(d["n_topo"])):
        et = [pg.EdgeType(a, b) for a, b in d[f"topo{i}_edges"].tolist()]
        assert list(pg.toposort_edges(et)) == d[f"topo{i}_order"].tolist()


def test_assign_connections_to_instances(pg):
    EdgeType, EdgeConnection, PeakID = pg.EdgeType, pg.EdgeConnection, pg.PeakID
    connections = {
        EdgeType(5, 7): [EdgeConnection(0, 0, 1.0465653)], EdgeType(5, 8): [EdgeConnection(0, 0, 1.0607507)],
        EdgeType(5, 9): [EdgeConnection(0, 0, 0.9563284)], EdgeType(5, 6): [EdgeConnection(0, 1, 0.5797864)],
        EdgeType(5, 11): [EdgeConnection(0, 0, 0.9892818)], EdgeType(5, 12): [EdgeConnection(0, 0, 0.7557168)],
        EdgeType(1, 0): [], EdgeType(1, 3): [], EdgeType(1, 2): [], EdgeType(1, 10): [], EdgeType(1, 13): [],
        EdgeType(1, 14): [], EdgeType(4, 5): [EdgeConnection(0, 0, 0.9735552)],
        EdgeType(4, 1): [EdgeConnection(0, 0, 0.31536198)],
    }
    got = pg.assign_connections_to_instances(connections, min_instance_peaks=0, n_nodes=15)
    want = {PeakID(5, 0): 0, PeakID(7, 0): 0, PeakID(8, 0): 0, PeakID(9, 0): 0, PeakID(6, 1): 0, PeakID(11, 0): 0,
            PeakID(12, 0): 0, PeakID(4, 0): 1, PeakID(1, 0): 1}
    assert got == want and list(got) == list(want)  # same mapping AND same insertion order
    edge_types = list(connections.keys())
    order = pg.toposort_edges(edge_types)
    got = pg.assign_connections_to_instances({edge_types[i]: connections[edge_types[i]] for i in order}, 0, 15)
    assert all(x == 0 for x in got.values())
    connections = {
        EdgeType(0, 1): [EdgeConnection(0, 0, 1.0), EdgeConnection(1, 1, 1.0)],
        EdgeType(1, 2): [EdgeConnection(0, 0, 1.0)],
        EdgeType(2, 3): [EdgeConnection(1, 1, 1.0)],
    }
    got = pg.assign_connections_to_instances(connections, min_instance_peaks=0.5, n_nodes=4)
    assert got == {PeakID(0, 0): 0, PeakID(1, 0): 0, PeakID(0, 1): 1, PeakID(1, 1): 1, PeakID(2, 0): 0,
                   PeakID(2, 1): 2, PeakID(3, 1): 2}


def test_make_predicted_instances(pg):
    peaks = np.array([[[0, 0], [1, 1]], [[2, 2], [3, 3]]])
    peak_scores = np.array([[0.9, 0.8], [0.7, 0.6]])
    connections = {pg.EdgeType(0, 1): [pg.EdgeConnection(0, 0, 0.5), pg.EdgeConnection(1, 1, 0.4)]}
    assign = {pg.PeakID(0, 0): 0, pg.PeakID(0, 1): 1, pg.PeakID(1, 0): 0, pg.PeakID(1, 1): 1}
    inst, pv, sc = pg.make_predicted_instances(peaks, peak_scores, connections, assign)
    np.testing.assert_array_almost_equal(inst, np.array([[[0, 0], [2, 2]], [[1, 1], [3, 3]]]))
    np.testing.assert_array_almost_equal(pv, np.array([[0.9, 0.7], [0.8, 0.6]]))
    np.testing.assert_array_almost_equal(sc, np.array([0.5, 0.4]))


def _group_inputs():
    return dict(
        peaks=torch.arange(10, dtype=torch.float32).reshape(5, 2), vals=torch.arange(5, dtype=torch.float32),
        ch=torch.tensor([0, 1, 2, 0, 1], dtype=torch.int32), me=torch.tensor([0, 1, 0], dtype=torch.int32),
        ms=torch.tensor([0, 0, 1], dtype=torch.int32), md=torch.tensor([0, 0, 1], dtype=torch.int32),
        msc=torch.ones(3, dtype=torch.float32),
    )


def test_group_instances_sample_batch_and_scorer(pg):
    g = _group_inputs()
    et = [pg.EdgeType(0, 1), pg.EdgeType(1, 2)]
    want_inst = [[[0.0, 1.0], [2.0, 3.0], [4.0, 5.0]], [[6.0, 7.0], [8.0, 9.0], [np.nan, np.nan]]]
    inst, pv, sc = pg.group_instances_sample(g["peaks"], g["vals"], g["ch"], g["me"], g["ms"], g["md"], g["msc"], 3,
                                             (0, 1), et, 0)
    assert isinstance(inst, np.ndarray)
    assert_array_equal(inst, want_inst)
    assert_array_equal(pv, [[0.0, 1.0, 2.0], [3.0, 4.0, np.nan]])
    assert_array_equal(sc, [2.0, 1.0])
    b = pg.group_instances_batch([g["peaks"]], [g["vals"]], [g["ch"]], [g["me"]], [g["ms"]], [g["md"]], [g["msc"]], 3,
                                 (0, 1), et, 0)
    assert all(isinstance(x, list) and len(x) == 1 for x in b)
    assert_array_equal(b[0][0].numpy(), want_inst)
    assert_array_equal(b[2][0].numpy(), [2.0, 1.0])

    class Cfg:  # attribute-access stand-in for the OmegaConf head config
        class confmaps:
            part_names = ["a", "b"]

        class pafs:
            edges = [("a", "b")]
            output_stride = 1

    scorer = pg.PAFScorer.from_config(config=Cfg)
    assert scorer and scorer.n_nodes == 2 and scorer.sorted_edge_inds == (0,)
    scorer.n_nodes, scorer.sorted_edge_inds, scorer.edge_types, scorer.min_instance_peaks = 3, (0, 1), et, 0
    r = scorer.group_instances([g["peaks"]], [g["vals"]], [g["ch"]], [g["me"]], [g["ms"]], [g["md"]], [g["msc"]])
    assert_array_equal(r[0][0].numpy(), want_inst)
    # fields are read at call time (tests/inference/test_paf_grouping.py:498-500)
    sc2 = pg.PAFScorer.from_config(config=Cfg, max_edge_length_ratio=2 / 12, n_points=3)
    sc2.edge_inds = torch.tensor([[0, 1], [1, 2], [2, 3]], dtype=torch.int32)
    sc2.pafs_stride, sc2.n_nodes = 2, 4
    e, p, s = sc2.score_paf_lines(torch.arange(48, dtype=torch.float32).reshape(1, 6, 4, 2),
                                  [torch.tensor([[0, 0], [4, 8]], dtype=torch.float32)],
                                  [torch.tensor([0, 1], dtype=torch.int32)])
    assert e[0].numpy().tolist() == [0] and p[0].numpy().tolist() == [[0, 1]]
    assert_close(s[0], torch.tensor([24.27]), rtol=8e-2, atol=8e-2)
    sc2.n_edges = 1
    m = sc2.match_candidates([torch.tensor([0, 0], dtype=torch.int32)], [torch.tensor([[0, 1], [2, 1]])],
                             [torch.tensor([-0.5, 1.0])])
    assert [x[0].numpy().tolist() for x in m] == [[0], [1], [0], [1.0]]


# ------------------------------------------------------------------- goldens from the reference
def test_line_subscripts_bit_exact(pg):
    d = golden("ref_paf_units.npz")
    for stride, n in ((2, 10), (4, 10), (8, 5), (3, 7), (1, 3)):
        hw = (200 // stride, 300 // stride)
        got = pg.make_line_subs(T(d["ls_peaks"]).cuda(), T(d["ls_epi"]).cuda(), T(d["ls_ei"]).cuda(), n, stride, hw)
        assert got.dtype == torch.int32 and got.is_cuda
        eq(npy(got), d[f"ls_s{stride}_n{n}"])


def test_scores_and_matches_random_field(pg):
    d = golden("ref_paf_units.npz")
    field = T(d["sc_field"]).cuda()
    nchw = field.permute(2, 0, 1).contiguous()[None]          # (1, 2E, H, W) as a model would emit
    view = nchw.permute(0, 2, 3, 1)                            # the (B,H,W,2E) VIEW, layers/bottomup.py:103
    assert not view.is_contiguous()
    e, p, s = pg.score_paf_lines_batch(view, [T(d["sc_peaks"]).cuda()], [T(d["sc_ch"]).cuda()], d["sc_edges"].tolist(),
                                       10, 2, 0.25, 1.0, 6)
    assert s[0].is_cuda
    ref_tab = candidate_table(d["sc_ei"], d["sc_epi"], d["sc_scores"])
    got_tab = candidate_table(e[0], p[0], s[0])
    assert ref_tab.keys() == got_tab.keys()
    for k in ref_tab:
        assert abs(ref_tab[k] - got_tab[k]) <= SCORE_ATOL + SCORE_RTOL * abs(ref_tab[k]), k
    m = pg.match_candidates_sample(T(d["sc_ei"]), T(d["sc_epi"]), T(d["sc_scores"]), 5)
    eq(npy(m[0]), d["mt_e"]); eq(npy(m[1]), d["mt_s"]); eq(npy(m[2]), d["mt_d"]); eq(npy(m[3]), d["mt_sc"])
    m = pg.match_candidates_sample(e[0], p[0], s[0], 5)       # from our own (canonical-order) candidates
    eq(npy(m[0]), d["mt_e"]); eq(npy(m[1]), d["mt_s"]); eq(npy(m[2]), d["mt_d"])
    close(npy(m[3]), d["mt_sc"], rtol=SCORE_RTOL, atol=SCORE_ATOL)


def test_assembly_cases_bit_exact(pg):
    d = golden("ref_assembly.npz")
    for c in range(int(d["n_cases"])):
        pre = f"c{c}_"
        mip = float(d[pre + "mip"]) if bool(d[pre + "mip_is_float"]) else int(d[pre + "mip"])
        edges = d[pre + "edges"].tolist()
        et = [pg.EdgeType(a, b) for a, b in edges]
        n_nodes = max(n for e in edges for n in e) + 1
        res = pg.group_instances_sample(T(d[pre + "pk"]), T(d[pre + "pv"]), T(d[pre + "ch"]), T(d[pre + "me"]),
                                        T(d[pre + "ms"]), T(d[pre + "md"]), T(d[pre + "msc"]), n_nodes,
                                        tuple(d[pre + "sorted"].tolist()), et, mip, 0.25)
        eq(res[0], d[pre + "inst"]); eq(res[1], d[pre + "inst_pv"]); eq(res[2], d[pre + "inst_sc"])


@pytest.mark.parametrize("name", ["ref_pipeline_mice.npz", "ref_pipeline_tree.npz"])
def test_pipeline_golden(pg, name):
    """Full bottom-up post-processing on rendered frames vs the reference's recorded outputs."""
    from sleap_nn_b200.inference import peak_finding as pf

    d = golden(name)
    cms, pafs = T(d["cms"]).cuda(), T(d["pafs"]).cuda()
    stride, n_nodes = int(d["stride"]), int(d["n_nodes"])
    edges = d["edges"].tolist()
    mip = float(d["min_instance_peaks"])
    mip = int(mip) if mip == int(mip) else mip
    pts, vals, si, ci = pf.find_local_peaks(cms, threshold=0.2, refinement="integral")
    eq(npy(si), d["pk_s"]); eq(npy(ci), d["pk_c"]); eq(npy(vals), d["pk_vals"]); close(npy(pts), d["pk_pts"], atol=2e-5)
    B = cms.shape[0]
    pts = pts * stride
    peaks = [pts[si == b] for b in range(B)]
    pv = [vals[si == b] for b in range(B)]
    pc = [ci[si == b] for b in range(B)]
    scorer = pg.PAFScorer(part_names=[str(i) for i in range(n_nodes)], edges=[(str(a), str(b)) for a, b in edges],
                          pafs_stride=stride, min_instance_peaks=mip)
    assert list(scorer.sorted_edge_inds) == d["sorted_edge_inds"].tolist()
    inst, ipv, isc, e, p, s = scorer.predict(pafs.permute(0, 2, 3, 1), peaks, pv, pc)
    we, wp, ws = ragged(d, "cand_e"), ragged(d, "cand_p"), ragged(d, "cand_s")
    want_inst, want_pv, want_sc = ragged(d, "inst"), ragged(d, "inst_pv"), ragged(d, "inst_sc")
    m = scorer.match_candidates(e, p, s)
    wm = [ragged(d, k) for k in ("m_e", "m_s", "m_d", "m_sc")]
    for b in range(B):
        a, g_ = candidate_table(we[b], wp[b], ws[b]), candidate_table(e[b], p[b], s[b])
        assert a.keys() == g_.keys()
        for k in a:
            assert abs(a[k] - g_[k]) <= SCORE_ATOL + SCORE_RTOL * abs(a[k])
        for j in range(3):
            eq(npy(m[j][b]), npy(wm[j][b]))
        close(npy(m[3][b]), npy(wm[3][b]), rtol=SCORE_RTOL, atol=SCORE_ATOL)
        # instance membership (which nodes are present) bit-exact; coordinates within refine tolerance
        assert inst[b].shape == want_inst[b].shape
        eq(np.isnan(npy(inst[b])), np.isnan(npy(want_inst[b])))
        close(npy(inst[b]), npy(want_inst[b]), atol=1e-4)
        eq(npy(ipv[b]), npy(want_pv[b]))
        close(npy(isc[b]), npy(want_sc[b]), rtol=SCORE_RTOL, atol=1e-6)


# ------------------------------------------------------------------- randomized vs the oracle
def test_lsap_matches_scipy_on_random_matrices(pg):
    """Rectangular, tied, partially-infinite cost matrices through the generic device matcher."""
    from scipy.optimize import linear_sum_assignment

    g = np.random.default_rng(3)
    for trial in range(60):
        ns, nd = int(g.integers(1, 12)), int(g.integers(1, 12))
        if trial == 0:
            ns, nd = 70, 45  # beyond the shared-memory solver size
        scores = g.normal(size=(ns, nd)).astype(np.float32)
        if trial % 3 == 0:
            scores = g.integers(0, 3, size=(ns, nd)).astype(np.float32)  # ties
        if trial % 4 == 0:
            scores[g.integers(0, ns), g.integers(0, nd)] = np.nan  # NaN -> +inf cost
        src = np.repeat(np.arange(ns) * 3 + 1, nd)  # sparse peak ids: ranks != ids
        dst = np.tile(np.arange(nd) * 2 + 100, ns)
        cost = -scores.astype(np.float64)
        cost[np.isnan(cost)] = np.inf
        want_r, want_c = linear_sum_assignment(cost)
        m = pg.match_candidates_sample(torch.zeros(ns * nd, dtype=torch.int32),
                                       torch.from_numpy(np.stack([src, dst], 1)), torch.from_numpy(scores.reshape(-1)), 1)
        eq(npy(m[1]), want_r.astype(np.int32)); eq(npy(m[2]), want_c.astype(np.int32))
        eq(npy(m[3]), (-cost[want_r, want_c]).astype(np.float32))
    with pytest.raises(ValueError, match="infeasible"):
        pg.match_candidates_sample(torch.zeros(2, dtype=torch.int32), torch.tensor([[0, 2], [1, 2]]),
                                   torch.tensor([float("nan"), float("nan")]), 1)


def test_empty_inputs(pg):
    e, p = pg.get_connection_candidates(torch.zeros(0, dtype=torch.int32), torch.tensor([[0, 1]]), 2)
    assert e.shape == (0,) and p.shape == (0, 2)
    pafs = torch.zeros((2, 8, 8, 2), device="cuda")
    z2, z1, zi = torch.zeros((0, 2)), torch.zeros((0,)), torch.zeros((0,), dtype=torch.int32)
    scorer = pg.PAFScorer(part_names=["a", "b"], edges=[("a", "b")], pafs_stride=2)
    out = scorer.predict(pafs, [z2, z2], [z1, z1], [zi, zi])
    assert [tuple(t.shape) for t in out[0]] == [(0, 2, 2), (0, 2, 2)]
    assert [tuple(t.shape) for t in out[3]] == [(0,), (0,)] and [tuple(t.shape) for t in out[4]] == [(0, 2), (0, 2)]


def test_interp1d(pg):
    from sleap_nn_b200.inference.utils import interp1d

    g = torch.Generator().manual_seed(0)
    y = torch.rand((7, 2), generator=g) * 100
    x = torch.tensor([0, 1]).repeat(7, 1)
    t = torch.linspace(0, 1, 10).repeat(7, 1)
    got = interp1d(x, y, t)
    slope = (y[:, 1:] - y[:, :1]) / (torch.tensor(torch.finfo(torch.float32).eps) + 1)
    eq(npy(got), npy(y[:, :1] + slope * t))
    xs = torch.tensor([0.0, 1.0, 2.5, 4.0])
    ys = torch.tensor([1.0, 3.0, 0.0, 8.0])
    q = torch.tensor([-1.0, 0.5, 2.5, 3.0, 9.0])
    want = np.array([-1.0, 2.0, 0.0, 8 / 3, 8 + 5 * 16 / 3], np.float32)  # linear extrapolation at both ends
    close(npy(interp1d(xs, ys, q)), want, rtol=1e-6, atol=1e-6)  # eps in the slope denominator


def test_warp_lsap_structured_matcher_matches_scipy(pg):
    """The warp-cooperative solver (one free column per lane, dims <= 32) through snb_match_structured:
    rectangular, heavily tied and partially infinite problems, 200 frames in one launch, vs scipy."""
    from scipy.optimize import linear_sum_assignment

    from sleap_nn_b200 import _native as N

    dev = torch.device("cuda", 0)
    g = np.random.default_rng(11)
    B = 200
    dims = [(int(g.integers(1, 33)), int(g.integers(1, 33))) for _ in range(B)]
    dims[0], dims[1], dims[2] = (32, 32), (1, 32), (32, 1)
    cand_stride, match_stride = 32 * 32, 32
    score = np.zeros((B, cand_stride), np.float32)
    node_start = np.zeros((B, 3), np.int32)
    edge_off = np.zeros((B, 2), np.int32)
    match_off = np.zeros((B, 2), np.int32)
    want = []
    for b, (ns, nd) in enumerate(dims):
        s = g.normal(size=(ns, nd)).astype(np.float32)
        if b % 3 == 0:
            s = g.integers(0, 3, size=(ns, nd)).astype(np.float32)  # many ties
        if b % 5 == 0 and ns * nd > 1:
            s[g.integers(0, ns), g.integers(0, nd)] = np.nan  # NaN -> +inf cost (still feasible)
        if b % 7 == 0:
            s[:] = 1.0  # everything tied
        score[b, : ns * nd] = s.reshape(-1)
        node_start[b] = [0, ns, ns + nd]
        edge_off[b] = [0, ns * nd]
        match_off[b] = [0, min(ns, nd)]
        cost = -s.astype(np.float64)
        cost[np.isnan(cost)] = np.inf
        want.append(linear_sum_assignment(cost) + (s,))
    t = lambda a: torch.from_numpy(a).to(dev)
    edges = torch.tensor([[0, 1]], dtype=torch.int32, device=dev)
    m_edge = torch.full((B, match_stride), -7, dtype=torch.int32, device=dev)
    m_src, m_dst = torch.full_like(m_edge, -7), torch.full_like(m_edge, -7)
    m_score = torch.zeros((B, match_stride), dtype=torch.float32, device=dev)
    m_count = torch.zeros((B,), dtype=torch.int32, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    d_score, d_ns, d_eo, d_mo = t(score), t(node_start), t(edge_off), t(match_off)
    N.check(N.lib.snb_match_structured(N.ptr(d_score), None, cand_stride, N.ptr(edges), 2, 1, N.ptr(d_ns), N.ptr(d_eo),
                                       N.ptr(d_mo), None, match_stride, None, 32, B, N.ptr(m_edge), N.ptr(m_src),
                                       N.ptr(m_dst), N.ptr(m_score), N.ptr(m_count), N.ptr(status), N.stream_ptr(dev)),
            "snb_match_structured")
    assert int(status.item()) == 0
    ms, md, msc, mc = npy(m_src), npy(m_dst), npy(m_score), npy(m_count)
    for b, (r, c, s) in enumerate(want):
        k = len(r)
        assert mc[b] == k
        eq(ms[b, :k], r.astype(np.int32)); eq(md[b, :k], c.astype(np.int32))
        eq(msc[b, :k], s[r, c])


@pytest.mark.parametrize("seed", range(6))
def test_assembly_randomised_busy_frames_vs_oracle(pg, seed):
    """The chunk-parallel greedy assembly against the (golden-pinned) oracle on inputs the hand-made cases do not
    reach: up to 45 peaks per node (several 32-wide chunks per edge), skeletons with cross edges (instances met again
    through another path: the merge / move-first branch and its fallback), proper matchings, partial matchings and,
    every third case, matchings that reuse a peak (the repeated-peak fallback), int and fractional min_instance_peaks."""
    from oracle import paf as opaf

    g = np.random.default_rng(1000 + seed)
    for case in range(12):
        n_nodes = int(g.integers(3, 9))
        edges = [(int(g.integers(0, k)), k) for k in range(1, n_nodes)]          # a random tree ...
        for _ in range(int(g.integers(0, 3))):                                   # ... plus cross edges
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
            if case % 3 == 2:   # not a matching: peaks may repeat
                src, dst = g.integers(0, n_per[a], n_m), g.integers(0, n_per[b], n_m)
            else:               # a proper (partial) matching, as the assignment produces
                src, dst = g.permutation(n_per[a])[:n_m], g.permutation(n_per[b])[:n_m]
            me += [k] * n_m; ms += src.tolist(); md += dst.tolist(); msc += g.uniform(0.0, 1.0, n_m).tolist()
        if case % 2:            # the match list need not be grouped by edge at the API level
            order = g.permutation(len(me))
            me, ms, md, msc = ([x[i] for i in order] for x in (me, ms, md, msc))
        et = [pg.EdgeType(a, b) for a, b in edges]
        sorted_inds = pg.toposort_edges(et)
        mip = [0, 3, 0.5][case % 3]
        i32 = lambda x: torch.tensor(x, dtype=torch.int32)
        args = (T(pk), T(pv), T(ch), i32(me), i32(ms), i32(md), torch.tensor(msc, dtype=torch.float32))
        try:
            want = opaf.group_sample(*args, n_nodes, sorted_inds, edges, mip, 0.25)
        except (AssertionError, KeyError) as err:  # the reference's sanity check on improper matchings (paf.py:866-873)
            with pytest.raises(type(err)):
                pg.group_instances_sample(*args, n_nodes, sorted_inds, et, mip, 0.25)
            continue
        got = pg.group_instances_sample(*args, n_nodes, sorted_inds, et, mip, 0.25)
        for a_, b_ in zip(got, want):
            eq(a_, b_)


def test_interp1d_vs_oracle():
    """Every broadcasting mode of interp1d (incl. the reference's flat-slope quirk for one row of knots) and queries outside
    the knots, against the oracle that tests/test_oracle_fuzz_vs_reference.py pins to the live reference."""
    from oracle.interp import interp1d as want_fn
    from sleap_nn_b200.inference.utils import interp1d
    from tests.test_oracle_fuzz_vs_reference import interp_cases

    for x, y, xq in interp_cases():
        want = want_fn(x, y, xq)
        got = interp1d(x, y, xq)
        assert tuple(got.shape) == tuple(want.shape)
        eq(npy(got), npy(want))
        eq(npy(interp1d(x.cuda(), y.cuda(), xq.cuda())), npy(want))
