"""GPU parity: confidence-map / PAF target kernels vs reference goldens and the oracle.

Bar (north_star): targets within 1e-5 relative; the absolute floor is FLT_MIN because fp32 exp
in the denormal band (arg in [-104, -87]) has no meaningful relative error.
The ZERO SET must be identical: the kernels skip work exactly where the reference yields 0.
"""

import numpy as np
import pytest
import torch

from tests.helpers import FLT_MIN, T, close, eq, golden, npy

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-5, atol=FLT_MIN)
# Sums over instances (make_multi_pafs) can cancel (+u and -u contributions): the error of each
# term is relative to ITS magnitude (<= 2 ulp of a value <= 1), so the sum needs an absolute floor
# of a few fp32 ulps at unit scale (|paf| within 1.2e-7 abs of make_pafs per instance).
TOL_SUM = dict(rtol=1e-5, atol=1e-6)


@pytest.fixture(scope="module")
def cm():
    from sleap_nn_b200.data import confidence_maps

    return confidence_maps


@pytest.fixture(scope="module")
def em():
    from sleap_nn_b200.data import edge_maps

    return edge_maps


def test_confmaps_golden(cm):
    d = golden("ref_targets.npz")
    xv, yv = T(d["xv"]), T(d["yv"])
    got = cm.make_confmaps(T(d["cm_pts"]), xv, yv, 3.0)
    assert got.device.type == "cpu" and got.dtype == torch.float32
    close(npy(got), d["cm"], **TOL)
    close(npy(cm.make_confmaps(T(d["cm_pts"]).cuda(), xv, yv, 3.0)), d["cm"], **TOL)
    close(npy(cm.make_multi_confmaps(T(d["mc_pts"]), xv, yv, 3.0)), d["mc"], **TOL)
    close(npy(cm.generate_multiconfmaps(T(d["mc_pts"]), (48, 64), 3, 1.5, 2)), d["gmc"], **TOL)
    close(npy(cm.generate_multiconfmaps(T(d["mc_pts"])[:, :, 0], (48, 64), 4, 1.5, 2, True)), d["gmc_centroids"], **TOL)
    close(npy(cm.generate_confmaps(T(d["mc_pts"])[:, 0], (48, 64), 1.5, 2)), d["gc"], **TOL)
    far = npy(cm.make_confmaps(T(d["far_pts"]), T(d["xv2"]), T(d["yv2"]), 5.0))
    close(far, d["far_cm"], **TOL)
    eq(far == 0, d["far_cm"] == 0)  # identical zero set (support-window skipping is exact)
    assert (d["far_cm"] == 0).mean() > 0.5 and ((d["far_cm"] > 0) & (d["far_cm"] < FLT_MIN)).any()


def test_edge_maps_and_pafs_golden(em):
    from sleap_nn_b200.data.utils import gaussian_pdf, make_grid_vectors

    d = golden("ref_targets.npz")
    xv3, yv3 = make_grid_vectors(3, 3, 1)
    s3, d3 = T(d["k_src"]), T(d["k_dst"])
    yy, xx = torch.meshgrid(yv3, xv3, indexing="ij")
    dist = em.distance_to_edge(torch.stack((xx, yy), -1), s3, d3)
    eq(npy(dist), d["k_dist"])
    assert npy(dist)[0].tolist() == [[1.25, 0.0], [0.25, 0.5], [1.25, 2.0]]  # tests/data/test_edge_maps.py:16-37
    close(npy(em.make_edge_maps(xv3, yv3, s3, d3, 1.0)), d["k_em"], **TOL)
    close(npy(gaussian_pdf(dist, 1.0)), d["k_em"], **TOL)
    close(npy(em.make_pafs(xv3, yv3, s3, d3, 1.0)), d["k_paf"], **TOL)
    close(npy(em.make_multi_pafs(xv3, yv3, torch.stack([s3, s3]), torch.stack([d3, d3]), 1.0)), d["k_mpaf"], **TOL_SUM)
    np.testing.assert_allclose(npy(em.make_pafs(xv3, yv3, s3, d3, 1.0))[0, 1],
                               [[0.4578, 0.9692, 0.4578], [0.6065, 1.0, 0.6065], [0.4578, 0.9692, 0.4578]], atol=1e-3)
    xv, yv = T(d["xv"]), T(d["yv"])
    src, dst = em.get_edge_points(T(d["pf_inst"]), T(d["pf_edges"]))
    eq(npy(src), d["pf_src"]); eq(npy(dst), d["pf_dst"])
    close(npy(em.make_pafs(xv, yv, src[0], dst[0], 1.5)), d["pf_single"], **TOL)
    deg = npy(em.make_pafs(xv, yv, src[1], dst[1], 1.5))
    close(deg, d["pf_degenerate"], **TOL)  # NaN planes of the src == dst edge are kept
    assert np.isnan(deg).any()
    multi = npy(em.make_multi_pafs(xv, yv, src, dst, 1.5))
    close(multi, d["pf_multi"], **TOL_SUM)
    eq(multi == 0, d["pf_multi"] == 0)
    close(npy(em.make_edge_maps(xv, yv, src[0], dst[0], 1.5)), d["pf_em"], **TOL)
    close(npy(em.generate_pafs(T(d["gp_inst"]), (48, 64), 1.5, 2, T(d["pf_edges"]), True)), d["gp"], **TOL_SUM)


@pytest.mark.parametrize("hw,stride,n_inst,n_nodes", [((96, 130), 2, 3, 6), ((64, 64), 1, 5, 4), ((50, 70), 4, 2, 3)])
def test_targets_vs_oracle_random(cm, em, hw, stride, n_inst, n_nodes):
    """Odd widths (scalar stores), several bands, NaN points; parity + identical zero sets."""
    from oracle import synth, targets as ot

    g = torch.Generator().manual_seed(hw[0] + n_inst)
    pts = torch.rand((1, n_inst, n_nodes, 2), generator=g) * torch.tensor([float(hw[1]), float(hw[0])])
    pts[0, 0, 1] = float("nan")
    xv, yv = ot.grid_vectors(hw[0], hw[1], stride)
    want = ot.multi_confmaps(pts, xv, yv, 1.5 * stride)
    got = npy(cm.make_multi_confmaps(pts.cuda(), xv, yv, 1.5 * stride))
    close(got, npy(want), **TOL); eq(got == 0, npy(want) == 0)
    edges = torch.tensor(synth.chain_edges(n_nodes))
    s, d = ot.edge_points(pts[0], edges)
    want = ot.multi_pafs(xv, yv, s, d, 2.5)
    got = npy(em.make_multi_pafs(xv, yv, s.cuda(), d.cuda(), 2.5))
    close(got, npy(want), **TOL_SUM); eq(got == 0, npy(want) == 0)
    want1 = ot.pafs(xv, yv, s[1], d[1], 2.5)
    close(npy(em.make_pafs(xv, yv, s[1], d[1], 2.5)), npy(want1), **TOL)


def test_multi_sample_quirk_and_bf16(cm, em):
    from oracle import targets as ot

    g = torch.Generator().manual_seed(1)
    pts = torch.rand((2, 2, 3, 2), generator=g) * 40
    xv, yv = ot.grid_vectors(40, 40, 2)
    want = ot.multi_confmaps(pts, xv, yv, 3.0)
    got = cm.make_multi_confmaps(pts, xv, yv, 3.0)
    assert got.shape == (2, 3, 20, 20)
    close(npy(got), npy(want), **TOL)  # both samples hold the reduction over ALL samples (reference quirk)
    b = cm.make_multi_confmaps(pts.cuda(), xv, yv, 3.0, out_dtype=torch.bfloat16)
    assert b.dtype == torch.bfloat16
    close(npy(b.float()), npy(want.bfloat16().float()), rtol=1e-2, atol=1e-3)


def test_full_size_flies_frame(cm, em):
    """BASELINE cfg4 size: 32 nodes / 31 edges / 8 instances, 1024^2 image, stride 2 (512^2 maps).

    Size-independent properties: every planted keypoint is the arg-max of its channel's blob
    (value 1 at a grid-aligned point), maps are bounded by [0, 1], PAF magnitude along an edge's
    midpoint is ~1 in the edge direction, and the result is idempotent (bit-identical re-run).
    """
    from oracle import synth

    n_nodes, n_inst, hw, stride = 32, 8, (1024, 1024), 2
    edges = synth.chain_edges(n_nodes)
    poses = synth.make_poses(7, 1, n_inst, n_nodes, hw, margin=150.0, step=30.0, edges=edges)
    poses = (poses / stride).round() * stride  # grid aligned -> peak value exactly exp(0) = 1
    maps = cm.generate_multiconfmaps(poses.cuda(), hw, n_inst, sigma=2.5, output_stride=stride)
    assert maps.shape == (1, 32, 512, 512) and maps.is_cuda
    m = maps[0]
    assert float(m.max()) == 1.0 and float(m.min()) >= 0.0
    idx = (poses[0] / stride).long()
    vals = m[torch.arange(n_nodes)[None, :].expand(n_inst, -1).cuda(), idx[..., 1].cuda(), idx[..., 0].cuda()]
    assert bool((vals == 1.0).all())
    assert torch.equal(maps, cm.generate_multiconfmaps(poses.cuda(), hw, n_inst, sigma=2.5, output_stride=stride))
    pafs = em.generate_pafs(poses.cuda(), hw, sigma=2.5, output_stride=stride, edge_inds=torch.tensor(edges))
    assert pafs.shape == (31, 2, 512, 512)
    src, dst = poses[0, 0, 0], poses[0, 0, 1]
    mid = ((src + dst) / 2 / stride).round().long()
    v = pafs[0, :, mid[1], mid[0]].cpu()
    u = (dst - src) / torch.linalg.vector_norm(dst - src)
    assert float((v * u).sum()) > 0.5
    assert float((pafs == 0).float().mean()) > 0.9  # narrow support: almost all zeros


def test_hoisted_reciprocal_division_is_bit_exact():
    """K7 replaces the per-pixel div.rn by the same FMA sequence with the divisor-only part hoisted; it must agree
    with __fdiv_rn bit for bit over the whole range the kernel feeds it ([0, 105 * den]) for awkward divisors."""
    from sleap_nn_b200 import _native as N

    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    n = 1 << 24
    dens = [50.0, 12.5, 4.5, 18.0, 8.0, 2.0, 1.0, 3.0, 0.02, 7.0e-5, 1.2345678e7, float(np.float32(1.9999999)),
            float(np.nextafter(np.float32(2.0), np.float32(3.0))), float(np.float32(1.0000001)), 1e-30, 3e30,
            float(np.float32(0.99999994)), 5.0e-38, 1.0e38]
    for den in dens:
        scale = min(105.0 * den, 3e38)
        a = torch.rand((n,), generator=g, device=dev) * scale
        a[: 1 << 20] = a[: 1 << 20] * 1e-6           # small numerators
        lo = 1 << 20
        a[lo : 2 * lo] = torch.exp(torch.rand((lo,), generator=g, device=dev) * 101.0 - 101.0).clamp(max=scale)  # 1e-44 .. 1
        a[2 * lo : 2 * lo + 4] = torch.tensor([0.0, scale, 1e-45, 1.1754944e-38], device=dev)
        fast, exact = torch.empty_like(a), torch.empty_like(a)
        N.check(N.lib.snb_debug_neg_div(N.ptr(a), n, den, N.ptr(fast), N.ptr(exact), N.stream_ptr(dev)), "dbg")
        # a quotient below 2^-30 only ever feeds exp(), which is exactly 1 there; everything else must agree exactly
        matters = exact.abs() >= 2.0 ** -30
        bad = (fast.view(torch.int32) != exact.view(torch.int32)) & matters
        assert int(bad.sum()) == 0, (den, int(bad.sum()), a[bad][:4].tolist())
        assert bool((torch.exp(fast[~matters]) == 1).all()) and bool((torch.exp(exact[~matters]) == 1).all())
