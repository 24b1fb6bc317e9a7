"""GPU parity: CUDA peak-finding kernels vs the reference goldens and the CPU oracle.

Calls go through the reference-shaped Python API, which calls the C ABI (ctypes).
Bars: integer positions, ordering, sample / channel indices and peak values bit-exact;
refined coordinates within 1e-4 px (north_star), asserted here at 2e-5.
"""

import numpy as np
import pytest
import torch

from tests.helpers import T, close, eq, golden, npy

pytestmark = pytest.mark.gpu

REFINE_ATOL = 2e-5


@pytest.fixture(scope="module")
def pf():
    from sleap_nn_b200.inference import peak_finding

    return peak_finding


@pytest.fixture(scope="module")
def opeaks():
    from oracle import peaks

    return peaks


def test_local_peaks_minimal_golden(pf):
    d = golden("ref_peaks_minimal.npz")
    for dev in ("cuda", "cpu"):  # CPU tensors are staged to the GPU and results come back on CPU
        cms = T(d["cms"]).to(dev)
        pts, vals, s, c = pf.find_local_peaks_rough(cms)
        assert pts.device.type == dev and s.dtype == torch.int32 and c.dtype == torch.int32
        eq(npy(pts), d["lr_pts"]); eq(npy(vals), d["lr_vals"]); eq(npy(s), d["lr_s"]); eq(npy(c), d["lr_c"])
        pts, vals, s, c = pf.find_local_peaks(cms, refinement="integral")
        close(npy(pts), d["li_pts"], atol=REFINE_ATOL); eq(npy(vals), d["li_vals"]); eq(npy(c), d["lr_c"])
        pts, *_ = pf.find_local_peaks(cms, refinement="invalid_input")
        eq(npy(pts), d["lr_pts"])
        pts, *_ = pf.find_local_peaks(cms)
        eq(npy(pts), d["lr_pts"])


def test_global_peaks_minimal_golden(pf):
    d = golden("ref_peaks_minimal.npz")
    cms = T(d["cms"]).cuda()
    pts, vals = pf.find_global_peaks_rough(cms, threshold=0.1)
    eq(npy(pts), d["gr_pts"]); eq(npy(vals), d["gr_vals"])
    pts, vals = pf.find_global_peaks(cms, threshold=0.2)
    eq(npy(pts), d["gr_pts"])
    pts, vals = pf.find_global_peaks(cms, threshold=0.2, refinement="invalid_input")
    eq(npy(pts), d["gr_pts"])
    pts, vals = pf.find_global_peaks(cms, threshold=0.2, refinement="integral")
    close(npy(pts), d["gi_pts"], atol=REFINE_ATOL); eq(npy(vals), d["gi_vals"])


def test_crops_integral_dilation_minimal_golden(pf):
    d = golden("ref_peaks_minimal.npz")
    planes = T(d["cms"]).reshape(13, 1, 80, 80).cuda()
    crops = pf.crop_bboxes(planes, T(d["bboxes"]).cuda(), torch.arange(13).cuda())
    assert crops.shape == (13, 1, 5, 5) and crops.dtype == torch.float32
    eq(npy(crops), d["crops"])
    gv = torch.arange(5, dtype=torch.float32) - 2
    dx, dy = pf.integral_regression(crops, xv=gv, yv=gv)
    assert dx.shape == dy.shape == (13, 1)
    close(npy(dx), d["ir_dx"], atol=1e-6); close(npy(dy), d["ir_dy"], atol=1e-6)
    eq(npy(pf.morphological_dilation(planes, torch.ones(3, 3))), d["dil"])


def test_random_maps_nan_thresholds_patch_sizes(pf):
    d = golden("ref_peaks_random.npz")
    cms = T(d["cms"]).cuda()
    for tag, thr in (("t02", 0.2), ("t09", 0.9)):
        r = pf.find_local_peaks_rough(cms, threshold=thr)
        for a, k in zip(r, ("pts", "vals", "s", "c")):
            eq(npy(a), d[f"lr_{tag}_{k}"])
    for size in (3, 4, 5, 7):
        r = pf.find_local_peaks(cms, threshold=0.9, refinement="integral", integral_patch_size=size)
        close(npy(r[0]), d[f"li_t09_p{size}_pts"], atol=REFINE_ATOL)
    clean = T(d["clean"]).cuda()
    for tag, thr in (("t01", 0.1), ("t0999", 0.999)):
        pts, vals = pf.find_global_peaks_rough(clean, threshold=thr)
        eq(npy(pts), d[f"gr_{tag}_pts"]); eq(npy(vals), d[f"gr_{tag}_vals"])
    for size in (3, 4, 5):
        pts, vals = pf.find_global_peaks(clean, threshold=0.2, refinement="integral", integral_patch_size=size)
        close(npy(pts), d[f"gi_p{size}_pts"], atol=REFINE_ATOL); eq(npy(vals), d[f"gi_p{size}_vals"])


def test_crop_rounding_oob_dtypes(pf):
    from sleap_nn_b200.data.instance_cropping import make_centered_bboxes

    d = golden("ref_peaks_random.npz")
    imgs = T(d["imgs"]).cuda()
    for tag, (bh, bw) in (("6x6", (6, 6)), ("5x7", (5, 7))):
        bb = make_centered_bboxes(T(d["cents"]).cuda(), bh, bw)
        eq(npy(bb), d[f"bb_{tag}"])
        eq(npy(pf.crop_bboxes(imgs, bb, T(d["sidx"]).cuda())), d[f"crops_{tag}"])
    # uint8 images (top-down crop pickup) and the empty case
    u8 = (imgs * 255).to(torch.uint8)
    bb = make_centered_bboxes(T(d["cents"]).cuda(), 6, 6)
    got = pf.crop_bboxes(u8, bb, T(d["sidx"]).cuda())
    from oracle import peaks as op

    eq(npy(got), npy(op.crop_patches(u8.cpu(), bb.cpu(), T(d["sidx"]))))
    e = pf.crop_bboxes(torch.zeros(1, 1, 20, 20, device="cuda"), torch.empty(0, 4, 2), torch.empty(0, dtype=torch.long))
    assert e.shape[0] == 0 and e.shape[1] == 1


@pytest.mark.parametrize("shape,thr", [((2, 5, 96, 128), 0.2), ((1, 3, 37, 53), 0.5), ((3, 2, 64, 260), 0.3)])
def test_local_peaks_vs_oracle_random(pf, opeaks, shape, thr):
    """Vector (W % 4 == 0) and scalar (odd W) detect paths, dense random peaks, refinement."""
    g = torch.Generator().manual_seed(sum(shape))
    cms = torch.rand(shape, generator=g)
    want = opeaks.local_peaks(cms, thr, "integral")
    got = pf.find_local_peaks(cms.cuda(), thr, "integral")
    assert got[0].shape == want[0].shape and want[0].shape[0] > 100
    close(npy(got[0]), npy(want[0]), atol=REFINE_ATOL)
    for a, b in zip(got[1:], want[1:]):
        eq(npy(a), npy(b))


def test_local_peaks_strided_view_and_overflow_regrow(pf, opeaks, monkeypatch):
    g = torch.Generator().manual_seed(5)
    base = torch.rand((2, 40, 48, 6), generator=g).cuda()
    view = base.permute(0, 3, 1, 2)  # (B,C,H,W) channels-last VIEW: generic-stride path
    want = opeaks.local_peaks_rough(view.cpu().contiguous(), 0.2)
    import sleap_nn_b200.inference.ops.peaks as P

    monkeypatch.setattr(P, "DEFAULT_PEAK_CAP", 16)  # force the capacity-overflow re-run
    got = pf.find_local_peaks_rough(view, 0.2)
    for a, b in zip(got, want):
        eq(npy(a), npy(b))
    assert want[0].shape[0] > 16 * 2


def test_plateaus_edges_and_empty(pf):
    cms = torch.zeros((1, 1, 8, 8))
    cms[0, 0, 0, 0] = 1.0        # corner peak
    cms[0, 0, 7, 3] = 0.9        # edge peak
    cms[0, 0, 3, 3] = cms[0, 0, 3, 4] = 0.8  # plateau: strict test -> no peak
    cms[0, 0, 5, 6] = 0.2        # == threshold: float32(0.2) > 0.2 is False
    pts, vals, s, c = pf.find_local_peaks_rough(cms.cuda(), 0.2)
    assert npy(pts).tolist() == [[0.0, 0.0], [3.0, 7.0]]
    out = pf.find_local_peaks(torch.zeros((2, 3, 16, 16), device="cuda"), 0.2, "integral")
    assert [tuple(o.shape) for o in out] == [(0, 2), (0,), (0,), (0,)]
    pts, vals = pf.find_global_peaks_rough(torch.zeros((2, 3, 16, 16), device="cuda"), 0.1)
    assert np.isnan(npy(pts)).all() and (npy(vals) == 0).all()


@pytest.mark.parametrize("shape", [(4, 13, 80, 80), (1, 2, 192, 192), (2, 2, 300, 517)])
def test_global_peaks_vs_oracle(pf, opeaks, shape):
    """cfg2 / cfg1 shapes plus a multi-chunk, odd-width plane; ties resolved like torch (first index)."""
    g = torch.Generator().manual_seed(shape[2])
    cms = torch.rand(shape, generator=g)
    cms[0, 0] = 0.01                      # below threshold -> NaN
    cms[0, 1, 5, 7] = cms[0, 1, 2, 9] = 2.0  # tie: x = min col, y = min row (two independent arg-maxes)
    want = opeaks.global_peaks(cms, 0.2, "integral")
    got = pf.find_global_peaks(cms.cuda(), 0.2, "integral")
    close(npy(got[0]), npy(want[0]), atol=REFINE_ATOL); eq(npy(got[1]), npy(want[1]))
    want = opeaks.global_peaks_rough(cms, 0.1)
    got = pf.find_global_peaks_rough(cms.cuda(), 0.1)
    eq(npy(got[0]), npy(want[0])); eq(npy(got[1]), npy(want[1]))
    assert npy(got[0])[0, 1].tolist() == [7.0, 2.0]


def test_global_peaks_small_plane_edge_cases(pf, opeaks):
    """The warp-per-plane kernel (planes <= 16K elements): ties inside one lane's load sequence, across lanes and
    inside one 16-byte word, plateaus, NaN, -inf, a ragged tail (H*W/4 not a multiple of 32) and strided views."""
    g = torch.Generator().manual_seed(7)
    cms = torch.rand((3, 6, 44, 36), generator=g)
    cms[0, 0, 30, 5] = cms[0, 0, 2, 33] = 3.0            # tie in different lanes
    cms[0, 1, 3, 4] = cms[0, 1, 35, 0] = 3.0             # word 31 and word 319 = 31 + 32*9: the same lane
    cms[0, 2, 9, 17] = cms[0, 2, 9, 18] = 3.0            # tie inside one 16-byte word
    cms[0, 3] = 0.7                                      # plateau above the threshold -> (0, 0)
    cms[0, 4, 20, 20] = float("nan")                     # NaN is the greatest value
    cms[0, 5, 10, 3] = float("nan"); cms[0, 5, 4, 30] = float("nan")
    cms[1, 0] = float("-inf")
    cms[1, 1] = 0.05                                     # plateau below the threshold -> NaN, 0
    cms[1, 2, 43, 35] = 9.0                              # last element of the ragged tail
    cms[1, 3, 0, 0] = 9.0
    cms[1, 4, 17, 8] = float("inf")
    for thr in (0.2, float("-inf")):
        want = opeaks.global_peaks_rough(cms, thr)
        got = pf.find_global_peaks_rough(cms.cuda(), thr)
        eq(npy(got[0]), npy(want[0])); eq(npy(got[1]), npy(want[1]))
    want = opeaks.global_peaks(cms, 0.2, "integral")
    got = pf.find_global_peaks(cms.cuda(), 0.2, "integral")
    close(npy(got[0]), npy(want[0]), atol=REFINE_ATOL); eq(npy(got[1]), npy(want[1]))
    # strided views: a row-cropped window (row stride != W) and a channel slice of a wider tensor
    big = torch.rand((2, 8, 50, 48), generator=g)
    big[0, 2, 11, 12] = big[0, 2, 40, 9] = 5.0
    view = big[:, 1:7, 3:47, 8:44]
    assert not view.is_contiguous()
    want = opeaks.global_peaks(view.contiguous(), 0.2, "integral")
    got = pf.find_global_peaks(big.cuda()[:, 1:7, 3:47, 8:44], 0.2, "integral")
    close(npy(got[0]), npy(want[0]), atol=REFINE_ATOL); eq(npy(got[1]), npy(want[1]))


def test_full_size_properties(pf):
    """BASELINE cfg3 map size (5 x 512 x 512 per frame): size-independent properties.

    Planted, well-separated impulses must come back exactly, in (sample, y, x, channel) order,
    with sum(counts) == number planted; refinement of a symmetric blob is the identity.
    """
    B, C, H, W = 8, 5, 512, 512
    g = torch.Generator().manual_seed(0)
    cms = torch.rand((B, C, H, W), generator=g) * 1e-3
    ys = torch.randint(2, H // 8 - 1, (B, C, 6), generator=g) * 8
    xs = torch.randint(2, W // 8 - 1, (B, C, 6), generator=g) * 8
    planted = set()
    for b in range(B):
        for c in range(C):
            for k in range(6):
                y, x = int(ys[b, c, k]), int(xs[b, c, k])
                cms[b, c, y - 1 : y + 2, x - 1 : x + 2] = 0.5
                cms[b, c, y, x] = 0.9
                planted.add((b, y, x, c))
    pts, vals, s, c = pf.find_local_peaks(cms.cuda(), 0.2, "integral")
    got = [(int(b_), int(round(p[1])), int(round(p[0])), int(c_)) for p, b_, c_ in zip(npy(pts), npy(s), npy(c))]
    assert got == sorted(planted)
    assert (npy(vals) == np.float32(0.9)).all()
    ipts = np.asarray([[t[2], t[1]] for t in got], np.float32)
    close(npy(pts), ipts, atol=2e-3)  # symmetric 3x3 blob on a <=1e-3 noise floor: offsets ~ noise only


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("thr", [0.3, 0.2, 0.7])
def test_half_precision_maps_threshold_in_the_maps_dtype(pf, opeaks, dt, thr):
    """fp16 / bf16 maps: `cms > threshold` and `max < threshold` compare in the maps' dtype (0.3 rounds UP to 0.30005 in
    fp16, 0.2 rounds DOWN to 0.19995), so an element equal to the ROUNDED threshold decides differently from a comparison
    with the fp32 threshold.  Integer results and values follow the reference exactly (the oracle runs the same torch ops on
    the half tensor); refined coordinates are computed in fp32 on the exact up-cast values, where the reference accumulates
    in the maps' dtype (differs by up to 4e-4 px fp16 / 3e-3 px bf16)."""
    g = torch.Generator().manual_seed(21)
    cms = torch.rand((2, 3, 40, 48), generator=g).to(dt)
    edge = torch.tensor(thr, dtype=dt)
    cms[0, 0, 8:13, 8:13] = 0.05
    cms[0, 0, 10, 10] = edge                      # an isolated local maximum exactly AT the rounded threshold
    cms[1, 2] = torch.minimum(cms[1, 2], edge)
    cms[1, 2, 5, 7] = edge                        # a plane whose maximum IS the rounded threshold
    want = opeaks.local_peaks_rough(cms, thr)
    got = pf.find_local_peaks_rough(cms.cuda(), threshold=thr)
    assert got[1].dtype == dt
    for a, b in zip(got, want):
        eq(npy(a.float()), npy(b.float()))
    gw = opeaks.global_peaks_rough(cms, thr)
    gg = pf.find_global_peaks_rough(cms.cuda(), threshold=thr)
    eq(npy(gg[0]), npy(gw[0]))
    eq(npy(gg[1].float()), npy(gw[1].float()))
    ri = pf.find_local_peaks(cms.cuda(), threshold=thr, refinement="integral")
    wi = opeaks.local_peaks(cms, thr, "integral", 5)
    for a, b in zip(ri[1:], wi[1:]):
        eq(npy(a.float()), npy(b.float()))
    close(npy(ri[0]), npy(wi[0].float()), atol=5e-3)
