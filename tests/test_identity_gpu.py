"""GPU parity of the multi-class identity path: sleap_nn_b200.inference.ops.identity and
sleap_nn_b200.data.identity against golden vectors from the unmodified reference and against the CPU oracle.
Integer outputs (peak / class indices) and gathered values are bit-exact; class maps within 1e-6 relative
(ATen's reduction order over instances depends on the host's vector width)."""

import numpy as np
import pytest
import torch

from tests.helpers import FLT_MIN, T, close, eq, golden, npy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ident():
    from sleap_nn_b200.inference.ops import identity

    return identity


@pytest.fixture(scope="module")
def dident():
    from sleap_nn_b200.data import identity

    return identity


def test_classify_peaks_from_maps_golden(ident):
    d = golden("ref_f3_identity.npz")
    for dev in ("cuda", "cpu"):  # CPU tensors are staged to the device and results come back on the CPU
        f = lambda k: T(d[k]).to(dev)
        pts, vals, probs = ident.classify_peaks_from_maps(f("cl_maps"), f("cl_pts"), f("cl_vals"), f("cl_si"), f("cl_ci"), 4)
        assert pts.device.type == dev and pts.dtype == torch.float32
        eq(npy(pts), d["cl_out_pts"]); eq(npy(vals), d["cl_out_vals"]); eq(npy(probs), d["cl_out_probs"])
    # a permuted (non-contiguous) view of the class maps is read through its strides
    maps = T(d["cl_maps"]).cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    assert not maps.is_contiguous()
    pts, vals, probs = ident.classify_peaks_from_maps(maps, T(d["cl_pts"]).cuda(), T(d["cl_vals"]).cuda(), T(d["cl_si"]).cuda(),
                                                      T(d["cl_ci"]).cuda(), 4)
    eq(npy(pts), d["cl_out_pts"]); eq(npy(probs), d["cl_out_probs"])


def test_group_class_peaks_and_vectors_golden(ident):
    d = golden("ref_f3_identity.npz")
    pi, ci = ident.group_class_peaks(T(d["gr_probs"]).cuda(), T(d["gr_si"]).cuda(), T(d["gr_ci"]).cuda(), 3, 2)
    assert pi.dtype == torch.int64 and ci.dtype == torch.int64 and pi.is_cuda
    eq(npy(pi), d["gr_peak_inds"]); eq(npy(ci), d["gr_class_inds"])
    for tag in ("tall", "wide", "square", "big"):
        inds, vals = ident.get_class_inds_from_vectors(T(d[f"vec_{tag}_probs"]).cuda())
        assert inds.device.type == "cpu" and inds.dtype == torch.int64
        eq(npy(inds), d[f"vec_{tag}_inds"]); eq(npy(vals), d[f"vec_{tag}_vals"])


def test_identity_vs_oracle_random_and_errors(ident):
    from oracle import identity as oid

    g = torch.Generator().manual_seed(5)
    for S, K, Cn, per in ((2, 2, 3, 3), (4, 6, 2, 9), (1, 40, 1, 50), (3, 5, 4, 0)):
        H, W = 24, 32
        maps = torch.softmax(torch.randn((S, K, H, W), generator=g) * 3, dim=1)
        P = S * Cn * per
        si = torch.randint(0, S, (P,), generator=g, dtype=torch.int32)
        ci = torch.randint(0, Cn, (P,), generator=g, dtype=torch.int32)
        pts = torch.rand((P, 2), generator=g) * torch.tensor([W - 1.0, H - 1.0])
        vals = torch.rand((P,), generator=g)
        want = oid.classify_peaks_from_maps(maps, pts, vals, si, ci, Cn)
        got = ident.classify_peaks_from_maps(maps.cuda(), pts.cuda(), vals.cuda(), si.cuda(), ci.cuda(), Cn)
        for a, b in zip(got, want):
            eq(npy(a), npy(b))
        probs = oid.class_probs_at_peaks(maps, pts, si) if P else torch.zeros((0, K))
        wp, wc = oid.group_class_peaks(probs, si, ci, S, Cn)
        gp, gc = ident.group_class_peaks(probs.cuda(), si.cuda(), ci.cuda(), S, Cn)
        eq(npy(gp), npy(wp)); eq(npy(gc), npy(wc))
    # ties: identical probability rows -> scipy's tie rules decide, bit for bit
    probs = torch.tensor([[0.5, 0.5, 0.1], [0.5, 0.5, 0.1], [0.5, 0.5, 0.9]])
    z = torch.zeros(3, dtype=torch.int32)
    wp, wc = oid.group_class_peaks(probs, z, z, 1, 1)
    gp, gc = ident.group_class_peaks(probs.cuda(), z.cuda(), z.cuda(), 1, 1)
    eq(npy(gp), npy(wp)); eq(npy(gc), npy(wc))
    wi, wv = oid.class_inds_from_vectors(probs)
    gi, gv = ident.get_class_inds_from_vectors(probs.cuda())
    eq(npy(gi), npy(wi)); eq(npy(gv), npy(wv))
    # scipy raises ValueError on NaN costs; so does the device path (reported through the status word)
    bad = probs.clone(); bad[1, 1] = float("nan")
    with pytest.raises(ValueError):
        ident.group_class_peaks(bad.cuda(), z.cuda(), z.cuda(), 1, 1)
    with pytest.raises(ValueError):
        ident.get_class_inds_from_vectors(bad.cuda())
    e = ident.group_class_peaks(torch.zeros((0, 3)).cuda(), z[:0].cuda(), z[:0].cuda(), 2, 2)
    assert e[0].shape == (0,) and e[1].shape == (0,) and e[0].dtype == torch.int64


def test_class_vectors_and_maps_golden(dident):
    d = golden("ref_f3_identity.npz")
    cv = dident.make_class_vectors(torch.Tensor([0, 2, 1, -1]).cuda(), 3)
    assert cv.dtype == torch.int32
    eq(npy(cv), d["cv_float"])
    eq(npy(dident.make_class_vectors(torch.tensor([3, -1, 0, 0, 1], dtype=torch.int32), 5)), d["cv_int"])
    small = dident.make_class_maps(T(d["cm_small_cms"]).cuda(), class_inds=torch.Tensor([1, 0]), n_classes=2, threshold=0.2)
    close(npy(small), d["cm_small"], rtol=1e-6)
    assert npy(small)[0, :, [6, 24], [4, 18]].tolist() == [[0.0, 1.0], [1.0, 0.0]]  # reference tests/data/test_identity.py:21-33
    multi = dident.make_class_maps(T(d["cm_multi_cms"]), torch.tensor([2, -1, 0], dtype=torch.int32), 4, 0.1)
    assert multi.device.type == "cpu"
    close(npy(multi), d["cm_multi"], rtol=1e-6)
    inst = T(d["gen_inst"])
    ci = torch.tensor([1, 0, 2], dtype=torch.int32)
    close(npy(dident.generate_class_maps(inst.cuda(), (64, 96), 3, ci, 3, output_stride=2)), d["gen_nodes"], rtol=1e-5, atol=FLT_MIN)
    close(npy(dident.generate_class_maps(inst[:, :, 0, :], (64, 96), 3, ci, 3, class_map_threshold=0.3, sigma=2.0,
                                         output_stride=4, is_centroids=True)), d["gen_centroids"], rtol=1e-5, atol=FLT_MIN)
    with pytest.raises(RuntimeError):
        dident.make_class_vectors(torch.tensor([5], dtype=torch.int32), 3)


def test_class_maps_vs_oracle_random(dident):
    from oracle import identity as oid

    g = torch.Generator().manual_seed(9)
    for I, K, h, w in ((1, 1, 16, 16), (4, 4, 64, 48), (6, 3, 33, 47), (3, 7, 128, 128)):
        cms = torch.rand((1, I, h, w), generator=g) ** 4
        cms[0, 0, 0, :4] = 0.0
        if I > 1:
            cms[0, 1:, 0, :4] = 0.0  # all-zero pixels: 0/0 = NaN under the threshold mask -> 0
        ci = torch.randint(-1, K, (I,), generator=g, dtype=torch.int32)
        if (I * K) % I:  # never: the reshape needs I*K elements, always true
            continue
        want = oid.class_maps(cms, ci, K, 0.2)
        got = dident.make_class_maps(cms.cuda(), ci.cuda(), K, 0.2)
        assert tuple(got.shape) == (1, K, h, w)
        close(npy(got), npy(want), rtol=1e-6)


def test_bottomup_multiclass_layer_golden():
    """BottomUpMultiClassPostproc == the reference layer's postprocess (peaks -> class assignment -> scale ladder ->
    nanmean scores -> cap), slot k == class k; which slots survive is bit-exact, coordinates within 1e-4 px."""
    from sleap_nn_b200.inference.layers import BottomUpMultiClassPostproc

    d = golden("ref_f3_multiclass_layer.npz")
    cms, cmaps = T(d["cms"]).cuda(), T(d["class_maps"]).cuda()
    for tag, scale, eff, cap in (("plain", 1.0, [1.0, 1.0, 1.0], None), ("scaled", 0.5, [1.0, 0.8, 1.25], None),
                                 ("cap2", 1.0, [1.0, 1.0, 1.0], 2), ("cap1", 0.5, [2.0, 1.0, 1.0], 1)):
        post = BottomUpMultiClassPostproc(0.2, "integral", 5, cms_output_stride=2, class_maps_output_stride=4, max_instances=cap)
        k, v, s, t = post(cms, cmaps, input_scale=scale, eff_scale=torch.tensor(eff))
        post.check()
        assert k.is_cuda and tuple(k.shape) == d[f"{tag}_kpts"].shape
        eq(np.isnan(npy(k)), np.isnan(d[f"{tag}_kpts"]))
        close(npy(k), d[f"{tag}_kpts"], atol=1e-4); eq(npy(v), d[f"{tag}_vals"])
        close(npy(s), d[f"{tag}_scores"], rtol=1e-6); close(npy(t), d[f"{tag}_tracking"], rtol=1e-6)
    # a permuted class-map view is read in place; a NaN probability is reported like scipy's ValueError
    post = BottomUpMultiClassPostproc(0.2, "integral", 5, 2, 4)
    view = cmaps.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    k2, *_ = post(cms, view)
    close(npy(k2), d["plain_kpts"], atol=1e-4)
    bad = cmaps.clone(); bad[:, 1] = float("nan")
    post(cms, bad)
    with pytest.raises(ValueError):
        post.check()


def test_bottomup_multiclass_layer_random_vs_oracle():
    """The generator of tests/test_oracle_fuzz_vs_reference.py::test_bottomup_multiclass_layer_fuzz (which pins the oracle
    to the unmodified reference layer) through the device path: several peaks per node, empty frames, K from 1 to 4,
    class-map strides 1 / 2 / 4, scales, instance caps."""
    from oracle import identity as oid
    from sleap_nn_b200.inference.layers import BottomUpMultiClassPostproc
    from tests.test_oracle_fuzz_vs_reference import multiclass_cases

    for cms, class_maps, cs, scale, eff, cap in multiclass_cases():
        want = oid.bottomup_multiclass_postprocess(cms, class_maps, 2, cs, scale, eff, cap, threshold=0.3)
        post = BottomUpMultiClassPostproc(0.3, "integral", 5, cms_output_stride=2, class_maps_output_stride=cs, max_instances=cap)
        k, v, s, t = post(cms.cuda(), class_maps.cuda(), input_scale=scale, eff_scale=eff)
        post.check()
        eq(np.isnan(npy(k)), np.isnan(npy(want[0])))
        close(npy(k), npy(want[0]), atol=1e-4)
        eq(npy(v), npy(want[1]))
        close(npy(s), npy(want[2]), rtol=1e-6, atol=1e-7)
        close(npy(t), npy(want[3]), rtol=1e-6, atol=1e-7)
