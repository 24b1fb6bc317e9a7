"""GPU parity of the layer-level epilogues against vectors composed from the
reference's own ops in the order its layers call them (tests/golden/make_golden.py::f2_layers)."""

import numpy as np
import pytest
import torch

from tests.helpers import T, close, eq, golden, npy

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["dyn", "top3", "pad7"])
def test_centroid_postproc(tag):
    from sleap_nn_b200.inference.layers import CentroidPostproc

    d = golden("ref_f2_layers.npz")
    mi = int(d[f"cen_{tag}_max"])
    post = CentroidPostproc(peak_threshold=0.2, refinement="integral", max_instances=None if mi < 0 else mi)
    xy, val = post(T(d["cen_cms"]).cuda(), output_stride=int(d["cen_stride"]), input_scale=float(d[f"cen_{tag}_scale"]),
                   eff_scale=T(d["cen_eff"]))
    want_xy, want_val = d[f"cen_{tag}_xy"], d[f"cen_{tag}_val"]
    assert tuple(xy.shape) == want_xy.shape
    eq(np.isnan(npy(xy)), np.isnan(want_xy))
    close(npy(xy), want_xy, atol=1e-4)
    eq(npy(val), want_val)  # values and their order (top-k by value when truncating) are exact
    assert int(post.last_status.item()) == 0


def test_centered_instance_postproc_and_topdown_uncrop():
    from sleap_nn_b200.inference.layers import CenteredInstancePostproc

    d = golden("ref_f2_layers.npz")
    cms = T(d["ci_cms"]).cuda()
    post = CenteredInstancePostproc(peak_threshold=0.2, refinement="integral")
    xy, val = post(cms, output_stride=2, input_scale=float(d["ci_scale"]), eff_scale=T(d["ci_eff"]))
    assert tuple(xy.shape) == d["ci_xy"].shape and tuple(val.shape) == d["ci_val"].shape
    eq(np.isnan(npy(xy)), np.isnan(d["ci_xy"]))
    close(npy(xy), d["ci_xy"], atol=1e-4)
    eq(npy(val), d["ci_val"])
    # top-down: crop offset + per-crop scale + scatter into (B, max_inst, N, 2) in the same launch
    vi = d["td_valid_idx"]
    rows = torch.from_numpy(vi[:, 0] * 3 + vi[:, 1]).to(torch.int32)
    xy2, val2 = post(cms, output_stride=2, crop_topleft=T(d["td_topleft"]), per_crop_eff_scale=T(d["td_eff"])[vi[:, 0]],
                     scatter_rows=rows, out_shape=(3, 3))
    assert tuple(xy2.shape) == d["td_xy"].shape
    eq(np.isnan(npy(xy2)), np.isnan(d["td_xy"]))
    close(npy(xy2), d["td_xy"], atol=1e-4)
    eq(npy(val2), d["td_val"])


def test_coord_ops_match_torch_expressions():
    """The four ops of ops/coord.py are single tensor expressions; ours must reproduce them bit for bit."""
    from sleap_nn_b200.inference.ops import coord

    g = torch.Generator().manual_seed(0)
    xy = (torch.rand((4, 3, 5, 2), generator=g) * 500).cuda()
    xy[1, 2, 3] = float("nan")
    eff = torch.tensor([1.0, 0.7, 1.3, 2.0])
    eq(npy(coord.undo_stride(xy, 4)), npy(xy * 4))
    assert coord.undo_stride(xy, 1) is xy
    # the parity target is the reference's CPU arithmetic: ATen's CPU `tensor / scalar` is a true fp32 division
    # (its CUDA kernel multiplies by the reciprocal instead, which differs in the last bit)
    eq(npy(coord.undo_input_scale(xy, 0.3)), npy(xy.cpu() / 0.3))
    assert coord.undo_input_scale(xy, 1.0) is xy
    eq(npy(coord.undo_eff_scale(xy, eff)), npy(xy / eff.view(4, 1, 1, 1).cuda()))
    assert coord.undo_eff_scale(xy, torch.ones(4)) is xy
    tl = torch.rand((12, 2), generator=g) * 100
    eq(npy(coord.add_crop_offset(xy, tl)), npy((xy.reshape(12, 5, 2) + tl.cuda().view(-1, 1, 2)).reshape(4, 3, 5, 2)))
    eq(npy(coord.add_crop_offset(xy.reshape(12, 5, 2), tl)), npy(xy.reshape(12, 5, 2) + tl.cuda().view(-1, 1, 2)))
    cpu = coord.undo_input_scale(xy.cpu(), 0.5)  # CPU tensors are staged to the device and come back on the CPU
    assert cpu.device.type == "cpu"
    eq(npy(cpu), npy(xy.cpu() / 0.5))


def test_peaks_topk_ties_and_overflow():
    from sleap_nn_b200 import _native as N

    dev = torch.device("cuda", 0)
    val = torch.tensor([[0.3, 0.9, 0.9, 0.1, 0.5, 0.0, 0.0, 0.0]], device=dev)
    xy = torch.arange(16, dtype=torch.float32, device=dev).reshape(1, 8, 2)
    cnt = torch.tensor([5], dtype=torch.int32, device=dev)
    o_xy = torch.empty((1, 3, 2), device=dev); o_val = torch.empty((1, 3), device=dev)
    N.check(N.lib.snb_peaks_topk(N.ptr(cnt), 1, 8, N.ptr(xy), N.ptr(val), 3, 1.0, None, N.ptr(o_xy), N.ptr(o_val),
                                 N.stream_ptr(dev)), "topk")
    assert o_val[0].tolist() == pytest.approx([0.9, 0.9, 0.5]) and o_xy[0, :, 0].tolist() == [2.0, 4.0, 8.0]


# ---------------------------------------------------------------------------------------------------------------
# Top-down composition (stage B + stage 2), multi-class per-frame assignment, single-instance layer
# ---------------------------------------------------------------------------------------------------------------
def _topdown_case(g, tag):
    from tests.helpers import topdown_model

    gain, pattern, cgain = (T(g[k]).cuda() for k in ("gain", "pattern", "cgain"))
    nms, multiclass, thr, stride, scale = (float(v) for v in g[f"{tag}_knobs"])
    model = (lambda c: topdown_model(c, gain, pattern, cgain)) if multiclass else (lambda c: topdown_model(c, gain, pattern))
    return model, bool(nms), bool(multiclass), thr, int(stride), scale


@pytest.mark.parametrize("tag", ["plain", "nms", "mc", "mcnms"])
def test_topdown_postproc_golden(tag):
    """TopDownPostproc vs TopDownLayer._centroid_nms_mask / _run_stage_2 run on the unmodified reference classes."""
    from sleap_nn_b200.inference.layers import TopDownPostproc

    g = golden("ref_f2_topdown.npz")
    model, nms, multiclass, thr, stride, scale = _topdown_case(g, tag)
    post = TopDownPostproc(tuple(int(v) for v in g["crop_hw"]), peak_threshold=0.2, refinement="integral", centroid_nms=nms,
                           centroid_nms_threshold=thr, return_crops=True, return_class_vectors=True)
    o = post(T(g["image"]).cuda(), T(g["centroids"]).cuda(), T(g["centroid_vals"]).cuda(), model, eff_scale=T(g["eff"]).cuda(),
             output_stride=stride, input_scale=scale)
    post.check()
    eq(npy(o["valid_mask"]), g[f"{tag}_valid"])
    eq(npy(o["crops"]), g[f"{tag}_crops"])                      # integer crop pickup: bit-exact
    eq(npy(o["pred_peak_values"]), g[f"{tag}_vals"])            # arg-max values: bit-exact
    eq(np.isnan(npy(o["pred_keypoints"])), np.isnan(g[f"{tag}_kpts"]))
    close(npy(o["pred_keypoints"]), g[f"{tag}_kpts"], atol=1e-4)      # refined coordinates: 1e-4 px
    close(npy(o["pred_crop_keypoints"]), g[f"{tag}_crop_kpts"], atol=1e-4)
    eq(npy(o["pred_centroids"]), g[f"{tag}_centroids"])         # (c * eff) / eff, separately rounded
    eq(npy(o["instance_bboxes"]), g[f"{tag}_bboxes"])
    eq(npy(o["instance_scores"]), g[f"{tag}_scores"])
    if multiclass:
        eq(npy(o["pred_class_inds"]), g[f"{tag}_class_inds"])   # per-frame optimal assignment: bit-exact
        eq(npy(o["instance_tracking_scores"]), g[f"{tag}_tracking"])
        eq(npy(o["pred_class_vectors"]), g[f"{tag}_class_vectors"])
    else:
        assert "pred_class_inds" not in o


def test_topdown_postproc_no_valid_centroid():
    """All-NaN centroids: no crop, no model call, all-NaN outputs of the right shape (topdown.py:218-233)."""
    from sleap_nn_b200.inference.layers import TopDownPostproc

    g = golden("ref_f2_topdown.npz")
    post = TopDownPostproc((24, 32), n_nodes=3)
    cen = torch.full((2, 3, 2), float("nan"), device="cuda")

    def model(_):
        raise AssertionError("the network must not run when there is nothing to crop")

    o = post(T(g["image"])[:2].cuda(), cen, torch.full((2, 3), float("nan"), device="cuda"), model)
    eq(npy(o["pred_keypoints"]), g["empty_kpts"])
    eq(npy(o["pred_peak_values"]), g["empty_vals"])
    assert not npy(o["valid_mask"]).any() and "instance_bboxes" not in o


def test_standalone_multiclass_and_single_instance_layers():
    from sleap_nn_b200.inference.layers import CenteredInstanceMultiClassPostproc, SingleInstancePostproc
    from tests.helpers import topdown_model

    g = golden("ref_f2_topdown.npz")
    gain, pattern, cgain = (T(g[k]).cuda() for k in ("gain", "pattern", "cgain"))
    cms, vec = topdown_model(T(g["sa_crops"]).cuda(), gain, pattern, cgain)
    o = CenteredInstanceMultiClassPostproc(0.2, "integral", 5).classify(cms, vec, output_stride=2, input_scale=0.5,
                                                                         eff_scale=T(g["sa_eff"]))
    close(npy(o["pred_keypoints"]), g["sa_kpts"], atol=1e-4)
    eq(npy(o["pred_peak_values"]), g["sa_vals"])
    eq(npy(o["pred_class_inds"]), g["sa_class_inds"])
    eq(npy(o["pred_class_probs"]), g["sa_class_probs"])
    eq(npy(o["instance_tracking_scores"]), g["sa_tracking"])
    xy, val = SingleInstancePostproc(0.2, "integral", 5)(cms, output_stride=2, input_scale=0.5, eff_scale=T(g["sa_eff"]))
    close(npy(xy), g["si_kpts"], atol=1e-4)
    eq(npy(val), g["si_vals"])


@pytest.mark.parametrize("seed,nms", [(0, False), (1, True), (2, True)])
def test_topdown_postproc_random_vs_oracle(seed, nms):
    """Seeded larger case (B=16, up to 12 centroids per frame, 5 classes) against the CPU oracle."""
    from oracle import topdown as otd
    from sleap_nn_b200.inference.layers import TopDownPostproc
    from tests.helpers import topdown_model

    gen = torch.Generator().manual_seed(seed)
    B, I, H, W, crop_hw, Nn, K = 16, 12, 160, 192, (32, 40), 4, 5
    cen = torch.rand((B, I, 2), generator=gen) * torch.tensor([W - 1.0, H - 1.0])
    cen[torch.rand((B, I), generator=gen) < 0.3] = float("nan")
    cen[3] = float("nan")                      # a frame without any centroid
    cen[5, 1:] = float("nan")                  # a frame with exactly one
    val = torch.rand((B, I), generator=gen)
    eff = torch.rand((B,), generator=gen) * 0.5 + 0.75
    img = (torch.rand((B, 1, H, W), generator=gen) * 255).to(torch.uint8)
    gain = torch.rand((Nn, *crop_hw), generator=gen) * 0.5 + 0.5
    pattern = torch.rand((Nn, *crop_hw), generator=gen) * 1e-3
    cgain = torch.rand((K, *crop_hw), generator=gen) * 0.5 + 0.5
    want = otd.stage_2(img, cen, val, eff, crop_hw, lambda c: topdown_model(c, gain, pattern, cgain), nms=nms,
                       nms_threshold=0.2, output_stride=2, input_scale=0.5)
    same_size = tuple(want["crops"].shape[-2:]) == crop_hw  # bbox 0 can round one short (ops/crops.py:66-67); the
    post = TopDownPostproc(crop_hw, centroid_nms=nms, centroid_nms_threshold=0.2, return_crops=same_size,  # reference
                           return_class_vectors=True)                                      # then cannot return crops
    gd, pd, cd = gain.cuda(), pattern.cuda(), cgain.cuda()
    o = post(img.cuda(), cen.cuda(), val.cuda(), lambda c: topdown_model(c, gd, pd, cd), eff_scale=eff.cuda(), output_stride=2,
             input_scale=0.5)
    post.check()
    eq(npy(o["valid_mask"]), want["valid"])
    if nms:
        assert want["valid"].sum() < (~np.isnan(npy(cen)).any(-1)).sum(), "the case must exercise the suppression"
    if same_size:
        eq(npy(o["crops"]), want["crops"])
    eq(npy(o["pred_peak_values"]), want["vals"])
    close(npy(o["pred_keypoints"]), want["kpts"], atol=1e-4)
    close(npy(o["pred_crop_keypoints"]), want["crop_kpts"], atol=1e-4)
    eq(npy(o["pred_centroids"]), want["centroids"])
    eq(npy(o["instance_bboxes"]), want["bboxes"])
    eq(npy(o["pred_class_inds"]), want["class_inds"])
    eq(npy(o["instance_tracking_scores"]), want["tracking"])
    eq(npy(o["pred_class_vectors"]), want["class_vectors"])


@pytest.mark.parametrize("shape,scale", [((2, 1, 8, 8), 0.5), ((1, 3, 37, 53), 0.75), ((2, 2, 20, 31), 2.0),
                                         ((1, 1, 384, 384), 0.5), ((4, 1, 1024, 1024), 0.25), ((1, 1, 5, 7), 0.3)])
def test_apply_input_scale_matches_oracle(shape, scale):
    """apply_input_scale (ops/coord.py:93-109) vs the numpy restatement of F.interpolate's bilinear rule."""
    from oracle.resize import apply_input_scale as want_fn
    from sleap_nn_b200.inference.ops.coord import apply_input_scale

    img = torch.rand(shape, generator=torch.Generator().manual_seed(7))
    got = apply_input_scale(img.cuda(), scale)
    want = want_fn(img.numpy(), scale)
    assert got.is_cuda and tuple(got.shape) == want.shape
    close(npy(got), want, rtol=1e-6, atol=1e-7)
    # CPU tensors are accepted (staged to the device, result returned on the CPU); a strided view is read in place
    close(npy(apply_input_scale(img, scale)), want, rtol=1e-6, atol=1e-7)
    view = torch.rand((shape[0], shape[1], shape[2], shape[3] * 2), generator=torch.Generator().manual_seed(8)).cuda()[..., ::2]
    close(npy(apply_input_scale(view, scale)), want_fn(npy(view), scale), rtol=1e-6, atol=1e-7)
    assert apply_input_scale(img, 1.0) is img
    half = apply_input_scale(img.cuda().half(), scale)
    assert half.dtype == torch.float16
    close(npy(half.float()), want_fn(img.half().float().numpy(), scale), rtol=2e-3, atol=1e-3)


def test_topdown_postproc_wide_frames_and_many_classes():
    """More slots than a warp (I = 40) and more classes than a warp (K = 40): the strided loops of the select kernel and the
    serial assignment path of the class kernel, against the oracle."""
    from oracle import topdown as otd
    from sleap_nn_b200.inference.layers import TopDownPostproc
    from tests.helpers import topdown_model

    gen = torch.Generator().manual_seed(11)
    B, I, H, W, crop_hw, Nn, K = 3, 40, 128, 160, (16, 24), 2, 40
    cen = torch.rand((B, I, 2), generator=gen) * torch.tensor([W - 1.0, H - 1.0])
    cen[torch.rand((B, I), generator=gen) < 0.15] = float("nan")
    val = torch.rand((B, I), generator=gen)
    img = (torch.rand((B, 1, H, W), generator=gen) * 255).to(torch.uint8)
    gain = torch.rand((Nn, *crop_hw), generator=gen) * 0.5 + 0.5
    pattern = torch.rand((Nn, *crop_hw), generator=gen) * 1e-3
    cgain = torch.rand((K, *crop_hw), generator=gen) * 0.5 + 0.5
    want = otd.stage_2(img, cen, val, None, crop_hw, lambda c: topdown_model(c, gain, pattern, cgain), nms=True, nms_threshold=0.1)
    post = TopDownPostproc(crop_hw, centroid_nms=True, centroid_nms_threshold=0.1, return_class_vectors=True)
    gd, pd, cd = gain.cuda(), pattern.cuda(), cgain.cuda()
    o = post(img.cuda(), cen.cuda(), val.cuda(), lambda c: topdown_model(c, gd, pd, cd))
    post.check()
    eq(npy(o["valid_mask"]), want["valid"])
    assert 0 < want["valid"].sum() < (~np.isnan(npy(cen)).any(-1)).sum()
    eq(npy(o["pred_peak_values"]), want["vals"])
    close(npy(o["pred_keypoints"]), want["kpts"], atol=1e-4)
    eq(npy(o["pred_class_inds"]), want["class_inds"])
    eq(npy(o["instance_tracking_scores"]), want["tracking"])
    eq(npy(o["pred_class_vectors"]), want["class_vectors"])


def test_topdown_select_degenerate_shapes_and_bad_class_vectors():
    from sleap_nn_b200.inference.layers import TopDownPostproc

    post = TopDownPostproc((8, 8), n_nodes=2)
    img = torch.zeros((0, 1, 16, 16), dtype=torch.uint8, device="cuda")
    o = post(img, torch.zeros((0, 3, 2), device="cuda"), torch.zeros((0, 3), device="cuda"), lambda c: c.float())
    assert tuple(o["pred_keypoints"].shape) == (0, 3, 2, 2) and tuple(o["valid_mask"].shape) == (0, 3)
    img = torch.zeros((2, 1, 16, 16), dtype=torch.uint8, device="cuda")
    o = post(img, torch.zeros((2, 0, 2), device="cuda"), torch.zeros((2, 0), device="cuda"), lambda c: c.float())
    assert tuple(o["pred_keypoints"].shape) == (2, 0, 2, 2)
    # a NaN class probability is an invalid cost matrix: ValueError, like scipy inside get_class_inds_from_vectors
    cen = torch.tensor([[[8.0, 8.0], [4.0, 4.0]]], device="cuda")
    model = lambda c: (c.float().repeat(1, 2, 1, 1) + 0.5, torch.full((c.shape[0], 3), float("nan"), device=c.device))
    o = post(img[:1], cen, torch.tensor([[0.9, 0.8]], device="cuda"), model)
    with pytest.raises(ValueError):
        post.check()
    assert (npy(o["pred_class_inds"]) == -1).all()


def test_topdown_crop_size_is_read_off_the_first_box_like_the_reference():
    """crop_bboxes takes the crop size from bbox 0 as int(|BL.y - TL.y|) + 1 (ops/crops.py:66-67).  For a centroid whose
    +/- half lands in another fp32 binade that is one less than the configured size; the reference then crops the smaller
    window for EVERY crop of the batch.  Reproduced (and `return_crops` raises, as the reference's scatter does)."""
    from oracle import topdown as otd
    from sleap_nn_b200.inference.layers import TopDownPostproc
    from tests.helpers import topdown_model

    gen = torch.Generator().manual_seed(3)
    crop_hw, Nn = (20, 28), 2
    # y = 56.7: fl(fl(56.7 + 10) - 0.5) - fl(fl(56.7 - 10) + 0.5) = 18.999996 -> int() = 18 -> 19 rows
    cen = torch.tensor([[[40.25, 56.7], [80.0, 30.0]]])
    val = torch.tensor([[0.9, 0.8]])
    img = (torch.rand((1, 1, 96, 120), generator=gen) * 255).to(torch.uint8)
    gain = torch.rand((Nn, *crop_hw), generator=gen) * 0.5 + 0.5
    pattern = torch.rand((Nn, *crop_hw), generator=gen) * 1e-3
    want = otd.stage_2(img, cen, val, None, crop_hw, lambda c: topdown_model(c, gain, pattern))
    assert want["crops"].shape[-2:] == (19, 28), "the case must exercise the short first box"
    gd, pd = gain.cuda(), pattern.cuda()
    o = TopDownPostproc(crop_hw)(img.cuda(), cen.cuda(), val.cuda(), lambda c: topdown_model(c, gd, pd))
    eq(npy(o["pred_peak_values"]), want["vals"])
    close(npy(o["pred_keypoints"]), want["kpts"], atol=1e-4)
    eq(npy(o["instance_bboxes"]), want["bboxes"])
    with pytest.raises(RuntimeError, match="shape mismatch"):
        TopDownPostproc(crop_hw, return_crops=True)(img.cuda(), cen.cuda(), val.cuda(), lambda c: topdown_model(c, gd, pd))


def test_single_stage_layers_random_vs_oracle():
    """CentroidPostproc / CenteredInstancePostproc / SingleInstancePostproc on the generator that pins oracle.layers to
    the unmodified reference layers: inferred and fixed max_instances, every ladder step, empty frames."""
    from oracle import layers as olay
    from sleap_nn_b200.inference.layers import CenteredInstancePostproc, CentroidPostproc, SingleInstancePostproc
    from tests.test_oracle_fuzz_vs_reference import single_stage_cases

    for cms, stride, scale, eff, cap in single_stage_cases():
        want_xy, want_val = olay.centroid_postprocess(cms[:, :1], stride, scale, eff, cap, threshold=0.25)
        xy, val = CentroidPostproc(0.25, "integral", 5, max_instances=cap)(cms[:, :1].cuda(), output_stride=stride,
                                                                           input_scale=scale, eff_scale=eff)
        assert tuple(xy.shape) == want_xy.shape
        eq(np.isnan(npy(xy)), np.isnan(want_xy))
        close(npy(xy), want_xy, atol=1e-4)
        eq(npy(val), want_val)
        wk, wv = olay.global_postprocess(cms, stride, scale, eff, threshold=0.25)
        for post in (CenteredInstancePostproc(0.25, "integral", 5), SingleInstancePostproc(0.25, "integral", 5)):
            k, v = post(cms.cuda(), output_stride=stride, input_scale=scale, eff_scale=eff)
            eq(np.isnan(npy(k)), np.isnan(wk))
            close(npy(k), wk, atol=1e-4)
            eq(npy(v), wv)
