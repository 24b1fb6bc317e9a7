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
