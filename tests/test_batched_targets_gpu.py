"""GPU parity of the dataset call-site row: `BatchedTargets` against the reference's per-frame
generate_* calls on a collated batch (golden vectors) and against the CPU oracle on random batches.
Targets are floating point: rtol 1e-5 (north_star), atol 1.2e-38 (denormal band)."""

import numpy as np
import pytest
import torch

from tests.helpers import FLT_MIN, T, close, golden, npy

pytestmark = pytest.mark.gpu
RTOL = 1e-5
# Sums over instances (make_multi_pafs) can cancel: each term is within 2 ulp of a value <= 1, so the SUM needs an
# absolute floor of a few fp32 ulps at unit scale (same convention as tests/test_targets_gpu.py:TOL_SUM).
ATOL_SUM = 1e-6


def _bt(hw=(64, 96), **kw):
    from sleap_nn_b200.data.batched_targets import BatchedTargets

    return BatchedTargets(hw, device=torch.device("cuda", 0), **kw)


def test_batched_targets_golden():
    d = golden("ref_f4_batched_targets.npz")
    bt = _bt()
    inst, num, edges, tracks = T(d["instances"]), T(d["num_instances"]), d["edges"].tolist(), T(d["tracks"])
    bu = bt.bottomup(inst, num, edges, confmap_sigma=1.5, confmap_stride=2, paf_sigma=4.0, paf_stride=4)
    assert bu["confidence_maps"].is_cuda and tuple(bu["confidence_maps"].shape) == d["confidence_maps"].shape
    close(npy(bu["confidence_maps"]), d["confidence_maps"], rtol=RTOL, atol=FLT_MIN)
    close(npy(bu["part_affinity_fields"]), d["part_affinity_fields"], rtol=RTOL, atol=ATOL_SUM)
    cen = bt.centroid(inst[:, 0, :, 0, :], num, sigma=2.0, output_stride=2)["centroids_confidence_maps"]
    close(npy(cen), d["centroid_maps"], rtol=RTOL, atol=FLT_MIN)
    single = bt.centered_instance(inst[:, :, 0], sigma=1.5, output_stride=2)["confidence_maps"]
    close(npy(single), d["single_maps"], rtol=RTOL, atol=FLT_MIN)
    # class maps: the reference cannot build them for an empty frame (torch.max over no instances raises); frames 0-2
    mc = bt.bottomup_multiclass(inst[:3], num[:3], tracks[:3], 4, confmap_sigma=1.5, confmap_stride=2,
                                class_map_threshold=0.2, class_map_sigma=3.0, class_map_stride=2)
    close(npy(mc["class_maps"]), d["class_maps"], rtol=RTOL, atol=FLT_MIN)
    close(npy(mc["confidence_maps"]), d["confidence_maps"][:3], rtol=RTOL, atol=FLT_MIN)
    cc = bt.class_maps(inst[:3, 0, :, 0, :], num[:3], tracks[:3], 4, class_map_threshold=0.1, sigma=3.0, output_stride=4,
                       is_centroids=True)
    close(npy(cc), d["class_maps_centroids"], rtol=RTOL, atol=FLT_MIN)
    # un-flattened PAFs and bf16 outputs
    p5 = bt.pafs(inst, edges, sigma=4.0, output_stride=4, flatten_channels=False)
    assert tuple(p5.shape) == (4, 4, 2, 16, 24)
    close(npy(p5).reshape(4, 8, 16, 24), d["part_affinity_fields"], rtol=RTOL, atol=ATOL_SUM)
    b16 = _bt(out_dtype=torch.bfloat16).bottomup(inst, num, edges, 1.5, 2, 4.0, 4)
    assert b16["confidence_maps"].dtype == torch.bfloat16
    close(npy(b16["confidence_maps"].float()), d["confidence_maps"], rtol=2.0 ** -8, atol=1e-30)
    close(npy(b16["part_affinity_fields"].float()), d["part_affinity_fields"], rtol=2.0 ** -8, atol=1e-30)


@pytest.mark.parametrize("hw,B,I,Nn", [((40, 52), 3, 2, 3), ((128, 160), 5, 6, 7), ((30, 34), 2, 1, 1)])
def test_batched_targets_vs_oracle_random(hw, B, I, Nn):
    """Includes a width that is not a multiple of 4 after striding (generic kernels) and single-node skeletons."""
    from oracle import identity as oid
    from oracle import targets as ot

    g = torch.Generator().manual_seed(B * 100 + I)
    H, W = hw
    inst = torch.rand((B, I, Nn, 2), generator=g) * torch.tensor([W * 1.2, H * 1.2]) - torch.tensor([W * 0.1, H * 0.1])
    inst[torch.rand((B, I, Nn), generator=g) < 0.15] = float("nan")
    num = torch.randint(1, I + 1, (B,), generator=g)
    edges = [[k, k + 1] for k in range(Nn - 1)] or [[0, 0]]
    tracks = torch.randint(-1, 3, (B, I), generator=g, dtype=torch.int32)
    bt = _bt(hw)
    for stride in (1, 2):
        cm = bt.multi_confmaps(inst, num, sigma=2.0, output_stride=stride)
        pf = bt.pafs(inst, edges, sigma=3.0, output_stride=stride)
        cl = bt.class_maps(inst, num, tracks, 3, class_map_threshold=0.2, sigma=2.0, output_stride=stride)
        one = bt.confmaps(inst[:, 0], sigma=2.0, output_stride=stride, filter_oob=True)
        for b in range(B):
            n = int(num[b])
            fr = inst[b : b + 1]
            close(npy(cm[b]), npy(ot.generate_multiconfmaps(fr, hw, n, 2.0, stride)), rtol=RTOL, atol=FLT_MIN)
            close(npy(pf[b]), npy(ot.generate_pafs(fr, hw, 3.0, stride, torch.tensor(edges), True)), rtol=RTOL, atol=ATOL_SUM)
            close(npy(cl[b]), npy(oid.generate_class_maps(fr, hw, n, tracks[b, :n], 3, 0.2, 2.0, stride)), rtol=RTOL,
                  atol=FLT_MIN)
            close(npy(one[b]), npy(ot.generate_confmaps(ot.filter_oob_points(fr[:, 0], H, W), hw, 2.0, stride)), rtol=RTOL,
                  atol=FLT_MIN)


def test_batched_targets_full_size_properties():
    """cfg4 size (32 nodes / 31 edges / 8 instances, 1024^2, stride 2), 8 frames in one launch per target: the
    batched result equals the per-frame public API frame by frame, bit for bit."""
    from sleap_nn_b200 import synthetic
    from sleap_nn_b200.data import confidence_maps as cmod
    from sleap_nn_b200.data import edge_maps as emod

    edges = synthetic.chain_edges(32)
    poses = synthetic.random_poses(3, 8, 8, 32, (1024, 1024), edges, margin=200.0, step=24.0).cuda()
    bt = _bt((1024, 1024))
    num = torch.tensor([8, 8, 5, 8, 1, 8, 8, 3])
    cm = bt.multi_confmaps(poses, num, sigma=2.5, output_stride=2)
    pf = bt.pafs(poses, edges, sigma=2.5, output_stride=2)
    assert tuple(cm.shape) == (8, 1, 32, 512, 512) and tuple(pf.shape) == (8, 62, 512, 512)
    for b in (0, 2, 4, 7):
        want = cmod.generate_multiconfmaps(poses[b : b + 1], (1024, 1024), int(num[b]), sigma=2.5, output_stride=2)
        assert torch.equal(cm[b], want)
        want = emod.generate_pafs(poses[b : b + 1], (1024, 1024), sigma=2.5, output_stride=2, edge_inds=torch.tensor(edges),
                                  flatten_channels=True)
        assert torch.equal(pf[b], want)
    assert float(cm.max()) > 0.99 and float(cm.min()) == 0.0


def _random_batch(seed, B, I, Nn, hw):
    g = torch.Generator().manual_seed(seed)
    inst = torch.rand((B, I, Nn, 2), generator=g) * torch.tensor([hw[1], hw[0]], dtype=torch.float32)
    inst[0, 0, 0] = float("nan")                       # a missing keypoint
    inst[-1, -1] = float("nan")                        # a missing instance
    inst[0, -1, :, 0] += hw[1]                         # an instance entirely outside the image (generate_pafs drops it)
    num = torch.randint(1, I + 1, (B,), generator=g)
    return inst, num


@pytest.mark.parametrize("hw,B,I,Nn,s7,s8", [((64, 96), 3, 2, 3, 2, 4), ((128, 160), 5, 6, 7, 2, 2), ((1024, 1024), 1, 8, 32, 2, 2),
                                             ((256, 256), 9, 3, 5, 4, 2)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_fused_bottomup_targets_equal_the_two_kernels(hw, B, I, Nn, s7, s8, dt):
    """snb_bottomup_targets (both targets as an overlapping programmatic-dependent-launch pair) == snb_confmaps_ex +
    snb_pafs_from_instances, bit for bit, fp32 and bf16, single frames and batches, different strides per head; repeated
    back-to-back calls (each pair's successor overlaps its tail) stay correct."""
    bt = _bt(hw, out_dtype=dt)
    inst, num = _random_batch(B * 7 + I, B, I, Nn, hw)
    edges = [(k, k + 1) for k in range(Nn - 1)] or [(0, 0)]
    sep = bt.bottomup(inst, num, edges, 2.5, s7, 2.5, s8, fused=False)
    for rep in range(3):
        fu = bt.bottomup(inst, num, edges, 2.5, s7, 2.5, s8, fused=True)
        assert fu["confidence_maps"].dtype == dt and fu["confidence_maps"].shape == sep["confidence_maps"].shape
        assert torch.equal(fu["confidence_maps"].view(torch.int16 if dt == torch.bfloat16 else torch.int32),
                           sep["confidence_maps"].view(torch.int16 if dt == torch.bfloat16 else torch.int32)), rep
        assert torch.equal(fu["part_affinity_fields"].view(torch.int16 if dt == torch.bfloat16 else torch.int32),
                           sep["part_affinity_fields"].view(torch.int16 if dt == torch.bfloat16 else torch.int32)), rep


def test_bf16_confmaps_separable_form_vs_exact_fp32():
    """bf16 confidence maps use exp(-dx^2/den) * exp(-dy^2/den) (a table per instance and band): after rounding to bf16
    they must sit within ONE bf16 ulp of the exactly-computed fp32 maps rounded to bf16, zeros where those are zero."""
    hw = (1024, 1024)
    inst, num = _random_batch(3, 2, 8, 32, hw)
    exact = _bt(hw).multi_confmaps(inst, num, 2.5, 2)
    b16 = _bt(hw, out_dtype=torch.bfloat16).multi_confmaps(inst, num, 2.5, 2)
    want = exact.bfloat16()
    a, b = b16.view(torch.int16).int(), want.view(torch.int16).int()
    assert int((a - b).abs().max()) <= 1                     # adjacent bf16 values at most (bit patterns of non-negative floats)
    big = exact > 1e-30
    assert bool(((b16.float() - exact).abs()[big] <= exact[big] * 2.0 ** -8).all())
    assert float((a != b).float().mean()) < 0.02             # and almost always the very same value
