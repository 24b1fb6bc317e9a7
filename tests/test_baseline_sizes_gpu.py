"""GPU parity AT THE SIZES BASELINE.json names (not reduced stand-ins), each against the CPU oracle on the same inputs:

  cfg1  single-instance find_global_peaks + integral refine, (1, 2, 192, 192)
  cfg2  top-down centred-instance find_global_peaks + refine, (256, 13, 80, 80)
  cfg3  bottom-up mice, batch 64 of 1024^2 frames (maps 512^2), full peak + PAF grouping
  cfg4  bottom-up flies targets: one 32-node / 31-edge / 8-instance frame of make_multi_confmaps + make_pafs

and the native half-precision path (ABI v5): fp16 / bf16 maps read in place must give bit-identical results to the
fp32 kernels run on the exact up-cast copy - which is what the reference computes, because its torch backend casts
autocast heads back with .float() before peak finding (inference/layers/backends/torch_backend.py:125-146).

Bars: integers (peak grid positions, order, channels, instance membership, NaN pattern) bit-exact; refined coordinates
<= 1e-4 px; instance scores within 1e-5 relative (+1e-6 absolute: the observed maximum is printed); targets rtol 1e-5.
"""

import numpy as np
import pytest
import torch

from tests.helpers import close, eq, npy

pytestmark = pytest.mark.gpu

HALF = [torch.float16, torch.bfloat16]


def _oracle_bottomup(cms_cpu, pafs_cpu, edges, n_nodes, stride):
    from oracle import paf as opaf
    from oracle import peaks as opeaks
    from oracle.synth import split_by_sample

    B = cms_cpu.shape[0]
    pts, vals, si, ci = opeaks.local_peaks(cms_cpu, 0.2, "integral")
    peaks, pvs, pcs = (split_by_sample(x, si, B) for x in (pts * stride, vals, ci))
    return (peaks, pvs, pcs), opaf.predict(pafs_cpu.permute(0, 2, 3, 1), peaks, pvs, pcs, edges, n_nodes, stride)


def _cfg3_batch(seed, B=64):
    from sleap_nn_b200 import synthetic

    Nn, hw, stride = 5, (1024, 1024), 2
    edges = synthetic.chain_edges(Nn)
    poses = synthetic.random_poses(seed, B, 2, Nn, hw, edges)
    cms, pafs = synthetic.render_batch(poses, hw, stride, edges, torch.device("cuda"), seed=seed)
    return edges, poses, cms, pafs


def test_cfg3_batch64_vs_oracle():
    """The timed bench configuration itself: 64 frames, (64,5,512,512) + (64,8,512,512), through the fused 2-launch chain."""
    from sleap_nn_b200.pipeline import BottomUpPostproc

    B, Nn, stride = 64, 5, 2
    edges, poses, cms, pafs = _cfg3_batch(11, B)
    pipe = BottomUpPostproc(Nn, edges, B, (512, 512), cms_stride=stride, pafs_stride=stride)
    res = pipe(cms, pafs)
    inst, pv, sc = res.to_lists()
    (peaks, pvs, pcs), want = _oracle_bottomup(cms.cpu(), pafs.cpu(), edges, Nn, stride)
    max_px = max_sc = 0.0
    for b in range(B):
        n = int(res.n_peaks[b])
        assert n == peaks[b].shape[0]
        close(npy(res.peaks[b, :n]), npy(peaks[b]), atol=1e-4)
        eq(npy(res.peak_vals[b, :n]), npy(pvs[b]))
        eq(npy(res.peak_channels[b, :n]), npy(pcs[b]))
        assert inst[b].shape == want[0][b].shape
        eq(np.isnan(npy(inst[b])), np.isnan(npy(want[0][b])))
        close(npy(inst[b]), npy(want[0][b]), atol=1e-4)
        eq(npy(pv[b]), npy(want[1][b]))
        close(npy(sc[b]), npy(want[2][b]), rtol=1e-5, atol=1e-6)
        if inst[b].numel():
            max_px = max(max_px, float(np.nanmax(np.abs(npy(inst[b]) - npy(want[0][b])))))
            max_sc = max(max_sc, float(np.max(np.abs(npy(sc[b]) - npy(want[2][b])))))
    assert sum(len(x) for x in inst) == 2 * B
    print(f"cfg3 batch 64: max |dx| {max_px:.2e} px, max |dscore| {max_sc:.2e}")


@pytest.mark.parametrize("dt", HALF)
def test_cfg3_batch64_half_maps_read_natively(dt):
    """fp16 / bf16 heads read in place == the fp32 chain on the exact up-cast copy, bit for bit, on every output; and
    the up-cast result equals the oracle run on the up-cast maps."""
    from sleap_nn_b200.pipeline import BottomUpPostproc

    B, Nn, stride = 64, 5, 2
    edges, poses, cms, pafs = _cfg3_batch(12, B)
    ch, ph = cms.to(dt), pafs.to(dt)
    pipe = BottomUpPostproc(Nn, edges, B, (512, 512), cms_stride=stride, pafs_stride=stride)
    keys = ("frame_count", "peak_xy", "peak_val", "peak_chan", "n_inst", "inst_xy", "inst_val", "inst_score", "m_count")
    pipe(ch, ph.permute(0, 2, 3, 1))
    torch.cuda.synchronize()
    got = {k: pipe.buf[k].clone() for k in keys}
    pipe(ch.float(), ph.float())
    torch.cuda.synchronize()
    want = {k: pipe.buf[k].clone() for k in keys}
    n_pk, n_in = npy(want["frame_count"]), npy(want["n_inst"])
    eq(npy(got["frame_count"]), n_pk)
    eq(npy(got["n_inst"]), n_in)
    eq(npy(got["m_count"]), npy(want["m_count"]))
    for b in range(B):
        for k in ("peak_xy", "peak_val", "peak_chan"):
            eq(npy(got[k][b, : n_pk[b]]), npy(want[k][b, : n_pk[b]]))
        for k in ("inst_xy", "inst_val", "inst_score"):
            eq(npy(got[k][b, : n_in[b]]), npy(want[k][b, : n_in[b]]))
    # the first 8 frames against the oracle on the up-cast maps
    res = pipe(ch, ph)
    inst, pv, sc = res.to_lists()
    _, ow = _oracle_bottomup(ch[:8].float().cpu(), ph[:8].float().cpu(), edges, Nn, stride)
    for b in range(8):
        assert inst[b].shape == ow[0][b].shape
        close(npy(inst[b]), npy(ow[0][b]), atol=1e-4)
        eq(npy(pv[b]), npy(ow[1][b]))
        close(npy(sc[b]), npy(ow[2][b]), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("dt", HALF)
@pytest.mark.parametrize("shape", [(2, 3, 40, 52), (1, 2, 33, 48), (2, 2, 64, 64), (1, 1, 12, 512), (1, 1, 6, 1024)])
def test_half_maps_every_detect_path(dt, shape):
    """Narrow rows, rows that are not a multiple of the 8-element vector (scalar path), sliced (strided, unaligned)
    views: the native kernels agree with the fp32 kernels on the up-cast copy, order and values bit for bit."""
    from sleap_nn_b200.inference.ops.peaks import local_peaks_padded

    g = torch.Generator().manual_seed(5)
    base = torch.rand((shape[0], shape[1], shape[2] + 3, shape[3] + 16), generator=g)  # row stride stays a multiple of 8
    base[0, 0, 4, 6] = float("nan")
    base[-1, -1, 7, 9] = float("inf")
    for view in (lambda t: t[:, :, : shape[2], : shape[3]], lambda t: t[:, :, 1 : shape[2] + 1, 1 : shape[3] + 1],
                 lambda t: t[:, :, 2 : shape[2] + 2, 8 : shape[3] + 8]):
        h = view(base.to(dt).cuda())
        a = local_peaks_padded(h, 0.55, 5, 2.0, 2048)
        b = local_peaks_padded(h.float(), 0.55, 5, 2.0, 2048)
        torch.cuda.synchronize()
        eq(npy(a[0]), npy(b[0]))
        for f in range(shape[0]):
            n = int(a[0][f])
            assert 0 < n <= 2048
            for i in (1, 2, 3):
                eq(npy(a[i][f, :n]), npy(b[i][f, :n]))


@pytest.mark.parametrize("dt", [torch.float32] + HALF)
def test_cfg2_batch256_global_peaks_vs_oracle(dt):
    """(256, 13, 80, 80) centred-instance crops through K2 (warp-per-plane kernel), integral refinement."""
    from oracle import peaks as opeaks
    from sleap_nn_b200.data.confidence_maps import make_confmaps
    from sleap_nn_b200.inference import peak_finding as pf

    B, Cn, H, W = 256, 13, 80, 80
    g = torch.Generator().manual_seed(2)
    pts = torch.rand((B, Cn, 2), generator=g) * 60 + 10
    pts[3, 4] = float("nan")  # a missing node: an all-zero plane (below threshold -> NaN, 0)
    xv, yv = torch.arange(W, dtype=torch.float32), torch.arange(H, dtype=torch.float32)
    cms = make_confmaps(pts.cuda(), xv, yv, 3.0) + torch.rand((B, Cn, H, W), generator=g).cuda() * 1e-3
    cms = cms.to(dt)
    ref_in = cms.float().cpu()  # the reference's backend hands fp32 (up-cast) maps to the op
    wp, wv = opeaks.global_peaks(ref_in, 0.2, "integral")
    if dt == torch.float32:
        gp, gv = pf.find_global_peaks(cms, threshold=0.2, refinement="integral")
    else:  # native half: the layer-level entry point keeps the fp32 threshold (no rounding to the maps' dtype)
        from sleap_nn_b200.inference.layers import CenteredInstancePostproc

        k, v = CenteredInstancePostproc(0.2, "integral")(cms, output_stride=1)
        gp, gv = k[:, 0], v[:, 0]
    eq(np.isnan(npy(gp)), np.isnan(npy(wp)))
    close(npy(gp), npy(wp), atol=1e-4)
    eq(npy(gv.float()), npy(wv))
    assert int(np.isnan(npy(wp)).all(-1).sum()) == 1
    # rough positions are integers: bit-exact
    rp, _ = opeaks.global_peaks_rough(ref_in, 0.2)
    if dt == torch.float32:
        eq(npy(pf.find_global_peaks_rough(cms, threshold=0.2)[0]), npy(rp))


@pytest.mark.parametrize("dt", [torch.float32] + HALF)
def test_cfg1_single_instance_plane_vs_oracle(dt):
    """(1, 2, 192, 192): the register-resident chunked kernel (planes larger than 16 K elements), strided views too."""
    from oracle import peaks as opeaks
    from sleap_nn_b200.inference.layers import CenteredInstancePostproc

    g = torch.Generator().manual_seed(4)
    big = torch.rand((2, 2, 200, 208), generator=g) * 0.3
    big[0, 0, 77, 131] = 0.9; big[0, 1, 5, 7] = 0.8; big[1, 0, 191, 0] = 0.7; big[1, 1, 100, 100] = 0.1
    big[0, 0, 76:79, 130:133] += 0.05
    for view in (lambda t: t[:1, :, :192, :192], lambda t: t[:, :, 4:196, 8:200]):
        x = view(big.to(dt).cuda())
        wp, wv = opeaks.global_peaks(x.float().cpu().contiguous(), 0.2, "integral")
        k, v = CenteredInstancePostproc(0.2, "integral")(x, output_stride=1)
        eq(np.isnan(npy(k[:, 0])), np.isnan(npy(wp)))
        close(npy(k[:, 0]), npy(wp), atol=1e-4)
        eq(npy(v[:, 0]), npy(wv))


def test_cfg4_one_full_frame_of_targets_vs_oracle():
    """One cfg4 frame (32 nodes / 31 edges / 8 instances, 1024^2, stride 2, sigma 2.5) of make_multi_confmaps and
    make_pafs against the oracle's restatement of data/confidence_maps.py:132-166 and data/edge_maps.py:167-220
    (~5 s of CPU), plus the 8-frames-per-launch batched path on the same frame."""
    from oracle import synth
    from oracle import targets as ot
    from sleap_nn_b200.data import confidence_maps as cm
    from sleap_nn_b200.data import edge_maps as em
    from sleap_nn_b200.data.batched_targets import BatchedTargets

    n_nodes, n_inst, hw, stride = 32, 8, (1024, 1024), 2
    edges = synth.chain_edges(n_nodes)
    poses = synth.make_poses(17, 1, n_inst, n_nodes, hw, margin=150.0, step=30.0, edges=edges)
    want_cm = ot.generate_multiconfmaps(poses, hw, n_inst, 2.5, stride)
    want_pf = ot.generate_pafs(poses, hw, 2.5, stride, torch.tensor(edges), False)
    got_cm = cm.generate_multiconfmaps(poses.cuda(), hw, n_inst, sigma=2.5, output_stride=stride)
    got_pf = em.generate_pafs(poses.cuda(), hw, sigma=2.5, output_stride=stride, edge_inds=torch.tensor(edges))
    assert got_cm.shape == want_cm.shape == (1, 32, 512, 512) and got_pf.shape == want_pf.shape == (31, 2, 512, 512)
    close(npy(got_cm), npy(want_cm), rtol=1e-5, atol=1.2e-38)
    eq(npy(got_pf) == 0, npy(want_pf) == 0)
    close(npy(got_pf), npy(want_pf), rtol=1e-5, atol=1e-30)
    bt = BatchedTargets(hw, device=torch.device("cuda"))
    batch = poses.expand(8, -1, -1, -1).contiguous()
    tg = bt.bottomup(batch, torch.tensor([n_inst] * 8), edges, confmap_sigma=2.5, confmap_stride=stride, paf_sigma=2.5,
                     paf_stride=stride)
    for b in (0, 7):
        eq(npy(tg["confidence_maps"][b]).reshape(32, 512, 512), npy(got_cm[0]))
        eq(npy(tg["part_affinity_fields"][b]).reshape(31, 2, 512, 512), npy(got_pf))


def test_second_device_runs_on_that_device():
    """with cuda:0 current, a pipeline built for cuda:1 must LAUNCH on cuda:1 (not reach its tables over peer
    access from cuda:0): an event recorded on cuda:1's stream right after the call must complete only after the chain."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from sleap_nn_b200 import synthetic
    from sleap_nn_b200.pipeline import BottomUpPostproc

    d1 = torch.device("cuda", 1)
    Nn, hw, stride, B = 5, (256, 256), 2, 4
    edges = synthetic.chain_edges(Nn)
    poses = synthetic.random_poses(0, B, 2, Nn, hw, edges, margin=60.0, step=24.0)
    cms, pafs = synthetic.render_batch(poses, hw, stride, edges, d1)
    torch.cuda.set_device(0)
    pipe = BottomUpPostproc(Nn, edges, B, (128, 128), cms_stride=stride, pafs_stride=stride, device=d1)
    pipe.buf["n_inst"].fill_(-7)
    torch.cuda.synchronize(d1)
    assert torch.cuda.current_device() == 0
    res = pipe(cms, pafs)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(d1))
    ev.synchronize()  # orders with d1's stream ONLY: results must be there without a device-wide sync on cuda:0
    assert res.n_instances.device == d1 and bool((res.n_instances == 2).all())
    assert torch.cuda.current_device() == 0
