"""GPU parity of the fused bottom-up pipeline (one C call per batch) vs reference goldens and,
at BASELINE cfg3 map size, vs the CPU oracle run on the same device-rendered frames."""

import numpy as np
import pytest
import torch

from tests.helpers import T, close, eq, golden, npy, ragged

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ref_pipeline_mice.npz", "ref_pipeline_tree.npz"])
def test_fused_pipeline_matches_reference_golden(name):
    from sleap_nn_b200.pipeline import BottomUpPostproc

    d = golden(name)
    cms, pafs = T(d["cms"]).cuda(), T(d["pafs"]).cuda()
    B, Nn, H, W = cms.shape
    mip = float(d["min_instance_peaks"])
    mip = int(mip) if mip == int(mip) else mip
    kw = dict(cms_stride=int(d["stride"]), pafs_stride=int(d["stride"]), min_instance_peaks=mip)
    side = torch.cuda.Stream(priority=-1)
    variants = {
        "fused": BottomUpPostproc(Nn, d["edges"].tolist(), B, (H, W), **kw),
        "fused-lean-2streams": BottomUpPostproc(Nn, d["edges"].tolist(), B, (H, W), keep_tables=False,
                                                tail_stream=side, **kw),
        "unfused": BottomUpPostproc(Nn, d["edges"].tolist(), B, (H, W), fused_tail=False, **kw),
    }
    assert variants["fused"].fused and not variants["unfused"].fused
    assert variants["fused"].launches_per_call == 2 and variants["unfused"].launches_per_call == 6
    for name_, pipe in variants.items():
      assert list(pipe.sorted_edge_inds) == d["sorted_edge_inds"].tolist()
      for layout in ("nchw", "view"):
        res = pipe(cms, pafs if layout == "nchw" else pafs.permute(0, 2, 3, 1))
        inst, pv, sc = res.to_lists()
        eq(npy(res.n_peaks), np.bincount(d["pk_s"], minlength=B).astype(np.int32))
        want_inst, want_pv, want_sc = ragged(d, "inst"), ragged(d, "inst_pv"), ragged(d, "inst_sc")
        for b in range(B):
            assert inst[b].shape == want_inst[b].shape, name_
            eq(np.isnan(npy(inst[b])), np.isnan(npy(want_inst[b])))
            close(npy(inst[b]), npy(want_inst[b]), atol=1e-4)
            eq(npy(pv[b]), npy(want_pv[b]))
            close(npy(sc[b]), npy(want_sc[b]), rtol=1e-5, atol=1e-6)
    # fused and unfused tails must agree bit for bit on every table they both write
    a, u = variants["fused"], variants["unfused"]
    a(cms, pafs); u(cms, pafs)
    torch.cuda.synchronize()
    for k in ("frame_count", "peak_xy", "peak_val", "peak_chan", "n_inst", "m_count"):
        if k.startswith("peak"):
            n = npy(a.buf["frame_count"])
            for b in range(B):
                eq(npy(a.buf[k][b, : n[b]]), npy(u.buf[k][b, : n[b]]))
        else:
            eq(npy(a.buf[k]), npy(u.buf[k]))
    ni = npy(a.buf["n_inst"])
    for b in range(B):
        eq(npy(a.buf["inst_xy"][b, : ni[b]]), npy(u.buf["inst_xy"][b, : ni[b]]))
        eq(npy(a.buf["inst_score"][b, : ni[b]]), npy(u.buf["inst_score"][b, : ni[b]]))


def test_fused_pipeline_full_size_vs_oracle():
    """cfg3 geometry (5 nodes / 4 edges, 1024^2 frames, stride 2 -> 512^2 maps), 6 frames."""
    from oracle import paf as opaf
    from oracle import peaks as opeaks
    from oracle.synth import split_by_sample
    from sleap_nn_b200 import synthetic
    from sleap_nn_b200.pipeline import BottomUpPostproc

    B, n_inst, Nn, hw, stride = 6, 2, 5, (1024, 1024), 2
    edges = synthetic.chain_edges(Nn)
    poses = synthetic.random_poses(3, B, n_inst, Nn, hw, edges)
    dev = torch.device("cuda")
    cms, pafs = synthetic.render_batch(poses, hw, stride, edges, dev, seed=3)
    assert cms.shape == (B, Nn, 512, 512) and pafs.shape == (B, 8, 512, 512)
    pipe = BottomUpPostproc(Nn, edges, B, (512, 512), cms_stride=stride, pafs_stride=stride)
    res = pipe(cms, pafs)
    inst, pv, sc = res.to_lists()
    # oracle on the very same maps
    c_cpu, p_cpu = cms.cpu(), pafs.cpu()
    pts, vals, si, ci = opeaks.local_peaks(c_cpu, 0.2, "integral")
    eq(npy(res.n_peaks), np.bincount(npy(si), minlength=B).astype(np.int32))
    pts = pts * stride
    peaks, pvs, pcs = (split_by_sample(x, si, B) for x in (pts, vals, ci))
    want = opaf.predict(p_cpu.permute(0, 2, 3, 1), peaks, pvs, pcs, edges, Nn, stride)
    for b in range(B):
        n = int(res.n_peaks[b])
        close(npy(res.peaks[b, :n]), npy(peaks[b]), atol=1e-4)
        eq(npy(res.peak_vals[b, :n]), npy(pvs[b])); eq(npy(res.peak_channels[b, :n]), npy(pcs[b]))
        assert inst[b].shape == want[0][b].shape == (n_inst, Nn, 2)
        eq(np.isnan(npy(inst[b])), np.isnan(npy(want[0][b])))
        close(npy(inst[b]), npy(want[0][b]), atol=1e-4)
        eq(npy(pv[b]), npy(want[1][b]))
        close(npy(sc[b]), npy(want[2][b]), rtol=1e-5, atol=1e-6)
        # every planted animal is recovered within a pixel
        got = npy(inst[b])
        for a in range(n_inst):
            dmin = min(np.nanmax(np.abs(got[k] - npy(poses[b, a]))) for k in range(n_inst))
            assert dmin < 1.5


def test_capacity_overflow_is_reported_not_silent():
    from sleap_nn_b200.pipeline import BottomUpPostproc

    g = torch.Generator().manual_seed(0)
    cms = torch.rand((2, 3, 64, 64), generator=g).cuda()  # noise: hundreds of peaks per frame
    pafs = torch.zeros((2, 4, 64, 64)).cuda()
    pipe = BottomUpPostproc(3, [(0, 1), (1, 2)], 2, (64, 64), peak_cap=16, cand_cap=64, match_cap=16, inst_cap=4)
    res = pipe(cms, pafs)
    assert int(res.n_peaks.max()) > 16
    with pytest.raises(RuntimeError, match="overflowed"):
        res.to_lists()


# ----------------------------------------------------------------- group_scored_batch epilogue
def _f1_pipe(d, **kw):
    from sleap_nn_b200.pipeline import BottomUpPostproc

    cms = T(d["cms"]).cuda()
    B, Nn, H, W = cms.shape
    mip = float(d["min_instance_peaks"])
    return BottomUpPostproc(Nn, d["edges"].tolist(), B, (H, W), cms_stride=int(d["stride"]), pafs_stride=int(d["stride"]),
                            min_instance_peaks=int(mip) if mip == int(mip) else mip, **kw)


@pytest.mark.parametrize("fixed", [True, False])
def test_outputs_epilogue_matches_reference_group_scored_batch(fixed):
    """NaN padding, top-N by score, scale undo and the skip short-circuit vs the reference's own
    group_scored_batch (inference/streaming.py:147-255) run on the same frames (ref_f1_outputs.npz)."""
    d, f1 = golden("ref_pipeline_tree.npz"), golden("ref_f1_outputs.npz")
    cms, pafs = T(d["cms"]).cuda(), T(d["pafs"]).cuda()
    for tag in f1["cases"].tolist():
        mi = int(f1[f"{tag}_max_instances"])
        mi = None if mi < 0 else mi
        skip = bool(f1[f"{tag}_skip"])
        if fixed and mi is None:
            continue
        kw = dict(max_peaks_per_node=(int(f1["max_node_peaks"]) - 1 if skip else int(f1["max_node_peaks"])))
        pipe = _f1_pipe(d, max_instances=mi if fixed else None, **kw)
        res = pipe(cms, pafs, input_scale=float(f1[f"{tag}_input_scale"]), eff_scale=T(f1[f"{tag}_eff"]))
        k, v, s = res.outputs() if fixed else res.outputs(mi)
        assert bool(res.skip_flag.item()) == skip
        want_k, want_v, want_s = f1[f"{tag}_kpts"], f1[f"{tag}_vals"], f1[f"{tag}_scores"]
        assert tuple(k.shape) == want_k.shape, (tag, tuple(k.shape), want_k.shape)
        eq(np.isnan(npy(k)), np.isnan(want_k))
        close(npy(k), want_k, atol=1e-4)
        eq(npy(v), want_v)
        close(npy(s), want_s, rtol=1e-5, atol=1e-6)
        if fixed:
            assert pipe.launches_per_call == 3 and res.pred_keypoints is k


def test_outputs_topn_order_with_ties_and_nan():
    """np.argsort(scores)[::-1][:n]: NaN first, then descending, equal scores -> higher index first."""
    from sleap_nn_b200 import _native as N

    dev = torch.device("cuda", 0)
    scores = torch.tensor([[0.5, float("nan"), 0.9, 0.5, 0.1, 0.9]], device=dev)
    n = torch.tensor([6], dtype=torch.int32, device=dev)
    xy = torch.arange(6 * 2 * 2, dtype=torch.float32, device=dev).reshape(1, 6, 2, 2)
    val = torch.arange(6 * 2, dtype=torch.float32, device=dev).reshape(1, 6, 2)
    I = 4
    k = torch.empty((1, I, 2, 2), device=dev); v = torch.empty((1, I, 2), device=dev); s = torch.empty((1, I), device=dev)
    N.check(N.lib.snb_bottomup_outputs(N.ptr(n), N.ptr(xy), N.ptr(val), N.ptr(scores), 1, 6, 2, I, 1.0, None, None,
                                       N.ptr(k), N.ptr(v), N.ptr(s), N.stream_ptr(dev)), "outputs")
    order = np.argsort(scores[0].cpu().numpy(), kind="stable")[::-1][:I]
    assert order.tolist() == [1, 5, 2, 3]
    eq(npy(s[0]), npy(scores[0].cpu())[order])
    eq(npy(k[0]), npy(xy[0].cpu())[order])
    eq(npy(v[0]), npy(val[0].cpu())[order])


def test_host_stream_pipelines_batches_and_matches_run_host():
    """BottomUpHostStream (depth-2 H2D / compute / D2H pipeline) returns, in order, what run_host returns."""
    from sleap_nn_b200 import synthetic
    from sleap_nn_b200.pipeline import BottomUpHostStream, BottomUpPostproc

    dev = torch.device("cuda", 0)
    edges = synthetic.chain_edges(5)
    make = lambda: BottomUpPostproc(5, edges, 3, (128, 128), cms_stride=2, pafs_stride=2, device=dev)
    batches = []
    for s in range(5):
        poses = synthetic.random_poses(40 + s, 3, 2, 5, (256, 256), edges, margin=60.0, step=24.0)
        cms, pafs = synthetic.render_batch(poses, (256, 256), 2, edges, dev, seed=s)
        batches.append((cms.cpu().pin_memory(), pafs.cpu().pin_memory() if s % 2 == 0 else pafs.cpu()))  # pinned -> zero-copy
    ref = make()
    want = [ref.run_host(c, p) for c, p in batches]
    hs = BottomUpHostStream(make, depth=2)
    got = []
    for c, p in batches:
        r = hs.submit(c, p)
        if r is not None:
            got.append(r)
    got += hs.drain()
    assert len(got) == len(want) and hs.drain() == []
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            assert len(a) == len(b) == 3
            for x, y in zip(a, b):
                assert torch.equal(torch.nan_to_num(x, nan=-1.0), torch.nan_to_num(y, nan=-1.0))
    assert sum(len(x) for x in got[0][0]) == 6


def test_cuda_graph_rotation_replays_the_chain():
    """capture_rotation: a graph replay leaves in every pipe's tables exactly what the eager calls leave."""
    from sleap_nn_b200 import synthetic
    from sleap_nn_b200.pipeline import BottomUpPostproc, capture_rotation

    dev = torch.device("cuda", 0)
    edges = synthetic.chain_edges(5)
    inputs = []
    for s in range(3):
        poses = synthetic.random_poses(70 + s, 4, 2, 5, (256, 256), edges, margin=60.0, step=24.0)
        inputs.append(synthetic.render_batch(poses, (256, 256), 2, edges, dev, seed=s))
    pipes = [BottomUpPostproc(5, edges, 4, (128, 128), cms_stride=2, pafs_stride=2, device=dev, max_instances=4) for _ in range(2)]
    graph, n = capture_rotation(pipes, inputs)
    assert n == 6
    # the last step that touches pipe k in a rotation of 6 over 2 pipes / 3 inputs: step 4 -> pipe 0 / input 1, step 5 -> pipe 1 / input 2
    want = []
    for k, inp in ((0, 1), (1, 2)):
        r = pipes[k](*inputs[inp])
        want.append([t.clone() for t in (r.n_instances, r.instances, r.peak_scores, r.instance_scores, r.pred_keypoints)])
    for p in pipes:
        for name in ("n_inst", "inst_xy", "inst_val", "inst_score"):
            p.buf[name].zero_()
        p._out[0].zero_()
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    for k in range(2):
        b = pipes[k].buf
        got = [b["n_inst"], b["inst_xy"], b["inst_val"], b["inst_score"], pipes[k]._out[0]]
        assert int(got[0].sum()) == 8
        n_i = got[0].tolist()
        for g, w in zip(got[1:4], want[k][1:4]):
            for f in range(4):  # rows beyond the count are scratch
                assert torch.equal(torch.nan_to_num(g[f, : n_i[f]], nan=-1.0), torch.nan_to_num(w[f, : n_i[f]], nan=-1.0))
        assert torch.equal(got[0], want[k][0])
        assert torch.equal(torch.nan_to_num(got[4], nan=-1.0), torch.nan_to_num(want[k][4], nan=-1.0))


def test_group_scored_batch_seam_matches_reference():
    """The ScoredBatch -> group_scored_batch seam (inference/streaming.py:42-255) as a drop-in: the scored batch comes
    from this package's own peak / PAF-scoring kernels, the grouped outputs are the reference's (ref_f1_outputs.npz)."""
    import types

    from sleap_nn_b200.inference.paf_grouping import PAFScorer
    from sleap_nn_b200.inference.peak_finding import find_local_peaks
    from sleap_nn_b200.inference.streaming import GroupingParams, ScoredBatch, group_scored_batch

    d, f1 = golden("ref_pipeline_tree.npz"), golden("ref_f1_outputs.npz")
    cms, pafs = T(d["cms"]).cuda(), T(d["pafs"]).cuda()
    edges, n_nodes, stride = d["edges"].tolist(), int(d["n_nodes"]), int(d["stride"])
    B = cms.shape[0]
    pts, vals, si, ci = find_local_peaks(cms, threshold=0.2, refinement="integral")
    pts = pts * stride
    split = lambda x: [x[si == b] for b in range(B)]
    peaks, pvals, pch = split(pts), split(vals), split(ci)
    kwargs = dict(part_names=[str(i) for i in range(n_nodes)], edges=[(str(a), str(b)) for a, b in edges], pafs_stride=stride,
                  max_edge_length_ratio=0.25, dist_penalty_weight=1.0, n_points=10,
                  min_instance_peaks=int(d["min_instance_peaks"]), min_line_scores=0.25)
    ei, epi, ls = PAFScorer(**kwargs).score_paf_lines(pafs.permute(0, 2, 3, 1), peaks, pch)
    for tag in f1["cases"].tolist():
        mi = int(f1[f"{tag}_max_instances"])
        skip = bool(f1[f"{tag}_skip"])
        info = types.SimpleNamespace(eff_scale=T(f1[f"{tag}_eff"]), input_scale=float(f1[f"{tag}_input_scale"]))
        sb = ScoredBatch(cms_peaks=peaks, cms_peak_vals=pvals, cms_peak_channel_inds=pch, edge_inds=[] if skip else ei,
                         edge_peak_inds=[] if skip else epi, line_scores=[] if skip else ls, info=info, n_samples=B,
                         n_nodes=n_nodes, skip_paf=skip, cms=cms)
        res = group_scored_batch(sb.to_cpu(), GroupingParams(paf_scorer_kwargs=kwargs, max_instances=None if mi < 0 else mi,
                                                             return_confmaps=True, return_paf_graph=True))
        want_k, want_v, want_s = f1[f"{tag}_kpts"], f1[f"{tag}_vals"], f1[f"{tag}_scores"]
        assert not res.pred_keypoints.is_cuda and tuple(res.pred_keypoints.shape) == want_k.shape, tag
        eq(np.isnan(npy(res.pred_keypoints)), np.isnan(want_k))
        close(npy(res.pred_keypoints), want_k, atol=1e-4)
        eq(npy(res.pred_peak_values), want_v)
        close(npy(res.instance_scores), want_s, rtol=1e-5, atol=1e-6)
        assert res.pred_confmaps is not None and len(res.pred_paf_graph) == 4
        assert res.pred_paf_graph[0].shape[0] == sum(p.shape[0] for p in peaks)


def test_fused_pipeline_busy_flies_frames_vs_oracle():
    """cfg4 geometry (32 nodes / 31 edges / 8 animals per 1024^2 frame: 256 peaks, ~2 k candidates, 248 connections per
    frame) through the fused 2-launch chain vs the oracle on the same maps: the busy-frame paths of the tail (4 peaks
    refined at once, per-edge match ranges, chunk-parallel assembly)."""
    from oracle import paf as opaf
    from oracle import peaks as opeaks
    from oracle.synth import split_by_sample
    from sleap_nn_b200 import synthetic
    from sleap_nn_b200.pipeline import BottomUpPostproc

    B, n_inst, Nn, hw, stride = 3, 8, 32, (1024, 1024), 2
    edges = synthetic.chain_edges(Nn)
    poses = synthetic.random_poses(5, B, n_inst, Nn, hw, edges, margin=200.0, step=24.0, min_limb=8.0, min_sep=10.0)
    dev = torch.device("cuda")
    cms, pafs = synthetic.render_batch(poses, hw, stride, edges, dev, seed=5)
    pipe = BottomUpPostproc(Nn, edges, B, (512, 512), cms_stride=stride, pafs_stride=stride, peak_cap=512, cand_cap=4096,
                            match_cap=512, inst_cap=32)
    assert pipe.fused and pipe.launches_per_call == 2
    res = pipe(cms, pafs)
    inst, pv, sc = res.to_lists()
    c_cpu, p_cpu = cms.cpu(), pafs.cpu()
    pts, vals, si, ci = opeaks.local_peaks(c_cpu, 0.2, "integral")
    eq(npy(res.n_peaks), np.bincount(npy(si), minlength=B).astype(np.int32))
    peaks, pvs, pcs = (split_by_sample(x, si, B) for x in (pts * stride, vals, ci))
    want = opaf.predict(p_cpu.permute(0, 2, 3, 1), peaks, pvs, pcs, edges, Nn, stride)
    for b in range(B):
        assert inst[b].shape == want[0][b].shape and inst[b].shape[0] == n_inst
        eq(np.isnan(npy(inst[b])), np.isnan(npy(want[0][b])))
        close(npy(inst[b]), npy(want[0][b]), atol=1e-4)
        eq(npy(pv[b]), npy(want[1][b]))
        close(npy(sc[b]), npy(want[2][b]), rtol=1e-5, atol=1e-6)  # a sum of 31 line scores: observed |d| < 2e-5 on scores of ~30


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["branching_tree", "children_first_order", "two_parents_all_edges"])
def test_fused_pipeline_assembly_paths_vs_oracle(case):
    """The tail's assembly has a parallel path for skeletons visited as a forest, parents first (instances = trees over
    the peaks: what toposort_edges' order gives), and the reference's sequential loop for everything else.  Same maps
    and the same visiting order through the oracle: a branching tree in the product's own order (forest path); the
    same tree visited children first, and a node with two parents with every edge visited (the connections then break
    the forest conditions - a destination used as a source earlier, a peak that is a destination twice - and the
    sequential loop runs)."""
    from oracle import paf as opaf
    from oracle import peaks as opeaks
    from oracle.synth import split_by_sample
    from sleap_nn_b200 import synthetic
    from sleap_nn_b200.pipeline import BottomUpPostproc

    # 14 nodes / 13 edges: from 12 visited edges on the tail tries the forest path first
    tree = [(0, 1), (0, 2), (1, 3), (1, 4), (2, 5), (5, 6), (5, 7), (3, 8), (3, 9), (6, 10), (6, 11), (7, 12), (7, 13)]
    edges = tree + [(4, 13)] if case == "two_parents_all_edges" else tree
    B, n_inst, Nn, hw, stride = 4, 4, 14, (512, 512), 2
    poses = synthetic.random_poses(11, B, n_inst, Nn, hw, tree, margin=140.0, step=26.0, min_limb=10.0, min_sep=14.0)
    dev = torch.device("cuda")
    cms, pafs = synthetic.render_batch(poses, hw, stride, edges, dev, seed=11)
    order = None
    if case == "children_first_order":
        order = tuple(reversed(BottomUpPostproc(Nn, edges, B, (256, 256)).sorted_edge_inds))
    elif case == "two_parents_all_edges":
        order = tuple(range(len(edges)))  # (7, 13) and (4, 13) are both visited
    pipe = BottomUpPostproc(Nn, edges, B, (256, 256), cms_stride=stride, pafs_stride=stride, sorted_edge_inds=order)
    assert pipe.fused
    res = pipe(cms, pafs)
    c_cpu, p_cpu = cms.cpu(), pafs.cpu()
    pts, vals, si, ci = opeaks.local_peaks(c_cpu, 0.2, "integral")
    peaks, pvs, pcs = (split_by_sample(x, si, B) for x in (pts * stride, vals, ci))
    try:
        want = opaf.predict(p_cpu.permute(0, 2, 3, 1), peaks, pvs, pcs, edges, Nn, stride, sorted_edge_inds=order)
    except (AssertionError, KeyError) as e:  # the reference's own sanity check (ops/paf.py:866-873) rejects this order
        with pytest.raises(type(e)):
            res.to_lists()
        return
    inst, pv, sc = res.to_lists()
    for b in range(B):
        assert inst[b].shape == want[0][b].shape
        eq(np.isnan(npy(inst[b])), np.isnan(npy(want[0][b])))
        close(npy(inst[b]), npy(want[0][b]), atol=1e-4)
        eq(npy(pv[b]), npy(want[1][b]))
        close(npy(sc[b]), npy(want[2][b]), rtol=1e-5, atol=1e-6)
    if case == "branching_tree":
        assert all(x.shape[0] == n_inst for x in inst)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process():
    """Inputs on cuda:1 while cuda:0 is the current device: kernels run on the input's device and stream, results come
    back on that device and equal the cuda:0 results (per-device function attributes, workspaces and tables)."""
    from sleap_nn_b200 import synthetic
    from sleap_nn_b200.data.confidence_maps import make_multi_confmaps
    from sleap_nn_b200.inference import peak_finding as pf
    from sleap_nn_b200.pipeline import BottomUpPostproc

    B, n_inst, Nn, hw, stride = 4, 2, 5, (512, 512), 2
    edges = synthetic.chain_edges(Nn)
    poses = synthetic.random_poses(9, B, n_inst, Nn, hw, edges, margin=60.0, step=24.0)
    d0, d1 = torch.device("cuda", 0), torch.device("cuda", 1)
    torch.cuda.set_device(d0)
    cms0, pafs0 = synthetic.render_batch(poses, hw, stride, edges, d0, seed=9)
    cms1, pafs1 = cms0.to(d1), pafs0.to(d1)
    a, b = pf.find_local_peaks(cms0, 0.2, "integral"), pf.find_local_peaks(cms1, 0.2, "integral")
    for x, y in zip(a, b):
        assert y.device == d1
        eq(npy(x), npy(y))
    g0, g1 = pf.find_global_peaks(cms0[:, :, :100, :100], 0.2, "integral"), pf.find_global_peaks(cms1[:, :, :100, :100], 0.2, "integral")
    assert g1[0].device == d1
    eq(npy(g0[0]), npy(g1[0])); eq(npy(g0[1]), npy(g1[1]))
    r0 = BottomUpPostproc(Nn, edges, B, tuple(cms0.shape[-2:]), cms_stride=stride, pafs_stride=stride, device=d0)(cms0, pafs0)
    r1 = BottomUpPostproc(Nn, edges, B, tuple(cms1.shape[-2:]), cms_stride=stride, pafs_stride=stride, device=d1)(cms1, pafs1)
    assert r1.instances.device == d1 and torch.cuda.current_device() == 0
    for x, y in zip(r0.to_lists(), r1.to_lists()):
        for u, v in zip(x, y):
            eq(npy(u), npy(v))
    xv, yv = torch.arange(0, 64, dtype=torch.float32), torch.arange(0, 48, dtype=torch.float32)
    pts = torch.rand((1, 2, 3, 2)) * 40
    t0, t1 = make_multi_confmaps(pts.to(d0), xv, yv, 2.0), make_multi_confmaps(pts.to(d1), xv, yv, 2.0)
    assert t1.device == d1
    eq(npy(t0), npy(t1))


@pytest.mark.parametrize("B", [1, 5, 16])
def test_cluster_tail_equals_single_cta_tail_and_oracle(B):
    """Small batches of skeletons with >= 8 edges run the tail as a 4-CTA cluster per frame (refined peaks, line scores and
    matches exchanged through distributed shared memory): every table must equal the one-CTA-per-frame tail's bit for
    bit, and the instances the oracle's."""
    from oracle import paf as opaf
    from oracle import peaks as opeaks
    from oracle.synth import split_by_sample
    from sleap_nn_b200 import synthetic
    from sleap_nn_b200.pipeline import BottomUpPostproc

    n_inst, Nn, hw, stride = 4, 12, (384, 384), 2
    edges = synthetic.chain_edges(Nn)
    poses = synthetic.random_poses(40 + B, B, n_inst, Nn, hw, edges, margin=90.0, step=20.0, min_limb=8.0, min_sep=10.0)
    cms, pafs = synthetic.render_batch(poses, hw, stride, edges, torch.device("cuda"), seed=B)
    kw = dict(cms_stride=stride, pafs_stride=stride, peak_cap=256, cand_cap=2048, match_cap=256, inst_cap=32)
    clu = BottomUpPostproc(Nn, edges, B, (192, 192), **kw)
    one = BottomUpPostproc(Nn, edges, B, (192, 192), tail_cluster=False, **kw)
    r_c, r_o = clu(cms, pafs), one(cms, pafs)
    torch.cuda.synchronize()
    npk, nin = npy(one.buf["frame_count"]), npy(one.buf["n_inst"])
    eq(npy(clu.buf["frame_count"]), npk); eq(npy(clu.buf["n_inst"]), nin); eq(npy(clu.buf["m_count"]), npy(one.buf["m_count"]))
    for b in range(B):
        for k in ("peak_xy", "peak_val", "peak_chan"):
            eq(npy(clu.buf[k][b, : npk[b]]), npy(one.buf[k][b, : npk[b]]))
        for k in ("inst_xy", "inst_val", "inst_score"):
            eq(npy(clu.buf[k][b, : nin[b]]), npy(one.buf[k][b, : nin[b]]))
        mc = int(one.buf["m_count"][b])
        o = b * one.caps["match_cap"]
        for k in ("m_edge", "m_src", "m_dst", "m_score"):
            eq(npy(clu.buf[k][o : o + mc]), npy(one.buf[k][o : o + mc]))
        no = b * one.caps["cand_cap"]
        nc = int(one.buf["edge_off"][b, -1])
        eq(npy(clu.buf["cand_score"][no : no + nc]), npy(one.buf["cand_score"][no : no + nc]))
    inst, pv, sc = r_c.to_lists()
    pts, vals, si, ci = opeaks.local_peaks(cms.cpu(), 0.2, "integral")
    peaks, pvs, pcs = (split_by_sample(x, si, B) for x in (pts * stride, vals, ci))
    want = opaf.predict(pafs.cpu().permute(0, 2, 3, 1), peaks, pvs, pcs, edges, Nn, stride)
    for b in range(B):
        assert inst[b].shape == want[0][b].shape
        eq(np.isnan(npy(inst[b])), np.isnan(npy(want[0][b])))
        close(npy(inst[b]), npy(want[0][b]), atol=1e-4)
        eq(npy(pv[b]), npy(want[1][b]))
        close(npy(sc[b]), npy(want[2][b]), rtol=1e-5, atol=1e-6)
