"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every array under tests/golden/*.npz is an input fed to, or an output produced by,
the reference's own functions loaded in place by oracle/ref_loader.py.  The GPU box
has no reference tree, so these files are what pins both the oracle and the CUDA
path there.  Inputs are stored next to the outputs (never re-synthesised on the box:
libm / SLEEF differences between hosts could move an input by an ulp).
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader, synth  # noqa: E402

REF_ASSETS = os.path.join(ref_loader.REF_ROOT, "tests/assets/inference")


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: npy(v) for k, v in arrays.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(arrays)} arrays")


def ragged(prefix, tensors, out):
    """Store a list of per-sample arrays as prefix_0, prefix_1, ... plus prefix_n."""
    out[f"{prefix}_n"] = np.int64(len(tensors))
    for i, t in enumerate(tensors):
        out[f"{prefix}_{i}"] = npy(t)


def peaks_minimal(R):
    cms = torch.load(os.path.join(REF_ASSETS, "minimal_cms.pt")).unsqueeze(0)
    bboxes = torch.load(os.path.join(REF_ASSETS, "minimal_bboxes.pt"))
    out = dict(cms=cms, bboxes=bboxes)
    lr = R.peaks.find_local_peaks_rough(cms)
    out.update(lr_pts=lr[0], lr_vals=lr[1], lr_s=lr[2], lr_c=lr[3])
    li = R.peaks.find_local_peaks(cms, refinement="integral")
    out.update(li_pts=li[0], li_vals=li[1])
    gr = R.peaks.find_global_peaks_rough(cms, threshold=0.1)
    out.update(gr_pts=gr[0], gr_vals=gr[1])
    gi = R.peaks.find_global_peaks(cms, threshold=0.2, refinement="integral")
    out.update(gi_pts=gi[0], gi_vals=gi[1])
    planes = cms.reshape(13, 1, 80, 80)
    crops = R.crops.crop_bboxes(planes, bboxes, torch.arange(13))
    gv = torch.arange(5, dtype=torch.float32) - 2
    dx, dy = R.peaks.integral_regression(crops, gv, gv)
    out.update(crops=crops, ir_dx=dx, ir_dy=dy, dil=R.peaks.morphological_dilation(planes, None))
    save("ref_peaks_minimal.npz", **out)


def peaks_random(R):
    g = torch.Generator().manual_seed(1234)
    cms = torch.rand((3, 4, 24, 20), generator=g)
    cms[1, 2, 5, 7] = float("nan")  # NaN neighbours suppress peaks, NaN itself is never one
    cms[2, 0] = 0.05  # an all-below-threshold plane -> NaN global peak
    out = dict(cms=cms)
    for tag, thr in (("t02", 0.2), ("t09", 0.9)):
        r = R.peaks.find_local_peaks_rough(cms, threshold=thr)
        out.update({f"lr_{tag}_pts": r[0], f"lr_{tag}_vals": r[1], f"lr_{tag}_s": r[2], f"lr_{tag}_c": r[3]})
    for size in (3, 4, 5, 7):
        r = R.peaks.find_local_peaks(cms, threshold=0.9, refinement="integral", integral_patch_size=size)
        out[f"li_t09_p{size}_pts"] = r[0]
    clean = torch.nan_to_num(cms, nan=0.3)
    for tag, thr in (("t01", 0.1), ("t0999", 0.999)):
        r = R.peaks.find_global_peaks_rough(clean, threshold=thr)
        out.update({f"gr_{tag}_pts": r[0], f"gr_{tag}_vals": r[1]})
    for size in (3, 4, 5):
        r = R.peaks.find_global_peaks(clean, threshold=0.2, refinement="integral", integral_patch_size=size)
        out[f"gi_p{size}_pts"] = r[0]
        out[f"gi_p{size}_vals"] = r[1]
    out["clean"] = clean
    # crop_bboxes on fractional / negative / far-out-of-bounds centroids, even size, 2 channels
    imgs = torch.rand((2, 2, 30, 26), generator=g)
    cents = torch.tensor([[20.3, 18.7], [2.4, 2.4], [0.6, 3.0], [10.0, 10.0], [1.4, 1.4], [1000.0, 1000.0],
                          [-1000.0, -3.0], [25.0, 29.0]])
    sidx = torch.tensor([0, 1, 0, 1, 1, 0, 1, 0])
    for (bh, bw) in ((6, 6), (5, 7)):
        bb = R.instance_cropping.make_centered_bboxes(cents, bh, bw)
        out[f"bb_{bh}x{bw}"] = bb
        out[f"crops_{bh}x{bw}"] = R.crops.crop_bboxes(imgs, bb, sidx)
    out.update(imgs=imgs, cents=cents, sidx=sidx)
    save("ref_peaks_random.npz", **out)


def paf_units(R):
    P = R.paf
    out = {}
    g = torch.Generator().manual_seed(7)
    # line subscripts: random sub-pixel peaks, several (stride, n_points)
    peaks = torch.rand((60, 2), generator=g) * torch.tensor([300.0, 200.0])
    epi = torch.randint(0, 60, (400, 2), generator=g)
    ei = torch.randint(0, 5, (400,), generator=g).to(torch.int32)
    out.update(ls_peaks=peaks, ls_epi=epi, ls_ei=ei)
    for stride, n in ((2, 10), (4, 10), (8, 5), (3, 7), (1, 3)):
        hw = (200 // stride, 300 // stride)
        out[f"ls_s{stride}_n{n}"] = P.make_line_subs(peaks, epi, ei, n, stride, hw)
    # the reference's own unit-test vectors (tests/inference/test_paf_grouping.py)
    pafs = torch.arange(6 * 4 * 2).view(6, 4, 2).float()
    pk = torch.tensor([[0, 0], [4, 8]], dtype=torch.float32)
    e1 = torch.tensor([[0, 1]], dtype=torch.int32)
    z = torch.tensor([0], dtype=torch.int32)
    lines = P.get_paf_lines(pafs, pk, e1, z, 3, 2)
    out.update(ut_lines=lines, ut_score=P.score_paf_lines(lines, pk, e1, max_edge_length=2))
    out["ut_pen"] = P.compute_distance_penalty(torch.tensor([1, 2, 3, 4.0]), 2, dist_penalty_weight=2)
    # scoring on a random field
    field = torch.randn((40, 50, 10), generator=g)
    ch = torch.randint(0, 6, (30,), generator=g).to(torch.int32)
    pk2 = torch.rand((30, 2), generator=g) * torch.tensor([98.0, 78.0])
    edges = [(0, 1), (1, 2), (2, 3), (1, 4), (4, 5)]
    r = P.score_paf_lines_batch(field[None], [pk2], [ch], torch.tensor(edges, dtype=torch.int32), 10, 2, 0.25, 1.0, 6)
    out.update(sc_field=field, sc_ch=ch, sc_peaks=pk2, sc_edges=np.asarray(edges), sc_ei=r[0][0], sc_epi=r[1][0], sc_scores=r[2][0])
    m = P.match_candidates_sample(r[0][0], r[1][0], r[2][0], 5)
    out.update(mt_e=m[0], mt_s=m[1], mt_d=m[2], mt_sc=m[3])
    save("ref_paf_units.npz", **out)


def paf_pipeline(R, name, seed, n_frames, n_inst, n_nodes, img_hw, edges, step, min_instance_peaks=0):
    P = R.paf
    stride = 2
    poses = synth.make_poses(seed, n_frames, n_inst, n_nodes, img_hw, margin=40.0, step=step, edges=edges)
    cms, pafs = synth.render(poses, img_hw, stride, edges, seed=seed)
    good = synth.certify(cms, pafs, edges, n_nodes, stride)
    out = dict(poses=poses, cms=cms, pafs=pafs, edges=np.asarray(edges), certified=good,
               stride=np.int64(stride), n_nodes=np.int64(n_nodes), min_instance_peaks=np.float64(min_instance_peaks))
    pts, vals, si, ci = R.peaks.find_local_peaks(cms, threshold=0.2, refinement="integral")
    out.update(pk_pts=pts, pk_vals=vals, pk_s=si, pk_c=ci)
    pts = pts * stride
    peaks = synth.split_by_sample(pts, si, n_frames)
    pvals = synth.split_by_sample(vals, si, n_frames)
    pch = synth.split_by_sample(ci, si, n_frames)
    names = [str(i) for i in range(n_nodes)]
    scorer = P.PAFScorer(part_names=names, edges=[(str(a), str(b)) for a, b in edges], pafs_stride=stride,
                         min_instance_peaks=min_instance_peaks)
    out["sorted_edge_inds"] = np.asarray(scorer.sorted_edge_inds)
    res = scorer.predict(pafs.permute(0, 2, 3, 1), peaks, pvals, pch)
    ragged("inst", res[0], out)
    ragged("inst_pv", res[1], out)
    ragged("inst_sc", res[2], out)
    ragged("cand_e", res[3], out)
    ragged("cand_p", res[4], out)
    ragged("cand_s", res[5], out)
    m = scorer.match_candidates(res[3], res[4], res[5])
    ragged("m_e", m[0], out)
    ragged("m_s", m[1], out)
    ragged("m_d", m[2], out)
    ragged("m_sc", m[3], out)
    print(name, "certified frames:", good.tolist(), "instances/frame:", [len(x) for x in res[0]])
    save(name, **out)


def assembly_cases(R):
    """Hand-made match lists that hit every branch of the greedy assembly."""
    P = R.paf
    out = {}
    g = np.random.default_rng(5)
    n_cases = 0
    for case in range(24):
        n_nodes = int(g.integers(3, 7))
        edges = synth.star_chain_edges(n_nodes, fan=2) if case % 2 else synth.chain_edges(n_nodes)
        if case % 5 == 0:
            edges = edges[::-1]
        n_per = g.integers(1, 4, n_nodes)
        ch = np.concatenate([np.full(n, k) for k, n in enumerate(n_per)]).astype(np.int32)
        perm = g.permutation(len(ch))
        ch = ch[perm]
        pk = g.uniform(0, 100, (len(ch), 2)).astype(np.float32)
        pv = g.uniform(0.2, 1, len(ch)).astype(np.float32)
        me, ms, md, msc = [], [], [], []
        for k, (a, b) in enumerate(edges):
            # random partial matchings, occasionally re-using a peak twice (case-3 branch)
            for _ in range(int(g.integers(0, 4))):
                me.append(k)
                ms.append(int(g.integers(0, n_per[a])))
                md.append(int(g.integers(0, n_per[b])))
                msc.append(float(g.uniform(0.0, 1.0)))
        order = g.permutation(len(me))
        me, ms, md, msc = (np.asarray(x)[order] for x in (me, ms, md, msc))
        et = [P.EdgeType(a, b) for a, b in edges]
        sorted_inds = P.toposort_edges(et)
        mip = [0, 2, 0.5][case % 3]
        res = P.group_instances_sample(
            torch.from_numpy(pk), torch.from_numpy(pv), torch.from_numpy(ch),
            torch.as_tensor(me, dtype=torch.int32), torch.as_tensor(ms, dtype=torch.int32),
            torch.as_tensor(md, dtype=torch.int32), torch.as_tensor(msc, dtype=torch.float32),
            n_nodes, sorted_inds, et, mip, 0.25)
        pre = f"c{case}_"
        out.update({pre + "edges": np.asarray(edges), pre + "ch": ch, pre + "pk": pk, pre + "pv": pv,
                    pre + "me": me.astype(np.int32), pre + "ms": ms.astype(np.int32), pre + "md": md.astype(np.int32),
                    pre + "msc": msc.astype(np.float32), pre + "sorted": np.asarray(sorted_inds, dtype=np.int64),
                    pre + "mip": np.float64(mip), pre + "mip_is_float": np.bool_(isinstance(mip, float)),
                    pre + "inst": res[0], pre + "inst_pv": res[1], pre + "inst_sc": res[2]})
        n_cases += 1
    out["n_cases"] = np.int64(n_cases)
    # toposort on the reference's two test skeletons + a forest + a DAG with a cross edge
    skels = [
        [(5, 7), (5, 8), (5, 9), (5, 6), (5, 11), (5, 12), (1, 0), (1, 3), (1, 2), (1, 10), (1, 13), (1, 14), (4, 5), (4, 1)],
        [(1, 4), (1, 5), (6, 8), (6, 7), (6, 9), (9, 10), (1, 0), (1, 3), (1, 2), (6, 1)],
        [(0, 1), (2, 3)],
        [(0, 1), (0, 2), (1, 2)],
        [(2, 1), (1, 0)],
    ]
    for i, sk in enumerate(skels):
        out[f"topo{i}_edges"] = np.asarray(sk)
        out[f"topo{i}_order"] = np.asarray(P.toposort_edges([P.EdgeType(a, b) for a, b in sk]), dtype=np.int64)
    out["n_topo"] = np.int64(len(skels))
    save("ref_assembly.npz", **out)


def targets(R):
    C, E, U = R.confidence_maps, R.edge_maps, R.data_utils
    out = {}
    g = torch.Generator().manual_seed(99)
    hw, stride, sigma = (48, 64), 2, 1.5
    xv, yv = U.make_grid_vectors(hw[0], hw[1], stride)
    pts = torch.rand((2, 5, 2), generator=g) * torch.tensor([64.0, 48.0])
    pts[0, 1] = float("nan")
    pts[1, 3, 0] = float("nan")
    out.update(xv=xv, yv=yv, cm_pts=pts, cm=C.make_confmaps(pts, xv, yv, sigma * stride))
    inst = torch.rand((1, 4, 5, 2), generator=g) * torch.tensor([64.0, 48.0])
    inst[0, 2, 1] = float("nan")
    inst[0, 3] = float("nan")
    out.update(mc_pts=inst, mc=C.make_multi_confmaps(inst, xv, yv, sigma * stride))
    out["gmc"] = C.generate_multiconfmaps(inst, hw, 3, sigma, stride)
    out["gmc_centroids"] = C.generate_multiconfmaps(inst[:, :, 0], hw, 4, sigma, stride, is_centroids=True)
    out["gc"] = C.generate_confmaps(inst[:, 0], hw, sigma, stride)
    # wide-range confmaps: far tails reach the denormal band and exact zero
    xv2, yv2 = U.make_grid_vectors(256, 256, 1)
    far = torch.tensor([[[3.25, 7.5], [250.0, 128.75]]])
    out.update(xv2=xv2, yv2=yv2, far_pts=far, far_cm=C.make_confmaps(far, xv2, yv2, 5.0))
    # the reference's 3x3 known-answer vectors (tests/data/test_edge_maps.py)
    xv3, yv3 = U.make_grid_vectors(3, 3, 1)
    s3 = torch.tensor([[1, 0.5], [0, 0]], dtype=torch.float32)
    d3 = torch.tensor([[1, 1.5], [2, 2]], dtype=torch.float32)
    yy, xx = torch.meshgrid(yv3, xv3, indexing="ij")
    out.update(k_src=s3, k_dst=d3, k_dist=E.distance_to_edge(torch.stack((xx, yy), -1), s3, d3),
               k_em=E.make_edge_maps(xv3, yv3, s3, d3, 1.0), k_paf=E.make_pafs(xv3, yv3, s3, d3, 1.0),
               k_mpaf=E.make_multi_pafs(xv3, yv3, torch.stack([s3, s3]), torch.stack([d3, d3]), 1.0))
    # random edges incl. a degenerate (src == dst) edge, a sub-unit-length edge and a NaN edge
    edges = torch.tensor([[0, 1], [1, 2], [2, 3], [1, 4]])
    inst2 = inst.clone()[0]
    inst2[1, 2] = inst2[1, 1]  # degenerate edge (1,2) in instance 1
    inst2[0, 4] = inst2[0, 1] + torch.tensor([0.3, 0.4])  # |v| < 1 -> edge_length clamps to 1
    src, dst = E.get_edge_points(inst2, edges)
    out.update(pf_inst=inst2, pf_edges=edges, pf_src=src, pf_dst=dst,
               pf_single=E.make_pafs(xv, yv, src[0], dst[0], sigma),
               pf_degenerate=E.make_pafs(xv, yv, src[1], dst[1], sigma),
               pf_multi=E.make_multi_pafs(xv, yv, src, dst, sigma),
               pf_em=E.make_edge_maps(xv, yv, src[0], dst[0], sigma))
    inst3 = inst2.clone()[None]
    inst3[0, 1] = inst3[0, 1] + 500.0  # wholly outside the image -> filtered by generate_pafs
    out.update(gp_inst=inst3, gp=E.generate_pafs(inst3, hw, sigma, stride, edges, flatten_channels=True))
    save("ref_targets.npz", **out)


def f1_outputs(R):
    """group_scored_batch (inference/streaming.py:147-255) on the frames of ref_pipeline_tree.npz: NaN padding,
    top-N by score, input / effective scale undo, the skip_paf short-circuit."""
    d = np.load(os.path.join(HERE, "ref_pipeline_tree.npz"))
    cms, pafs = torch.from_numpy(d["cms"]), torch.from_numpy(d["pafs"])
    edges, n_nodes, stride = d["edges"].tolist(), int(d["n_nodes"]), int(d["stride"])
    B = cms.shape[0]
    pts, vals, si, ci = R.peaks.find_local_peaks(cms, threshold=0.2, refinement="integral")
    pts = pts * stride
    peaks, pvals, pch = (synth.split_by_sample(x, si, B) for x in (pts, vals, ci))
    names = [str(i) for i in range(n_nodes)]
    kwargs = dict(part_names=names, edges=[(str(a), str(b)) for a, b in edges], pafs_stride=stride,
                  max_edge_length_ratio=0.25, dist_penalty_weight=1.0, n_points=10,
                  min_instance_peaks=int(d["min_instance_peaks"]), min_line_scores=0.25)  # int = a count (float = fraction)
    scorer = R.paf.PAFScorer(**kwargs)
    ei, epi, ls = scorer.score_paf_lines(pafs.permute(0, 2, 3, 1), peaks, pch)
    out = {}
    cases = [
        ("dyn", None, 1.0, [1.0, 1.0, 1.0], False),
        ("top2", 2, 0.5, [1.0, 0.75, 1.25], False),
        ("pad6", 6, 1.0, [0.5, 0.5, 0.5], False),
        ("one", 1, 2.0, [1.0, 1.0, 1.0], False),
        ("skip3", 3, 1.0, [1.0, 1.0, 1.0], True),
        ("skipdyn", None, 1.0, [1.0, 1.0, 1.0], True),
    ]
    for tag, max_inst, scale, eff, skip in cases:
        info = R.preprocess_info.PreprocInfo(eff_scale=torch.tensor(eff), input_scale=scale)
        sb = R.streaming.ScoredBatch(cms_peaks=peaks, cms_peak_vals=pvals, cms_peak_channel_inds=pch,
                                     edge_inds=[] if skip else ei, edge_peak_inds=[] if skip else epi,
                                     line_scores=[] if skip else ls, info=info, n_samples=B, n_nodes=n_nodes,
                                     skip_paf=skip)
        with ref_loader.reference_imports():  # group_scored_batch imports PAFScorer lazily by its real module name
            res = R.streaming.group_scored_batch(sb, R.streaming.GroupingParams(paf_scorer_kwargs=kwargs, max_instances=max_inst))
        out.update({f"{tag}_kpts": res.pred_keypoints, f"{tag}_vals": res.pred_peak_values, f"{tag}_scores": res.instance_scores,
                    f"{tag}_max_instances": np.int64(-1 if max_inst is None else max_inst), f"{tag}_input_scale": np.float64(scale),
                    f"{tag}_eff": np.asarray(eff, np.float32), f"{tag}_skip": np.bool_(skip)})
    out["cases"] = np.asarray([c[0] for c in cases])
    # the smallest per-node peak count that trips / does not trip the max_peaks_per_node guard on these frames
    out["max_node_peaks"] = np.int64(max(int((c == k).sum()) for c in pch for k in range(n_nodes)))
    save("ref_f1_outputs.npz", **out)


def f2_layers(R):
    """CentroidLayer / CenteredInstanceLayer / TopDownLayer post-model arithmetic, composed from the reference's own
    ops in the order the layers call them (layers/centroid.py:196-258, layers/centered_instance.py:199-230,
    layers/topdown.py:259-289)."""
    Cd = R.coord
    out = {}
    g = torch.Generator().manual_seed(321)
    # ---- centroid: 3 frames with 5 / 2 / 0 blobs on a (64, 80) stride-2 map
    hw, stride = (128, 160), 2
    xv, yv = R.data_utils.make_grid_vectors(hw[0], hw[1], stride)
    pts = torch.full((3, 5, 1, 2), float("nan"))
    pts[0, :, 0] = torch.tensor([[20.0, 30.0], [60.5, 90.25], [120.0, 40.0], [140.0, 110.0], [75.0, 20.0]])
    pts[1, :2, 0] = torch.tensor([[33.0, 64.0], [100.0, 100.0]])
    amp = torch.tensor([0.9, 0.5, 0.7, 0.95, 0.6])
    cms = torch.zeros((3, 1, len(yv), len(xv)))
    for b in range(3):
        for i in range(5):
            cms[b] = torch.maximum(cms[b], amp[i] * R.confidence_maps.make_multi_confmaps(pts[b:b + 1, i:i + 1], xv, yv, 3.0)[0])
    cms = cms + torch.rand(cms.shape, generator=g) * 1e-3
    out["cen_cms"] = cms
    eff = torch.tensor([1.0, 0.8, 1.25])
    for tag, max_inst, scale in (("dyn", None, 1.0), ("top3", 3, 0.5), ("pad7", 7, 2.0)):
        peaks, vals, si, _ = R.peaks.find_local_peaks(cms, threshold=0.2, refinement="integral", integral_patch_size=5)
        peaks = Cd.undo_input_scale(Cd.undo_stride(peaks, stride), scale)
        B = cms.shape[0]
        mi = max_inst or int(torch.bincount(si.long(), minlength=B).max())
        padded = torch.full((B, mi, 2), float("nan"))
        pvals = torch.full((B, mi), float("nan"))
        for b in range(B):
            sp, sv = peaks[si == b], vals[si == b]
            if sp.numel() == 0:
                continue
            if sp.shape[0] > mi:
                sv, idx = torch.topk(sv, mi)
                sp = sp[idx]
            padded[b, : sp.shape[0]] = sp
            pvals[b, : sp.shape[0]] = sv
        padded = Cd.undo_eff_scale(padded, eff)
        out.update({f"cen_{tag}_xy": padded, f"cen_{tag}_val": pvals, f"cen_{tag}_max": np.int64(-1 if max_inst is None else max_inst),
                    f"cen_{tag}_scale": np.float64(scale)})
    out["cen_eff"] = eff
    out["cen_stride"] = np.int64(stride)
    # ---- centered instance: 6 crops, 4 nodes, (40, 40) stride-2 maps; one node below threshold
    xv2, yv2 = R.data_utils.make_grid_vectors(80, 80, 2)
    cpts = torch.rand((6, 4, 2), generator=g) * 60 + 10
    ccms = R.confidence_maps.make_confmaps(cpts, xv2, yv2, 3.0) + torch.rand((6, 4, 40, 40), generator=g) * 1e-3
    ccms[2, 1] *= 0.1
    out["ci_cms"] = ccms
    pk, pv = R.peaks.find_global_peaks(ccms, threshold=0.2, refinement="integral", integral_patch_size=5)
    eff6 = torch.tensor([1.0, 0.5, 1.0, 1.5, 0.75, 1.0])
    lad = Cd.undo_eff_scale(Cd.undo_input_scale(Cd.undo_stride(pk, 2), 0.5), eff6)
    out.update(ci_xy=lad.unsqueeze(1), ci_val=pv.unsqueeze(1), ci_eff=eff6, ci_scale=np.float64(0.5))
    # ---- top-down un-crop: the 6 crops belong to (b, i) slots of a (B=3, max_inst=3) layout
    valid_idx = torch.tensor([[0, 0], [0, 2], [1, 0], [1, 1], [2, 1], [2, 2]])
    cents = torch.rand((6, 2), generator=g) * 300 + 50
    bboxes = R.instance_cropping.make_centered_bboxes(cents, 80, 80)
    eff3 = torch.tensor([1.0, 0.8, 1.25])
    per_crop_eff = eff3[valid_idx[:, 0]]
    k3 = Cd.undo_input_scale(Cd.undo_stride(pk, 2), 1.0)  # the inner layer's own ladder (eff_scale all ones there)
    sized = Cd.add_crop_offset(k3, bboxes[:, 0, :])
    img = sized / per_crop_eff.view(-1, 1, 1)
    full = torch.full((3, 3, 4, 2), float("nan"))
    fullv = torch.full((3, 3, 4), float("nan"))
    full[valid_idx[:, 0], valid_idx[:, 1]] = img
    fullv[valid_idx[:, 0], valid_idx[:, 1]] = pv
    out.update(td_valid_idx=valid_idx, td_topleft=bboxes[:, 0, :], td_eff=eff3, td_xy=full, td_val=fullv)
    save("ref_f2_layers.npz", **out)


def f3_identity(R):
    """Multi-class identity grouping (inference/ops/identity.py) and class-map targets (data/identity.py)."""
    I, D = R.identity, R.data_identity
    g = torch.Generator().manual_seed(303)
    out = {}
    # --- classify_peaks_from_maps: S=3 samples, K=3 classes, C=4 channels, 40x48 maps
    S, K, Cn, H, W = 3, 3, 4, 40, 48
    class_maps = torch.softmax(torch.randn((S, K, H, W), generator=g) * 2.0, dim=1)
    pts, vals, si, ci = [], [], [], []
    for s in range(S):
        for c in range(Cn):
            n = [2, 3, 5, 1, 0, 4, 3, 2, 3, 1, 2, 6][s * Cn + c]  # fewer, equal and more peaks than classes
            for _ in range(n):
                pts.append([float(torch.rand((), generator=g)) * (W + 6) - 3, float(torch.rand((), generator=g)) * (H + 6) - 3])
                vals.append(float(torch.rand((), generator=g)))
                si.append(s); ci.append(c)
    pts = torch.tensor(pts, dtype=torch.float32)
    pts[0] = torch.tensor([10.5, 7.5]); pts[1] = torch.tensor([11.5, 8.5])  # round-half-even: 10, 8 / 12, 8
    pts[2] = torch.tensor([-0.5, 39.5]); pts[3] = torch.tensor([47.5, 0.5])
    vals = torch.tensor(vals, dtype=torch.float32)
    si = torch.tensor(si, dtype=torch.int32); ci = torch.tensor(ci, dtype=torch.int32)
    p, v, cp = I.classify_peaks_from_maps(class_maps, pts, vals, si, ci, Cn)
    out.update(cl_maps=class_maps, cl_pts=pts, cl_vals=vals, cl_si=si, cl_ci=ci, cl_out_pts=p, cl_out_vals=v, cl_out_probs=cp)
    # --- group_class_peaks on explicit probabilities, peaks NOT sorted by sample, int64 index tensors
    P = 23
    probs = torch.rand((P, 4), generator=g)
    probs[5] = probs[6]  # identical rows: the assignment decides
    si2 = torch.randint(0, 3, (P,), generator=g); ci2 = torch.randint(0, 2, (P,), generator=g)
    pi, cl = I.group_class_peaks(probs, si2, ci2, 3, 2)
    out.update(gr_probs=probs, gr_si=si2, gr_ci=ci2, gr_peak_inds=pi, gr_class_inds=cl)
    # --- get_class_inds_from_vectors: tall, wide, square
    for tag, shape in (("tall", (6, 3)), ("wide", (3, 5)), ("square", (4, 4)), ("big", (40, 7))):
        m = torch.softmax(torch.randn(shape, generator=g) * 1.5, dim=1)
        inds, pr = I.get_class_inds_from_vectors(m)
        out.update({f"vec_{tag}_probs": m, f"vec_{tag}_inds": inds, f"vec_{tag}_vals": pr})
    # --- class vectors / class maps (the reference's own known answers, tests/data/test_identity.py:13-33)
    out["cv_float"] = D.make_class_vectors(torch.Tensor([0, 2, 1, -1]), 3)
    out["cv_int"] = D.make_class_vectors(torch.tensor([3, -1, 0, 0, 1], dtype=torch.int32), 5)
    xv, yv = R.data_utils.make_grid_vectors(32, 32, output_stride=1)
    cms = R.confidence_maps.make_confmaps(torch.Tensor([[[4, 6], [18, 24]]]).to(torch.float32), xv, yv, sigma=2)
    out.update(cm_small_cms=cms, cm_small=D.make_class_maps(cms, class_inds=torch.Tensor([1, 0]), n_classes=2, threshold=0.2))
    # overlapping blobs, 3 instances / 4 classes with one unlabeled instance (the reshape quirk shows when I != K)
    xv, yv = R.data_utils.make_grid_vectors(64, 80, output_stride=2)
    ptsm = torch.tensor([[[20.0, 20.0], [30.0, 26.0], [60.0, 50.0]]])
    cms3 = R.confidence_maps.make_confmaps(ptsm, xv, yv, sigma=6.0)
    out.update(cm_multi_cms=cms3,
               cm_multi=D.make_class_maps(cms3, class_inds=torch.tensor([2, -1, 0], dtype=torch.int32), n_classes=4, threshold=0.1))
    inst = torch.rand((1, 4, 5, 2), generator=g) * torch.tensor([96.0, 64.0])
    inst[0, 1, 2] = float("nan")
    inst[0, 3] = float("nan")  # a padded (missing) instance
    out.update(gen_inst=inst,
               gen_nodes=D.generate_class_maps(inst, (64, 96), 3, torch.tensor([1, 0, 2], dtype=torch.int32), 3, output_stride=2),
               gen_centroids=D.generate_class_maps(inst[:, :, 0, :], (64, 96), 3, torch.tensor([1, 0, 2], dtype=torch.int32), 3,
                                                   class_map_threshold=0.3, sigma=2.0, output_stride=4, is_centroids=True))
    save("ref_f3_identity.npz", **out)


def f4_batched_targets(R):
    """Dataset call sites (data/custom_datasets.py:1305-1327, 1489-1511, 1788, 2835, 2986): the reference's
    per-frame generate_* calls on a collated batch of frames, stacked like the default collate does."""
    g = torch.Generator().manual_seed(404)
    B, I, Nn, H, W = 4, 4, 5, 64, 96
    edges = [[0, 1], [1, 2], [1, 3], [3, 4]]
    inst = torch.rand((B, 1, I, Nn, 2), generator=g) * torch.tensor([W + 10.0, H + 10.0]) - 5.0  # some nodes off-frame
    num = torch.tensor([4, 2, 3, 0])
    inst[1, 0, 2:] = float("nan")          # padded slots of frame 1
    inst[2, 0, 3] = float("nan")
    inst[2, 0, 0, 2] = float("nan")        # a missing node
    inst[0, 0, 3] = inst[0, 0, 3] * 0 + torch.tensor([-3.0, 200.0])  # an instance entirely outside the frame
    inst[3, 0] = float("nan")              # an empty frame
    tracks = torch.tensor([[0, 2, 1, 3], [1, 0, -1, -1], [2, -1, 0, -1], [-1, -1, -1, -1]], dtype=torch.int32)
    out = dict(instances=inst, num_instances=num, edges=torch.tensor(edges), tracks=tracks)
    CM, EM, ID = R.confidence_maps, R.edge_maps, R.data_identity
    cms, pafs, cen, single, cls, cls_cen = [], [], [], [], [], []
    for b in range(B):
        n = int(num[b])
        cms.append(CM.generate_multiconfmaps(inst[b], img_hw=(H, W), num_instances=n, sigma=1.5, output_stride=2))
        pafs.append(EM.generate_pafs(inst[b], img_hw=(H, W), sigma=4.0, output_stride=4, edge_inds=torch.Tensor(edges),
                                     flatten_channels=True))
        cen.append(CM.generate_multiconfmaps(inst[b][:, :, 0, :], img_hw=(H, W), num_instances=n, sigma=2.0, output_stride=2,
                                             is_centroids=True))
        one = R.providers.filter_oob_points(inst[b][:, 0], H, W)  # (1, N, 2): the first instance as a "crop" sample
        single.append(CM.generate_confmaps(one, img_hw=(H, W), sigma=1.5, output_stride=2))
        if n > 0:
            cls.append(ID.generate_class_maps(inst[b], (H, W), n, tracks[b, :n], 4, class_map_threshold=0.2, sigma=3.0,
                                              output_stride=2))
            cls_cen.append(ID.generate_class_maps(inst[b][:, :, 0, :], (H, W), n, tracks[b, :n], 4, class_map_threshold=0.1,
                                                  sigma=3.0, output_stride=4, is_centroids=True))
    out.update(confidence_maps=torch.stack(cms), part_affinity_fields=torch.stack(pafs), centroid_maps=torch.stack(cen),
               single_maps=torch.stack(single), class_maps=torch.stack(cls), class_maps_centroids=torch.stack(cls_cen))
    save("ref_f4_batched_targets.npz", **out)


FILTER_CASES = {
    "peak": dict(min_peak_value=0.3),
    "nodes": dict(min_visible_nodes=3, min_visible_node_fraction=0.5),
    "scores": dict(min_instance_score=0.4, min_mean_node_score=0.45),
    "iou": dict(overlapping=True, overlapping_threshold=0.5, overlapping_method="iou"),
    "oks": dict(overlapping=True, overlapping_threshold=0.3, overlapping_method="oks"),
    "all": dict(min_peak_value=0.25, min_visible_nodes=2, min_visible_node_fraction=0.3, min_instance_score=0.2,
                min_mean_node_score=0.3, overlapping=True, overlapping_threshold=0.4, overlapping_method="oks"),
}


def filter_inputs(seed: int):
    """(B, I, N) = (6, 8, 6) poses with near-duplicates, missing nodes, empty slots and a score spread."""
    g = torch.Generator().manual_seed(seed)
    B, I, Nn = 6, 8, 6
    kpts = torch.full((B, I, Nn, 2), float("nan"))
    for b in range(B):
        base = torch.rand((3, 1, 2), generator=g) * 300 + 50 + torch.cumsum(torch.rand((3, Nn, 2), generator=g) * 30 - 10, dim=1)
        for i in range(I - 1):  # slot I-1 stays empty
            jitter = [0.0, 0.0, 0.0, 1.5, 6.0, 14.0, 40.0][i]
            kpts[b, i] = base[i % 3] + (torch.rand((Nn, 2), generator=g) - 0.5) * 2 * jitter
    drop = torch.rand((B, I, Nn), generator=g) < 0.2
    kpts[drop] = float("nan")
    kpts[0, 1, :, 0] = float("nan")  # x missing everywhere, y present: "any coordinate" vs "both coordinates"
    vals = torch.rand((B, I, Nn), generator=g)
    vals[torch.isnan(kpts).any(-1)] = float("nan")
    scores = torch.rand((B, I), generator=g)
    scores[:, I - 1] = float("nan")
    scores[2, 3] = float("nan")  # a NaN score on a live instance sorts first
    return kpts, vals, scores


def f4_filters(R):
    """FilterPipeline (inference/filters.py) on synthetic Outputs; every decision is certified to have a margin."""
    from oracle import filters as ofil

    FP, FC, Out = R.filters.FilterPipeline, R.filters.FilterConfig, R.outputs.Outputs
    seed = 500
    while True:  # re-draw until no similarity / score sits within 1e-4 of a threshold it is compared with
        kpts, vals, scores = filter_inputs(seed)
        k = kpts.numpy()
        ok = True
        for b in range(k.shape[0]):
            for i in range(k.shape[1]):
                for j in range(k.shape[1]):
                    if i != j:
                        ok &= all(abs(ofil.bbox_iou(k[b, i], k[b, j]) - t) > 1e-4 for t in (0.5,))
                        ok &= all(abs(ofil.oks(k[b, i], k[b, j]) - t) > 1e-4 for t in (0.3, 0.4))
        ok &= bool((torch.nan_to_num(scores - 0.4, nan=1.0).abs() > 1e-4).all() and (torch.nan_to_num(scores - 0.2, nan=1.0).abs() > 1e-4).all())
        ok &= bool((torch.nan_to_num(vals - 0.3, nan=1.0).abs() > 1e-5).all() and (torch.nan_to_num(vals - 0.25, nan=1.0).abs() > 1e-5).all())
        m = torch.nan_to_num(torch.nanmean(vals, dim=-1), nan=0.0)
        ok &= bool(((m - 0.45).abs() > 1e-4).all() and ((m - 0.3).abs() > 1e-4).all())
        if ok:
            break
        seed += 1
    out = dict(kpts=kpts, vals=vals, scores=scores, seed=np.int64(seed))
    for tag, kw in FILTER_CASES.items():
        r = FP.run(Out(pred_keypoints=kpts, pred_peak_values=vals, instance_scores=scores), FC(**kw))
        out.update({f"{tag}_kpts": r.pred_keypoints, f"{tag}_vals": r.pred_peak_values, f"{tag}_scores": r.instance_scores})
    # no instance scores: the overlap NMS visits the slots in index order
    r = FP.run(Out(pred_keypoints=kpts, pred_peak_values=vals), FC(overlapping=True, overlapping_threshold=0.5))
    out.update(noscore_kpts=r.pred_keypoints, noscore_vals=r.pred_peak_values)
    # centroid-only outputs: score gate + distance NMS
    g = torch.Generator().manual_seed(77)
    cen = torch.rand((5, 9, 2), generator=g) * 200
    cen[:, 4] = cen[:, 1] + torch.tensor([3.0, -4.0])      # 5 px from slot 1
    cen[:, 6] = cen[:, 2] + torch.tensor([9.0, 12.0])      # 15 px from slot 2
    cen[1, 3] = float("nan"); cen[3, 0, 1] = float("nan")
    cenv = torch.rand((5, 9), generator=g)
    cenv[1, 3] = float("nan"); cenv[4, 7] = float("nan")
    out.update(cen=cen, cenv=cenv)
    r = FP.run(Out(pred_centroids=cen, pred_centroid_values=cenv), FC(min_instance_score=0.3, min_centroid_distance=12.0))
    out.update(cen_a=r.pred_centroids, cenv_a=r.pred_centroid_values)
    r = FP.run(Out(pred_centroids=cen, pred_centroid_values=cenv, instance_scores=1 - cenv), FC(min_centroid_distance=20.0))
    out.update(cen_b=r.pred_centroids, cenv_b=r.pred_centroid_values, cens_b=r.instance_scores)
    r = FP.run(Out(pred_centroids=cen, instance_scores=cenv), FC(min_instance_score=0.5, min_centroid_distance=6.0))
    out.update(cen_c=r.pred_centroids, cens_c=r.instance_scores)
    save("ref_f4_filters.npz", **out)


def f3_multiclass_layer(R):
    """BottomUpMultiClassLayer.postprocess (inference/layers/bottomup_multiclass.py:75-190) called on the unmodified
    reference class with a stand-in `self`."""
    import types

    L = R.bottomup_multiclass.BottomUpMultiClassLayer
    g = torch.Generator().manual_seed(606)
    B, Nn, K, H, W = 3, 4, 4, 96, 128
    edges = synth.chain_edges(Nn)
    poses = synth.make_poses(61, B, 3, Nn, (H, W), edges=edges, margin=24.0, step=14.0)
    xv, yv = R.data_utils.make_grid_vectors(H, W, 2)
    cms = torch.stack([R.confidence_maps.make_multi_confmaps(poses[b : b + 1], xv, yv, 3.0)[0] for b in range(B)])
    cms = cms + torch.rand(cms.shape, generator=g) * 1e-3
    # class maps at stride 4: each animal's neighbourhood votes for "its" class, plus noise; frame 2 has an animal no
    # class claims strongly
    logits = torch.randn((B, K, H // 4, W // 4), generator=g) * 0.3
    yy, xx = torch.meshgrid(torch.arange(H // 4) * 4.0, torch.arange(W // 4) * 4.0, indexing="ij")
    for b in range(B):
        for i in range(3):
            c = poses[b, i].mean(0)
            logits[b, (i + b) % K] += 4.0 * torch.exp(-((xx - c[0]) ** 2 + (yy - c[1]) ** 2) / (2 * 18.0**2))
    class_maps = torch.softmax(logits, dim=1)
    out = dict(cms=cms, class_maps=class_maps)
    for tag, scale, eff, cap in (("plain", 1.0, [1.0, 1.0, 1.0], None), ("scaled", 0.5, [1.0, 0.8, 1.25], None),
                                 ("cap2", 1.0, [1.0, 1.0, 1.0], 2), ("cap1", 0.5, [2.0, 1.0, 1.0], 1)):
        cfg = types.SimpleNamespace(peak_threshold=0.2, effective_refinement="integral", integral_patch_size=5,
                                    max_instances=None, return_confmaps=False, return_class_maps=False)
        me = types.SimpleNamespace(postprocess_config=cfg, cms_output_stride=2, class_maps_output_stride=4, max_instances=cap,
                                   _cap_instances_by_score=L._cap_instances_by_score)
        info = R.preprocess_info.PreprocInfo(eff_scale=torch.tensor(eff), input_scale=scale)
        o = L.postprocess(me, {"MultiInstanceConfmapsHead": cms, "ClassMapsHead": class_maps}, info)
        out.update({f"{tag}_kpts": o.pred_keypoints, f"{tag}_vals": o.pred_peak_values, f"{tag}_scores": o.instance_scores,
                    f"{tag}_tracking": o.instance_tracking_scores})
    save("ref_f3_multiclass_layer.npz", **out)


def labels_fixture(seed: int):
    """Plain-array description of a small Labels object: per frame a list of (kind, points (N,2) f64, score,
    point_scores (N,) or None); kind 1 = predicted, 0 = user instance."""
    g = np.random.default_rng(seed)
    frames = []
    for f in range(6):
        n_pred = [0, 1, 3, 5, 7, 4][f]
        base = g.uniform(40, 300, (3, 1, 2)) + np.cumsum(g.uniform(-12, 25, (3, 6, 2)), axis=1)
        insts = []
        for i in range(n_pred):
            jitter = [0.0, 1.0, 5.0, 12.0, 30.0, 2.5, 60.0][i]
            pts = base[i % 3] + g.uniform(-1, 1, (6, 2)) * jitter
            pts[g.uniform(size=6) < 0.2] = np.nan
            ps = g.uniform(0.05, 1.0, 6)
            ps[np.isnan(pts).any(1)] = np.nan
            insts.append((1, pts, float(g.uniform(0.05, 1.0)), ps if i % 4 != 3 else None))
        if f in (2, 4):
            insts.insert(1, (0, base[0] + 3.0, 1.0, None))  # a user instance between the predictions
        frames.append(insts)
    frames[3][1][1][:] = np.nan  # a prediction without any visible node
    return frames


def f4_labels_filters(R):
    """The Labels-level filters of inference/ops/filters.py on fake Labels (sleap_io is a stub in this image: the fake
    prediction class derives from the stub's PredictedInstance so the reference's isinstance checks pass)."""
    import types

    OF = R.ops_filters
    Pred = type("PredictedInstance", (OF.sio.PredictedInstance,), {})

    def build(frames):
        lfs = []
        for insts in frames:
            objs = []
            for uid, (kind, pts, score, ps) in enumerate(insts):
                cls = Pred if kind else types.SimpleNamespace
                o = cls()
                o.uid, o._pts, o.score = uid, pts, score
                o.numpy = (lambda p: (lambda: p))(pts)
                o.skeleton = types.SimpleNamespace(nodes=list(range(pts.shape[0])))
                o.points = {"score": ps} if ps is not None else {}
                objs.append(o)
            lfs.append(types.SimpleNamespace(instances=objs))
        return types.SimpleNamespace(labeled_frames=lfs)

    frames = labels_fixture(808)
    out = {}
    ragged("lab_n", [np.int64(len(f)) for f in frames], out)
    flat = [x for f in frames for x in f]
    out["lab_kind"] = np.array([x[0] for x in flat])
    out["lab_pts"] = np.stack([x[1] for x in flat])
    out["lab_score"] = np.array([x[2] for x in flat])
    out["lab_ps"] = np.stack([x[3] if x[3] is not None else np.full(6, -1.0) for x in flat])
    out["lab_has_ps"] = np.array([x[3] is not None for x in flat])
    cases = {
        "count": lambda L: OF.filter_by_node_count(L, min_visible_nodes=4, min_visible_node_fraction=0.6),
        "conf": lambda L: OF.filter_by_node_confidence(L, min_mean_node_score=0.5, min_instance_score=0.3),
        "iou": lambda L: OF.filter_overlapping_instances(L, threshold=0.45, method="iou"),
        "oks": lambda L: OF.filter_overlapping_instances(L, threshold=0.3, method="oks"),
    }
    for tag, fn in cases.items():
        L = fn(build(frames))
        ragged(f"lab_{tag}", [np.array([o.uid for o in lf.instances], dtype=np.int64) for lf in L.labeled_frames], out)
    # the numeric cores directly
    pts = [x[1] for x in frames[4] if x[0]]
    sc = np.array([x[2] for x in frames[4] if x[0]])
    out["core_iou_keep"] = np.array(OF._nms_greedy_iou(np.array([OF._instance_bbox(types.SimpleNamespace(numpy=(lambda p: (lambda: p))(p)))
                                                                 for p in pts]), sc, 0.45))
    out["core_oks_keep"] = np.array(OF._nms_greedy_oks(pts, sc, 0.3))
    out["core_oks_matrix"] = np.array([[OF._compute_oks(a, b) for b in pts] for a in pts])
    save("ref_f4_labels_filters.npz", **out)


def topdown_model_tables(seed: int, n_nodes: int, n_classes: int, crop_hw):
    """Fixed tables of the stand-in centred-instance "network": every op it applies is an exactly rounded elementwise
    fp32 op (or a max), so the CPU reference run and the CUDA test see bit-identical confidence maps."""
    g = torch.Generator().manual_seed(seed)
    gain = torch.rand((n_nodes, *crop_hw), generator=g) * 0.5 + 0.5
    pattern = torch.rand((n_nodes, *crop_hw), generator=g) * 1e-3
    cgain = torch.rand((n_classes, *crop_hw), generator=g) * 0.5 + 0.5
    return gain, pattern, cgain


from tests.helpers import topdown_model  # noqa: E402  (shared with the tests)


def f2_topdown(R):
    """TopDownLayer._centroid_nms_mask / _run_stage_2 (layers/topdown.py:186-466), CenteredInstanceLayer.postprocess,
    CenteredInstanceMultiClassLayer.postprocess (layers/topdown_multiclass.py:79-145) and SingleInstanceLayer.postprocess
    (layers/single_instance.py:71-106) called on the unmodified reference classes with stand-in `self` objects; the
    centred-instance "network" is `topdown_model`."""
    import types

    T, TM = R.topdown.TopDownLayer, R.topdown_multiclass.CenteredInstanceMultiClassLayer
    CI, SI = R.centered_instance.CenteredInstanceLayer, R.single_instance.SingleInstanceLayer
    P = R.preprocess_info.PreprocInfo
    g = torch.Generator().manual_seed(909)
    B, I, H, W, crop_hw, Nn, K = 4, 4, 96, 128, (24, 32), 3, 3
    cen = torch.full((B, I, 2), float("nan"))
    cen[0, :3] = torch.tensor([[30.2, 20.7], [90.5, 60.1], [60.0, 80.9]])
    cen[1] = torch.tensor([[40.0, 40.0], [44.5, 42.25], [100.3, 30.6], [15.1, 70.4]])  # 0 and 1 overlap heavily
    cen[2, 2] = torch.tensor([64.0, 48.0])
    cen_val = torch.where(torch.isnan(cen[..., 0]), torch.tensor(float("nan")), torch.rand((B, I), generator=g) * 0.7 + 0.3)
    cen_val[1, 0], cen_val[1, 1] = 0.61, 0.83  # the later slot wins the NMS
    eff = torch.tensor([1.0, 0.8, 1.25, 1.0])
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    img = torch.rand((B, 1, H, W), generator=g) * 40.0
    for b in range(B):
        for i in range(I):
            if torch.isnan(cen[b, i, 0]):
                continue
            for k in range(3):  # three blobs scattered around each (sized-space) centroid
                c = cen[b, i] * eff[b] + (torch.rand(2, generator=g) - 0.5) * 16.0
                img[b, 0] += 200.0 * torch.exp(-((xx - c[0]) ** 2 + (yy - c[1]) ** 2) / (2 * 2.5**2))
    img_u8 = img.clamp(0, 255).to(torch.uint8)
    gain, pattern, cgain = topdown_model_tables(910, Nn, K, crop_hw)
    out = dict(image=img_u8, centroids=cen, centroid_vals=cen_val, eff=eff, gain=gain, pattern=pattern, cgain=cgain,
               crop_hw=np.array(crop_hw))
    cfg = types.SimpleNamespace(peak_threshold=0.2, effective_refinement="integral", integral_patch_size=5,
                                return_confmaps=False, return_class_vectors=True)
    extract = lambda raw: raw["CenteredInstanceConfmapsHead"]

    def ci_predict(multiclass, stride, scale):
        def predict(crops):
            info = P(eff_scale=torch.ones(crops.shape[0]), input_scale=scale, output_stride=stride)
            me2 = types.SimpleNamespace(postprocess_config=cfg, _extract_confmaps=extract)
            if multiclass:
                cms, vec = topdown_model(crops, gain, pattern, cgain)
                return TM.postprocess(me2, {"CenteredInstanceConfmapsHead": cms, "ClassVectorsHead": vec}, info)
            return CI.postprocess(me2, {"CenteredInstanceConfmapsHead": topdown_model(crops, gain, pattern)}, info)
        return predict

    cases = (("plain", False, False, 0.5, 1, 1.0), ("nms", True, False, 0.3, 2, 0.5), ("mc", False, True, 0.5, 1, 1.0),
             ("mcnms", True, True, 0.3, 1, 1.0))
    for tag, nms, multiclass, thr, stride, scale in cases:
        inner = types.SimpleNamespace(predict=ci_predict(multiclass, stride, scale), postprocess_config=cfg)
        me = types.SimpleNamespace(crop_size=crop_hw, centered_instance_layer=inner, return_crops=True, centroid_nms=nms,
                                   centroid_nms_threshold=thr, _bbox_iou=T._bbox_iou, _infer_n_nodes=lambda: Nn)
        with ref_loader.reference_imports():
            valid = ~torch.isnan(cen).any(dim=-1)  # topdown.py:101
            if nms:
                valid = valid & T._centroid_nms_mask(me, cen, cen_val, valid)
            o = T._run_stage_2(me, img_u8, cen * eff.view(-1, 1, 1), cen_val, valid, eff_scale=eff)  # topdown.py:147-150
        out.update({f"{tag}_valid": valid, f"{tag}_kpts": o.pred_keypoints, f"{tag}_crop_kpts": o.pred_crop_keypoints,
                    f"{tag}_vals": o.pred_peak_values, f"{tag}_centroids": o.pred_centroids,
                    f"{tag}_scores": o.instance_scores, f"{tag}_bboxes": o.instance_bboxes, f"{tag}_crops": o.crops,
                    f"{tag}_knobs": np.array([float(nms), float(multiclass), thr, float(stride), scale])})
        if multiclass:
            out.update({f"{tag}_class_inds": o.pred_class_inds, f"{tag}_tracking": o.instance_tracking_scores,
                        f"{tag}_class_vectors": o.pred_class_vectors})
    # no valid centroid at all: the early return (topdown.py:218-233)
    me = types.SimpleNamespace(crop_size=crop_hw, centered_instance_layer=None, return_crops=False, _infer_n_nodes=lambda: Nn)
    nan_cen = torch.full((2, 3, 2), float("nan"))
    o = T._run_stage_2(me, img_u8[:2], nan_cen, torch.full((2, 3), float("nan")), torch.zeros((2, 3), dtype=torch.bool))
    out.update(empty_kpts=o.pred_keypoints, empty_vals=o.pred_peak_values)
    # stand-alone multi-class layer: ONE assignment over all crops; single-instance layer: the plain ladder
    crops6 = img_u8[:1, :, :24, :32].repeat(6, 1, 1, 1).clone()
    for r in range(6):
        crops6[r] = img_u8[r % B, :, 8 * r : 8 * r + 24, 10 * r : 10 * r + 32]
    cms6, vec6 = topdown_model(crops6, gain, pattern, cgain)
    info6 = P(eff_scale=torch.tensor([1.0, 0.5, 1.0, 1.5, 0.75, 1.0]), input_scale=0.5, output_stride=2)
    me2 = types.SimpleNamespace(postprocess_config=cfg, _extract_confmaps=extract)
    with ref_loader.reference_imports():
        o = TM.postprocess(me2, {"CenteredInstanceConfmapsHead": cms6, "ClassVectorsHead": vec6}, info6)
    out.update(sa_crops=crops6, sa_eff=info6.eff_scale, sa_kpts=o.pred_keypoints, sa_vals=o.pred_peak_values,
               sa_class_inds=o.pred_class_inds, sa_class_probs=o.pred_class_probs, sa_tracking=o.instance_tracking_scores)
    me3 = types.SimpleNamespace(postprocess_config=cfg, _extract_confmaps=lambda raw: raw["SingleInstanceConfmapsHead"])
    o = SI.postprocess(me3, {"SingleInstanceConfmapsHead": cms6}, info6)
    out.update(si_kpts=o.pred_keypoints, si_vals=o.pred_peak_values)
    save("ref_f2_topdown.npz", **out)


def main():
    R = ref_loader.ref()
    if len(sys.argv) > 1:  # regenerate only the named files' functions, e.g. `make_golden.py f2_topdown`
        torch.set_num_threads(1)
        for name in sys.argv[1:]:
            globals()[name](R)
        return
    torch.set_num_threads(1)  # reductions are then run-to-run reproducible
    peaks_minimal(R)
    peaks_random(R)
    paf_units(R)
    paf_pipeline(R, "ref_pipeline_mice.npz", seed=11, n_frames=3, n_inst=2, n_nodes=5, img_hw=(128, 160),
                 edges=synth.chain_edges(5), step=24.0)
    paf_pipeline(R, "ref_pipeline_tree.npz", seed=23, n_frames=3, n_inst=4, n_nodes=8, img_hw=(192, 192),
                 edges=synth.star_chain_edges(8, fan=2), step=26.0, min_instance_peaks=2)
    assembly_cases(R)
    targets(R)
    f1_outputs(R)
    f2_layers(R)
    f2_topdown(R)
    f3_identity(R)
    f3_multiclass_layer(R)
    f4_batched_targets(R)
    f4_filters(R)
    f4_labels_filters(R)


if __name__ == "__main__":
    main()
