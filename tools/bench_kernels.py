#!/usr/bin/env python
"""Per-kernel roofline lines for the rows of SURVEY.md section 8(d) that bench.py's headline does not cover.

    python tools/bench_kernels.py [--iters 200] [--only k1_cfg4,k2_cfg2,k7_cfg4,k8_cfg4,chain_cfg4]

One JSON line per kernel: algorithmic bytes per launch, average launch time over `iters`
launches (CUDA events on the launching stream, after warm-up, inputs/outputs rotating over buffers whose
total exceeds the 126 MB L2), achieved GB/s and its fraction of MEASURED_PEAKS.json's hbm_gbs.
Not part of the product; results are copied into profiles/ by hand.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from sleap_nn_b200 import _native as N  # noqa: E402
from sleap_nn_b200 import synthetic  # noqa: E402
from sleap_nn_b200.data.utils import make_grid_vectors  # noqa: E402


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


GRAPH = os.environ.get("SNB_BENCH_GRAPH", "1") != "0"


def timed(fn, iters, warm=10):
    """ms per launch of `iters` back-to-back launches on one stream, CUDA events around the whole run.

    The launches are replayed from a CUDA graph (captured once, replayed warm) so that short kernels are timed at the
    GPU's pace: issued one by one from this Python loop, a ctypes call + cudaLaunchKernelEx costs ~12 us of host time,
    more than a single-frame target kernel runs.  SNB_BENCH_GRAPH=0 times the eager loop instead; both are reported
    when they differ by more than 10 %."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()

    def run_eager():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(iters):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    eager = run_eager()
    timed.last = {"eager_ms": eager, "graph_ms": None}
    if not GRAPH:
        return eager
    try:
        st = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(st):
            with torch.cuda.graph(g, stream=st):
                for i in range(iters):
                    fn(i)
            g.replay()
            st.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            g.replay()
            b.record(st)
            st.synchronize()
        ms = a.elapsed_time(b) / iters
        timed.last["graph_ms"] = ms
        return min(ms, eager)
    except Exception as e:  # noqa: BLE001 - capture not possible for this launcher: keep the eager figure
        timed.last["graph_error"] = f"{type(e).__name__}: {e}"
        torch.cuda.synchronize()
        return eager


timed.last = {}


RESULTS = []  # every line() of this process, in order (bench.py's `extra` block reads it)


def line(name, kernel, algo_bytes, ms, units, unit_name, extra=None):
    peak, src = peak_gbs()
    ach = algo_bytes / (ms / 1e3) / 1e9
    d = {"bench": name, "kernel": kernel, "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": ms,
         "achieved_GBps": ach, "peak_GBps": peak, "peak_source": src, "frac": ach / peak,
         f"{unit_name}_per_s": units / (ms / 1e3)}
    if timed.last.get("graph_ms") is not None:
        d["launch_loop"] = ("CUDA graph replay" if timed.last["graph_ms"] <= timed.last["eager_ms"] else "eager Python loop")
        d["eager_loop_ms"], d["graph_replay_ms"] = timed.last["eager_ms"], timed.last["graph_ms"]
    if extra:
        d.update(extra)
    RESULTS.append(d)
    if not QUIET:
        print(json.dumps(d), flush=True)
    return d


QUIET = False


def k2_cfg2(dev, iters, B=256, dtype=torch.float32):
    """cfg2: find_global_peaks + integral refine on (B,13,80,80) centred-instance crops (B = 256 in BASELINE.json)."""
    Cn, H, W = 13, 80, 80
    g = torch.Generator(device=dev).manual_seed(0)
    bufs = []
    esz = 4 if dtype == torch.float32 else 2
    n_bufs = max(4, -(-400_000_000 // (esz * B * Cn * H * W)))  # rotate over > 3 x L2
    from sleap_nn_b200.data.confidence_maps import _confmaps
    for _ in range(n_bufs):
        pts = torch.rand((B, 1, Cn, 2), generator=g, device=dev) * 60 + 10
        xv, yv = make_grid_vectors(H, W, 1)
        cms = _confmaps(pts, xv, yv, 3.0, torch.float32, dev)
        cms += torch.rand(cms.shape, generator=g, device=dev) * 1e-3
        bufs.append(cms.to(dtype))
    rpc, nch, nbytes = C.c_int(), C.c_int(), C.c_longlong()
    N.check(N.lib.snb_global_peaks_workspace(B, Cn, H, W, C.byref(rpc), C.byref(nch), C.byref(nbytes)), "ws")
    ws = torch.zeros(((nbytes.value + 3) // 4,), dtype=torch.int32, device=dev)
    pts_o = torch.empty((B, Cn, 2), device=dev)
    val_o = torch.empty((B, Cn), device=dev)
    st = lambda: N.stream_ptr(dev)  # the CURRENT stream at launch time (the capture stream during graph capture)
    dt = N.dtype_code(dtype)

    def fn(i):
        x = bufs[i % len(bufs)]
        N.check(N.lib.snb_global_peaks_t(N.ptr(x), dt, B, Cn, H, W, *x.stride(), 0.2, 5, N.ptr(ws), None, N.ptr(pts_o),
                                         N.ptr(val_o), st()), "k2")

    ms = timed(fn, iters)
    tag = ("" if B == 256 else f"_b{B}") + ("" if dtype == torch.float32 else "_" + str(dtype).split(".")[-1])
    return line("k2_cfg2" + tag, "global_peaks_warp_kernel", esz * B * Cn * H * W, ms, B, "crops",
                {"shape": [B, Cn, H, W], "dtype": str(dtype), "valid_peaks": int((val_o > 0).sum())})


def k1_cfg3(dev, iters, dtype=torch.float32):
    """cfg3 detect kernel alone on (64,5,512,512) maps of the given element type (rotating batches > L2)."""
    Bn, Cn, H, W = 64, 5, 512, 512
    g = torch.Generator(device=dev).manual_seed(1)
    esz = 4 if dtype == torch.float32 else 2
    bufs = [(torch.rand((Bn, Cn, H, W), generator=g, device=dev) * 0.15).to(dtype) for _ in range(4 if esz == 4 else 6)]
    for x in bufs:  # a few hundred blobs per batch so that the rare path is exercised as in real maps
        ys = torch.randint(2, H - 2, (Bn, Cn, 2), generator=g, device=dev)
        xs = torch.randint(2, W - 2, (Bn, Cn, 2), generator=g, device=dev)
        bi = torch.arange(Bn, device=dev)[:, None, None].expand_as(ys)
        ci = torch.arange(Cn, device=dev)[None, :, None].expand_as(ys)
        x[bi, ci, ys, xs] = 0.9
    cap = 256
    count = torch.empty((Bn,), dtype=torch.int32, device=dev)
    keys = torch.empty((Bn * cap,), dtype=torch.int32, device=dev)
    st, dt = (lambda: N.stream_ptr(dev)), N.dtype_code(dtype)

    def fn(i):
        x = bufs[i % len(bufs)]
        N.check(N.lib.snb_local_peaks_detect_t(N.ptr(x), dt, Bn, Cn, H, W, *x.stride(), 0.2, cap, N.ptr(count), N.ptr(keys),
                                               None, None, st()), "k1")

    ms = timed(fn, iters)
    return line("k1_cfg3_" + str(dtype).split(".")[-1], "local_peaks_detect_vec", esz * Bn * Cn * H * W, ms, Bn, "frames",
                {"shape": [Bn, Cn, H, W], "dtype": str(dtype), "note": "includes the 256-byte counter memset node"})


def _flies_poses(n_frames, seed=0):
    edges = synthetic.chain_edges(32)
    return edges, synthetic.random_poses(seed, n_frames, 8, 32, (1024, 1024), edges, margin=200.0, step=24.0,
                                         min_limb=8.0, min_sep=10.0)


def k7_cfg4(dev, iters, bf16=False, G=8):
    """cfg4 targets: make_multi_confmaps, 32 nodes, 8 instances, sigma 2.5 x stride 2, out (G,32,512,512) per launch
    (G = 1 is the shape BottomUpDataset.__getitem__ issues, data/custom_datasets.py:1305-1327)."""
    edges, poses = _flies_poses(G)
    xv, yv = make_grid_vectors(1024, 1024, 2)
    xd, yd, pts = xv.to(dev), yv.to(dev), poses.to(dev).contiguous()
    dt = torch.bfloat16 if bf16 else torch.float32
    n_out = max(3, -(-400_000_000 // (G * 32 * 512 * 512 * (2 if bf16 else 4))))  # rotate over > 3 x L2
    outs = [torch.empty((G, 32, 512, 512), dtype=dt, device=dev) for _ in range(n_out)]
    st = lambda: N.stream_ptr(dev)  # the CURRENT stream at launch time (the capture stream during graph capture)
    den = float(2 * (2.5 * 2) ** 2)

    def fn(i):
        N.check(N.lib.snb_confmaps(N.ptr(pts), G, 8, 32, N.ptr(xd), N.ptr(yd), 512, 512, den, int(bf16),
                                   N.ptr(outs[i % n_out]), st()), "k7")

    ms = timed(fn, iters)
    esz = 2 if bf16 else 4
    return line("k7_cfg4" + ("" if G == 8 else f"_g{G}") + ("_bf16" if bf16 else ""), "confmaps_rows2_kernel", esz * G * 32 * 512 * 512, ms, G, "frames",
         {"frames_per_launch": G, "out_dtype": str(dt)})


def k8_cfg4(dev, iters, bf16=False, G=1):
    """cfg4 targets: make_multi_pafs, 31 edges, 8 instances, sigma 2.5, out (G,31,2,512,512) per launch."""
    n_sets = max(2, -(-400_000_000 // (G * 62 * 512 * 512 * (2 if bf16 else 4))))  # rotate over > 3 x L2
    edges, poses = _flies_poses(G * n_sets)
    e = torch.tensor(edges, dtype=torch.int64)
    xv, yv = make_grid_vectors(1024, 1024, 2)
    xd, yd = xv.to(dev), yv.to(dev)
    pd = poses.to(dev).reshape(n_sets, G, 8, 32, 2)
    srcs = [pd[f][:, :, e[:, 0]].contiguous() for f in range(n_sets)]
    dsts = [pd[f][:, :, e[:, 1]].contiguous() for f in range(n_sets)]
    dt = torch.bfloat16 if bf16 else torch.float32
    outs = [torch.empty((G, 31, 2, 512, 512), dtype=dt, device=dev) for _ in range(n_sets)]
    st = lambda: N.stream_ptr(dev)  # the CURRENT stream at launch time (the capture stream during graph capture)
    den = float(2 * 2.5 ** 2)

    def fn(i):
        k = i % n_sets
        N.check(N.lib.snb_pafs(N.ptr(srcs[k]), N.ptr(dsts[k]), G, 8, 31, N.ptr(xd), N.ptr(yd), 512, 512, den, 1,
                               int(bf16), N.ptr(outs[k]), st()), "k8")

    ms = timed(fn, iters)
    esz = 2 if bf16 else 4
    return line(f"k8_cfg4_g{G}" + ("_bf16" if bf16 else ""), "pafs_rows_kernel", esz * G * 31 * 2 * 512 * 512, ms, G, "frames",
         {"frames_per_launch": G, "out_dtype": str(dt)})


def targets_cfg4_fused(dev, iters, bf16=False, G=1):
    """cfg4 targets of G frames (default ONE, the granularity of BottomUpDataset.__getitem__) in one launch:
    confidence maps (G,32,512,512) + PAFs (G,62,512,512) through snb_bottomup_targets."""
    from sleap_nn_b200.data.batched_targets import BatchedTargets
    esz = 2 if bf16 else 4
    per_launch = esz * G * (32 + 62) * 512 * 512
    n_sets = max(3, -(-400_000_000 // per_launch))
    edges, poses = _flies_poses(G * n_sets)
    pd = poses.to(dev).reshape(n_sets, G, 8, 32, 2).contiguous()
    bt = BatchedTargets((1024, 1024), device=dev, out_dtype=torch.bfloat16 if bf16 else torch.float32)
    e = torch.tensor(edges, dtype=torch.int32, device=dev)
    xv, yv = bt._grid(2)
    dt = torch.bfloat16 if bf16 else torch.float32
    cms = [torch.empty((G, 32, 512, 512), dtype=dt, device=dev) for _ in range(n_sets)]
    pafs = [torch.empty((G, 31, 2, 512, 512), dtype=dt, device=dev) for _ in range(n_sets)]
    st = lambda: N.stream_ptr(dev)  # the CURRENT stream at launch time (the capture stream during graph capture)
    den7, den8 = float(2 * (2.5 * 2) ** 2), float(2 * 2.5 ** 2)

    def fn(i):
        k = i % n_sets
        N.check(N.lib.snb_bottomup_targets(N.ptr(pd[k]), G, 8, 32, None, 0.0, 0.0, N.ptr(e), 31, 1022.0, 1022.0, N.ptr(xv),
                                           N.ptr(yv), 512, 512, den7, N.ptr(xv), N.ptr(yv), 512, 512, den8, int(bf16),
                                           N.ptr(cms[k]), N.ptr(pafs[k]), st()), "fused targets")

    ms = timed(fn, iters)
    return line(f"targets_cfg4_fused_g{G}" + ("_bf16" if bf16 else ""), "confmaps_rows2 / confmaps_sep + pafs_rows (PDL pair)", per_launch, ms, G, "frames",
                {"frames_per_launch": G, "out_dtype": str(dt)})


def memset_ref(dev, iters):
    """Reference point for the store-bound kernels: cudaMemsetAsync (torch zero_) of 268 MB, rotating buffers."""
    bufs = [torch.empty((268435456 // 4,), dtype=torch.float32, device=dev) for _ in range(3)]
    ms = timed(lambda i: bufs[i % 3].zero_(), iters)
    return line("memset_268MB", "cudaMemsetAsync (reference, not ours)", 268435456, ms, 1, "launches")


def chain_cfg4(dev, iters, Bn=64, n_streams=None):
    """cfg4 inference: 32 nodes / 31 edges / 8 instances per frame (256 peaks, ~2 k candidates per frame), full chain.

    Batch 64 on two streams: the per-frame tail (one CTA per frame, ~230 us for such busy frames) needs at least as
    many frames in flight as it takes to hide it under the next batch's detect pass (2.1 GB of maps, ~340 us); at
    batch 8 the tail would run on 8 of the 148 SMs and dominate.
    """
    from sleap_nn_b200.pipeline import BottomUpPostproc
    n_streams = n_streams or int(os.environ.get("SNB_CHAIN_STREAMS", "4"))
    Bn = int(os.environ.get("SNB_CHAIN_BATCH", Bn))
    inputs = []
    for s in range(2):  # 2 x 64 x (33.5 + 65.0 MB) = 12.6 GB
        edges, poses = _flies_poses(Bn, seed=s)
        inputs.append(synthetic.render_batch(poses, (1024, 1024), 2, edges, dev, seed=s))
    tail = torch.cuda.Stream(device=dev, priority=-1)  # shared high-priority tail stream (see pipeline.PipelineRing)
    pipes = [BottomUpPostproc(32, edges, Bn, (512, 512), cms_stride=2, pafs_stride=2, device=dev, peak_cap=512,
                              cand_cap=4096, match_cap=512, inst_cap=32, keep_tables=False, tail_stream=tail)
             for _ in range(n_streams)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    res = pipes[0](*inputs[0])
    inst, _, _ = res.to_lists()
    n_inst = sum(len(x) for x in inst)
    n_peaks = int(res.n_peaks.sum())
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    main = torch.cuda.current_stream(dev)
    for a, b in evs:
        a.record(main); b.record(main)
    torch.cuda.synchronize()

    def step(i):
        with torch.cuda.stream(streams[i % n_streams]):
            pipes[i % n_streams](*inputs[i % 2], detect_events=evs[i % iters])

    for s_ in streams:
        s_.wait_stream(main)
    for i in range(3 * n_streams):
        step(i)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(main)
    for s_ in streams + [tail]:
        s_.wait_stream(main)
    for i in range(iters):
        step(i)
    for s_ in streams + [tail]:
        main.wait_stream(s_)
    t1.record(main)
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / iters
    det = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    algo = 4 * Bn * 32 * 512 * 512
    line("chain_cfg4", "detect+tail (whole chain, batch %d, %d streams)" % (Bn, n_streams), algo, ms, Bn, "frames",
         {"frames_per_launch": Bn, "instances_found": n_inst, "instances_planted": Bn * 8, "peaks": n_peaks,
          "launches_per_call": pipes[0].launches_per_call, "fused_tail": pipes[0].fused})
    line("k1_cfg4", "local_peaks_detect_vec4 (in situ)", algo, det, Bn, "frames", {"frames_per_launch": Bn})


ALL = {"targets_cfg4_fused": targets_cfg4_fused, "targets_cfg4_fused_bf16": lambda d, i: targets_cfg4_fused(d, i, True),
       "targets_cfg4_fused_g8": lambda d, i: targets_cfg4_fused(d, i, False, 8),
       "k7_cfg4_g1": lambda d, i: k7_cfg4(d, i, False, 1), "k7_cfg4_g1_bf16": lambda d, i: k7_cfg4(d, i, True, 1),
       "k2_cfg2": k2_cfg2, "k2_cfg2_b1024": lambda d, i: k2_cfg2(d, i, 1024),
       "k2_cfg2_f16": lambda d, i: k2_cfg2(d, i, 256, torch.float16),
       "k1_cfg3_f32": k1_cfg3, "k1_cfg3_f16": lambda d, i: k1_cfg3(d, i, torch.float16),
       "k1_cfg3_bf16": lambda d, i: k1_cfg3(d, i, torch.bfloat16), "k7_cfg4": k7_cfg4, "k7_cfg4_bf16": lambda d, i: k7_cfg4(d, i, True),
       "k8_cfg4": k8_cfg4, "k8_cfg4_g8": lambda d, i: k8_cfg4(d, i, False, 8),
       "k8_cfg4_bf16": lambda d, i: k8_cfg4(d, i, True), "k8_cfg4_g8_bf16": lambda d, i: k8_cfg4(d, i, True, 8),
       "memset_ref": memset_ref, "chain_cfg4": chain_cfg4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--only", default=",".join(ALL))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    for name in args.only.split(","):
        try:
            ALL[name](dev, args.iters)
        except Exception as e:  # keep going: one failing config must not hide the others
            print(json.dumps({"bench": name, "error": f"{type(e).__name__}: {e}"}), flush=True)


if __name__ == "__main__":
    main()
