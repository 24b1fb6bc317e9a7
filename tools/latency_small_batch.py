#!/usr/bin/env python
"""Single-call latency of the bottom-up post-processing chain for SMALL batches (what a streaming predictor issues):
cfg3 (5 nodes, 2 animals) at batch 1 and 8, cfg4 (32 nodes, 8 animals) at batch 8 - with the 4-CTA cluster per frame
(the default up to 16 frames) and with one CTA per frame, next to the reference's own CPU chain on the same frames
(the unmodified files staged under baseline/_ref, else the oracle port).

    python tools/latency_small_batch.py            # one JSON line per configuration

Latency = CUDA events around ONE call (detect + tail) on an idle GPU, median of 200 calls, inputs rotating over four
batches.  `tail_us` = the tail kernel alone (events around it are not available from outside, so it is the chain time
minus the detect kernel's time measured by the C ABI's own events).
"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sleap_nn_b200 import synthetic  # noqa: E402
from sleap_nn_b200.pipeline import BottomUpPostproc  # noqa: E402


def one(dev, name, Bn, Nn, n_inst, pipe_kw, pose_kw, cpu=True):
    edges = synthetic.chain_edges(Nn)
    inputs = []
    for s in range(4):
        poses = synthetic.random_poses(s, Bn, n_inst, Nn, (1024, 1024), edges, **pose_kw)
        inputs.append(synthetic.render_batch(poses, (1024, 1024), 2, edges, dev, seed=s))
    out = {"config": name, "batch": Bn, "n_nodes": Nn, "instances_per_frame": n_inst}
    for label, cluster in (("cluster4", True), ("one_cta_per_frame", False)):
        pipe = BottomUpPostproc(Nn, edges, Bn, (512, 512), device=dev, keep_tables=False, tail_cluster=cluster, **pipe_kw)
        res = pipe(*inputs[0])
        n_found = sum(len(x) for x in res.to_lists()[0])
        chain, det = [], []
        for i in range(int(os.environ.get("SNB_LATENCY_CALLS", 210))):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for e in (d0, d1):
                e.record()
            torch.cuda.synchronize()
            e0.record()
            pipe(*inputs[i % 4], detect_events=(d0, d1))
            e1.record()
            torch.cuda.synchronize()
            if i >= 10:
                chain.append(e0.elapsed_time(e1) * 1e3)
                det.append(d0.elapsed_time(d1) * 1e3)
        out[label] = {"chain_us": statistics.median(chain), "detect_us": statistics.median(det),
                      "tail_us": statistics.median(chain) - statistics.median(det), "instances_found": n_found}
    if cpu and not os.environ.get("SNB_LATENCY_NO_CPU"):
        import bench

        bench.N_NODES, bench.STRIDE = Nn, 2
        fn, kind, what = bench.cpu_arm(edges)
        c, p = inputs[0][0].cpu(), inputs[0][1].cpu()
        torch.set_num_threads(os.cpu_count() or 1)
        fn(c, p)
        t0 = time.perf_counter()
        for _ in range(3):
            fn(c, p)
        out["cpu_reference"] = {"ms_per_call": (time.perf_counter() - t0) / 3 * 1e3, "kind": kind, "threads": torch.get_num_threads(),
                                "what": what}
    print(json.dumps(out), flush=True)


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    flies = dict(margin=200.0, step=24.0, min_limb=8.0, min_sep=10.0)
    big = dict(peak_cap=512, cand_cap=4096, match_cap=512, inst_cap=32)
    one(dev, "cfg3 batch 1", 1, 5, 2, {}, {})
    one(dev, "cfg3 batch 8", 8, 5, 2, {}, {})
    one(dev, "cfg4 batch 8", 8, 32, 8, big, flies)
    one(dev, "cfg4 batch 1", 1, 32, 8, big, flies)


if __name__ == "__main__":
    main()
