#!/usr/bin/env python
"""Per-phase cycle counts of the fused per-frame tail kernel (profiling build, not the product).

    SNB_LIB_NAME=libsleapnn_b200_timing.so SNB_NVCC_EXTRA=-DSNB_TAIL_TIMING bash sleap_nn_b200/csrc/build.sh
    SLEAPNN_B200_LIB=sleap_nn_b200/lib/libsleapnn_b200_timing.so python tools/tail_phases.py [cfg3|cfg4]

Phases: sort | refine | group | score | match | assemble (clock64 deltas of thread 0, median over frames).
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sleap_nn_b200 import synthetic
from sleap_nn_b200.pipeline import BottomUpPostproc

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
dev = torch.device("cuda", 0)
if cfg == "cfg3":
    Bn, Nn, n_inst, kw = 64, 5, 2, {}
    pose_kw = {}
else:
    Bn, Nn, n_inst, kw = 8, 32, 8, dict(peak_cap=512, cand_cap=4096, match_cap=512, inst_cap=32)
    pose_kw = dict(margin=200.0, step=24.0, min_limb=8.0, min_sep=10.0)
edges = synthetic.chain_edges(Nn)
poses = synthetic.random_poses(0, Bn, n_inst, Nn, (1024, 1024), edges, **pose_kw)
cms, pafs = synthetic.render_batch(poses, (1024, 1024), 2, edges, dev, seed=0)
pipe = BottomUpPostproc(Nn, edges, Bn, (512, 512), device=dev, keep_tables=False, **kw)
for _ in range(3):
    res = pipe(cms, pafs)
torch.cuda.synchronize()
raw = pipe.buf["asm_ws"][: Bn * 16].reshape(Bn, 16).cpu().double()
t = raw[:, :6]
d = torch.cat([t[:, :1], t[:, 1:] - t[:, :-1]], dim=1)
names = ["sort", "refine", "group", "score", "match", "assemble"]
sub = raw[:, [4, 6, 7, 8, 9, 5]]  # match end | greedy loop | compaction | NaN fill + rank pass | score sums | scatter (= assemble end)
sd = sub[:, 1:] - sub[:, :-1]
for k, nm in enumerate(["asm: greedy loop", "asm: count+compact", "asm: fill+rank pass", "asm: score sums", "asm: scatter"]):
    print(f"{nm:20s} {float(sd[:, k].median()):10.0f} cycles  ~{float(sd[:, k].median()) / 1900.0:8.1f} us at 1.9 GHz")
ns = float(raw[:, 10].median())  # globaltimer span of the same region (ns): the SM clock is not fixed
mhz = float(t[:, 5].median()) / (ns / 1e3) if ns > 0 else 1900.0
print(f"wall {ns / 1e3:.1f} us for {float(t[:, 5].median()):.0f} cycles -> SM clock {mhz:.0f} MHz during the tail")
print(cfg, "peaks/frame", float(res.n_peaks.float().mean()), "instances/frame", float(res.n_instances.float().mean()))
for k, nm in enumerate(names):
    c = float(d[:, k].median())
    print(f"{nm:9s} {c:10.0f} cycles  ~{c / mhz:8.1f} us")
print(f"total     {float(t[:, 5].median()):10.0f} cycles  ~{float(t[:, 5].median()) / mhz:8.1f} us")
