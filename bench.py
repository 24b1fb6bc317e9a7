#!/usr/bin/env python
"""Benchmark of the bottom-up post-processing hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype f32|f16|bf16]

Workload (`config.workload`): BASELINE cfg3 "bottom-up mice" - 1024x1024 frames, 5 nodes /
4 edges, stride-2 confidence maps (64,5,512,512) + PAFs (64,8,512,512), batch 64, full peak +
PAF grouping.  One STEP = one pass of the whole hot path (K1 peaks + refinement -> K4 line
scores -> K5 assignment -> K6 assembly) over one batch of 64 synthetic frames.

  value     frames/s with the maps already resident in HBM (as they are after the backbone),
            batches pipelined over `--streams` CUDA streams, inputs rotated over several distinct
            batches (each 872 MB > 126 MB L2, so nothing is served from cache);
  e2e       the same metric through the public API with HOST buffers: per step a pinned-host ->
            device copy of the maps and a device -> host read of the grouped instances;
  roofline  the dominant kernel (streaming NMS detect) alone: average launch duration from two CUDA
            events around 120 back-to-back launches (the launch-by-launch event-pair figure is
            reported next to it); algorithmic bytes = esz*C*H*W*B per launch; `traffic` from the
            committed ncu capture (profiles/traffic.json, keyed by kernel + hash of the source);
  cpu_baseline  the reference's own unmodified files (staged under baseline/_ref by build(); kind
            "reference") - or, where they are absent, the oracle port - on a bounded sample of the
            same frames, all host threads and one thread, per stage.  N=1, rank 0 only;
  parity    the GPU results of that same sample compared with the oracle's, in this run;
  extra     (N=1) the other kernels / configs of BASELINE.json, each with algorithmic bytes,
            average launch time and roofline fraction: K2 cfg2, the cfg4 chain, K7 / K8 cfg4
            (8 frames and 1 frame per launch, fp32 and bf16), K1a on f16 maps;
  gather    (N>1) the same steps through ShardRunner + the end-of-shard gather_packed collective.

`--impl reference` times the reference's CPU implementation alone, on the same config, all 64
frames per step.  Multi-GPU: frames shard, no collective on the hot path; one process per GPU
(torchrun), barrier + synchronize around the timed region, max over ranks.
"""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = dict(
    workload="cfg3 bottom-up mice: 1024x1024 frames, 5 nodes/4 edges, stride-2 cms (64,5,512,512) + pafs (64,8,512,512), "
             "batch 64, peaks+integral refine+PAF score+match+group",
    batch=64, n_nodes=5, n_edges=4, n_instances=2, img_hw=[1024, 1024], stride=2, maps_hw=[512, 512],
)
B, N_NODES, N_INST, IMG_HW, STRIDE = 64, 5, 2, (1024, 1024), 2
METRIC, UNIT = "bottom-up post-proc frames/s", "frames/s"
DTYPES = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}
emit = lambda obj: print(json.dumps(obj), flush=True)  # replaced in main() by a writer on the saved stdout descriptor


def shared_config(world: int, dtype: str = "f32"):
    """The `config` object BOTH arms print (the driver compares them for `same_config`)."""
    return dict(WORKLOAD, parallelism=f"frame-sharded x{world}, no collective on the hot path", dtype=dtype,
                l2="inputs larger than L2: each batch is 872 MB (fp32) and batches rotate")


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"  # B200_PROFILING.md fallback


def recorded_traffic(kernel: str):
    """dram bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json), with whether the
    capture was taken on the current source of the kernel (sha256 of csrc/peaks.cu)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            rec = json.load(f)[kernel]
        with open(os.path.join(ROOT, "sleap_nn_b200", "csrc", "peaks.cu"), "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()[:16]
        return rec["dram_bytes_read"] + rec["dram_bytes_write"], dict(rec, source_matches=(rec.get("source_sha16") == sha))
    except Exception:
        return None, None


class ClockSampler:
    """SM clocks / throttle reasons of the job's GPUs sampled while the timed region runs.

    In-process NVML (nvidia-ml-py) when available - a query costs microseconds and takes no driver-wide lock;
    falls back to spawning `nvidia-smi` (the recipe's clocks line).  Only rank 0 samples, for all GPUs of the job:
    eight ranks each forking nvidia-smi five times a second measurably slowed the other ranks' kernel launches.
    """

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, indices, enabled: bool = True):
        self.indices, self.enabled = list(indices), enabled
        self.sm, self.sm_max, self.reasons, self.n = [], None, set(), 0
        self._stop, self._t, self._nvml = threading.Event(), None, None
        if enabled:
            try:
                import pynvml

                pynvml.nvmlInit()
                self._nvml = pynvml
                self._handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in self.indices]
            except Exception:
                self._nvml = None

    def _sample_nvml(self):
        nv = self._nvml
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        for h in self._handles:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            try:
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            self.reasons |= {name for bit, name in bits.items() if mask & bit}
        self.n += 1

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                              ",".join(str(i) for i in self.indices)], capture_output=True, text=True, timeout=5).stdout
        for line in out.strip().splitlines():
            r = [x.strip() for x in line.split(",")]
            if r and r[0].replace(".", "").isdigit():
                self.sm.append(float(r[0]))
                self.sm_max = float(r[1])
                self.reasons |= {n for n, v in zip(self.NAMES, r[2:6]) if v.lower().startswith("active")}
        self.n += 1

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample_nvml() if self._nvml is not None else self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.02 if self._nvml is not None else 0.3)

    def __enter__(self):
        if self.enabled:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": self.n, "gpus": self.indices, "source": "nvml" if self._nvml is not None else "nvidia-smi"}


# --------------------------------------------------------------------------- the CPU arms (test infrastructure)
def oracle_postproc(cms_cpu, pafs_cpu, edges):
    """The reference's bottom-up post-processing chain, as restated by oracle/ (CPU, torch ops)."""
    from oracle import paf as opaf
    from oracle import peaks as opeaks
    from oracle.synth import split_by_sample

    nb = cms_cpu.shape[0]
    pts, vals, si, ci = opeaks.local_peaks(cms_cpu, 0.2, "integral")
    peaks, pvs, pcs = (split_by_sample(x, si, nb) for x in (pts * STRIDE, vals, ci))
    return opaf.predict(pafs_cpu.permute(0, 2, 3, 1), peaks, pvs, pcs, edges, N_NODES, STRIDE)


class ReferenceChain:
    """The UNMODIFIED reference files, exec'd in place by oracle/ref_loader.py from baseline/_ref (staged by build(),
    git-ignored, travels with the snapshot) or, in the build container, from /root/reference.

    One call = what BottomUpLayer runs after the backbone on one batch (layers/bottomup.py:95-236): find_local_peaks
    (ops/peaks.py:221-259) -> x stride -> per-sample split -> PAFScorer.predict (ops/paf.py:1456-1532).
    """

    def __init__(self, edges):
        self.root = None
        for root in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
            if os.path.isfile(os.path.join(root, "sleap_nn", "inference", "ops", "paf.py")):
                self.root = root
                break
        if self.root is None:
            raise FileNotFoundError("no reference files (baseline/_ref not staged and /root/reference absent)")
        os.environ["SLEAPNN_REFERENCE_ROOT"] = self.root
        from oracle import ref_loader

        self.R = ref_loader.ref()
        names = [str(i) for i in range(N_NODES)]
        self.scorer = self.R.paf.PAFScorer(part_names=names, edges=[(str(a), str(b)) for a, b in edges],
                                           pafs_stride=STRIDE)
        self.stage_s = {}

    def __call__(self, cms, pafs, stages: bool = False):
        R, sc = self.R, self.scorer
        t = [time.perf_counter()]
        pts, vals, si, ci = R.peaks.find_local_peaks(cms, threshold=0.2, refinement="integral", integral_patch_size=5)
        pts = pts * STRIDE
        peaks, pvals, pch = [], [], []
        for b in range(cms.shape[0]):
            mask = si == b
            peaks.append(pts[mask]); pvals.append(vals[mask].to(torch.float32)); pch.append(ci[mask])
        t.append(time.perf_counter())
        view = pafs.permute(0, 2, 3, 1)
        if not stages:
            return sc.predict(view, peaks, pvals, pch)
        e, ep, ls = sc.score_paf_lines(view, peaks, pch)
        t.append(time.perf_counter())
        m = sc.match_candidates(e, ep, ls)
        t.append(time.perf_counter())
        out = sc.group_instances(peaks, pvals, pch, *m)
        t.append(time.perf_counter())
        for name, a, b_ in zip(("find_local_peaks+split", "score_paf_lines", "match_candidates", "group_instances"),
                               t[:-1], t[1:]):
            self.stage_s[name] = b_ - a
        return tuple(out) + (e, ep, ls)


class PortChain:
    """Fallback CPU arm when the reference files are nowhere to be found: oracle/, the port of the same chain."""

    def __init__(self, edges):
        self.edges, self.stage_s = edges, {}

    def __call__(self, cms, pafs, stages: bool = False):
        return oracle_postproc(cms, pafs, self.edges)


def cpu_arm(edges):
    """(callable(cms, pafs) -> predict-shaped tuple, kind, description)."""
    try:
        chain = ReferenceChain(edges)
        where = os.path.relpath(chain.root, ROOT) if chain.root.startswith(ROOT) else chain.root
        return chain, "reference", f"unmodified reference files exec'd from {where}"
    except Exception as e:  # noqa: BLE001 - fall back to the port, and say why
        return PortChain(edges), "port", f"oracle/ port (reference files unavailable: {type(e).__name__}: {e})"


def time_cpu(fn, cms_cpu, pafs_cpu, calls: int, threads: int, warm: int = 1):
    torch.set_num_threads(threads)
    for _ in range(warm):
        fn(cms_cpu, pafs_cpu)
    t0 = time.perf_counter()
    for _ in range(calls):
        fn(cms_cpu, pafs_cpu)
    dt = time.perf_counter() - t0
    return cms_cpu.shape[0] * calls / dt, dt / calls


def make_inputs(dev, n_batches: int, seed0: int, dtype=torch.float32):
    from sleap_nn_b200 import synthetic

    edges = synthetic.chain_edges(N_NODES)
    out = []
    for i in range(n_batches):
        poses = synthetic.random_poses(seed0 + i, B, N_INST, N_NODES, IMG_HW, edges)
        cms, pafs = synthetic.render_batch(poses, IMG_HW, STRIDE, edges, dev, seed=seed0 + i)
        out.append((cms.to(dtype), pafs.to(dtype)))  # f16 / bf16: the heads an autocast backbone emits, read natively
    return edges, out


# --------------------------------------------------------------------------- reference arm
def run_reference(args, rank: int):
    """The reference's own CPU implementation on the host cores: same config, all 64 frames of a batch per step.
    Inputs come from the oracle's CPU renderer, so this process never maps libsleapnn_b200.so."""
    if rank != 0:
        return
    from oracle import synth as osynth

    edges = osynth.chain_edges(N_NODES)
    poses = osynth.make_poses(1000, B, N_INST, N_NODES, IMG_HW, edges=edges)
    cms_cpu, pafs_cpu = osynth.render(poses, IMG_HW, STRIDE, edges, seed=1000)
    fn, kind, what = cpu_arm(edges)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    frames = B
    t1 = time.perf_counter()
    fn(cms_cpu, pafs_cpu)
    t1 = time.perf_counter() - t1
    # keep the whole --steps + --warmup run within a few minutes whatever K the caller picks
    while frames > 1 and t1 * (args.steps + args.warmup) > 240.0:
        frames //= 2
        t1 /= 2
    c, p = cms_cpu[:frames], pafs_cpu[:frames]
    for _ in range(max(args.warmup - 1, 0)):
        fn(c, p)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = fn(c, p)
    dt = time.perf_counter() - t0
    fps = frames * args.steps / dt
    n_inst = sum(len(x) for x in out[0])
    fn(c, p, stages=True)
    stages = dict(fn.stage_s)
    fps1, _ = time_cpu(fn, cms_cpu[:8], pafs_cpu[:8], 2, 1, warm=1)
    torch.set_num_threads(cores)
    sample = (f"{frames} frames/step" + (" (the whole cfg3 batch)" if frames == B else " (reduced to fit the time budget)")
              + f", {args.steps} steps, {what}, {cores} threads")
    emit({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": shared_config(args.gpus),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "one_thread": {"value": fps1, "unit": UNIT, "sample": "2 x 8 frames, torch.set_num_threads(1)"},
                         "stage_seconds_per_step": stages, "frames_per_step": frames,
                         "instances_found_last_step": n_inst, "instances_planted_per_step": frames * N_INST},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# --------------------------------------------------------------------------- extras (N = 1): the other BASELINE configs
def run_extras(dev, iters: int):
    """Per-kernel roofline lines for the BASELINE configs the headline does not cover, measured in THIS run (outside the
    headline's timed region): CUDA events on the launching stream, rotating buffers larger than L2 (tools/bench_kernels.py)."""
    from tools import bench_kernels as bk

    bk.QUIET = True
    out = {}
    plan = [
        ("k1_cfg3_f16", lambda: bk.k1_cfg3(dev, iters, torch.float16)),
        ("k1_cfg3_bf16", lambda: bk.k1_cfg3(dev, iters, torch.bfloat16)),
        ("k2_cfg2", lambda: bk.k2_cfg2(dev, iters)),
        ("k2_cfg2_b1024", lambda: bk.k2_cfg2(dev, iters, 1024)),
        ("k2_cfg2_f16", lambda: bk.k2_cfg2(dev, iters, 256, torch.float16)),
        ("k7_cfg4_g8", lambda: bk.k7_cfg4(dev, iters)),
        ("k7_cfg4_g8_bf16", lambda: bk.k7_cfg4(dev, iters, True)),
        ("k7_cfg4_g1", lambda: bk.k7_cfg4(dev, iters, False, 1)),
        ("k7_cfg4_g1_bf16", lambda: bk.k7_cfg4(dev, iters, True, 1)),
        ("k8_cfg4_g8", lambda: bk.k8_cfg4(dev, iters, False, 8)),
        ("k8_cfg4_g8_bf16", lambda: bk.k8_cfg4(dev, iters, True, 8)),
        ("k8_cfg4_g1", lambda: bk.k8_cfg4(dev, iters, False, 1)),
        ("k8_cfg4_g1_bf16", lambda: bk.k8_cfg4(dev, iters, True, 1)),
    ]
    if hasattr(bk, "targets_cfg4_fused"):
        plan += [("targets_cfg4_fused_g1", lambda: bk.targets_cfg4_fused(dev, iters)),
                 ("targets_cfg4_fused_g1_bf16", lambda: bk.targets_cfg4_fused(dev, iters, True))]
    keep = ("kernel", "algorithmic_bytes_per_launch", "avg_launch_ms", "achieved_GBps", "frac")
    for name, fn in plan:
        try:
            d = fn()
            out[name] = {k: d[k] for k in keep}
        except Exception as e:  # noqa: BLE001 - one failing config must not hide the others
            out[name] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
    try:
        n0 = len(bk.RESULTS)
        bk.chain_cfg4(dev, min(iters, 60))
        for d in bk.RESULTS[n0:]:
            if d["bench"] == "k1_cfg4":  # the detect kernel's in-situ time while other chains share the GPU: not a roofline line
                continue
            out[d["bench"]] = {k: d[k] for k in d if k in keep + ("frames_per_s", "instances_found", "instances_planted")}
    except Exception as e:  # noqa: BLE001
        out["chain_cfg4"] = {"error": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------- our arm
def run_ours(args, rank: int, world: int):
    import torch.distributed as dist

    from sleap_nn_b200 import _native as NN
    from sleap_nn_b200.pipeline import BottomUpHostStream, BottomUpPostproc

    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    host_cpus = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # one rank per GPU: keep this rank's pinned staging buffers on the GPU's own NUMA node (e2e leg)
        from sleap_nn_b200.sharding import bind_host_to_gpu

        host_cpus = None if args.no_numa_bind else bind_host_to_gpu(local)
    n_bufs, n_streams = args.buffers, args.streams
    dtype = DTYPES[args.dtype]
    esz = 4 if dtype == torch.float32 else 2
    algo_bytes_per_frame = esz * N_NODES * 512 * 512  # K1 reads every confidence-map element once
    edges, inputs = make_inputs(dev, n_bufs, seed0=100 * (rank + 1), dtype=dtype)
    make_pipe = lambda **kw: BottomUpPostproc(N_NODES, edges, B, (512, 512), cms_stride=STRIDE, pafs_stride=STRIDE,
                                              device=dev, **kw)
    # `--streams` pipeline instances, each with its own tables: instance i's detect kernel runs on the
    # (single) detect stream, its per-frame tail on the high-priority tail stream, so tail(i) overlaps
    # detect(i+1) and the step time tends to the HBM time of the confidence maps.
    # default: one detect stream per pipeline instance + ONE shared high-priority tail stream (see PipelineRing);
    # --graph needs single-stream pipelines, --tail-stream is the older one-detect-stream variant
    tail_priority = not (args.no_tail_priority or args.graph or args.tail_stream)
    tail_stream = torch.cuda.Stream(device=dev, priority=-1) if (args.tail_stream or tail_priority) else None
    pipes = [make_pipe(tail_stream=tail_stream, keep_tables=not args.lean) for _ in range(n_streams)]
    if args.tail_stream:
        det = torch.cuda.Stream(device=dev)
        streams = [det for _ in range(n_streams)]
    else:  # one detect stream per pipeline instance (default: plus ONE shared high-priority tail stream)
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    main = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    from sleap_nn_b200.pipeline import PipelineRing

    ring = PipelineRing(pipes, streams, stagger=True if args.fill_stagger else None) if not args.tail_stream else None

    def run_steps(n, events=None):
        if events is None and ring is not None:  # the product's own multi-stream loop (staggered start, explicit streams)
            ring.reset()
            for i in range(n):
                ring.submit(*inputs[i % n_bufs])
            return
        for i in range(n):
            s = i % n_streams
            cms, pafs = inputs[i % n_bufs]
            with torch.cuda.stream(streams[s]):
                pipes[s](cms, pafs, detect_events=None if events is None else events[i])

    # ---- correctness guard on the timed configuration.  fp32: every planted animal comes back.  Half-precision maps:
    # quantisation can turn a blob's top into a two-pixel plateau (no strict maximum - in the reference too), so the
    # guard is that the natively-read chain equals the fp32 chain on the exact up-cast copy, bit for bit.
    res = pipes[0](*inputs[0])
    inst, _, sc0 = res.to_lists()
    if dtype == torch.float32:
        assert sum(len(x) for x in inst) == B * N_INST, "pipeline did not recover the planted instances"
    else:
        up = make_pipe()
        inst32, _, sc32 = up(inputs[0][0].float(), inputs[0][1].float()).to_lists()
        same_xy = all(a.shape == b_.shape and bool(((a == b_) | (a.isnan() & b_.isnan())).all()) for a, b_ in zip(inst, inst32))
        assert same_xy and all(torch.equal(a, b_) for a, b_ in zip(sc0, sc32)), \
            "native half-precision chain differs from the fp32 chain on the up-cast maps"
        del up

    # ---- device-resident timed region (value).  Default: the eager multi-stream loop (one ctypes call per step).
    # --graph: rotations of the pipeline are captured in a CUDA graph and replayed (one host launch per replay).
    # Exactly --steps steps either way (the remainder of a partial replay runs eagerly).
    run_steps(max(args.warmup, 3))
    barrier()
    graph, per_replay = None, 1
    if args.graph and tail_stream is None:
        from sleap_nn_b200.pipeline import capture_rotation

        graph, per_replay = capture_rotation(pipes, inputs, streams, repeats=args.graph_repeats)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize(dev)
    n_replays, n_rest = (args.steps // per_replay, args.steps % per_replay) if graph is not None else (0, args.steps)
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    all_streams = list({id(s): s for s in streams + ([tail_stream] if tail_stream is not None else [])}.values())

    def timed_steps():
        for _ in range(n_replays):
            graph.replay()
        if n_rest:
            for s in all_streams:
                s.wait_stream(main)
            run_steps(n_rest)
            for s in all_streams:
                main.wait_stream(s)

    with ClockSampler(range(world) if world > 1 else [local], enabled=(rank == 0)) as clocks:
        barrier()
        t_begin.record(main)
        h0 = time.perf_counter()
        timed_steps()
        host_issue_us = (time.perf_counter() - h0) * 1e6 / max(args.steps, 1)  # host time to ISSUE one step (no sync)
        t_end.record(main)
        barrier()
        # keep the GPU busy a little longer if the region was too short for nvidia-smi to sample it
        if t_begin.elapsed_time(t_end) < 600 and rank == 0:
            t_fill = time.perf_counter()
            while time.perf_counter() - t_fill < 0.8:
                timed_steps() if graph is not None else run_steps(n_streams * 8)
                torch.cuda.synchronize(dev)
    # in-situ timing of the detect kernel (events recorded by the C ABI around it) needs the eager loop: a separate,
    # untimed pass of up to 200 steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 200))]
    for a_, b_ in ev:  # torch creates the cudaEvent lazily on first record; the C side re-records them in situ
        a_.record(main)
        b_.record(main)
    torch.cuda.synchronize(dev)
    for s in all_streams:
        s.wait_stream(main)
    run_steps(len(ev), ev)
    torch.cuda.synchronize(dev)
    ms_local = t_begin.elapsed_time(t_end)
    ms_total = ms_local
    detect_ms = [a.elapsed_time(b) for a, b in ev]
    per_rank_ms = [ms_local / args.steps]
    if world > 1:
        t = torch.tensor([ms_local], device=dev)
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        per_rank_ms = [float(g.item()) / args.steps for g in gathered]
        ms_total = max(float(g.item()) for g in gathered)
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- the dominant kernel alone (same process, same inputs, rotating batches): roofline.achieved
    pipe0 = pipes[0]
    iso = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(min(max(args.steps, 50), 200))]
    for a_, b_ in iso:
        a_.record(main); b_.record(main)
    torch.cuda.synchronize(dev)

    def detect_only(i, evs=None):
        cms = inputs[i % n_bufs][0]
        sb_, sc_, sh_, sw_ = cms.stride()
        NN.check(NN.lib.snb_local_peaks_detect_t(NN.ptr(cms), NN.dtype_code(cms.dtype), B, N_NODES, 512, 512, sb_, sc_,
                                                 sh_, sw_, 0.2, pipe0.caps["peak_cap"], NN.ptr(pipe0.buf["frame_count"]),
                                                 NN.ptr(pipe0.buf["keys"]), evs[0].cuda_event if evs else None,
                                                 evs[1].cuda_event if evs else None, NN.stream_ptr(dev)), "detect")

    for i in range(5):
        detect_only(i)
    torch.cuda.synchronize(dev)
    for i, e_ in enumerate(iso):
        detect_only(i, e_)
    torch.cuda.synchronize(dev)
    iso_ms = [a_.elapsed_time(b_) for a_, b_ in iso]
    # what an event pair costs by itself (nothing between the two records): 2.7-2.9 us on a B200, i.e. 5 % of a 52 us
    # kernel and 9 % of a 29 us one - why the per-launch figure is reported but not used for the roofline
    for a_, b_ in iso[:50]:
        a_.record(main); b_.record(main)
    torch.cuda.synchronize(dev)
    empty_pair_ms = sorted(a_.elapsed_time(b_) for a_, b_ in iso[:50])[25]
    # the figure the roofline uses: the AVERAGE launch duration over a run of back-to-back launches, two events around
    # the whole run (replayed from a CUDA graph so the host's launch rate does not enter).  Each launch is the ABI call
    # as it stands - a 256-byte counter memset node + the kernel - so this is an upper bound of the kernel's own time.
    b2b_ms = None
    if rank == 0:
        from tools import bench_kernels as bk_

        b2b_ms = bk_.timed(lambda i: detect_only(i), 120)

    # ---- end-to-end timed region (host buffers in, host results out), same steps.  The confidence maps are
    # copied pinned-host -> device every step; the PAF tensor is only sampled (20 taps per candidate), so it is
    # read in place from pinned host memory over PCIe (zero-copy) and only the sampled 32-byte sectors cross the link.
    host = [(c.cpu().pin_memory(), p.cpu().pin_memory()) for c, p in inputs[: min(2, n_bufs)]]
    hs = BottomUpHostStream(lambda: make_pipe(keep_tables=True), depth=args.e2e_depth,
                            zero_copy_pafs=not args.copy_pafs, zero_copy_cms=args.zero_copy_cms)
    want_per_host = []  # what each host batch must yield: exactly what the resident chain yields for it
    for k in range(len(host)):
        r_ = pipes[0](*inputs[k]).to_lists()
        want_per_host.append(sum(len(x) for x in r_[0]))
    for i in range(3):
        hs.submit(*host[i % len(host)])
    hs.drain()
    torch.cuda.synchronize(dev)
    p0 = hs.pipes[0]
    n_cand = int(p0.buf["edge_off"][:, -1].sum().item()) if p0._args.edge_off else 16 * B
    paf_sector_bytes = n_cand * p0.n_points * 2 * 32 if hs.last_zero_copy else 0
    h2d = hs.h2d_bytes + paf_sector_bytes
    d2h = hs.d2h_bytes
    barrier()
    e2e_steps = min(args.steps, args.e2e_steps)
    # N > 1: the world * e2e_steps batches of the job sit in ONE queue and every rank pulls the next batch index when it
    # has a free slot (a counter in torch.distributed's store, ~0.1 ms per pull against ~6 ms per batch) - what a
    # multi-GPU predictor fed from host memory does.  On this pool's 8-GPU boxes the GPUs sit behind two host bridges
    # of unequal speed (tools/h2d_scaling_probe.py: 20.8 vs 35.7 GB/s per GPU when all eight copy at once), and a fixed
    # equal split is paced by the slow group.  --e2e-static keeps the fixed split.  (The protocol of
    # sleap_nn_b200.sharding.SharedBatchQueue, which has its own world-size-2 gloo test, spelled out inline.)
    queue = None
    if world > 1 and not args.e2e_static:
        try:
            queue = dist.distributed_c10d._get_default_store()
            if rank == 0:
                queue.set("snb_e2e_next", "0")
        except Exception:  # noqa: BLE001 - no store: fall back to the fixed split
            queue = None
    barrier()
    total_steps = world * e2e_steps

    def next_index(i):
        if queue is None:
            return i if i < e2e_steps else None
        k = int(queue.add("snb_e2e_next", 1)) - 1
        return k if k < total_steps else None

    n_got, n_want, my_steps = 0, 0, 0
    t0 = time.perf_counter()
    while True:  # every step: H2D of its inputs, the chain, D2H of its results; `depth` steps in flight
        k = next_index(my_steps)
        if k is None:
            break
        out = hs.submit(*host[k % len(host)])
        n_want += want_per_host[k % len(host)]
        my_steps += 1
        if out is not None:
            n_got += sum(len(x) for x in out[0])
    for out in hs.drain():
        n_got += sum(len(x) for x in out[0])
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3  # host clock: the region ends when the last result is unpacked on the host
    barrier()
    assert n_got == n_want, "host path lost instances"
    e2e_steps_per_rank = [my_steps]
    if world > 1:
        t = torch.tensor([e2e_ms, float(my_steps)], device=dev, dtype=torch.float64)
        got = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(got, t)
        e2e_ms = max(float(g[0]) for g in got)
        e2e_steps_per_rank = [int(g[1]) for g in got]
        assert sum(e2e_steps_per_rank) == total_steps
    e2e_value = B * sum(e2e_steps_per_rank) / (e2e_ms / 1e3)
    zero_copy_pafs, zero_copy_cms = hs.last_zero_copy, getattr(hs, "last_zero_copy_cms", False)
    paf_host_bytes = host[0][1].numel() * esz
    del hs, host

    # ---- N > 1: the same steps through ShardRunner (results packed on the device, no host sync inside the shard) and
    # the one collective of the design, the end-of-shard gather of the variable-length instance lists
    gather = None
    if world > 1:
        from sleap_nn_b200 import sharding

        total_frames = world * B * args.steps
        src = lambda s, e: inputs[((s // B) % n_bufs)]
        # warm-up: pack kernel load, and NCCL's lazy all_gather channel setup (first call only)
        sharding.gather_packed(sharding.ShardRunner(pipes[0], world * B * 4, rank, world).run(src).finish())
        runner = sharding.ShardRunner(pipes[0], total_frames, rank, world, rows_cap=(total_frames // world) * 8)
        barrier()
        g0, g1, g2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        g0.record(main)
        runner.run(src)
        g1.record(main)
        local_res = runner.finish()
        merged = sharding.gather_packed(local_res)
        g2.record(main)
        torch.cuda.synchronize(dev)
        t = torch.tensor([g0.elapsed_time(g1), g1.elapsed_time(g2)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        payload = merged.rows * (N_NODES * 2 * 4 + N_NODES * 4 + 4 + 4) + merged.n_frames * 4
        gather = {"shard_run_ms": float(t[0]), "gather_ms": float(t[1]), "rows_gathered": merged.rows,
                  "frames": merged.n_frames, "payload_bytes": payload,
                  "frames_per_s_incl_pack_and_gather": total_frames / ((float(t[0]) + float(t[1])) / 1e3),
                  "note": "ShardRunner on one stream: chain + snb_pack_instances per batch; then finish() (the one host "
                          "sync) + gather_packed (all_gather of counts, then of the padded payloads, NCCL)"}

    if rank == 0:
        peak, which = measured_peaks()
        per_launch_event_ms = sum(iso_ms) / len(iso_ms)
        avg_detect_ms = b2b_ms if b2b_ms else per_launch_event_ms
        insitu_ms = sum(detect_ms) / len(detect_ms)
        achieved = algo_bytes_per_frame * B / (avg_detect_ms / 1e3) / 1e9
        kernel = {"f32": "local_peaks_detect_vec<float,4,1,6,1>", "f16": "local_peaks_detect_vec<__half,2,1,8,1>",
                  "bf16": "local_peaks_detect_vec<__nv_bfloat16,2,1,8,1>"}[args.dtype]  # <T, loads, rows, CTAs/SM, EXACT>
        traffic, traffic_rec = recorded_traffic(kernel)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": shared_config(world, args.dtype),
            "run": dict(streams=n_streams, input_batches=n_bufs,
                        tail="fused per-frame tail kernel" + (" on one shared high-priority stream" if tail_stream is not None else ""),
                        intermediate_tables_written=not args.lean,
                        launch=("CUDA graph: %d steps per replay, remainder eager" % per_replay) if graph is not None else "eager Python loop",
                        per_rank_ms_per_step={"min": min(per_rank_ms), "median": statistics.median(per_rank_ms),
                                              "max": max(per_rank_ms), "all": per_rank_ms}),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "in_flight": args.e2e_depth,
                    "steps_per_rank": e2e_steps_per_rank,
                    "split": ("one shared batch queue, ranks pull" if (world > 1 and not args.e2e_static) else "fixed, equal per rank"),
                    "cms": ("streamed by the detect kernel straight from pinned host memory (zero-copy)"
                            if zero_copy_cms else "cudaMemcpyAsync pinned host -> HBM staging buffer"),
                    "host_cpus_rank0": ("all" if host_cpus is None else f"{len(host_cpus)} CPUs local to the GPU"),
                    "pafs": ("sampled in place from pinned host memory (zero-copy): "
                             f"{paf_sector_bytes} B of 32-byte sectors per step instead of {paf_host_bytes} B"
                             if zero_copy_pafs else "copied to the device every step")},
            "gpu_launches": pipes[0].launches_per_call * args.steps, "host_issue_us_per_step": host_issue_us,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_record": traffic_rec, "kernel": kernel, "peak_source": which,
                         "algorithmic_bytes_per_launch": algo_bytes_per_frame * B, "avg_launch_ms": avg_detect_ms,
                         "how": "average launch duration: two CUDA events around 120 back-to-back launches of the kernel "
                                "alone on one stream (each = the ABI call: a 256-byte counter memset node + the kernel), "
                                "replayed from a CUDA graph, rotating batches larger than L2; traffic = dram__bytes_read.sum + "
                                "dram__bytes_write.sum per launch from the committed ncu --set full capture named in traffic_record",
                         "per_launch_event_pair_ms": per_launch_event_ms,
                         "per_launch_event_pair_note": "the same kernel bracketed launch by launch by events recorded in the C ABI "
                                                       "(round 1's method): includes what an event pair costs by itself",
                         "empty_event_pair_ms": empty_pair_ms,
                         "ncu_frac": ((algo_bytes_per_frame * B / (traffic_rec["ncu_time_us"] * 1e-6) / 1e9) / peak
                                      if traffic_rec and traffic_rec.get("ncu_time_us") else None),
                         "ncu_note": "the same bytes over gpu__time_duration of the committed ncu capture (cold cache, "
                                     "serialised; no event records around the kernel) - for comparison only",
                         "in_situ_avg_launch_ms": insitu_ms,
                         "in_situ_note": "same events inside the timed region; inflated when two streams overlap two detect kernels",
                         "whole_step_frac": (algo_bytes_per_frame * B / (ms_total / args.steps / 1e3) / 1e9) / peak},
            "clocks": clocks.summary(),
        }
        if gather is not None:
            line["gather"] = gather
        if world == 1 and not args.no_cpu_baseline:
            # the CPU arm on 16 frames of the batch the GPU just processed, and - for free - the parity of the two
            n_chk = 16
            cms_cpu, pafs_cpu = inputs[0][0][:n_chk].float().cpu(), inputs[0][1][:n_chk].float().cpu()
            fn, kind, what = cpu_arm(edges)
            cores = os.cpu_count() or 1
            fps, per_call = time_cpu(fn, cms_cpu, pafs_cpu, args.cpu_calls, cores)
            fps1, _ = time_cpu(fn, cms_cpu[:8], pafs_cpu[:8], 1, 1, warm=0)
            torch.set_num_threads(cores)
            fn(cms_cpu, pafs_cpu, stages=True)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"{args.cpu_calls} x {n_chk} frames of the same cfg3 batch ({per_call:.2f} s/call), {what}",
                                    "one_thread": {"value": fps1, "unit": UNIT, "sample": "8 frames, torch.set_num_threads(1)"},
                                    "stage_seconds_per_call": dict(fn.stage_s)}
            want = oracle_postproc(cms_cpu, pafs_cpu, edges)
            got = pipes[0](*inputs[0]).to_lists()
            max_px = max_sc = 0.0
            same = True
            for b in range(n_chk):
                g_xy, w_xy = got[0][b], want[0][b]
                g_pv, w_pv = got[1][b], want[1][b]  # peak values: bit-exact, NaN (a missing node) == NaN
                ok = (g_xy.shape == w_xy.shape and bool((g_xy.isnan() == w_xy.isnan()).all())
                      and bool(((g_pv == w_pv) | (g_pv.isnan() & w_pv.isnan())).all()))
                same = same and ok
                if ok and g_xy.numel():
                    max_px = max(max_px, float((g_xy - w_xy).abs().nan_to_num(0.0).max()))
                    max_sc = max(max_sc, float((got[2][b] - want[2][b]).abs().max()))
            line["parity"] = {"parity_checked_frames": n_chk, "instances_and_peak_values_identical": same,
                              "max_abs_px": max_px, "max_score_err": max_sc,
                              "against": "oracle/ (the CPU restatement pinned to the reference's goldens) on the same maps"}
            assert same and max_px <= 1e-4 and max_sc <= 1e-5, f"GPU result differs from the oracle: {line['parity']}"
        if world == 1 and not args.no_extras:
            torch.cuda.empty_cache()
            line["extra"] = run_extras(dev, args.extra_iters)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f32", choices=list(DTYPES),
                    help="element type of the maps: f32 (the reference's, default) or the f16 / bf16 heads of an autocast backbone")
    ap.add_argument("--streams", type=int, default=3)
    ap.add_argument("--buffers", type=int, default=6)
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--e2e-depth", type=int, default=2, help="batches in flight in the host-buffer pipeline (BottomUpHostStream)")
    ap.add_argument("--cpu-calls", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the `extra` block (the other BASELINE configs)")
    ap.add_argument("--extra-iters", type=int, default=100)
    ap.add_argument("--graph-repeats", type=int, default=20, help="rotations of the pipeline captured per CUDA graph")
    ap.add_argument("--graph", action="store_true",
                    help="replay CUDA graphs of --graph-repeats pipeline rotations instead of the eager launch loop (one host "
                         "launch per replay; measured slower at N=1: the graph's three chains run in lockstep, see DESIGN.md)")
    ap.add_argument("--tail-stream", dest="tail_stream", action="store_true",
                    help="one detect stream + one high-priority tail stream instead of one stream per pipeline instance")
    ap.add_argument("--no-tail-priority", dest="no_tail_priority", action="store_true",
                    help="A/B: per-instance streams WITHOUT the shared high-priority tail stream (tails run on the detect streams)")
    ap.add_argument("--fill-stagger", dest="fill_stagger", action="store_true",
                    help="A/B: after a sync the first detect kernels of the ring's chains start one after another instead of together")
    ap.add_argument("--zero-copy-cms", action="store_true",
                    help="e2e: the detect kernel streams the pinned host confidence maps itself (no cudaMemcpy + HBM staging)")
    ap.add_argument("--e2e-static", action="store_true",
                    help="N>1 e2e: a fixed equal number of batches per rank instead of one shared batch queue")
    ap.add_argument("--copy-pafs", action="store_true", help="e2e: stage the PAF tensor in HBM instead of sampling it in place")
    ap.add_argument("--no-numa-bind", action="store_true", help="N>1: do not pin each rank to the CPUs next to its GPU")
    ap.add_argument("--lean", action="store_true", help="do not write candidate / match tables to global memory")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    # stdout carries exactly ONE JSON line: anything a library writes to fd 1 (e.g. NCCL's version banner) goes to
    # stderr instead, and the line itself is written to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global emit
    emit = lambda obj: os.write(json_fd, (json.dumps(obj) + "\n").encode())
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: `python bench.py --gpus N` re-launches itself under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd, stdout=json_fd))
    run_ours(args, rank, world)


if __name__ == "__main__":
    main()
