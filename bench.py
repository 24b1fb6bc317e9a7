#!/usr/bin/env python
"""Benchmark of the bottom-up post-processing hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (`config.workload`): BASELINE cfg3 "bottom-up mice" - 1024x1024 frames, 5 nodes /
4 edges, stride-2 confidence maps (64,5,512,512) + PAFs (64,8,512,512), batch 64, full peak +
PAF grouping.  One STEP = one pass of the whole hot path (K1 peaks + refinement -> K4 line
scores -> K5 assignment -> K6 assembly) over one batch of 64 synthetic frames.

  value   frames/s with the maps already resident in HBM (as they are after the backbone),
          batches pipelined over `--streams` CUDA streams, inputs rotated over several distinct
          batches (each 872 MB > 126 MB L2, so nothing is served from cache);
  e2e     the same metric through the public API with HOST buffers: per step a pinned-host ->
          device copy of the maps and a device -> host read of the grouped instances;
  roofline  the dominant kernel (streaming NMS detect) timed in situ with CUDA events recorded
          around it inside the timed region; algorithmic bytes = 4*C*H*W*B per launch;
  cpu_baseline  the CPU oracle (a port of the reference's op chain, torch CPU ops, all host
          threads) on a bounded sample of the same frames.  N=1, rank 0 only.

`--impl reference` times that CPU port alone (the reference is pure Python / ATen and cannot
travel to the GPU box).  Multi-GPU: frames shard, no collective on the hot path;
one process per GPU (torchrun), barrier + synchronize around the timed region, max over ranks.
"""

from __future__ import annotations

import argparse
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
ALGO_BYTES_PER_FRAME = 4 * N_NODES * 512 * 512  # K1 reads every confidence-map element once
METRIC, UNIT = "bottom-up post-proc frames/s", "frames/s"
emit = lambda obj: print(json.dumps(obj), flush=True)  # replaced in main() by a writer on the saved stdout descriptor


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"  # B200_PROFILING.md fallback


class ClockSampler:
    """SM clocks / throttle reasons of the job's GPUs sampled every 100 ms while the timed region runs.

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


# --------------------------------------------------------------------------- CPU port (oracle)
def oracle_postproc(cms_cpu, pafs_cpu, edges):
    """The reference's bottom-up post-processing chain, as restated by oracle/ (CPU, torch ops)."""
    from oracle import paf as opaf
    from oracle import peaks as opeaks
    from oracle.synth import split_by_sample

    nb = cms_cpu.shape[0]
    pts, vals, si, ci = opeaks.local_peaks(cms_cpu, 0.2, "integral")
    peaks, pvs, pcs = (split_by_sample(x, si, nb) for x in (pts * STRIDE, vals, ci))
    return opaf.predict(pafs_cpu.permute(0, 2, 3, 1), peaks, pvs, pcs, edges, N_NODES, STRIDE)


def time_cpu_port(cms_cpu, pafs_cpu, edges, frames_per_call: int, calls: int, warm: int = 1):
    torch.set_num_threads(os.cpu_count() or 1)
    c, p = cms_cpu[:frames_per_call], pafs_cpu[:frames_per_call]
    for _ in range(warm):
        oracle_postproc(c, p, edges)
    t0 = time.perf_counter()
    for _ in range(calls):
        oracle_postproc(c, p, edges)
    dt = time.perf_counter() - t0
    return frames_per_call * calls / dt, dt / calls


DTYPES = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}


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
    if rank != 0:
        return
    from sleap_nn_b200 import synthetic

    frames_per_step = 16  # bounded sample of the workload: 16 of the batch's 64 frames per step
    edges = synthetic.chain_edges(N_NODES)
    if torch.cuda.is_available():
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
        poses = synthetic.random_poses(1000, frames_per_step, N_INST, N_NODES, IMG_HW, edges)
        cms, pafs = synthetic.render_batch(poses, IMG_HW, STRIDE, edges, dev, seed=1000)
        cms_cpu, pafs_cpu = cms.cpu(), pafs.cpu()
        del cms, pafs
    else:  # CPU-only box: render with the oracle's own target code
        from oracle import synth as osynth

        poses = osynth.make_poses(1000, frames_per_step, N_INST, N_NODES, IMG_HW, edges=edges)
        cms_cpu, pafs_cpu = osynth.render(poses, IMG_HW, STRIDE, edges, seed=1000)
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(max(args.warmup - 1, 0)):
        oracle_postproc(cms_cpu, pafs_cpu, edges)
    t1 = time.perf_counter()
    oracle_postproc(cms_cpu, pafs_cpu, edges)
    t1 = time.perf_counter() - t1
    # keep the whole --steps run within a few minutes whatever K the caller picks: halve the per-step sample if needed
    while frames_per_step > 1 and t1 * args.steps > 150.0:
        frames_per_step //= 2
        t1 /= 2
        cms_cpu, pafs_cpu = cms_cpu[:frames_per_step], pafs_cpu[:frames_per_step]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_postproc(cms_cpu, pafs_cpu, edges)
    dt = time.perf_counter() - t0
    fps = frames_per_step * args.steps / dt
    sample = f"{frames_per_step} frames/step of the cfg3 batch, {args.steps} steps, torch CPU ops, {torch.get_num_threads()} threads"
    emit(({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(WORKLOAD, frames_per_step=frames_per_step),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is pure Python/ATen with no native path and cannot be installed offline (needs sleap-io, "
                "lightning, omegaconf): this arm times oracle/, the CPU port of its op chain, on the host cores; in the build "
                "container (8 cores) the port runs this chain 2.9x FASTER than the unmodified reference files (99 vs 34 "
                "frames/s), so the baseline errs in the reference's favour",
    }))


# --------------------------------------------------------------------------- our arm
def run_ours(args, rank: int, world: int):
    import torch.distributed as dist

    from sleap_nn_b200.pipeline import BottomUpPostproc

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
    algo_bytes_per_frame = esz * N_NODES * 512 * 512
    edges, inputs = make_inputs(dev, n_bufs, seed0=100 * (rank + 1), dtype=dtype)
    # `--streams` pipeline instances, each with its own tables: instance i's detect kernel runs on the
    # (single) detect stream, its per-frame tail on the high-priority tail stream, so tail(i) overlaps
    # detect(i+1) and the step time tends to the HBM time of the confidence maps.
    tail_stream = torch.cuda.Stream(device=dev, priority=-1) if args.tail_stream else None
    pipes = [BottomUpPostproc(N_NODES, edges, B, (512, 512), cms_stride=STRIDE, pafs_stride=STRIDE, device=dev,
                              tail_stream=tail_stream, keep_tables=not args.lean) for _ in range(n_streams)]
    if tail_stream is not None:
        det = torch.cuda.Stream(device=dev)
        streams = [det for _ in range(n_streams)]
    else:
        streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    main = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def run_steps(n, events=None):
        for i in range(n):
            s = i % n_streams
            with torch.cuda.stream(streams[s]):
                cms, pafs = inputs[i % n_bufs]
                pipes[s](cms, pafs, detect_events=None if events is None else events[i])

    # ---- correctness guard: the timed configuration must produce the planted animals
    res = pipes[0](*inputs[0])
    inst, _, _ = res.to_lists()
    assert sum(len(x) for x in inst) == B * N_INST, "pipeline did not recover the planted instances"

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
    ms_total = t_begin.elapsed_time(t_end)
    detect_ms = [a.elapsed_time(b) for a, b in ev]
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = world * B * args.steps / (ms_total / 1e3)

    # ---- the dominant kernel alone (same process, same inputs, rotating batches): roofline.achieved
    from sleap_nn_b200 import _native as NN

    pipe0 = pipes[0]
    iso = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 200))]
    for a_, b_ in iso:
        a_.record(main); b_.record(main)
    torch.cuda.synchronize(dev)

    def detect_only(i, evs=None):
        cms = inputs[i % n_bufs][0]
        sb_, sc_, sh_, sw_ = cms.stride()
        NN.check(NN.lib.snb_local_peaks_detect_t(NN.ptr(cms), NN.dtype_code(cms.dtype), B, N_NODES, 512, 512, sb_, sc_, sh_, sw_, 0.2,
                                               pipe0.caps["peak_cap"], NN.ptr(pipe0.buf["frame_count"]),
                                               NN.ptr(pipe0.buf["keys"]), evs[0].cuda_event if evs else None,
                                               evs[1].cuda_event if evs else None, NN.stream_ptr(dev)), "detect")

    for i in range(5):
        detect_only(i)
    torch.cuda.synchronize(dev)
    for i, e_ in enumerate(iso):
        detect_only(i, e_)
    torch.cuda.synchronize(dev)
    iso_ms = [a_.elapsed_time(b_) for a_, b_ in iso]

    # ---- end-to-end timed region (host buffers in, host results out), same steps.  The confidence maps are
    # copied pinned-host -> device every step; the PAF tensor is only sampled (20 taps per candidate), so it is
    # read in place from pinned host memory over PCIe (zero-copy) and only the sampled 32-byte sectors cross the link.
    host = [(c.cpu().pin_memory(), p.cpu().pin_memory()) for c, p in inputs[: min(2, n_bufs)]]
    from sleap_nn_b200.pipeline import BottomUpHostStream

    hs = BottomUpHostStream(lambda: BottomUpPostproc(N_NODES, edges, B, (512, 512), cms_stride=STRIDE, pafs_stride=STRIDE,
                                                     device=dev, keep_tables=True), depth=args.e2e_depth,
                            zero_copy_pafs=not args.copy_pafs, zero_copy_cms=args.zero_copy_cms)
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
    n_got = 0
    t0 = time.perf_counter()
    for i in range(e2e_steps):  # every step: H2D of its inputs, the chain, D2H of its results; `depth` steps in flight
        out = hs.submit(*host[i % len(host)])
        if out is not None:
            n_got += sum(len(x) for x in out[0])
    for out in hs.drain():
        n_got += sum(len(x) for x in out[0])
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3  # host clock: the region ends when the last result is unpacked on the host
    barrier()
    assert n_got == e2e_steps * B * N_INST, "host path did not recover the planted instances"
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * B * e2e_steps / (e2e_ms / 1e3)

    if rank == 0:
        peak, which = measured_peaks()
        avg_detect_ms = sum(iso_ms) / len(iso_ms)
        insitu_ms = sum(detect_ms) / len(detect_ms)
        achieved = algo_bytes_per_frame * B / (avg_detect_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": dict(WORKLOAD, parallelism=f"frame-sharded x{world}, no collective", streams=n_streams,
                           input_batches=n_bufs, l2="inputs larger than L2: each batch is 872 MB and batches rotate",
                           tail="fused per-frame tail kernel" + (" on a high-priority second stream" if tail_stream is not None else ""),
                           intermediate_tables_written=not args.lean,
                           launch=("CUDA graph: %d steps per replay, remainder eager" % per_replay) if graph is not None else "eager Python loop"),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "in_flight": args.e2e_depth,
                    "cms": ("streamed by the detect kernel straight from pinned host memory (zero-copy)"
                            if getattr(hs, "last_zero_copy_cms", False) else "cudaMemcpyAsync pinned host -> HBM staging buffer"),
                    "host_cpus_rank0": ("all" if host_cpus is None else f"{len(host_cpus)} CPUs local to the GPU"),
                    "pafs": ("sampled in place from pinned host memory (zero-copy): "
                             f"{paf_sector_bytes} B of 32-byte sectors per step instead of {host[0][1].numel() * esz} B"
                             if hs.last_zero_copy else "copied to the device every step")},
            "gpu_launches": pipes[0].launches_per_call * args.steps, "host_issue_us_per_step": host_issue_us,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": 338793984, "kernel": "local_peaks_detect_vec4<4,1,6>", "peak_source": which,
                         "algorithmic_bytes_per_launch": algo_bytes_per_frame * B, "avg_launch_ms": avg_detect_ms,
                         "how": "CUDA events recorded by the C ABI right around the kernel, kernel running alone, "
                                "rotating 872 MB batches; traffic = dram__bytes_read.sum 335598848 + dram__bytes_write.sum 3195136 per "
                                "launch from one ncu --set full capture (profiles/r1_d_detect_tail_ncu_raw.txt)",
                         "in_situ_avg_launch_ms": insitu_ms,
                         "in_situ_note": "same events inside the timed region; inflated when two streams overlap two detect kernels",
                         "whole_step_frac": (algo_bytes_per_frame * B / (ms_total / args.steps / 1e3) / 1e9) / peak},
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            cms_cpu, pafs_cpu = inputs[0][0][:16].float().cpu(), inputs[0][1][:16].float().cpu()
            fps, per_call = time_cpu_port(cms_cpu, pafs_cpu, edges, 16, args.cpu_calls)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{args.cpu_calls} x 16 frames of the same cfg3 batch ({per_call:.2f} s/call), "
                                              "oracle/ torch-CPU port of the reference chain"}
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
    ap.add_argument("--cpu-calls", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph-repeats", type=int, default=20, help="rotations of the pipeline captured per CUDA graph")
    ap.add_argument("--graph", action="store_true",
                    help="replay CUDA graphs of --graph-repeats pipeline rotations instead of the eager launch loop (one host "
                         "launch per replay; measured slower at N=1: the graph's three chains run in lockstep, see DESIGN.md)")
    ap.add_argument("--tail-stream", dest="tail_stream", action="store_true",
                    help="one detect stream + one high-priority tail stream instead of one stream per pipeline instance")
    ap.add_argument("--zero-copy-cms", action="store_true",
                    help="e2e: the detect kernel streams the pinned host confidence maps itself (no cudaMemcpy + HBM staging)")
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
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU port)")
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: `python bench.py --gpus N` re-launches itself under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd, stdout=json_fd))
    run_ours(args, rank, world)


if __name__ == "__main__":
    main()
