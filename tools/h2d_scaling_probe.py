#!/usr/bin/env python
"""Probe for the e2e (host-buffer) leg: what limits host -> device feeding when N ranks stream at once?

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_scaling_probe.py

Pure transfers, no post-processing: every rank owns one cfg3 batch of confidence maps (336 MB fp32) in pinned host
memory and pushes it to its GPU repeatedly.  Measured per rank and in aggregate:

  solo        each rank in turn while the others idle (the link itself: PCIe Gen5 x16 ~ 55 GB/s)
  together    all ranks at once (shared switch uplinks / host memory feed show up here)
  together_thp   the same from a 2 MB-aligned, transparent-huge-page backed buffer registered with cudaHostRegister
  zero_copy   the detect kernel streaming the pinned maps itself through their device alias (no cudaMemcpy, no HBM staging)
  half        together, fp16 maps (half the bytes per frame)

Rank 0 also records the PCIe / NUMA topology (`nvidia-smi topo -m`, `lspci -tv`, /sys/devices/system/node).  One JSON
line on stdout (rank 0); frames/s equivalents assume 5.24 MB (fp32) per cfg3 frame.
"""
import ctypes
import json
import mmap
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from sleap_nn_b200 import _native as N  # noqa: E402

B, C, H, W = 64, 5, 512, 512


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout[-6000:]
    except Exception as e:  # noqa: BLE001
        return f"{type(e).__name__}: {e}"


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = B * C * H * W
    host = torch.rand((n,), dtype=torch.float32).mul_(0.15).pin_memory()
    host16 = host.to(torch.float16).pin_memory()
    dst = [torch.empty((n,), dtype=torch.float32, device=dev) for _ in range(2)]
    dst16 = [torch.empty((n,), dtype=torch.float16, device=dev) for _ in range(2)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def copy_rate(src, dsts, seconds=0.6):
        """GB/s of back-to-back H2D copies on two streams (the depth-2 pipeline's shape)."""
        for k in range(2):
            with torch.cuda.stream(streams[k]):
                dsts[k].copy_(src, non_blocking=True)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        it = 0
        while time.perf_counter() - t0 < seconds:
            for k in range(2):
                with torch.cuda.stream(streams[k]):
                    dsts[k].copy_(src, non_blocking=True)
            it += 2
            streams[0].synchronize()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        return it * src.numel() * src.element_size() / dt / 1e9

    def gather(x):
        if world == 1:
            return [x]
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    res = {}
    # ---- solo: one rank at a time
    solo = 0.0
    for r in range(world):
        barrier()
        if r == rank:
            solo = copy_rate(host, dst)
        barrier()
    res["solo_GBps_per_rank"] = gather(solo)
    # ---- together
    barrier()
    res["together_GBps_per_rank"] = gather(copy_rate(host, dst))
    barrier()
    res["together_half_GBps_per_rank"] = gather(copy_rate(host16, dst16))
    # ---- together, THP-backed 2 MB aligned buffer registered with the driver
    thp = None
    try:
        nbytes = ((n * 4 + (2 << 20) - 1) // (2 << 20)) * (2 << 20)
        mm = mmap.mmap(-1, nbytes + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        addr = ctypes.addressof(ctypes.c_char.from_buffer(mm))
        aligned = (addr + (2 << 20) - 1) & ~((2 << 20) - 1)
        libc = ctypes.CDLL("libc.so.6", use_errno=True)
        libc.madvise(ctypes.c_void_p(aligned), ctypes.c_size_t(nbytes), 14)  # MADV_HUGEPAGE
        arr = (ctypes.c_float * n).from_address(aligned)
        t = torch.frombuffer(arr, dtype=torch.float32)
        t.copy_(host)  # first touch
        rt = ctypes.CDLL("libcudart.so.12")
        rc = rt.cudaHostRegister(ctypes.c_void_p(aligned), ctypes.c_size_t(nbytes), ctypes.c_uint(0))
        if rc == 0:
            barrier()
            thp = copy_rate(t, dst)
            barrier()
            rt.cudaHostUnregister(ctypes.c_void_p(aligned))
        else:
            thp = -float(rc)
            barrier(); barrier()
    except Exception as e:  # noqa: BLE001
        res["thp_error"] = f"{type(e).__name__}: {e}"
        barrier(); barrier()
    res["together_thp_GBps_per_rank"] = gather(thp if thp is not None else 0.0)
    # ---- zero-copy: the detect kernel reads the pinned maps in place
    alias = ctypes.c_void_p()
    zc = 0.0
    if N.lib.snb_host_device_pointer(host.data_ptr(), ctypes.byref(alias)) == N.OK and alias.value:
        cap = 256
        count = torch.empty((B,), dtype=torch.int32, device=dev)
        keys = torch.empty((B * cap,), dtype=torch.int32, device=dev)
        st = N.stream_ptr(dev)

        def detect():
            N.check(N.lib.snb_local_peaks_detect_t(alias.value, 0, B, C, H, W, C * H * W, H * W, W, 1, 0.2, cap, N.ptr(count),
                                                   N.ptr(keys), None, None, st), "detect")

        detect()
        barrier()
        t0 = time.perf_counter()
        it = 0
        while time.perf_counter() - t0 < 0.6:
            detect()
            it += 1
            torch.cuda.synchronize(dev)
        zc = it * n * 4 / (time.perf_counter() - t0) / 1e9
        barrier()
    res["together_zero_copy_detect_GBps_per_rank"] = gather(zc)
    if rank == 0:
        for k in list(res):
            if k.endswith("_per_rank"):
                res[k.replace("_per_rank", "_sum")] = sum(res[k])
        res["frames_per_s_equiv_together_fp32"] = res["together_GBps_sum"] * 1e9 / (4 * C * H * W)
        res["frames_per_s_equiv_together_fp16"] = res["together_half_GBps_sum"] * 1e9 / (2 * C * H * W)
        res.update({"tool": "h2d_scaling_probe", "n_gpus": world, "bytes_per_copy": n * 4,
                    "cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)),
                    "topo": sh("nvidia-smi topo -m"), "lspci_tree": sh("lspci -tv 2>/dev/null | head -80"),
                    "numa_nodes": sh("ls -d /sys/devices/system/node/node* 2>/dev/null; cat /sys/devices/system/node/node*/meminfo 2>/dev/null | grep MemTotal"),
                    "thp": sh("cat /sys/kernel/mm/transparent_hugepage/enabled")})
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
