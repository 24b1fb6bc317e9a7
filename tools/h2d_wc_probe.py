#!/usr/bin/env python
"""Probe: host -> device copy rate of one cfg3 batch of confidence maps (336 MB) from (a) ordinary pinned memory
(torch pin_memory = cudaHostAlloc default) and (b) write-combined pinned memory (cudaHostAllocWriteCombined, not
snooped during PCIe reads).  Scratch tool for the e2e leg; prints GB/s for both."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

torch.cuda.init()
dev = torch.device("cuda", 0)
n = 64 * 5 * 512 * 512
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else None
rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
dst = [torch.empty((n,), dtype=torch.float32, device=dev) for _ in range(2)]


def rate(host, label, iters=30):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(3):
            dst[i % 2].copy_(host, non_blocking=True)
        s.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for i in range(iters):
            dst[i % 2].copy_(host, non_blocking=True)
        b.record(s)
        s.synchronize()
    ms = a.elapsed_time(b) / iters
    print(f"{label:28s} {n * 4 / ms / 1e6:7.2f} GB/s  ({ms:.3f} ms per 336 MB)  pinned={host.is_pinned()}", flush=True)


plain = torch.rand((n,), dtype=torch.float32).pin_memory()
rate(plain, "pinned (default)")
for flags, label in ((0x4, "pinned write-combined"), (0x1, "pinned portable"), (0x4 | 0x2, "pinned WC + mapped")):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), n * 4, flags)
    if rc != 0:
        print(label, "cudaHostAlloc failed", rc)
        continue
    arr = np.ctypeslib.as_array((ctypes.c_float * n).from_address(p.value))
    t0 = time.perf_counter()
    arr[:] = plain.numpy()
    fill = time.perf_counter() - t0
    host = torch.from_numpy(arr)
    rate(host, label)
    print(f"   host fill of the buffer: {n * 4 / fill / 1e9:.1f} GB/s")
# two copies in flight on two streams (the depth-2 pipeline)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for i in range(15):
    for k, s in enumerate((s1, s2)):
        s.wait_stream(torch.cuda.current_stream()) if i == 0 else None
        with torch.cuda.stream(s):
            dst[k].copy_(plain, non_blocking=True)
for s in (s1, s2):
    torch.cuda.current_stream().wait_stream(s)
b.record()
torch.cuda.synchronize()
print(f"two streams, default pinned: {30 * n * 4 / a.elapsed_time(b) / 1e6:.2f} GB/s")
