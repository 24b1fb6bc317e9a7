"""Scratch micro-benchmark of the K1 chain (detect + finalize) at cfg3 size. Not part of the product."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sleap_nn_b200.inference.ops import peaks as P

dev = torch.device("cuda")
B, C, H, W = 64, 5, 512, 512
nbuf = 6
g = torch.Generator(device="cuda").manual_seed(0)
bufs = []
for i in range(nbuf):
    x = torch.rand((B, C, H, W), device=dev, generator=g) * 1e-3
    ys = torch.randint(8, H - 8, (B, C, 2), device=dev, generator=g)
    xs = torch.randint(8, W - 8, (B, C, 2), device=dev, generator=g)
    bi = torch.arange(B, device=dev)[:, None, None].expand(B, C, 2)
    ci = torch.arange(C, device=dev)[None, :, None].expand(B, C, 2)
    x[bi, ci, ys, xs] = 0.9
    bufs.append(x)
for cap in (1024,):
    for it in range(3):
        P.local_peaks_padded(bufs[it % nbuf], 0.2, 5, 2.0, cap)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 30
    ev[0].record()
    for it in range(n):
        P.local_peaks_padded(bufs[it % nbuf], 0.2, 5, 2.0, cap)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    gb = B * C * H * W * 4 / 1e9
    print(json.dumps({"cap": cap, "ms_per_batch": ms, "GBps": gb / (ms / 1e3), "frames_per_s": B / (ms / 1e3)}))
fc, xy, val, chan, status, cap = P.local_peaks_padded(bufs[0], 0.2, 5, 2.0, 1024)
print("counts", fc[:8].tolist(), "status", status.item())
