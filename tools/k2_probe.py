"""Probe: how much of K2's 16.8 us at cfg2 is the stream itself (all planes below threshold), how much the refinement
epilogue adds, and what a plain copy of the same 85 MB achieves at this size.  Scratch tool; numbers quoted in DESIGN.md."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sleap_nn_b200 import _native as N
from sleap_nn_b200.data.utils import make_grid_vectors
from sleap_nn_b200.data.confidence_maps import _confmaps
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from bench_kernels import timed
dev = torch.device("cuda", 0)
B, Cn, H, W = 256, 13, 80, 80
g = torch.Generator(device=dev).manual_seed(0)
bufs = []
for _ in range(4):
    pts = torch.rand((B, 1, Cn, 2), generator=g, device=dev) * 60 + 10
    xv, yv = make_grid_vectors(H, W, 1)
    cms = _confmaps(pts, xv, yv, 3.0, torch.float32, dev)
    cms += torch.rand(cms.shape, generator=g, device=dev) * 1e-3
    bufs.append(cms)
rpc, nch, nbytes = C.c_int(), C.c_int(), C.c_longlong()
N.check(N.lib.snb_global_peaks_workspace(B, Cn, H, W, C.byref(rpc), C.byref(nch), C.byref(nbytes)), "ws")
ws = torch.zeros(((nbytes.value + 3) // 4,), dtype=torch.int32, device=dev)
pts_o = torch.empty((B, Cn, 2), device=dev); val_o = torch.empty((B, Cn), device=dev)
st = N.stream_ptr(dev)
for thr, ref, label in ((0.2, 5, "refine 5x5"), (0.2, 0, "no refinement"), (5.0, 0, "all below threshold (stream + reductions only)")):
    def fn(i):
        x = bufs[i % 4]
        N.check(N.lib.snb_global_peaks(N.ptr(x), B, Cn, H, W, *x.stride(), thr, ref, N.ptr(ws), N.ptr(pts_o), N.ptr(val_o), N.stream_ptr(dev)), "k2")
    print(label, round(timed(fn, 400) * 1e3, 2), "us")
# a plain streaming read of the same bytes by torch (max over all) for reference
print("torch.amax over the same 85 MB:", round(timed(lambda i: bufs[i % 4].amax(dim=(2, 3)), 200) * 1e3, 2), "us")
big = [torch.empty(85196800 // 4, device=dev) for _ in range(4)]
dst = torch.empty(85196800 // 4, device=dev)
print("torch copy 85 MB (r+w):", round(timed(lambda i: dst.copy_(big[i % 4]), 200) * 1e3, 2), "us")

# ---- per-plane globaltimer stamps (A/B build only: SLEAPNN_B200_LIB=.../libsleapnn_b200_ab.so) ----
if hasattr(N.lib, "snb_ab_k2_stamps"):
    import json
    import numpy as np
    P = B * Cn
    host = (C.c_ulonglong * (3 * P))()
    for wpc in (1, 4):
        os.environ["SNB_K2_WPC"] = str(wpc)
        x = bufs[0]

        def one(thr=0.2, ref=5):
            N.check(N.lib.snb_global_peaks(N.ptr(x), B, Cn, H, W, *x.stride(), thr, ref, N.ptr(ws), N.ptr(pts_o), N.ptr(val_o), N.stream_ptr(dev)), "k2")

        for _ in range(3):
            one()
        N.lib.snb_ab_k2_stamps(1, None, 0)
        big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        big.zero_()  # flush L2, then ONE stamped launch, alone
        torch.cuda.synchronize()
        one()
        N.lib.snb_ab_k2_stamps(0, host, P)
        t = np.frombuffer(host, dtype=np.uint64).reshape(P, 3).astype(np.int64)
        t0 = t[:, 0].min()
        q = lambda a: [round(float(v) / 1e3, 2) for v in np.percentile(a, [0, 10, 50, 90, 100])]
        print(json.dumps({"k2_stamps_us": {"wpc": wpc, "planes": P, "pct": [0, 10, 50, 90, 100],
                                           "start": q(t[:, 0] - t0), "stream_end": q(t[:, 1] - t0), "end": q(t[:, 2] - t0),
                                           "stream_dur": q(t[:, 1] - t[:, 0]), "epilogue_dur": q(t[:, 2] - t[:, 1])}}), flush=True)
