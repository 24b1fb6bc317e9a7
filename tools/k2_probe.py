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
