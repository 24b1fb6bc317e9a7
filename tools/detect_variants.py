"""Scratch: time the detect kernel alone for the variant selected by SNB_DETECT_VARIANT, plus streams sweep."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sleap_nn_b200 import _native as NN
dev = torch.device("cuda", 0)
edges, inputs = bench.make_inputs(dev, 4, 100)
B = bench.B
fc = torch.empty(B, dtype=torch.int32, device=dev); keys = torch.empty(B * 256, dtype=torch.int32, device=dev)
def run(i):
    cms = inputs[i % 4][0]
    sb, sc, sh, sw = cms.stride()
    NN.check(NN.lib.snb_local_peaks_detect(NN.ptr(cms), B, 5, 512, 512, sb, sc, sh, sw, 0.2, 256, NN.ptr(fc), NN.ptr(keys), None, None, NN.stream_ptr(dev)), "d")
for i in range(5): run(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 200
a.record()
for i in range(n): run(i)
b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) / n * 1e3
print(os.environ.get("SNB_DETECT_VARIANT", "default"), os.environ.get("SNB_DETECT_BULK", ""), f"{us:.1f} us  {B*5*512*512*4/us/1e3:.0f} GB/s  counts ok={int(fc.sum())}")
