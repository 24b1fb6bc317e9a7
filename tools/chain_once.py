"""Scratch: run the fused chain a few times at cfg3 size (for ncu captures). Not part of the product."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sleap_nn_b200.pipeline import BottomUpPostproc
dev = torch.device("cuda", 0)
edges, inputs = bench.make_inputs(dev, 2, 100)
pipe = BottomUpPostproc(bench.N_NODES, edges, bench.B, (512, 512), cms_stride=2, pafs_stride=2, device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for i in range(n):
    pipe(*inputs[i % 2])
torch.cuda.synchronize()
print("done")
