#!/usr/bin/env python
"""A/B of the streaming detect kernel's launch shapes and of the cp.async.bulk (1-D TMA) ring variant.

Needs the A/B build of the library (the product build compiles the variants and their getenv switches out):

    SNB_LIB_NAME=libsleapnn_b200_ab.so SNB_NVCC_EXTRA=-DSNB_AB_VARIANTS bash sleap_nn_b200/csrc/build.sh
    SLEAPNN_B200_LIB=sleap_nn_b200/lib/libsleapnn_b200_ab.so [SNB_DETECT_VARIANT=k | SNB_DETECT_BULK=1 | SNB_DETECT_TMA=1] \
        python tools/detect_variants.py [f32|f16|bf16]

Times the kernel alone on the bench's cfg3 batches (real blobs, rotating 6 batches > L2), CUDA events around 200 launches.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from sleap_nn_b200 import _native as NN  # noqa: E402

dt = bench.DTYPES[sys.argv[1] if len(sys.argv) > 1 else "f32"]
dev = torch.device("cuda", 0)
edges, inputs = bench.make_inputs(dev, 6, 100, dtype=dt)
B = bench.B
fc = torch.empty(B, dtype=torch.int32, device=dev)
keys = torch.empty(B * 256, dtype=torch.int32, device=dev)


def run(i):
    cms = inputs[i % 6][0]
    NN.check(NN.lib.snb_local_peaks_detect_t(NN.ptr(cms), NN.dtype_code(cms.dtype), B, 5, 512, 512, *cms.stride(), 0.2, 256,
                                             NN.ptr(fc), NN.ptr(keys), None, None, NN.stream_ptr(dev)), "detect")


for i in range(6):
    run(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 200
a.record()
for i in range(n):
    run(i)
b.record()
torch.cuda.synchronize()
us = a.elapsed_time(b) / n * 1e3
nbytes = B * 5 * 512 * 512 * inputs[0][0].element_size()
print(json.dumps({"dtype": str(dt), "variant": os.environ.get("SNB_DETECT_VARIANT", "default"),
                  "bulk_ring_single_thread_refill": bool(os.environ.get("SNB_DETECT_BULK")),
                  "tma_ring_producer_consumer": bool(os.environ.get("SNB_DETECT_TMA")),
                  "tma_ctas_per_sm": os.environ.get("SNB_DETECT_TMA_CTAS", "3") if os.environ.get("SNB_DETECT_TMA") else None,
                  "us_per_launch_incl_memset": us,
                  "GBps": nbytes / us / 1e3, "peaks_last_batch": int(fc.sum())}))
