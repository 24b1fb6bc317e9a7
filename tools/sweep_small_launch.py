#!/usr/bin/env python
"""Launch-shape sweeps for the short (one-to-three-wave) launches: K2 at cfg2 batch 256, K7 / K8 / the PDL pair on ONE frame.

Needs the A/B build (its getenv switches are compiled out of the product library):

    SNB_NVCC_EXTRA=-DSNB_AB_VARIANTS SNB_LIB_NAME=libsleapnn_b200_ab.so bash sleap_nn_b200/csrc/build.sh
    SLEAPNN_B200_LIB=sleap_nn_b200/lib/libsleapnn_b200_ab.so python tools/sweep_small_launch.py > gpurun_out/sweep.jsonl

One process: the A/B library re-reads SNB_K2_WPC / SNB_K2_U / SNB_PAF_RPB / SNB_K7_ROWS_PER_BAND at every launch, and
`bench_kernels.timed` captures its launches into a graph right after the variable is set.  Also prints what a plain
cudaMemsetAsync / torch reduction reaches at the same byte counts: the practical ceiling of a launch that short.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import bench_kernels as bk  # noqa: E402

PEAK = bk.peak_gbs()[0]


def out(tag, knobs, d):
    print(json.dumps({"sweep": tag, **knobs, "us": round(d["avg_launch_ms"] * 1e3, 2), "frac": round(d["frac"], 4)}), flush=True)


def with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    try:
        for k, v in env.items():
            os.environ[k] = str(v)
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    bk.QUIET = True
    iters = 300
    only = set(sys.argv[1:])

    def want(name):
        return not only or name in only

    if want("ceil"):
        for mb, nbytes in (("k7_g1_bf16", 16777216), ("k7_g1", 33554432), ("k8_g1", 65011712), ("k2_cfg2", 85196800),
                           ("targets_g1", 98566144)):
            n = max(3, -(-400_000_000 // nbytes))
            bufs = [torch.empty((nbytes // 4,), dtype=torch.float32, device=dev) for _ in range(n)]
            ms = bk.timed(lambda i: bufs[i % n].zero_(), iters)
            print(json.dumps({"sweep": "memset", "like": mb, "bytes": nbytes, "us": round(ms * 1e3, 2),
                              "frac": round(nbytes / (ms * 1e-3) / 1e9 / PEAK, 4)}), flush=True)
            ms = bk.timed(lambda i: bufs[i % n].amax(), iters)
            print(json.dumps({"sweep": "torch_amax_2_kernels", "like": mb, "bytes": nbytes, "us": round(ms * 1e3, 2),
                              "frac": round(nbytes / (ms * 1e-3) / 1e9 / PEAK, 4)}), flush=True)
            del bufs
    if want("k2"):
        for wpc in (1, 2, 4, 8):
            for u in (6, 8, 12):
                env = {"SNB_K2_WPC": wpc, "SNB_K2_U": u}
                out("k2_cfg2", env, with_env(env, lambda: bk.k2_cfg2(dev, iters)))
        for wpc in (1, 4):
            for u in (8, 12):
                env = {"SNB_K2_WPC": wpc, "SNB_K2_U": u}
                out("k2_cfg2_f16", env, with_env(env, lambda: bk.k2_cfg2(dev, iters, dtype=torch.float16)))
                out("k2_cfg2_b1024", env, with_env(env, lambda: bk.k2_cfg2(dev, iters, B=1024)))
    if want("k8"):
        for rpb in (8, 11, 13, 16, 19, 24, 32, 37, 43):
            env = {"SNB_PAF_RPB": rpb}
            out("k8_cfg4_g1", env, with_env(env, lambda: bk.k8_cfg4(dev, iters, G=1)))
        for rpb in (16, 19, 37):
            env = {"SNB_PAF_RPB": rpb}
            out("k8_cfg4_g1_bf16", env, with_env(env, lambda: bk.k8_cfg4(dev, iters, bf16=True, G=1)))
            out("k8_cfg4_g8", env, with_env(env, lambda: bk.k8_cfg4(dev, iters, G=8)))
    if want("k7"):
        for rpb in (16, 24, 32, 48, 64):
            env = {"SNB_K7_ROWS_PER_BAND": rpb}
            out("k7_cfg4_g1", env, with_env(env, lambda: bk.k7_cfg4(dev, iters, G=1)))
    if want("k2u"):
        for rep in range(2):
            for u in (4, 5, 6, 8, 10):
                env = {"SNB_K2_WPC": 1, "SNB_K2_U": u}
                out("k2_cfg2", {**env, "rep": rep}, with_env(env, lambda: bk.k2_cfg2(dev, iters)))
                out("k2_cfg2_f16", {**env, "rep": rep}, with_env(env, lambda: bk.k2_cfg2(dev, iters, dtype=torch.float16)))
                if rep == 0:
                    out("k2_cfg2_b1024", env, with_env(env, lambda: bk.k2_cfg2(dev, iters, B=1024)))
                    out("k2_cfg2_b64", env, with_env(env, lambda: bk.k2_cfg2(dev, iters, B=64)))
    if want("k7g"):
        for bf16 in (False, True):
            for G in (1, 2, 4, 8):
                for rpb in (16, 32, 64):
                    env = {"SNB_K7_ROWS_PER_BAND": rpb}
                    out("k7_cfg4" + ("_bf16" if bf16 else ""), {**env, "G": G}, with_env(env, lambda: bk.k7_cfg4(dev, iters, bf16=bf16, G=G)))
    if want("k8g"):  # product sizing (wave fit) at several launch sizes, fp32 and bf16
        for bf16 in (False, True):
            for G in (1, 2, 4, 8):
                out("k8_cfg4" + ("_bf16" if bf16 else ""), {"G": G, "rows": "wave_fit"}, bk.k8_cfg4(dev, iters, bf16=bf16, G=G))
                env = {"SNB_PAF_RPB": 32}
                out("k8_cfg4" + ("_bf16" if bf16 else ""), {**env, "G": G}, with_env(env, lambda: bk.k8_cfg4(dev, iters, bf16=bf16, G=G)))
    if want("prod"):  # the library's own choices (no switch set)
        for G in (1, 2, 4, 8):
            for bf16 in (False, True):
                out("k7_cfg4" + ("_bf16" if bf16 else ""), {"G": G, "rows": "product"}, bk.k7_cfg4(dev, iters, bf16=bf16, G=G))
        for bf16 in (False, True):
            for G in (1, 8):
                out("targets_cfg4_fused" + ("_bf16" if bf16 else ""), {"G": G, "rows": "product"},
                    bk.targets_cfg4_fused(dev, iters, bf16=bf16, G=G))
        for B in (64, 256, 1024):
            for dt in (torch.float32, torch.float16):
                out("k2_cfg2_" + str(dt).split(".")[-1], {"B": B, "loads_in_flight": "product"}, bk.k2_cfg2(dev, iters, B=B, dtype=dt))
    if want("pair"):
        for r7 in (16, 32, 64):
            for r8 in (16, 19, 37):
                env = {"SNB_K7_ROWS_PER_BAND": r7, "SNB_PAF_RPB": r8}
                out("targets_cfg4_fused_g1", env, with_env(env, lambda: bk.targets_cfg4_fused(dev, iters, G=1)))


if __name__ == "__main__":
    main()
