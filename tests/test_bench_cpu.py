"""CPU checks of bench.py's reference arm (`--impl reference`): it runs without a GPU, prints ONE JSON line with the
contract's keys, times the UNMODIFIED reference files when they are available (staged under baseline/_ref by build(),
or live under /root/reference), emits exactly the config object the GPU arm emits, and never maps the product's .so."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    env = dict(os.environ, LD_DEBUG="files")  # the dynamic loader lists every shared object the process maps (stderr)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, cwd=ROOT, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    have_ref = any(os.path.isfile(os.path.join(p, "sleap_nn", "inference", "ops", "paf.py"))
                   for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"))
    assert cb["kind"] == ("reference" if have_ref else "port"), cb
    assert cb["frames_per_step"] == 64 and cb["instances_found_last_step"] == 128
    assert cb["one_thread"]["value"] > 0 and cb["cores"] >= 1
    if have_ref:
        assert set(cb["stage_seconds_per_step"]) == {"find_local_peaks+split", "score_paf_lines", "match_candidates",
                                                     "group_instances"}
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.shared_config(1), "both arms must print the same config object"
    assert "libsleapnn_b200" not in r.stderr, "the reference arm must not map the product's shared library"


def test_committed_traffic_record_matches_the_source():
    """`roofline.traffic` is read from profiles/traffic.json, keyed by kernel + the hash of csrc/peaks.cu at capture time:
    the committed record must belong to the committed source (re-capture with tools/final_run.sh + tools/ncu_traffic.py
    after editing the kernel)."""
    sys.path.insert(0, ROOT)
    import bench

    for kernel, lo, hi in (("local_peaks_detect_vec<float,4,1,6,1>", 335.5e6, 345e6),
                           ("local_peaks_detect_vec<__half,2,1,8,1>", 167.7e6, 175e6),
                           ("local_peaks_detect_vec<__nv_bfloat16,2,1,8,1>", 167.7e6, 175e6)):
        traffic, rec = bench.recorded_traffic(kernel)
        assert rec is not None and rec["source_matches"], kernel
        assert lo <= traffic <= hi, (kernel, traffic)  # = the algorithmic bytes (+ the key / counter writes)
