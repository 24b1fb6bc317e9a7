set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"confmaps|pafs" -s 20 -c 12 --csv --log-file gpurun_out/r2_h_targets_launches.csv python tools/bench_kernels.py --iters 6 --only k7_cfg4,k7_cfg4_bf16,k7_cfg4_g1,k7_cfg4_g1_bf16,k8_cfg4,k8_cfg4_g8,targets_cfg4_fused > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_h_targets_launches.csv')))
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        print(d['Kernel Name'][:60], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
PY
