set -x
SNB_LATENCY_NO_CPU=1 SNB_LATENCY_CALLS=40 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bottomup_tail|local_peaks_detect" -c 700 --csv --log-file gpurun_out/r2_tail_durations.csv python tools/latency_small_batch.py > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2_tail_durations.csv
python - <<'PY'
import csv, collections, statistics
rows=[r for r in csv.reader(open('gpurun_out/r2_tail_durations.csv')) if len(r)>10 and r[0].isdigit()]
by=collections.OrderedDict()
for r in rows:
    key=(r[4].split('(')[0][-40:], r[8], r[7])
    by.setdefault(key,[]).append(float(r[14].replace(',','')))
for k,v in by.items(): print(k, len(v), 'median_us', statistics.median(v)/1e3)
PY
