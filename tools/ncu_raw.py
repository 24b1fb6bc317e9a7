"""Print selected metrics of an .ncu-rep (via `ncu -i ... --page raw --csv`), one block per launch."""
import csv, subprocess, sys
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size',
        'launch__block_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'launch__occupancy_limit_blocks', 'smsp__inst_executed.sum',
        'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio' ]
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    for w in WANT + extra:
        for i, h in enumerate(hdr):
            if h == w or (w in extra and w in h):
                print(f"{h:75s} {r[i][:70]} {units[i]}")
    print('---')
