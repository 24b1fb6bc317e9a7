"""Summarise an ncu --csv launch list: per kernel count, mean duration, share, DRAM bytes."""
import csv, collections, sys
path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
by = collections.OrderedDict()
for r in rows:
    by.setdefault(int(r[0]), {"name": r[4].split("(")[0][-60:]})[r[12]] = float(r[14].replace(",", ""))
agg = collections.OrderedDict()
for i, m in by.items():
    a = agg.setdefault(m["name"], {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
    a["n"] += 1
    a["ns"] += m.get("gpu__time_duration.sum", 0)
    a["rd"] += m.get("dram__bytes_read.sum", 0)
    a["wr"] += m.get("dram__bytes_write.sum", 0)
tot = sum(a["ns"] for a in agg.values())
print(f"{'kernel':62s} {'n':>4s} {'mean_us':>9s} {'share':>7s} {'rd_MB/launch':>13s} {'wr_MB/launch':>13s}")
for k, a in agg.items():
    print(f"{k:62s} {a['n']:4d} {a['ns']/a['n']/1e3:9.2f} {a['ns']/tot*100:6.1f}% {a['rd']/a['n']/1e6:13.3f} {a['wr']/a['n']/1e6:13.3f}")
