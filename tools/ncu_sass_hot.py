"""Print the SASS of the first kernel in an .ncu-rep with executed-instruction counts and stall samples
(`ncu --page source --csv`), marking the hottest instructions.  Usage: ncu_sass_hot.py report.ncu-rep [min_share]"""
import csv, subprocess, sys
rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
body = []
for r in rows[2:]:
    if len(r) <= iex or r[0] == "Kernel Name":
        if r and r[0] == "Kernel Name" and body:
            break
        continue
    if r[0] == "Address":
        continue
    body.append((r[ia], r[isrc], int(r[iex] or 0), int(r[ismp] or 0)))
tot = sum(b[2] for b in body) or 1
tots = sum(b[3] for b in body) or 1
print(f"total executed warp-instructions {tot}, samples {tots}")
for i, (a, s, ex, smp) in enumerate(body):
    if ex / tot >= min_share or smp / tots >= min_share:
        print(f"{i:5d} {ex:10d} {100*ex/tot:5.1f}% smp {100*smp/tots:5.1f}%  {s.strip()}")
