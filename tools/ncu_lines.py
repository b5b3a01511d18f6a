"""Per-source-line summary of an ncu report's source page (needs -lineinfo):
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_cs.csv ; python tools/ncu_lines.py X_cs.csv [top]
Aggregates executed warp instructions, thread utilisation and stall samples per (file, line)."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None
hdr = None
agg = defaultdict(lambda: [0, 0, 0, ""])  # inst, thread inst, samples, text
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        hdr = None
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        continue
    if cur_file and hdr is None and r and r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if not hdr or len(r) < len(hdr):
        continue
    try:
        inst = int(r[ix["Instructions Executed"]])
        tinst = int(r[ix["Thread Instructions Executed"]])
        samp = int(r[ix["# Samples"]])
    except ValueError:
        continue
    if not r[0].strip():
        continue  # SASS rows under a source line: already included in the line's own row
    a = agg[(cur_file, r[0])]
    a[0] += inst
    a[1] += tinst
    a[2] += samp
    if r[1].strip():
        a[3] = r[1].strip()[:90]
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[2] for a in agg.values()) or 1
print(f"total warp instructions {tot_i}, samples {tot_s}")
by_file = defaultdict(lambda: [0, 0])
for (f, _), a in agg.items():
    by_file[f][0] += a[0]
    by_file[f][1] += a[2]
for f, (i, s) in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:24s} inst {100 * i / tot_i:5.1f}%  samples {100 * s / tot_s:5.1f}%")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    lanes = a[1] / a[0] if a[0] else 0
    print(f"{100 * a[0] / tot_i:5.1f}% inst {100 * a[2] / tot_s:5.1f}% samp lanes {lanes:4.1f}  {f}:{ln}  {a[3]}")
