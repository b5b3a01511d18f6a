"""Instruction mix of one kernel from `ncu --page source --csv`: warp-level instructions executed per opcode.
Usage: python tools/ncu_sass_mix.py src.csv [items]   (items = work items the launch processed, for per-item counts)"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if len(starts) > 1:  # several launches in one dump: keep the first
    rows = rows[: starts[1]]
hdr = rows[1]
si, ei, ti = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
items = float(sys.argv[2]) if len(sys.argv) > 2 else None
mix = collections.Counter()
thr = collections.Counter()
total = 0
for r in rows[2:]:
    if len(r) <= ti:
        continue
    sass = r[si].strip()
    parts = sass.split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("MUFU") and "." in op else "")
    n = int(r[ei])
    mix[op] += n
    thr[op] += int(r[ti])
    total += n
print(f"static SASS lines {len(rows) - 2}, warp instructions executed {total}")
for op, n in mix.most_common(40):
    extra = f"  {32 * n / items:8.1f} /item(warp-wide)  {thr[op] / items:8.1f} /item(active)" if items else ""
    print(f"{op:14s} {n:12d} {100 * n / total:6.2f}%  lanes {thr[op] / max(n, 1):5.1f}{extra}")
