"""One line per kernel of an `ncu --set full --page raw --csv` dump: duration, DRAM bytes and achieved bandwidth against
the measured copy bandwidth (MEASURED_PEAKS.json), issue-slot and FP32-pipe utilisation, lanes per instruction.
    python tools/ncu_stage_table.py raw.csv [more.csv ...]"""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
peaks = ROOT / "MEASURED_PEAKS.json"
hbm = float(json.loads(peaks.read_text())["hbm_gbs"]) if peaks.exists() else 6542.7
SCALE_B = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
SCALE_T = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
print(f"# HBM peak {hbm} GB/s (measured copy bandwidth)")
print(f"{'kernel':34s} {'us':>9s} {'DRAM MB':>9s} {'GB/s':>8s} {'of HBM':>7s} {'issue%':>7s} {'fma%':>6s} {'alu%':>6s} {'lanes':>6s} {'regs':>5s}")
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, name, table=None):
        if name not in ix:
            return float("nan")
        x = float(r[ix[name]].replace(",", ""))
        return x * table.get(units[ix[name]], 1) if table else x

    seen = set()
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        if name in seen:
            continue
        seen.add(name)
        t = val(r, "gpu__time_duration.sum", SCALE_T)
        b = val(r, "dram__bytes_read.sum", SCALE_B) + val(r, "dram__bytes_write.sum", SCALE_B)
        gbs = b / t / 1e9 if t > 0 else 0.0
        print(f"{name[:34]:34s} {t * 1e6:9.1f} {b / 1e6:9.1f} {gbs:8.0f} {gbs / hbm:7.2f} "
              f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):7.1f} "
              f"{val(r, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):6.1f} "
              f"{val(r, 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):6.1f} "
              f"{val(r, 'smsp__thread_inst_executed_per_inst_executed.ratio'):6.1f} "
              f"{val(r, 'launch__registers_per_thread'):5.0f}")
