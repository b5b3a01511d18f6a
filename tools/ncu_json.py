"""Turns one kernel's row of an `ncu --set full --page raw --csv` dump into the JSON bench.py reads for the roofline's
`traffic` and `issue_slots` (profiles/r2_ncu/<kernel>.json).  The JSON records the git head and the hash of
cornelis_b200/csrc it was taken from, so bench.py can tell when the figures no longer describe the running kernels.

    ncu -i X.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_json.py raw.csv k_persistent_queued --units <rays of the launch> --command "<what ran>" > out.json
"""
import argparse
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from cornelis_b200 import build  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("raw")
ap.add_argument("kernel")
ap.add_argument("--units", type=float, default=0.0, help="work items (rays) the captured launch processed")
ap.add_argument("--command", default="")
args = ap.parse_args()

rows = list(csv.reader(open(args.raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
row = next(r for r in rows[2:] if args.kernel in r[idx["Kernel Name"]])


def val(name, scale=None):
    if name not in idx:
        return None
    x = float(row[idx[name]].replace(",", ""))
    u = units[idx[name]]
    if scale == "bytes":
        x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    if scale == "ms":
        x *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1)
    return x


try:
    head = subprocess.run(["git", "-C", str(ROOT), "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
except OSError:
    head = ""
inst = val("smsp__inst_executed.sum")
out = {
    "kernel": row[idx["Kernel Name"]].split("(")[0], "command": args.command, "git_head": head,
    "csrc_sha256": build.csrc_hash(),
    "launch": {"duration_ms": val("gpu__time_duration.sum", "ms"), "grid": val("launch__grid_size"),
               "registers": val("launch__registers_per_thread"), "units": args.units or None},
    "dram_bytes": (val("dram__bytes_read.sum", "bytes") or 0) + (val("dram__bytes_write.sum", "bytes") or 0),
    "dram_bytes_read": val("dram__bytes_read.sum", "bytes"), "dram_bytes_write": val("dram__bytes_write.sum", "bytes"),
    "issue": {
        "busy_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_instructions": inst,
        "warp_instructions_per_unit": inst / args.units if inst and args.units else None,
        "lanes_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "no_instruction_stall_per_issue": val("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
        "pipe_fma_pct": val("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "pipe_alu_pct": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
    },
    "l2_hit_pct": val("lts__t_sector_hit_rate.pct"),
}
print(json.dumps(out, indent=1))
