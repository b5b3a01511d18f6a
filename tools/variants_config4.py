"""Runs tools/bench_config4.py (wavefront pipeline only, short) once per A/B build in cornelis_b200/lib/variants/."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for so in sorted((ROOT / "cornelis_b200" / "lib" / "variants").glob("*.so")):
    env = dict(os.environ, CORNELIS_CUDA_LIB=str(so))
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "bench_config4.py"), "--no-cpu", "--no-exhaustive", "--spp",
                        "256", *sys.argv[1:]], env=env, capture_output=True, text=True)
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        w = j["wavefront"]
        print(f"{so.stem:16s} wavefront {w['msamples_per_s']:8.1f} Msamples/s  intersect {w['stage_ms_per_pass']['intersect_ms']:.3f} ms/pass"
              f"  persistent {j['persistent']['msamples_per_s']:8.1f}", flush=True)
    except Exception:
        print(so.stem, "FAILED", r.stdout[-300:], r.stderr[-300:], flush=True)
