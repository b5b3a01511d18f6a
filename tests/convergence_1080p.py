"""One-off evidence run (not collected by pytest): BASELINE.json configs[1] at FULL size on both sides.

The reference's CPU loop (oracle/_ref) renders the Cornell box at 1920x1080, aspect 0.5625, --ref-spp samples per
pixel on all host threads (about nine minutes at 4096 spp on the GPU box's 16 threads); the GPU renders the same frame
at 4096 spp; the report holds the per-pixel 3-sigma fraction, RMSE, relative RMSE, energy ratio and z-score statistics.

    python tests/convergence_1080p.py [--ref-spp 4096] [--out profiles/.../convergence.json]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from cornelis_b200 import binding, scenes  # noqa: E402
from oracle import loader  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ref-spp", type=int, default=4096)
ap.add_argument("--spp", type=int, default=4096)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--out", default="")
args = ap.parse_args()
W, H = args.width, args.height
flat = scenes.cornell_box(aspect=H / W)

sc = binding.Scene(flat)
sc.render_accumulate(W, H, 16)
st = sc.render_accumulate(W, H, args.spp, variance=True)
mean, var = sc.resolve(args.spp, variance=True)
gpu_s = st["gpu_ms"] * 1e-3

ora = loader.best()
t0 = time.time()
ref = ora.scene(flat).render(W, H, args.ref_spp, tile=(40, 40), variance=True, stats=True)
cpu_s = time.time() - t0

sigma = np.sqrt(ref["variance"].astype(np.float64) / args.ref_spp + var.astype(np.float64) / args.spp)
diff = mean.astype(np.float64) - ref["mean"].astype(np.float64)
finite = np.isfinite(mean).all(axis=2) & np.isfinite(ref["mean"]).all(axis=2)
ok = np.abs(diff) <= 3.0 * sigma + 1e-4 * (1.0 + np.abs(ref["mean"]))
z = diff[finite] / np.maximum(sigma[finite], 1e-12)
z = z[sigma[finite] > 1e-6]
rmse = float(np.sqrt(np.mean(diff[finite] ** 2)))
report = {
    "workload": f"Cornell {W}x{H}, aspect {H / W:.4f}, GPU {args.spp} spp vs reference {args.ref_spp} spp (40x40 tiles)",
    "oracle": ora.kind, "host_threads": os.cpu_count(),
    "three_sigma_fraction": float(ok[finite].mean()), "pixels_not_finite": int((~finite).sum()),
    "rmse": rmse, "relative_rmse": rmse / float(np.sqrt(np.mean(ref["mean"][finite].astype(np.float64) ** 2))),
    "energy_ratio": float(mean[finite].mean() / ref["mean"][finite].mean()),
    "z_mean": float(z.mean()), "z_std": float(z.std()), "z_abs_gt_4": float((np.abs(z) > 4).mean()),
    "rays_per_sample_gpu": st["rays"] / st["pixel_samples"],
    "rays_per_sample_reference": ref["stats"]["rays"] / ref["stats"]["pixel_samples"],
    "gpu_seconds": gpu_s, "gpu_msamples_per_s": st["pixel_samples"] / gpu_s / 1e6,
    "cpu_seconds": cpu_s, "cpu_msamples_per_s": ref["stats"]["pixel_samples"] / cpu_s / 1e6,
    "speedup": cpu_s / gpu_s,
}
print(json.dumps(report))
if args.out:
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(report, indent=1) + "\n")
