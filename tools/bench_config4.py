"""Config 4 (BASELINE.json configs[3]): 10 000 random spheres over a 4000x4000 floor, 64 mixed Oren-Nayar / glossy
materials, 1920x1080 at 1024 spp, max depth 64 — rendered through the uniform grid.  Prints one JSON line with
pixel-samples/s and rays/s for both pipelines, the exhaustive-scan kernel on a bounded sample beside it, and the
reference's CPU loop on a bounded sample of the same frame, timed by `bench.py --impl reference --workload spheres`.

    python tools/bench_config4.py [--spp 1024] [--width 1920 --height 1080] [--no-cpu]
"""
import argparse
import json
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cornelis_b200 import binding, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--spp", type=int, default=1024)
ap.add_argument("--spheres", type=int, default=10000)
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--no-exhaustive", action="store_true")
ap.add_argument("--no-persistent", action="store_true")
args = ap.parse_args()

W, H = args.width, args.height
flat = scenes.many_spheres(args.spheres, aspect=H / W)
scene = binding.Scene(flat)
out = {"workload": f"{args.spheres} spheres + floor, 64 materials, {W}x{H}, {args.spp} spp, max depth 64",
       "acceleration": scene.acceleration()}
scene.render_accumulate(W, H, 4, max_depth=64)  # warm-up
for name, pipeline in (("persistent", binding.PIPELINE_PERSISTENT), ("wavefront", binding.PIPELINE_WAVEFRONT)):
    if name == "persistent" and args.no_persistent:
        continue
    st = scene.render_accumulate(W, H, args.spp, max_depth=64, pipeline=pipeline, stage_timing=True)
    out[name] = {"stage_ms_per_pass": {k: st[k] for k in ("raygen_ms", "intersect_ms", "shade_ms", "accumulate_ms")},"msamples_per_s": st["pixel_samples"] / st["gpu_ms"] / 1e3, "mrays_per_s": st["rays"] / st["gpu_ms"] / 1e3,
                 "gpu_ms": st["gpu_ms"], "rays_per_sample": st["rays"] / st["pixel_samples"], "max_depth": st["max_depth"],
                 "kernel_launches": st["kernel_launches"]}
if not args.no_exhaustive:
    # the reference's own strategy on the GPU (every sphere for every ray, tables in shared memory), bounded sample
    scene.set_acceleration(binding.ACCEL_NONE)
    spp_small = max(1, args.spp // 256)
    st = scene.render_accumulate(W, H, spp_small, max_depth=64)
    out["exhaustive_scan"] = {"msamples_per_s": st["pixel_samples"] / st["gpu_ms"] / 1e3,
                              "mrays_per_s": st["rays"] / st["gpu_ms"] / 1e3, "sample": f"{spp_small} spp"}
    scene.set_acceleration(binding.ACCEL_AUTO)
if not args.no_cpu:
    # the reference's CPU loop on a bounded sample of the same frame: bench.py's reference arm (the one place besides
    # the tests that may run the oracle), on a tenth of the frame in each dimension
    w, h = W // 10, H // 10
    r = subprocess.run([sys.executable, str(Path(__file__).resolve().parents[1] / "bench.py"), "--impl", "reference",
                        "--workload", "spheres", "--width", str(w), "--height", str(h), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True)
    try:
        line = json.loads(r.stdout.strip().splitlines()[-1])
        out["cpu_baseline"] = {"msamples_per_s": line["value"], "mrays_per_s": line["mrays_per_s"],
                               **{k: line["cpu_baseline"][k] for k in ("kind", "cores", "sample")}}
    except Exception:
        out["cpu_baseline"] = {"error": (r.stdout + r.stderr)[-300:]}
print(json.dumps(out))
