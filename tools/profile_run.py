"""Short steady-state run of the hot path for ncu: same kernels as bench.py, fewer samples.
Usage: python tools/profile_run.py [spp] [width height] [--scene cornell|config4] [--pipeline persistent|wavefront]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cornelis_b200 import binding, scenes  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
opts = {sys.argv[i][2:]: sys.argv[i + 1] for i in range(1, len(sys.argv) - 1) if sys.argv[i].startswith("--")}
args = [a for a in args if a not in opts.values()]
spp = int(args[0]) if len(args) > 0 else 32
W, H = (int(args[1]), int(args[2])) if len(args) > 2 else (1920, 1080)
which = opts.get("scene", "cornell")
pipeline = {"persistent": binding.PIPELINE_PERSISTENT, "wavefront": binding.PIPELINE_WAVEFRONT}[opts.get("pipeline", "persistent")]
flat = scenes.many_spheres(10000, aspect=H / W) if which == "config4" else scenes.cornell_box(aspect=H / W)
depth = 64 if which == "config4" else 0
scene = binding.Scene(flat)
st = scene.render_accumulate(W, H, spp, max_depth=depth, pipeline=pipeline)
st = scene.render_accumulate(W, H, spp, max_depth=depth, pipeline=pipeline)
print({k: st[k] for k in ("pixel_samples", "rays", "iterations", "kernel_launches", "gpu_ms")},
      "Msamples/s", st["pixel_samples"] / st["gpu_ms"] / 1e3)
