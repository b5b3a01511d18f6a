"""Short steady-state run of the hot path for ncu: Cornell 1080p at a few spp (same kernels, same pool size as
bench.py, fewer passes).  Usage: python tools/profile_run.py [spp] [width height]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cornelis_b200 import binding, scenes  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
scene = binding.Scene(scenes.cornell_box(aspect=H / W))
st = scene.render_accumulate(W, H, spp)
st = scene.render_accumulate(W, H, spp)
print({k: st[k] for k in ("pixel_samples", "rays", "iterations", "kernel_launches", "gpu_ms")},
      "Msamples/s", st["pixel_samples"] / st["gpu_ms"] / 1e3)
