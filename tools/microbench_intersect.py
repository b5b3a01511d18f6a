"""Config 3 (BASELINE.json configs[2]): 2^24 random rays against 1024 spheres + 6 planes, device-resident float4 rays,
the k_intersect_batch kernel through cornelis_cuda_intersect_device.  Prints one JSON line.
torch is only the device allocator here."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cornelis_b200 import binding, scenes  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24)
n_spheres = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
flat = scenes.microbench_scene(n_spheres)
scene = binding.Scene(flat)
# the microbench measures the reference's own strategy — every primitive for every ray from shared memory — against
# the FP32 roofline; the grid (what a render of this scene would use) is timed beside it
scene.set_acceleration(binding.ACCEL_NONE)
org, dirs = scenes.microbench_rays(n)
pad = np.zeros((n, 1), np.float32)
d_org = torch.from_numpy(np.concatenate([org, pad], 1)).cuda()
d_dir = torch.from_numpy(np.concatenate([dirs, pad], 1)).cuda()
d_hit = torch.empty((n, 2), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
scene.intersect_device(n, d_org.data_ptr(), d_dir.data_ptr(), d_hit.data_ptr(), repeats=3)  # warm-up
ms = scene.intersect_device(n, d_org.data_ptr(), d_dir.data_ptr(), d_hit.data_ptr(), repeats=10)
hits = d_hit.cpu().numpy()
prim = hits[:, 1].view(np.int32)
scene.set_acceleration(binding.ACCEL_GRID)
d_hit_grid = torch.empty_like(d_hit)
scene.intersect_device(n, d_org.data_ptr(), d_dir.data_ptr(), d_hit_grid.data_ptr(), repeats=3)
ms_grid = scene.intersect_device(n, d_org.data_ptr(), d_dir.data_ptr(), d_hit_grid.data_ptr(), repeats=10)
same = bool(torch.equal(d_hit.view(torch.int32), d_hit_grid.view(torch.int32)))
flop = n * (26.0 * n_spheres + 35.0 * 6)  # 4.502e11 for the named size
peak = 148 * 128 * 1.965e9
print(json.dumps({
    "workload": f"2^{int(np.log2(n))} rays x ({n_spheres} spheres + 6 planes)", "ms_per_launch": ms,
    "grays_per_s": n / ms / 1e6, "gtests_per_s": n * (n_spheres + 6) / ms / 1e6,
    "algorithmic_tflops": flop / ms / 1e9, "fp32_peak_tflops_nonfma": peak / 1e12,
    "frac_of_fp32_peak": flop / (ms * 1e-3) / peak, "sphere_hit_fraction": float((prim < n_spheres).mean()),
    "bytes_per_ray": 40, "hbm_gbs": 40.0 * n / ms / 1e6,
    "grid": {"ms_per_launch": ms_grid, "grays_per_s": n / ms_grid / 1e6, "bit_identical_to_exhaustive": same,
             **scene.acceleration()},
}))
