#!/usr/bin/env python
"""Headline benchmark: pixel-samples/s (and rays/s) of the batched render loop on the Cornell scene at 1080p.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path through the C-ABI
    python bench.py --impl reference [...]                          the reference's own CPU loop (oracle/_ref)

A "step" is one full render of the workload: BASELINE.json configs[1], the reference CLI's Cornell scene at
1920x1080 (camera aspect 0.5625 for square pixels), 4096 samples per pixel, Russian roulette, no depth cap.
With N GPUs the 4096 samples of every pixel are split by global sample index (rank r renders [r*4096/N,
(r+1)*4096/N)): total work is fixed — STRONG scaling, the north star's multi-GPU case — and the only exchange is one
NCCL all-reduce of the float4 accumulation images at the end of the step, issued by the library itself
(cornelis_cuda_allreduce_framebuffers); torch.distributed only carries the communicator id, the barriers and the
max-over-ranks of the timings.  The weak-scaling figure (every rank renders 4096 spp) is reported beside it under
"weak"; `--scaling weak` makes it the headline.

At N = 1 the line also carries, under "configs", BASELINE.json configs[2] (the 2^24-ray intersection microbench) and
configs[3] (the 10 000-sphere scene) with their own CPU baselines; at N > 1 configs[4] (3840x2160 at 16 384 spp split
over the N GPUs, one timed render) and, under "single_process", a render of the
same frame by ONE process driving all N GPUs through the C++ RenderSession (RenderOptions::devices = N, the CLI),
compared with its own 1-GPU render.

Prints ONE JSON line (rank 0).  See the task contract for the meaning of every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SM_COUNT = 148
FP32_LANES_PER_SM = 128
SM_MAX_GHZ = 1.965
# algorithmic work of the two compute kernels (reference source after trivial CSE)
FLOP_PER_SPHERE_TEST = 26
FLOP_PER_PLANE_TEST = 35
FLOP_PER_RAY_CORNELL = 4 * FLOP_PER_SPHERE_TEST + 5 * FLOP_PER_PLANE_TEST  # 279
FLOP_PER_SURVIVING_BOUNCE = 250
FLOP_PER_KILLED_HIT = 12
# Where the committed measurements that feed the roofline live (written by tools/ubench/fp32_peak.cu and
# tools/ncu_json.py on the GPU box; each carries the hash of cornelis_b200/csrc it was taken from)
FP32_PEAK_JSON = ROOT / "profiles" / "r2_peaks" / "fp32_peak.json"
NCU_JSON = {"k_persistent_queued": ROOT / "profiles" / "r2_ncu" / "k_persistent_queued.json"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=4096)
    ap.add_argument("--scaling", choices=["weak", "strong"], default="strong",
                    help="strong (default): --spp is the whole image's sample count, split across the GPUs; "
                         "weak: every GPU renders --spp samples")
    ap.add_argument("--pool", type=int, default=0, help="paths in flight per GPU (wavefront pipeline; 0 = default)")
    ap.add_argument("--pipeline", choices=["default", "wavefront", "persistent"], default="default")
    ap.add_argument("--workload", choices=["cornell", "spheres"], default="cornell",
                    help="spheres = BASELINE.json configs[3] (10 000 spheres) as the timed workload of the reference "
                         "arm; our arm reports it under configs.c4")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip configs[2] / configs[3] (N = 1)")
    ap.add_argument("--no-single-process", action="store_true", help="skip the devices = N RenderSession check (N > 1)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def load_peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        p = json.loads(path.read_text())
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_fp32_peak():
    """The FP32 ceiling of the parity build: one IEEE operation per instruction (--fmad=false), so the peak is the rate
    of scalar FMUL / FADD instructions.  Measured on this pool's B200 by tools/ubench/fp32_peak.cu (an alternating
    FMUL / FADD stream); FP32 is not in MEASURED_PEAKS.json.  The computed figure is reported beside it."""
    computed = SM_COUNT * FP32_LANES_PER_SM * SM_MAX_GHZ / 1e3  # 37.2 T lane-ops/s
    if FP32_PEAK_JSON.exists():
        d = json.loads(FP32_PEAK_JSON.read_text())
        m = d["modes"]["fmul_fadd_mix"]
        return {"peak": float(m["t_lane_ops_per_s"]), "peak_source": "measured",
                "peak_detail": f"tools/ubench/fp32_peak.cu on {d.get('gpu', 'B200')}: alternating FMUL/FADD, "
                               f"{m['warp_inst_per_clk_per_sm']:.2f} warp instructions / clk / SM at "
                               f"{d.get('sm_max_mhz', 1965.0):.0f} MHz ({FP32_PEAK_JSON.relative_to(ROOT)})",
                "peak_computed": computed}
    return {"peak": computed, "peak_source": "computed",
            "peak_detail": "148 SM x 128 FP32 lanes x 1.965 GHz, one non-FMA operation per lane per clock "
                           "(profiles/r2_peaks/fp32_peak.json missing)", "peak_computed": computed}


def load_ncu(kernel):
    """DRAM traffic and issue statistics of `kernel` from the committed ncu --set full capture, or None.  The capture
    records the hash of cornelis_b200/csrc it was taken from; a different hash today marks the figures stale."""
    path = NCU_JSON.get(kernel)
    if not path or not path.exists():
        return None
    from cornelis_b200 import build
    d = json.loads(path.read_text())
    d["stale"] = d.get("csrc_sha256") != build.csrc_hash()
    d["file"] = str(path.relative_to(ROOT))
    return d


# ------------------------------------------------------------------------------------------------- clocks --

class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region.

    nvidia-smi takes up to a second to print its first line (longer on an 8-GPU box), and at N = 8 the timed region of
    the strong-scaling run is 0.3 s: a sampler started with the region never saw it.  So the process is started once,
    as soon as the device is chosen, every line is stamped on arrival, and `summary()` keeps the lines that arrived
    inside the window `with sampler.window():` bracketed (grown by one sampling period on either side, which is the
    resolution of the stamps).  If the window was still shorter than the gaps between lines, the line nearest to it
    is used and the distance is reported (`nearest_sample_s`)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    PERIOD_S = 0.2

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.lines = []  # (arrival time, text)
        self.t0 = self.t1 = None
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", str(int(self.PERIOD_S * 1000)), "-i", str(self.device_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def window(self):
        return self

    def __enter__(self):
        self.t0 = time.monotonic()
        return self

    def __exit__(self, *exc):
        self.t1 = time.monotonic()

    def close(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.proc = None

    @staticmethod
    def _parse(line):
        parts = [p.strip() for p in line.split(",")]
        if len(parts) < 8:
            return None
        try:
            return float(parts[1]), float(parts[2]), float(parts[3]), parts[4:8]
        except ValueError:
            return None

    def summary(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        empty = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.t0 is None or self.t1 is None:
            return empty
        deadline = time.monotonic() + 2.0  # a line after the window, if none has come yet
        while self.proc and time.monotonic() < deadline and not any(t >= self.t1 for t, _ in list(self.lines)):
            time.sleep(0.05)
        parsed = [(t, self._parse(text)) for t, text in list(self.lines)]
        parsed = [(t, v) for t, v in parsed if v is not None]
        if not parsed:
            return empty
        inside = [(t, v) for t, v in parsed if self.t0 - self.PERIOD_S <= t <= self.t1 + self.PERIOD_S]
        out = {}
        if not inside:
            def distance(item):
                t = item[0]
                return self.t0 - t if t < self.t0 else t - self.t1
            nearest = min(parsed, key=distance)
            inside = [nearest]
            out["nearest_sample_s"] = round(distance(nearest), 3)
        sm = [v[0] for _, v in inside]
        reasons = {name for _, v in inside for name, val in zip(names, v[3]) if val.lower().startswith("active")}
        out.update({"sm_mhz": statistics.median(sm), "sm_max_mhz": max(v[1] for _, v in inside),
                    "reasons": sorted(reasons), "power_w_max": max(v[2] for _, v in inside), "samples": len(sm)})
        return out


# ------------------------------------------------------------------------------------------ reference (CPU) --

def time_reference(flat, W, H, seconds_target, threads=0):
    """Times the reference's own CPU loop (oracle/_ref: the unmodified sources compiled by oracle/build_ref.sh; the
    C restatement if that library is absent) on a bounded sample of the workload: the full frame at reduced spp
    (throughput is spp-independent: SURVEY.md section 6).  Tiles that divide the frame (40x40 at 1080p): 32x32 does
    not divide 1080 and trips the reference's FrameTiling spill bug (Tiles.cpp:21-24)."""
    from oracle import loader
    oracle = loader.best()
    scene = oracle.scene(flat)
    def dividing(n):  # a tile edge of 16..64 pixels that divides the frame edge (40 for 1920 and 1080), else the edge
        return 40 if n % 40 == 0 else min((d for d in range(16, 65) if n % d == 0), default=n)
    tile = (dividing(W), dividing(H))
    probe = scene.render(W, H, 1, tile=tile, threads=threads, stats=True)["stats"]
    rate = probe["pixel_samples"] / probe["seconds"]
    spp = max(1, min(64, int(round(seconds_target * rate / (W * H)))))
    return oracle, scene, tile, spp


def reference_arm(args, flat):
    """bench.py --impl reference: the reference CPU implementation, all host threads, same metric and config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    W, H = args.width, args.height
    cores = os.cpu_count() or 1
    total_budget = 150.0  # seconds for warm-up + timed steps
    per_step = max(2.0, min(12.0, total_budget / max(1, args.steps + args.warmup)))
    oracle, scene, tile, spp = time_reference(flat, W, H, per_step)
    for _ in range(args.warmup):
        scene.render(W, H, spp, tile=tile, stats=True)
    t0 = time.perf_counter()
    rays = 0.0
    for _ in range(args.steps):
        st = scene.render(W, H, spp, tile=tile, stats=True)["stats"]
        rays += st["rays"]
    elapsed = time.perf_counter() - t0
    samples = float(W) * H * spp * args.steps
    value = samples / elapsed / 1e6
    sample_desc = f"{W}x{H} at {spp} spp per step ({tile[0]}x{tile[1]} tiles), {args.steps} steps"
    line = {
        "impl": "reference", "metric": "pixel-samples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "mrays_per_s": rays / elapsed / 1e6,
        "config": workload_config(args),  # the same workload as our arm; the CPU renders a bounded sample of it:
        "note": f"CPU arm renders a bounded sample: {sample_desc}",
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": oracle.kind,
                         "sample": sample_desc},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_threads": cores,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- helpers --

def workload_config(args, note=None):
    scene = ("cornelis CLI Cornell box (5 planes, 4 spheres, layered Oren-Nayar+GGX material)" if args.workload == "cornell"
             else "10 000 random spheres over a floor, 64 mixed Oren-Nayar / glossy materials")
    cfg = {
        "workload": (f"{scene} "
                     f"{args.width}x{args.height}, camera aspect {args.height / args.width:.4f}, {args.spp} spp"
                     f"{' per GPU' if args.scaling == 'weak' else ' in the image (split across the GPUs)'}, "
                     f"Russian roulette, no depth cap "
                     f"(BASELINE.json configs[{1 if args.workload == 'cornell' else 3}])"),
        "width": args.width, "height": args.height, "spp": args.spp,
        "sharding": "contiguous global-sample-index ranges per GPU + ONE ncclAllReduce of the float4 accumulation "
                    "images, issued by the library (cornelis_cuda_allreduce_framebuffers)",
        "l2": "no L2 flush needed: the persistent pipeline keeps path state in registers (the only global traffic is "
              "scattered framebuffer atomics over a 33 MB image); the wavefront pipeline's pool + queues (2.9 GB) "
              "exceed the 126 MB L2",
        "seed": 19791102,
    }
    if note:
        cfg["note"] = note
    return cfg


def roofline(stats_list, hbm_peak, hbm_source, persistent):
    """Roofline of the dominant kernel.  Algorithmic work per SURVEY.md 8d: 279 flop per ray (4 sphere + 5 plane
    tests), 250 flop per surviving bounce, 12 per Russian-roulette-killed hit; duration = that kernel's mean launch
    duration measured with CUDA events inside the timed region."""
    tot = {k: sum(s[k] for s in stats_list) for k in ("rays", "pixel_samples", "shaded_hits", "iterations",
                                                       "kernel_launches", "contributions")}
    surviving = tot["rays"] - tot["pixel_samples"]          # every ray after the camera ray came from a surviving bounce
    killed = tot["shaded_hits"] - surviving
    fp32 = load_fp32_peak()
    extra = {}
    if persistent:
        launches = max(1, tot["kernel_launches"])
        name = "k_persistent_queued"
        ms = sum(s["gpu_ms"] for s in stats_list) / launches
        flop = (FLOP_PER_RAY_CORNELL * tot["rays"] + FLOP_PER_SURVIVING_BOUNCE * surviving +
                FLOP_PER_KILLED_HIT * killed) / launches
        # path state never leaves the registers: the only algorithmic traffic is the framebuffer read-modify-write
        # (16 B read + 16 B written) of paths that end with non-zero radiance
        nbytes = 32.0 * tot["contributions"] / launches
        extra["rays_per_launch"] = tot["rays"] / launches
        ncu = load_ncu(name)
        if ncu:
            # DRAM sees one read of the 33 MB image per launch whatever the sample count (it lives in the 126 MB L2),
            # so the captured figure is reported as captured, not scaled to this launch's sample count
            extra["traffic"] = ncu["dram_bytes"]
            extra["traffic_source"] = {k: ncu.get(k) for k in ("file", "command", "git_head", "csrc_sha256", "stale",
                                                               "launch")}
            extra["issue_slots"] = dict(ncu.get("issue", {}), stale=ncu["stale"])
    else:
        launches = max(1, tot["iterations"])
        stage = {k: statistics.mean(s[k] for s in stats_list)
                 for k in ("intersect_ms", "shade_ms", "raygen_ms", "accumulate_ms")}
        rays_per_launch = tot["rays"] / launches
        hits_per_launch = tot["shaded_hits"] / launches
        kernels = {
            "intersect": {
                "ms": stage["intersect_ms"], "flop": FLOP_PER_RAY_CORNELL * rays_per_launch,
                # 32 B ray read + 8 B hit written per ray; 4 B queue entry per hit; 16 B radiance read per miss
                "bytes": 40.0 * rays_per_launch + 4.0 * hits_per_launch + 16.0 * (rays_per_launch - hits_per_launch)},
            "shade": {
                "ms": stage["shade_ms"],
                "flop": (FLOP_PER_SURVIVING_BOUNCE * surviving + FLOP_PER_KILLED_HIT * killed) / launches,
                # per hit: 4 B queue + 64 B state + 8 B hit read; per survivor 64 B written
                "bytes": 76.0 * hits_per_launch + 64.0 * surviving / launches},
        }
        which = max(kernels, key=lambda k: kernels[k]["ms"])
        name, ms, flop, nbytes = f"k_{which}", kernels[which]["ms"], kernels[which]["flop"], kernels[which]["bytes"]
        extra["stage_ms_per_launch"] = stage
        extra["rays_per_launch"] = rays_per_launch
    dur = ms * 1e-3
    achieved = flop / dur / 1e12 if dur > 0 else 0.0
    hbm_achieved = nbytes / dur / 1e9 if dur > 0 else 0.0
    return {
        "kernel": name, "bound": "fp32", "achieved": achieved, "peak": fp32["peak"], "unit": "TFLOP/s",
        "frac": achieved / fp32["peak"], "peak_source": fp32["peak_source"], "peak_detail": fp32["peak_detail"],
        "peak_computed": fp32["peak_computed"], "frac_of_computed_peak": achieved / fp32["peak_computed"],
        "traffic": extra.pop("traffic", None),
        "launch_ms": ms, "flop_per_launch": flop,
        "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                "bytes_per_launch": nbytes, "peak_source": hbm_source},
        **extra,
    }


# ------------------------------------------------------------------------------ configs[2] and configs[3] (N = 1) --

def config3(args, device):
    """BASELINE.json configs[2]: 2^24 random rays against 1024 spheres + 6 planes, the reference's own strategy (every
    primitive for every ray, Render.cpp:115-140) in k_intersect_batch, device-resident float4 rays.  The CPU baseline leg
    runs the reference's intersect on a bounded sample of the same rays on all host threads and doubles as the check:
    hit ids and t must be bit-identical."""
    import numpy as np
    import torch

    from cornelis_b200 import binding, scenes

    n, n_spheres, n_planes = 1 << 24, 1024, 6
    flat = scenes.microbench_scene(n_spheres)
    scene = binding.Scene(flat, device=device)
    scene.set_acceleration(binding.ACCEL_NONE)
    org, dirs = scenes.microbench_rays(n)
    pad = np.zeros((n, 1), np.float32)
    dev = f"cuda:{device}"
    d_org = torch.from_numpy(np.concatenate([org, pad], 1)).to(dev)
    d_dir = torch.from_numpy(np.concatenate([dirs, pad], 1)).to(dev)
    d_hit = torch.empty((n, 2), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    scene.intersect_device(n, d_org.data_ptr(), d_dir.data_ptr(), d_hit.data_ptr(), repeats=3)  # warm-up
    ms = scene.intersect_device(n, d_org.data_ptr(), d_dir.data_ptr(), d_hit.data_ptr(), repeats=10)
    scene.set_acceleration(binding.ACCEL_GRID)
    d_hit_grid = torch.empty_like(d_hit)
    scene.intersect_device(n, d_org.data_ptr(), d_dir.data_ptr(), d_hit_grid.data_ptr(), repeats=2)
    ms_grid = scene.intersect_device(n, d_org.data_ptr(), d_dir.data_ptr(), d_hit_grid.data_ptr(), repeats=5)
    same_grid = bool(torch.equal(d_hit.view(torch.int32), d_hit_grid.view(torch.int32)))
    hits = d_hit.cpu().numpy()
    flop = n * float(FLOP_PER_SPHERE_TEST * n_spheres + FLOP_PER_PLANE_TEST * n_planes)  # 4.502e11, SURVEY.md 8d
    fp32 = load_fp32_peak()
    out = {
        "workload": f"2^24 rays x ({n_spheres} spheres + {n_planes} planes), exhaustive scan (BASELINE.json configs[2])",
        "ms": ms, "grays_per_s": n / ms / 1e6, "gtests_per_s": n * (n_spheres + n_planes) / ms / 1e6,
        "algorithmic_tflops": flop / ms / 1e9, "frac_of_fp32_peak": flop / (ms * 1e-3) / (fp32["peak"] * 1e12),
        "fp32_peak_source": fp32["peak_source"], "gpu_launches": 10,
        "grid": {"ms": ms_grid, "grays_per_s": n / ms_grid / 1e6, "bit_identical_to_exhaustive": same_grid},
    }
    if not args.no_cpu_baseline:
        from concurrent.futures import ThreadPoolExecutor

        from oracle import loader
        oracle = loader.best()
        ref = oracle.scene(flat)
        cores = os.cpu_count() or 1
        probe_n = 1 << 12
        t0 = time.perf_counter()
        ref.intersect(org[:probe_n], dirs[:probe_n])
        per_ray = (time.perf_counter() - t0) / probe_n
        sample = int(min(n, max(1 << 16, 6.0 * cores / per_ray)))  # about 6 s on all threads
        sample = 1 << (sample.bit_length() - 1)
        chunks = np.array_split(np.arange(sample), cores * 4)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as pool:  # ctypes releases the GIL: one chunk per host thread at a time
            parts = list(pool.map(lambda idx: ref.intersect(org[idx[0]:idx[-1] + 1], dirs[idx[0]:idx[-1] + 1]), chunks))
        seconds = time.perf_counter() - t0
        t_ref = np.concatenate([p["t"] for p in parts])
        prim_ref = np.concatenate([p["prim"] for p in parts])
        same = bool(np.array_equal(hits[:sample, 0].view(np.uint32), t_ref.view(np.uint32)) and
                    np.array_equal(hits[:sample, 1].view(np.int32), prim_ref))
        out["bit_identical"] = same
        out["cpu_baseline"] = {"value": sample / seconds / 1e6, "unit": "Mrays/s", "cores": cores, "kind": oracle.kind,
                               "sample": f"the first {sample} of the 2^24 rays, {seconds:.1f} s on all host threads; "
                                         "t and hit id compared bit for bit with the GPU's"}
    scene.close()
    return out


def config4(args, device):
    """BASELINE.json configs[3]: 10 000 spheres over a floor, 64 materials, 1920x1080, max depth 64, through the uniform
    grid (bit-identical to the exhaustive scan the reference runs, Render.cpp:115-140), at the configuration's own 1024
    spp: one render of ~1.2 s (the last ~20 passes of a render only carry the dwindling deep paths, so a 256-spp sample
    of it reads ~4 % low; stage times are the means over one pass in 32, 16 samples at 1024 spp)."""
    from cornelis_b200 import binding, scenes

    W, H, spp = args.width, args.height, 1024
    flat = scenes.many_spheres(10000, aspect=H / W)
    scene = binding.Scene(flat, device=device)
    # warm-up: enough paths (64 spp = 1.3e8) for the pool and queues to reach the size the timed render uses (2^26 paths,
    # 11.5 GB), so that their allocation and first use are not inside the timed render
    scene.render_accumulate(W, H, 64, max_depth=64)
    st = scene.render_accumulate(W, H, spp, max_depth=64, stage_timing=True)
    hbm_peak, _ = load_peaks()
    stage = {k: st[k] for k in ("raygen_ms", "intersect_ms", "shade_ms", "accumulate_ms")}
    rays_per_pass = st["rays"] / max(1, st["iterations"])
    hits_per_pass = st["shaded_hits"] / max(1, st["iterations"])
    surviving_per_pass = (st["rays"] - st["pixel_samples"]) / max(1, st["iterations"])
    # algorithmic bytes per pass of each stage over its measured duration, against the measured copy
    # bandwidth: raygen writes 64 B per new path; walk + compaction read 32 B and write 8 B per ray, then re-read the
    # 8 B and write 4 B per hit; shade reads 76 B per hit and writes 64 B per survivor
    new_per_pass = st["pixel_samples"] / max(1, st["iterations"])
    stage_bytes = {"raygen_ms": 64.0 * new_per_pass, "intersect_ms": 48.0 * rays_per_pass + 4.0 * hits_per_pass,
                   "shade_ms": 76.0 * hits_per_pass + 64.0 * surviving_per_pass}
    out = {
        "workload": f"10 000 spheres + floor, 64 materials, {W}x{H}, {spp} spp, max depth 64 "
                    "(BASELINE.json configs[3])",
        "pipeline": "wavefront + k_walk (uniform grid)", "acceleration": scene.acceleration(),
        "msamples_per_s": st["pixel_samples"] / st["gpu_ms"] / 1e3, "mrays_per_s": st["rays"] / st["gpu_ms"] / 1e3,
        "ms": st["gpu_ms"], "passes": st["iterations"], "rays_per_sample": st["rays"] / st["pixel_samples"],
        "max_depth": st["max_depth"],
        "gpu_launches": st["kernel_launches"], "stage_ms_per_pass": stage,
        "stage_hbm_frac": {k.replace("_ms", ""): (b / (stage[k] * 1e-3) / 1e9 / hbm_peak if stage[k] > 0 else None)
                           for k, b in stage_bytes.items()},
    }
    if not args.no_cpu_baseline:
        # the reference's CPU loop (every sphere for every ray) on a bounded sample: a tenth of the frame per dimension
        w, h = W // 10, H // 10
        oracle, ref_scene, tile, spp_cpu = time_reference(scenes.many_spheres(10000, aspect=h / w), w, h, 6.0)
        rs = ref_scene.render(w, h, spp_cpu, tile=tile, stats=True)["stats"]
        out["cpu_baseline"] = {"value": rs["pixel_samples"] / rs["seconds"] / 1e6, "unit": "Msamples/s",
                               "mrays_per_s": rs["rays"] / rs["seconds"] / 1e6, "cores": os.cpu_count() or 1,
                               "kind": oracle.kind,
                               "sample": f"{w}x{h} at {spp_cpu} spp ({tile[0]}x{tile[1]} tiles), {rs['seconds']:.1f} s "
                                         "on all host threads"}
    scene.close()
    return out


def config5(world, rank, local_rank, comm, dist, torch, binding, flat_4k):
    """BASELINE.json configs[4]: the Cornell scene at 3840x2160, 16 384 spp, sample-sharded across the GPUs of the box
    with the NCCL framebuffer sum at the end — one timed render (all ranks; device time, max over ranks)."""
    from cornelis_b200.sharding import sample_range
    W, H, spp = 3840, 2160, 16384
    first, count, total = sample_range(rank, world, spp, "strong")
    scene = binding.Scene(flat_4k, device=local_rank)
    stream = torch.cuda.Stream()
    scene.set_stream(stream.cuda_stream)

    def step(samples_first, samples_count):
        with torch.cuda.stream(stream):
            st = scene.render_accumulate(W, H, total, first_sample=samples_first, sample_count=samples_count)
            comm.allreduce_framebuffers([scene])
            scene.resolve_device(total)
        return st

    step(first, max(1, min(count, 16)))  # warm-up: allocations, the 133 MB all-reduce once
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    st = step(first, count)
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=f"cuda:{local_rank}")
    rays = torch.tensor([float(st["rays"])], dtype=torch.float64, device=f"cuda:{local_rank}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    scene.close()
    seconds = float(t[0])
    return {"workload": f"Cornell {W}x{H}, {spp} spp split over {world} GPUs by sample index, one ncclAllReduce of the "
                        f"{W * H * 4} accumulator floats (BASELINE.json configs[4])",
            "seconds": seconds, "msamples_per_s": float(W) * H * spp / seconds / 1e6,
            "mrays_per_s": float(rays[0]) / seconds / 1e6, "spp_per_gpu": count, "steps": 1}


def single_process_check(args, world):
    """The C++ path for one process driving several GPUs (RenderSession with RenderOptions::devices = N: one host
    thread per device, one grouped ncclReduce of the accumulation images onto device 0), through the CLI: the frame at
    256 spp on N devices and on one, compared pixel by pixel."""
    import numpy as np

    cli = ROOT / "cornelis_b200" / "lib" / "cornelis"
    if not cli.exists():  # normally built by __graft_entry__.build() and shipped with the snapshot
        from cornelis_b200 import build
        build.build_all()
    W, H, spp = args.width, args.height, 256
    out = {"devices": world, "frame": f"{W}x{H} at {spp} spp", "path": "cornelis CLI -> RenderSession(devices = N) -> "
           "cornelis_cuda_render_accumulate per device thread + cornelis_cuda_reduce_framebuffers (ncclReduce)"}
    images = {}
    with tempfile.TemporaryDirectory() as tmp:
        for n in (world, 1):
            raw = Path(tmp) / f"d{n}.f32"
            cmd = [str(cli), "--width", str(W), "--height", str(H), "--spp", str(spp), "--devices", str(n), "--no-save",
                   "--dump-raw", str(raw)]
            r = None
            for _ in range(2):  # the second run has the contexts, NCCL communicator set-up and allocations warm
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
                if r.returncode != 0:
                    break
            if r.returncode != 0:
                out["error"] = (r.stdout + r.stderr)[-400:]
                return out
            line = r.stdout.strip().splitlines()[-1]
            images[n] = np.fromfile(raw, np.float32)
            out[f"devices_{n}"] = line
            try:
                out[f"gpu_ms_{n}"] = 1e3 * float(line.split("s wall,")[1].split("s on the GPU")[0])
            except (IndexError, ValueError):
                pass
    a, b = images[world], images[1]
    finite = np.isfinite(a) & np.isfinite(b)
    out["max_abs_diff"] = float(np.abs(a[finite] - b[finite]).max())
    out["max_rel_diff"] = float((np.abs(a[finite] - b[finite]) / np.maximum(np.abs(b[finite]), 1e-3)).max())
    out["same_nonfinite_pixels"] = bool(np.array_equal(np.isfinite(a), np.isfinite(b)))
    if out.get("gpu_ms_1") and out.get(f"gpu_ms_{world}"):
        out["speedup"] = out["gpu_ms_1"] / out[f"gpu_ms_{world}"]
    return out


# ---------------------------------------------------------------------------------------------------- ours --

def ours(args, flat):
    import numpy as np
    import torch
    import torch.distributed as dist

    from cornelis_b200 import binding
    from cornelis_b200.sharding import sample_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    sampler = ClockSampler(local_rank)  # started early: its first line takes about a second
    comm = None
    if world > 1:
        # torch.distributed: rendezvous, barriers, max-over-ranks of the timings.  The data plane is the library's own
        # NCCL communicator: rank 0 creates the id, every rank joins with ncclCommInitRank.
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [binding.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = binding.Comm.init_rank(box[0], rank, world, local_rank)

    W, H = args.width, args.height
    npix = W * H
    pipeline = {"default": binding.PIPELINE_DEFAULT, "wavefront": binding.PIPELINE_WAVEFRONT,
                "persistent": binding.PIPELINE_PERSISTENT}[args.pipeline]
    persistent = args.pipeline == "persistent" or (args.pipeline == "default" and
                                                   os.environ.get("CORNELIS_PIPELINE", "") != "wavefront")
    stream = torch.cuda.Stream()
    scene = binding.Scene(flat, device=local_rank)
    scene.set_stream(stream.cuda_stream)
    pinned = torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True)
    pinned_np = pinned.numpy()
    launches = [0]
    debug = os.environ.get("CORNELIS_BENCH_DEBUG")

    def make_steps(scaling):
        first, count, total_spp = sample_range(rank, world, args.spp, scaling)

        def step_device(collect=None):
            """Inputs resident in HBM, result left in HBM."""
            with torch.cuda.stream(stream):
                st = scene.render_accumulate(W, H, total_spp, first_sample=first, sample_count=count,
                                             pool_paths=args.pool, stage_timing=True, pipeline=pipeline)
                if comm:
                    comm.allreduce_framebuffers([scene])  # on the scene's stream: the resolve is ordered behind it
                scene.resolve_device(total_spp)
            launches[0] += st["kernel_launches"] + 1
            if collect is not None:
                collect.append(st)
            return st

        def step_e2e():
            """Through the public call with HOST buffers: scene description up, framebuffer down, every step."""
            marks = [time.perf_counter()]
            with torch.cuda.stream(stream):
                sc = binding.Scene(flat, device=local_rank)
                sc.set_stream(stream.cuda_stream)
                marks.append(time.perf_counter())
                sc.render_accumulate(W, H, total_spp, first_sample=first, sample_count=count, pool_paths=args.pool,
                                     pipeline=pipeline)
                marks.append(time.perf_counter())
                if comm:
                    comm.allreduce_framebuffers([sc])
                marks.append(time.perf_counter())
                if rank == 0:
                    sc.resolve(total_spp, out=pinned_np)
                else:
                    stream.synchronize()
                marks.append(time.perf_counter())
                sc.close()
                marks.append(time.perf_counter())
            if debug:
                names = ["create", "render", "reduce", "resolve+d2h", "destroy"]
                print(f"[rank {rank}] e2e " + "  ".join(f"{n} {1e3 * (b - a):.1f} ms" for n, a, b in
                                                       zip(names, marks, marks[1:])), file=sys.stderr, flush=True)
            return sc.scene_bytes

        return step_device, step_e2e, total_spp

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        sync_all()
        dev = e0.elapsed_time(e1) * 1e-3
        t = torch.tensor([dev, wall], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    step_device, step_e2e, total_spp = make_steps(args.scaling)
    for _ in range(args.warmup):
        step_device()
    launches[0] = 0
    collected = []
    with sampler.window():
        dev_s, wall_s = timed(lambda: step_device(collected), args.steps)
    try:
        clocks_summary = sampler.summary()
    except Exception as e:  # the clocks block is evidence beside the measurement: never the reason a bench line is lost
        clocks_summary = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0,
                          "error": f"{type(e).__name__}: {e}"}
    sampler.close()
    timed_launches = launches[0]
    scene_bytes = step_e2e()  # warm the allocation path once
    e2e_dev_s, e2e_wall_s = timed(step_e2e, args.steps)

    other = None
    if world > 1:  # the other scaling mode beside the headline, a few steps
        other_mode = "weak" if args.scaling == "strong" else "strong"
        o_step, _, o_spp = make_steps(other_mode)
        o_steps = min(args.steps, 3)
        o_step()
        o_dev_s, _ = timed(o_step, o_steps)
        other = {"scaling": other_mode, "value": float(npix) * o_spp * o_steps / o_dev_s / 1e6, "unit": "Msamples/s",
                 "ms_per_step": 1e3 * o_dev_s / o_steps, "steps": o_steps, "spp_in_image": o_spp,
                 "spp_per_gpu": o_spp // world}

    c5 = None
    if world > 1 and not args.no_configs:
        from cornelis_b200 import scenes
        try:
            c5 = config5(world, rank, local_rank, comm, dist, torch, binding, scenes.cornell_box(aspect=2160 / 3840))
        except Exception as e:  # every rank takes the same path: the collectives inside stay matched
            c5 = {"error": f"{type(e).__name__}: {e}"}

    # whole-job totals: the ranks' sample ranges partition (strong) or extend (weak) the image's samples
    samples_per_step = float(npix) * total_spp
    rays_local = statistics.mean(s["rays"] for s in collected)
    rays_t = torch.tensor([rays_local], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(rays_t, op=dist.ReduceOp.SUM)
        dist.barrier()
    rays_per_step = float(rays_t[0])
    store = dist.distributed_c10d._get_default_store() if world > 1 else None
    if rank != 0:
        # rank 0 may now drive every GPU of the box from one process (single_process_check): wait on the rendezvous
        # store, not in an NCCL barrier — a barrier is a kernel spinning on this rank's GPU
        store.wait(["bench_rank0_done"], __import__("datetime").timedelta(minutes=20))
        comm.close()
        dist.destroy_process_group()
        return

    hbm_peak, hbm_source = load_peaks()
    value = samples_per_step * args.steps / dev_s / 1e6
    cores = os.cpu_count() or 1
    line = {
        "metric": "pixel-samples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "mrays_per_s": rays_per_step * args.steps / dev_s / 1e6,
        "rays_per_pixel_sample": rays_per_step / samples_per_step,
        "wall_ms_per_step": 1e3 * wall_s / args.steps,
        "e2e": {"value": samples_per_step * args.steps / e2e_wall_s / 1e6, "unit": "Msamples/s",
                "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": int(npix * 3 * 4),
                "ms_per_step": 1e3 * e2e_wall_s / args.steps,
                "path": "cornelis_cuda_scene_create + render_accumulate + (cornelis_cuda_allreduce_framebuffers) + "
                        "resolve into pinned host memory"},
        "gpu_launches": int(timed_launches),
        "roofline": roofline(collected, hbm_peak, hbm_source, persistent),
        "pipeline": "persistent" if persistent else "wavefront",
        "clocks": clocks_summary,
        "max_depth": max(s["max_depth"] for s in collected),
        "host_threads": cores,
    }
    if comm:
        line["comm"] = dict(comm.info(), library="cornelis_cuda_comm_init_rank (ncclCommInitRank) + "
                                                 "cornelis_cuda_allreduce_framebuffers (one ncclAllReduce, fp32 sum, "
                                                 f"{npix * 4} floats)")
        line[other["scaling"]] = other
        line["spp_per_gpu"] = total_spp // world if args.scaling == "strong" else args.spp
    if world == 1 and not args.no_cpu_baseline:
        oracle, ref_scene, tile, spp = time_reference(flat, W, H, args.cpu_seconds)
        st = ref_scene.render(W, H, spp, tile=tile, stats=True)["stats"]
        line["cpu_baseline"] = {
            "value": st["pixel_samples"] / st["seconds"] / 1e6, "unit": "Msamples/s", "cores": cores,
            "kind": oracle.kind, "mrays_per_s": st["rays"] / st["seconds"] / 1e6,
            "sample": f"{W}x{H} at {spp} spp ({tile[0]}x{tile[1]} tiles), {st['seconds']:.1f} s on all "
                      f"{cores} host threads",
        }
    if world == 1 and not args.no_configs:
        scene.close()
        binding.trim_memory()
        configs = {}
        for name, fn in (("c3", config3), ("c4", config4)):
            try:
                configs[name] = fn(args, local_rank)
            except Exception as e:  # the headline line must still be printed
                configs[name] = {"error": f"{type(e).__name__}: {e}"}
        line["configs"] = configs
    if c5 is not None:
        line["configs"] = {"c5": c5}
    if world > 1 and not args.no_single_process:
        try:
            line["single_process"] = single_process_check(args, world)
        except Exception as e:
            line["single_process"] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        store.set("bench_rank0_done", "1")
        comm.close()
        dist.destroy_process_group()


def main():
    args = parse_args()
    from cornelis_b200 import scenes
    if args.workload == "spheres":
        if args.impl != "reference":
            raise SystemExit("--workload spheres is the reference arm's; our arm reports the scene under configs.c4")
        flat = scenes.many_spheres(10000, aspect=args.height / args.width)
    else:
        flat = scenes.cornell_box(aspect=args.height / args.width)
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch under torch.distributed.run, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), __file__,
               *sys.argv[1:]]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        reference_arm(args, flat)
    else:
        ours(args, flat)


if __name__ == "__main__":
    main()
