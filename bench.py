#!/usr/bin/env python
"""Headline benchmark: pixel-samples/s (and rays/s) of the batched render loop on the Cornell scene at 1080p.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path through the C-ABI
    python bench.py --impl reference [...]                          the reference's own CPU loop (oracle/_ref)

A "step" is one full render of the workload: BASELINE.json configs[1], the reference CLI's Cornell scene at
1920x1080 (camera aspect 0.5625 for square pixels), 4096 samples per pixel, Russian roulette, no depth cap.
With N GPUs every rank renders its own 4096-sample range of every pixel (global sample indices
[rank*4096, (rank+1)*4096)): per-GPU work is fixed, the image gets N*4096 spp — weak scaling — and the only
exchange is one NCCL sum of the float4 framebuffers at the end of the step.  `--scaling strong` splits 4096 spp
across the ranks instead.

Prints ONE JSON line (rank 0).  See the task contract for the meaning of every key.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SM_COUNT = 148
FP32_LANES_PER_SM = 128
SM_MAX_GHZ = 1.965
# algorithmic work of the two compute kernels (reference source after trivial CSE)
FLOP_PER_RAY_CORNELL = 4 * 26 + 5 * 35        # 279: 4 sphere tests x 26 flop + 5 plane tests x 35 flop
FLOP_PER_SURVIVING_BOUNCE = 250
FLOP_PER_KILLED_HIT = 12
# bytes per item of the wavefront layout (DESIGN.md "Data layout"): see roofline() below
# DRAM traffic of the persistent kernel from the committed ncu capture (profiles/r1_queue/
# ncu_full_summary_k_persistent_queued_cornell_512spp.txt): dram__bytes_read.sum + dram__bytes_write.sum = 32.49 MB +
# 0.09 MB for one launch of 1920x1080 x 512 spp.  The only global traffic is framebuffer atomics over a 33 MB image
# that lives in the 126 MB L2, so DRAM sees one read of the image per launch whatever the sample count (a 64-spp
# launch: 31.65 MB + 0.02 MB): the figure is reported as captured, not scaled.
NCU_PERSISTENT_DRAM_BYTES = 32.489472e6 + 0.090624e6
NCU_PERSISTENT_ISSUE = {"busy_pct": 84.8, "warp_instructions_per_ray": 37.1, "lanes_per_instruction": 28.95,
                        "no_instruction_stall_per_issue": 0.64, "registers": 72, "ctas_per_sm": "7 x 128 threads"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=4096)
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak")
    ap.add_argument("--pool", type=int, default=0, help="paths in flight per GPU (wavefront pipeline; 0 = default)")
    ap.add_argument("--pipeline", choices=["default", "wavefront", "persistent"], default="default")
    ap.add_argument("--workload", choices=["cornell", "spheres"], default="cornell",
                    help="spheres = BASELINE.json configs[3] (10 000 spheres); only the reference arm takes it here — "
                         "tools/bench_config4.py measures the GPU side and calls this for its CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def load_peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        p = json.loads(path.read_text())
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------- clocks --

class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.lines = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


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
        "config": workload_config(args, note=f"CPU arm renders a bounded sample: {sample_desc}"),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": oracle.kind,
                         "sample": sample_desc},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- helpers --

def workload_config(args, note=None):
    scene = ("cornelis CLI Cornell box (5 planes, 4 spheres, layered Oren-Nayar+GGX material)" if args.workload == "cornell"
             else "10 000 random spheres over a floor, 64 mixed Oren-Nayar / glossy materials")
    cfg = {
        "workload": (f"{scene} "
                     f"{args.width}x{args.height}, camera aspect {args.height / args.width:.4f}, {args.spp} spp"
                     f"{' per GPU' if args.scaling == 'weak' else ' total'}, Russian roulette, no depth cap "
                     f"(BASELINE.json configs[{1 if args.workload == 'cornell' else 3}])"),
        "width": args.width, "height": args.height, "spp": args.spp,
        "sharding": "sample ranges per GPU + one NCCL sum of the float4 framebuffers",
        "l2": "no L2 flush needed: the persistent pipeline keeps path state in registers (the only global traffic is "
              "scattered framebuffer atomics over a 33 MB image); the wavefront pipeline's pool + queues (2.9 GB) "
              "exceed the 126 MB L2",
        "seed": 19791102,
    }
    if note:
        cfg["note"] = note
    return cfg


class DeviceArray:
    """Wraps a raw device pointer for torch (CUDA array interface) — torch is only the NCCL plumbing here."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def roofline(stats_list, hbm_peak, hbm_source, persistent):
    """Roofline of the dominant kernel.  Algorithmic work per SURVEY.md 8d: 279 flop per ray (4 sphere + 5 plane
    tests), 250 flop per surviving bounce, 12 per Russian-roulette-killed hit; duration = that kernel's mean launch
    duration measured with CUDA events inside the timed region."""
    tot = {k: sum(s[k] for s in stats_list) for k in ("rays", "pixel_samples", "shaded_hits", "iterations",
                                                       "kernel_launches", "contributions")}
    surviving = tot["rays"] - tot["pixel_samples"]          # every ray after the camera ray came from a surviving bounce
    killed = tot["shaded_hits"] - surviving
    fp32_peak = SM_COUNT * FP32_LANES_PER_SM * SM_MAX_GHZ / 1e3  # 37.2 Tflop/s: non-FMA FP32 instruction rate
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
        extra["traffic"] = NCU_PERSISTENT_DRAM_BYTES
        extra["traffic_source"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of a "
                                   "1920x1080 x 512-spp launch of the same kernel (profiles/r1_queue); not scaled: the "
                                   "33 MB framebuffer lives in L2 and DRAM sees one read of it per launch whatever the "
                                   "sample count, far below the algorithmic bytes of the framebuffer atomics")
        extra["issue_slots"] = dict(NCU_PERSISTENT_ISSUE,
                                    source="same capture: smsp__issue_active, smsp__inst_executed.sum / rays, "
                                           "smsp__thread_inst_executed_per_inst_executed, stall_no_instruction")
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
        "kernel": name, "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": achieved / fp32_peak,
        "peak_source": "148 SM x 128 FP32 lanes x 1.965 GHz, one non-FMA op per lane per clock (SURVEY.md 8d); "
                       "FP32 is not in MEASURED_PEAKS.json",
        "traffic": extra.pop("traffic", None),
        "launch_ms": ms, "flop_per_launch": flop,
        "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                "bytes_per_launch": nbytes, "peak_source": hbm_source},
        **extra,
    }


# ---------------------------------------------------------------------------------------------------- ours --

def ours(args, flat):
    import numpy as np
    import torch
    import torch.distributed as dist

    from cornelis_b200 import binding
    from cornelis_b200.sharding import sample_range, sum_framebuffers

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W, H = args.width, args.height
    npix = W * H
    first, count, total_spp = sample_range(rank, world, args.spp, args.scaling)
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

    def reduce_framebuffer(sc):
        if world > 1:
            ptr, n = sc.framebuffer_device()
            fb = torch.as_tensor(DeviceArray(ptr, n), device=f"cuda:{local_rank}")
            sum_framebuffers(fb)
            launches[0] += 1

    def step_device(collect=None):
        """Inputs resident in HBM, result left in HBM."""
        with torch.cuda.stream(stream):
            st = scene.render_accumulate(W, H, total_spp, first_sample=first, sample_count=count, pool_paths=args.pool,
                                         stage_timing=True, pipeline=pipeline)
            reduce_framebuffer(scene)
            stream.synchronize()
            scene.resolve_device(total_spp)
        launches[0] += st["kernel_launches"] + 1
        if collect is not None:
            collect.append(st)
        return st

    debug = os.environ.get("CORNELIS_BENCH_DEBUG")

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
            reduce_framebuffer(sc)
            stream.synchronize()
            torch.cuda.synchronize()
            marks.append(time.perf_counter())
            if rank == 0:
                sc.resolve(total_spp, out=pinned_np)
            marks.append(time.perf_counter())
            sc.close()
            marks.append(time.perf_counter())
        if debug:
            names = ["create", "render", "reduce", "resolve+d2h", "destroy"]
            print(f"[rank {rank}] e2e " + "  ".join(f"{n} {1e3 * (b - a):.1f} ms" for n, a, b in
                                                   zip(names, marks, marks[1:])), file=sys.stderr, flush=True)
        return sc.scene_bytes

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

    for _ in range(args.warmup):
        step_device()
    launches[0] = 0
    collected = []
    with ClockSampler(local_rank) as clocks:
        dev_s, wall_s = timed(lambda: step_device(collected), args.steps)
    timed_launches = launches[0]
    scene_bytes = step_e2e()  # warm the allocation path once
    e2e_dev_s, e2e_wall_s = timed(step_e2e, args.steps)

    # whole-job totals: every rank renders the same amount
    samples_per_step = float(npix) * total_spp
    rays_per_step = statistics.mean(s["rays"] for s in collected) * world
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, hbm_source = load_peaks()
    value = samples_per_step * args.steps / dev_s / 1e6
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
                "path": "cornelis_cuda_scene_create + render_accumulate + (NCCL sum) + resolve into pinned host memory"},
        "gpu_launches": int(timed_launches),
        "roofline": roofline(collected, hbm_peak, hbm_source, persistent),
        "pipeline": "persistent" if persistent else "wavefront",
        "clocks": clocks.summary(),
        "max_depth": max(s["max_depth"] for s in collected),
    }
    if world == 1 and not args.no_cpu_baseline:
        oracle, ref_scene, tile, spp = time_reference(flat, W, H, args.cpu_seconds)
        st = ref_scene.render(W, H, spp, tile=tile, stats=True)["stats"]
        line["cpu_baseline"] = {
            "value": st["pixel_samples"] / st["seconds"] / 1e6, "unit": "Msamples/s", "cores": os.cpu_count() or 1,
            "kind": oracle.kind, "mrays_per_s": st["rays"] / st["seconds"] / 1e6,
            "sample": f"{W}x{H} at {spp} spp ({tile[0]}x{tile[1]} tiles), {st['seconds']:.1f} s on all host threads",
        }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    from cornelis_b200 import scenes
    if args.workload == "spheres":
        if args.impl != "reference":
            raise SystemExit("--workload spheres: use tools/bench_config4.py for the GPU side")
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
