"""TEST INFRASTRUCTURE (it loads the oracle): walk statistics of the uniform grid on config 4's scene, on the CPU.

Builds tests/native/grid_host.cu (the grid builder and closestHitGrid compiled for the host: test infrastructure) once
per set of -D flags, follows camera rays through a few diffuse bounces (hit point + 1e-4 along a uniform hemisphere
direction, Render.cpp:207) and prints, per generation of rays, the mean number of cells a walk visits and of spheres it
tests.  With --check N the hits of the first N rays of every generation are compared bit for bit with the exhaustive scan
of the plain-C oracle.  Used to size changes of the walk (termination slack, registration radius, grid density) before
spending GPU time on them:

    python tests/grid_walk_stats.py --rays 200000 --variant base=-DCORNELIS_GRID_RAY_MARGIN=0 --variant ray=-DCORNELIS_GRID_RAY_MARGIN=1
"""
import argparse
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from cornelis_b200 import binding, scenes  # noqa: E402

SRC = ROOT / "tests" / "native" / "grid_host.cu"
OUT = ROOT / "tests" / "native" / "_build"


def build(name, flags):
    OUT.mkdir(parents=True, exist_ok=True)
    so = OUT / f"libgrid_host_{name}.so"
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "--fmad=false", "-std=c++17",
                    "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", *flags, "-o", str(so), str(SRC)], check=True)
    L = C.CDLL(str(so))
    vp, sz = C.c_void_p, C.c_size_t
    L.grid_host_intersect.argtypes = [C.POINTER(binding.CameraDesc), vp, sz, vp, sz, sz, vp, vp, vp, vp, vp, vp]
    return L


def walk(L, flat, org, dirs):
    cam, S, P, _, (sph, pl, _m) = binding.descriptors(flat)
    org = np.ascontiguousarray(org, np.float32)
    dirs = np.ascontiguousarray(dirs, np.float32)
    n = len(org)
    t, prim = np.empty(n, np.float32), np.empty(n, np.int32)
    stats, info = np.zeros((n, 2), np.uint32), np.zeros(6, np.uint64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.grid_host_intersect(C.byref(cam), S, len(sph), P, len(pl), n, p(org), p(dirs), p(t), p(prim), p(stats), p(info))
    assert rc == 0
    return t, prim, stats[:, 0], stats[:, 1], info


def bounce(rng, flat, org, dirs, t, prim):
    """Uniform hemisphere directions about the surface normal at every hit; origins 1e-4 along the new direction."""
    hit = prim >= 0
    o, d, t, prim = org[hit], dirs[hit], t[hit], prim[hit]
    P = (o + d * t[:, None]).astype(np.float32)
    sph = flat["spheres"]
    ns = len(sph)
    N = np.zeros_like(P)
    on_sphere = prim < ns
    c = sph[np.where(on_sphere, prim, 0), :3]
    N[on_sphere] = (P - c)[on_sphere]
    N[~on_sphere] = np.float32([0, 1, 0])
    N /= np.linalg.norm(N, axis=1, keepdims=True)
    g = rng.standard_normal(P.shape).astype(np.float32)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    g = np.where((g * N).sum(axis=1, keepdims=True) < 0, -g, g).astype(np.float32)
    return (P + g * np.float32(1e-4)).astype(np.float32), g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=100000)
    ap.add_argument("--spheres", type=int, default=10000)
    ap.add_argument("--bounces", type=int, default=3)
    ap.add_argument("--check", type=int, default=0, help="rays per generation compared with the oracle's exhaustive scan")
    ap.add_argument("--variant", action="append", default=[], help="name=flag[,flag...]")
    args = ap.parse_args()
    variants = [v.split("=", 1) for v in args.variant] or [["default", ""]]
    flat = scenes.many_spheres(args.spheres)
    oracle_scene = None
    if args.check:
        from oracle import loader
        oracle_scene = loader.load("port").scene(flat)
    rng = np.random.default_rng(1)
    pi, pj = rng.integers(0, 1920, args.rays), rng.integers(0, 1080, args.rays)
    from oracle import loader
    cam_scene = loader.load("port").scene(flat)
    org, dirs = cam_scene.pixel_rays(1920, 1080, pi, pj, rng.random(args.rays, dtype=np.float32),
                                     rng.random(args.rays, dtype=np.float32))
    libs = [(name, build(name, [f for f in flags.split(",") if f])) for name, flags in variants]
    for gen in range(args.bounces + 1):
        first = None
        for name, L in libs:
            t, prim, cells, tests, info = walk(L, flat, org, dirs)
            line = (f"gen {gen} {name:>10}: rays {len(org)} hit {np.mean(prim >= 0):.3f} cells mean {cells.mean():.2f} "
                    f"median {np.median(cells):.0f} tests mean {tests.mean():.2f} walked {np.mean(cells > 0):.3f} "
                    f"refs {int(info[3])} dims {tuple(int(x) for x in info[:3])}")
            if first is None:
                first = (t, prim)
            else:
                same = np.array_equal(first[1], prim) and np.array_equal(first[0].view(np.uint32), t.view(np.uint32))
                line += f" same_as_first {same}"
            if oracle_scene is not None:
                k = min(args.check, len(org))
                want = oracle_scene.intersect(org[:k], dirs[:k])
                ok = np.array_equal(want["prim"], prim[:k]) and np.array_equal(
                    np.ascontiguousarray(want["t"], np.float32).view(np.uint32), t[:k].view(np.uint32))
                line += f" oracle_ok({k}) {ok}"
            print(line, flush=True)
        org, dirs = bounce(rng, flat, org, dirs, *first)


if __name__ == "__main__":
    main()
