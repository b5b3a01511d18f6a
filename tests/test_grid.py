"""Uniform grid over the spheres (cornelis_b200/csrc/scene_tables.h buildGrid and
csrc/geometry.cuh closestHitGrid): CPU tests.

closestHitGrid is a __host__ __device__ function; tests/native/grid_host.cu compiles it and the grid builder for the
host (test infrastructure, not shipped) so that the claim "walking the grid returns the exhaustive scan's hit id and t
bit for bit" is checked against the oracle without a GPU.  The GPU side of the same claim is in
tests/test_gpu_parity.py.
"""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from cornelis_b200 import binding, scenes
from conftest import bit_equal

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "native" / "grid_host.cu"
OUT = ROOT / "tests" / "native" / "_build" / "libgrid_host.so"


@pytest.fixture(scope="session")
def grid_host():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    deps = [SRC, *(ROOT / "cornelis_b200" / "csrc").glob("*.h"), *(ROOT / "cornelis_b200" / "csrc").glob("*.cuh")]
    if not OUT.exists() or any(d.stat().st_mtime > OUT.stat().st_mtime for d in deps):
        OUT.parent.mkdir(parents=True, exist_ok=True)
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "--fmad=false", "-std=c++17",
                        "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o", str(OUT), str(SRC)],
                       check=True, capture_output=True)
    L = C.CDLL(str(OUT))
    vp, sz = C.c_void_p, C.c_size_t
    L.grid_host_intersect.argtypes = [C.POINTER(binding.CameraDesc), vp, sz, vp, sz, sz, vp, vp, vp, vp, vp, vp]
    L.grid_host_check_structure.argtypes = [C.POINTER(binding.CameraDesc), vp, sz, vp, sz]
    L.grid_host_check_structure.restype = C.c_uint64
    return L


def walk_grid(L, flat, org, dirs):
    cam, S, P, _, (sph, pl, _m) = binding.descriptors(flat)
    org = np.ascontiguousarray(org, np.float32)
    dirs = np.ascontiguousarray(dirs, np.float32)
    n = len(org)
    t = np.empty(n, np.float32)
    prim = np.empty(n, np.int32)
    walk = np.zeros((n, 2), np.uint32)
    info = np.zeros(6, np.uint64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.grid_host_intersect(C.byref(cam), S, len(sph), P, len(pl), n, p(org), p(dirs), p(t), p(prim), p(walk), p(info))
    assert rc == 0, "grid could not be built"
    return dict(t=t, prim=prim, cells=walk[:, 0], tests=walk[:, 1], dims=tuple(int(x) for x in info[:3]),
                references=int(info[3]), margin=np.uint32(info[4]).view(np.float32), ncells=int(info[5]))


def surface_rays(rng, flat, oracle_scene, n):
    """Bounce-like rays: origins 1e-4 off a surface along the new direction (Render.cpp:207), like the render loop's."""
    org, dirs = scenes.microbench_rays(n, seed=int(rng.integers(1 << 30)))
    cam = np.asarray(flat["camera"][:3], np.float32)
    span = np.float32(np.abs(flat["spheres"][:, :3]).max())
    org = (org / np.float32(1000.0) * span).astype(np.float32)
    h = oracle_scene.intersect(org, dirs)
    hit = h["prim"] >= 0
    g = rng.standard_normal((hit.sum(), 3)).astype(np.float32)
    g /= np.linalg.norm(g, axis=1, keepdims=True).astype(np.float32)
    new_org = (h["P"][hit] + g * np.float32(1e-4)).astype(np.float32)
    return np.concatenate([new_org, np.tile(cam, (8, 1))]), np.concatenate([g, dirs[:8]])


def check_same(got, want):
    assert np.array_equal(got["prim"], want["prim"]), \
        f"{(got['prim'] != want['prim']).sum()} hit ids differ from the exhaustive scan"
    assert bit_equal(got["t"], want["t"])


def test_grid_structure(grid_host):
    for flat in (scenes.many_spheres(2000), scenes.microbench_scene(512)):
        cam, S, P, _, (sph, pl, _m) = binding.descriptors(flat)
        assert grid_host.grid_host_check_structure(C.byref(cam), S, len(sph), P, len(pl)) == 0


@pytest.mark.parametrize("n_spheres", [10000, 300])
def test_grid_matches_exhaustive_scan_config4(grid_host, port_oracle, n_spheres):
    """Config 4's scene (BASELINE.json configs[3]): camera rays, random rays and bounce-like rays."""
    flat = scenes.many_spheres(n_spheres)
    ref = port_oracle.scene(flat)
    rng = np.random.default_rng(7)
    n = 6000 if n_spheres > 1000 else 20000
    pi, pj = rng.integers(0, 1920, n), rng.integers(0, 1080, n)
    o1, d1 = ref.pixel_rays(1920, 1080, pi, pj, rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32))
    o2, d2 = scenes.microbench_rays(n, seed=5)           # origins in [-1000, 1000]^3: in and around the spheres
    o3, d3 = surface_rays(rng, flat, ref, n)
    org, dirs = np.concatenate([o1, o2, o3]), np.concatenate([d1, d2, d3])
    want = ref.intersect(org, dirs)
    got = walk_grid(grid_host, flat, org, dirs)
    check_same(got, want)
    sphere_hit = (want["prim"] >= 0) & (want["prim"] < n_spheres)
    assert (want["prim"] >= 0).mean() > 0.3 and sphere_hit.sum() > 1000
    # the point of the grid: a ray meets a few dozen spheres, not all of them
    assert got["tests"].mean() < 0.02 * n_spheres + 40, got["tests"].mean()


def test_grid_matches_exhaustive_scan_microbench(grid_host, port_oracle):
    """Config 3's scene through the grid: 1024 spheres inside six box faces (planes bound every walk)."""
    flat = scenes.microbench_scene(1024)
    ref = port_oracle.scene(flat)
    org, dirs = scenes.microbench_rays(40000)
    check_same(walk_grid(grid_host, flat, org, dirs), ref.intersect(org, dirs))


def test_grid_odd_rays_and_ties(grid_host, port_oracle):
    """Rays outside the trusted region, degenerate / axis-parallel / non-finite rays, duplicate spheres (exact ties:
    the lowest index must win), spheres touching cell boundaries, origins inside spheres."""
    rng = np.random.default_rng(11)
    base = scenes.many_spheres(400)
    sph = base["spheres"].copy()
    sph[200:300] = sph[100:200]                      # exact duplicates: ties on every hit
    sph[300:350, 3] = 300.0                          # big spheres spanning many cells, containing many origins
    flat = dict(base, spheres=sph)
    ref = port_oracle.scene(flat)
    n = 4000
    org, dirs = scenes.microbench_rays(n, seed=3)
    org[:500] *= np.float32(50.0)                    # far outside the trusted region -> exhaustive path
    dirs[500:800, 0] = 0.0                           # axis-parallel directions
    dirs[800:900, :2] = 0.0
    dirs[900:950] = 0.0                              # degenerate: ignored by every primitive
    dirs[950:960] = np.float32(np.nan)
    org[960:970] = np.float32(np.inf)
    dirs[970:1000] *= np.float32(1e-3)               # non-unit directions
    dirs[1000:1030] *= np.float32(1e4)
    dirs[1030:1040, 1] = np.float32(1e-35)           # components too small to invert
    c = sph[rng.integers(0, 400, 500), :3]
    org[1100:1600] = c + rng.standard_normal((500, 3)).astype(np.float32)   # origins inside spheres
    # rays aimed exactly at cell-boundary-aligned coordinates
    org[1600:1700] = np.round(org[1600:1700] / 50) * 50
    want = ref.intersect(org, dirs)
    got = walk_grid(grid_host, flat, org, dirs)
    check_same(got, want)
    dup = (want["prim"] >= 100) & (want["prim"] < 300)
    assert dup.sum() > 20 and (want["prim"][dup] < 200).all()   # ties went to the lower index in the oracle too


def test_grid_terminates_early(grid_host):
    """Front-to-back walk: a ray that hits something close visits few cells; the margin is a small fraction of a cell."""
    flat = scenes.many_spheres(10000)
    org, dirs = scenes.microbench_rays(5000, seed=9)
    org[:, 1] = np.abs(org[:, 1])
    got = walk_grid(grid_host, flat, org, dirs)
    nx, ny, nz = got["dims"]
    assert got["ncells"] == nx * ny * nz and 10000 <= got["ncells"] <= 80000
    assert got["references"] < 40 * 10000
    hit = got["prim"] >= 0
    assert got["cells"][hit].mean() < 0.6 * got["cells"][~hit].mean() + 2
    assert got["cells"].max() <= nx + ny + nz + 3


def grazing_rays(rng, flat, n, spread):
    """Rays aimed at the silhouette of random spheres: the computed discriminant is zero up to rounding, which is the case
    the registration radius (eps) and the termination slack exist for.  `spread` perturbs the aim relative to the radius."""
    sph = flat["spheres"]
    pick = rng.integers(0, len(sph), n)
    c, r = sph[pick, :3].astype(np.float64), sph[pick, 3].astype(np.float64)
    lo = np.array([-1900.0, 1.0, -2400.0])
    hi = np.array([1900.0, 2020.0, 1900.0])
    org = rng.uniform(lo, hi, (n, 3))
    to_c = c - org
    dist = np.linalg.norm(to_c, axis=1)
    keep = dist > 1.5 * r
    axis = to_c / dist[:, None]
    # a unit vector perpendicular to the axis, then the tangent direction: angle asin(r / dist) off the axis
    rnd = rng.standard_normal((n, 3))
    perp = rnd - (rnd * axis).sum(axis=1, keepdims=True) * axis
    perp /= np.linalg.norm(perp, axis=1, keepdims=True)
    sin_a = np.clip(r / dist * (1.0 + spread * rng.standard_normal(n)), 0.0, 0.999)
    cos_a = np.sqrt(1.0 - sin_a * sin_a)
    d = axis * cos_a[:, None] + perp * sin_a[:, None]
    return org[keep].astype(np.float32), d[keep].astype(np.float32)


@pytest.mark.parametrize("n_spheres,scale", [(10000, 1.0), (1500, 1.0), (1500, 40.0)])
def test_grid_grazing_rays(grid_host, port_oracle, n_spheres, scale):
    """Tangent rays, from near and far origins, unit and scaled directions: the walk's per-ray termination slack
    (geometry.cuh) and the registration radius must still deliver the exhaustive scan's hit bit for bit."""
    flat = scenes.many_spheres(n_spheres)
    ref = port_oracle.scene(flat)
    rng = np.random.default_rng(23)
    parts = [grazing_rays(rng, flat, 6000, s) for s in (0.0, 1e-7, 1e-5, 1e-3)]
    org, dirs = np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
    dirs = (dirs * np.float32(scale)).astype(np.float32)
    want = ref.intersect(org, dirs)
    got = walk_grid(grid_host, flat, org, dirs)
    check_same(got, want)
    assert ((want["prim"] >= 0) & (want["prim"] < n_spheres)).mean() > 0.3
