"""GPU parity tests proper: the CUDA path (through the C-ABI) against the oracle on identical inputs.

Bars (BASELINE.json north_star):
  * intersection: hit primitive ids bit-exact, t within 2 ulp (we require bit-exact: 0 ulp)
  * material sample / eval on fixed random inputs: 1e-5 relative
  * converged images: per-pixel within 3 sigma of the Monte-Carlo estimator, RMSE stated
"""
import numpy as np
import pytest

from cornelis_b200 import scenes
from conftest import bit_equal, ulp_distance

pytestmark = pytest.mark.gpu

REL = 1e-5  # material parity tolerance (north_star)


@pytest.fixture(scope="module")
def binding():
    from cornelis_b200 import binding as b
    assert b.device_count() >= 1
    return b


@pytest.fixture(scope="module")
def oracle(port_oracle, ref_oracle):
    return ref_oracle if ref_oracle is not None else port_oracle


def unit(rng, n):
    v = rng.standard_normal((n, 3)).astype(np.float32)
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def rel_err(a, b, floor=1e-3):
    """|a-b| / max(|b|, floor) per vector (3-vectors are judged on their norm, scalars on their magnitude)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.ndim == 2:
        return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), floor)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


# ------------------------------------------------------------------------------------------------- camera rays --

def test_pixel_rays_bit_exact(binding, oracle, golden):
    g = golden("pixel_rays_1080p.npz")
    flat = scenes.cornell_box(aspect=0.5625)
    sc = binding.Scene(flat)
    o, d = sc.pixel_rays(int(g["W"]), int(g["H"]), g["pi"], g["pj"], g["phi1"], g["phi2"])
    assert bit_equal(o, g["org"]) and bit_equal(d, g["dir"])
    rng = np.random.default_rng(5)
    for W, H, aspect in ((512, 512, 1.0), (3840, 2160, 0.5625), (50, 30, 0.6)):
        flat = scenes.cornell_box(aspect=aspect)
        n = 50000
        pi, pj = rng.integers(0, W, n), rng.integers(0, H, n)
        p1, p2 = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
        a = binding.Scene(flat).pixel_rays(W, H, pi, pj, p1, p2)
        b = oracle.scene(flat).pixel_rays(W, H, pi, pj, p1, p2)
        assert bit_equal(a[0], b[0]) and bit_equal(a[1], b[1])


# ------------------------------------------------------------------------------------------------ intersection --

@pytest.mark.parametrize("name,flat", [("intersect_microbench.npz", scenes.microbench_scene(1024)),
                                       ("intersect_cornell.npz", scenes.cornell_box())])
def test_intersect_golden(binding, golden, name, flat):
    g = golden(name)
    h = binding.Scene(flat).intersect(g["org"], g["dir"])
    assert np.array_equal(h["prim"], g["prim"]), "hit primitive ids must be bit-exact"
    assert ulp_distance(h["t"], g["t"]).max() == 0
    assert bit_equal(h["P"], g["P"]) and bit_equal(h["N"], g["N"]) and np.array_equal(h["mat"], g["mat"])


def test_intersect_against_oracle_large(binding, oracle):
    """Config 3 at oracle-friendly size (2^20 rays x 1024 spheres + 6 planes) plus ragged / degenerate batches."""
    flat = scenes.microbench_scene(1024)
    sc, osc = binding.Scene(flat), oracle.scene(flat)
    org, dirs = scenes.microbench_rays(1 << 20)
    a, b = sc.intersect(org, dirs), osc.intersect(org, dirs)
    assert np.array_equal(a["prim"], b["prim"])
    assert ulp_distance(a["t"], b["t"]).max() == 0
    assert bit_equal(a["P"], b["P"]) and bit_equal(a["N"], b["N"]) and np.array_equal(a["mat"], b["mat"])
    assert (a["prim"] >= 0).all()  # closed box: every ray hits something
    # ragged sizes, non-unit and zero directions
    rng = np.random.default_rng(11)
    for n in (1, 31, 257, 1000):
        o = (rng.random((n, 3), dtype=np.float32) * 1800 - 900).astype(np.float32)
        d = (unit(rng, n) * rng.random((n, 1), dtype=np.float32) * 4).astype(np.float32)
        d[::7] = 0
        d[1::7] *= np.float32(1e-5)
        a, b = sc.intersect(o, d), osc.intersect(o, d)
        assert np.array_equal(a["prim"], b["prim"]) and bit_equal(a["t"], b["t"])
        assert (a["prim"][::7] == -1).all() and np.isinf(a["t"][::7]).all()
    empty = sc.intersect(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert len(empty["t"]) == 0


def test_intersect_full_size_properties(binding, oracle):
    """Config 3 at BASELINE size: 2^24 rays.  The oracle checks a strided 2^19 subset bit for bit; the whole batch
    is checked through size-independent properties (determinism, hits lie on their primitive)."""
    flat = scenes.microbench_scene(1024)
    sc = binding.Scene(flat)
    n = 1 << 24
    org, dirs = scenes.microbench_rays(n)
    a = sc.intersect(org, dirs, surface=False)
    again = sc.intersect(org, dirs, surface=False)
    assert np.array_equal(a["prim"], again["prim"]) and bit_equal(a["t"], again["t"])  # idempotent / deterministic
    sel = slice(0, n, 32)
    b = oracle.scene(flat).intersect(org[sel], dirs[sel])
    assert np.array_equal(a["prim"][sel], b["prim"]) and ulp_distance(a["t"][sel], b["t"]).max() == 0
    # every sphere hit lies on its sphere (|P - c| = r within fp32 slack), every plane hit on its plane
    t, prim = a["t"].astype(np.float64), a["prim"]
    P = org.astype(np.float64) + dirs.astype(np.float64) * t[:, None]
    sph = prim < 1024
    c = flat["spheres"][prim[sph]].astype(np.float64)
    r = np.linalg.norm(P[sph] - c[:, :3], axis=1)
    err = np.abs(r - c[:, 3])  # grazing hits amplify the fp32 rounding of the discriminant (reference algorithm)
    assert np.quantile(err, 0.999) < 0.02 and err.max() < 1.0
    pl = flat["planes"][prim[~sph] - 1024].astype(np.float64)
    dist = ((P[~sph] - pl[:, 3:6]) * pl[:, 0:3]).sum(1)
    assert np.abs(dist).max() < 0.05
    assert (t >= 0).all() and np.isfinite(t).all()
    assert 0.2 < sph.mean() < 0.35


def test_intersect_full_batch_against_oracle(binding, oracle):
    """Config 3 at BASELINE size against the oracle on EVERY ray: 2^24 rays x (1024 spheres + 6 planes), hit ids and
    t bit for bit (the oracle's 1.7e10 primitive tests run in chunks on all host threads)."""
    from concurrent.futures import ThreadPoolExecutor
    import os
    flat = scenes.microbench_scene(1024)
    sc, osc = binding.Scene(flat), oracle.scene(flat)
    n = 1 << 24
    org, dirs = scenes.microbench_rays(n)
    a = sc.intersect(org, dirs, surface=False)
    workers = os.cpu_count() or 1
    step = 1 << 16

    def check(first):
        b = osc.intersect(org[first:first + step], dirs[first:first + step])
        return (np.array_equal(a["prim"][first:first + step], b["prim"]) and
                bit_equal(a["t"][first:first + step], b["t"]))

    with ThreadPoolExecutor(workers) as pool:  # ctypes releases the GIL
        ok = list(pool.map(check, range(0, n, step)))
    assert all(ok), [k for k, good in enumerate(ok) if not good][:8]


@pytest.mark.parametrize("which", ["cornell", "config4-grid"])
def test_compaction_queue_contents(binding, oracle, which):
    """Compaction #1 (the reference's rebuilt activeList, Render.cpp:142-149) checked directly: after the intersect
    stage of one wavefront pass the hit queue holds, as a set and without duplicates, exactly the rays with t < INF
    according to the oracle, and the finished queue exactly the others.  Cornell runs k_intersect (block-level
    ballot + popc append), the 10 000-sphere scene k_walk + k_compact_hits."""
    rng = np.random.default_rng(17)
    if which == "cornell":
        flat = scenes.cornell_box()
        osc = oracle.scene(flat)
        n = 200_003  # not a multiple of the block size
        o1, d1 = osc.camera_rays(rng.random(n // 2, dtype=np.float32), rng.random(n // 2, dtype=np.float32))
        o2 = (rng.random((n - n // 2, 3), dtype=np.float32) * np.float32(700) + np.float32([-350, -50, -600]))
        d2 = unit(rng, n - n // 2)
    else:
        flat = scenes.many_spheres(10000)
        osc = oracle.scene(flat)
        n = 70_001
        o1, d1 = osc.camera_rays(rng.random(n // 2, dtype=np.float32), rng.random(n // 2, dtype=np.float32))
        o2 = (rng.random((n - n // 2, 3), dtype=np.float32) * np.float32([2400, 2400, 2400]) +
              np.float32([-1200, -100, -1200]))
        d2 = unit(rng, n - n // 2)
    org = np.concatenate([o1, o2.astype(np.float32)]).astype(np.float32)
    dirs = np.concatenate([d1, d2]).astype(np.float32)
    dirs[5::1013] = 0  # degenerate directions miss everything (Geometry.cpp:67-70)
    sc = binding.Scene(flat)
    assert sc.acceleration()["grid"] == (which != "cornell")
    hits, misses = sc.intersect_compact(org, dirs)
    t = osc.intersect(org, dirs)["t"]
    want_hits, want_misses = np.flatnonzero(t < np.inf), np.flatnonzero(~(t < np.inf))
    assert len(want_hits) > n // 4 and len(want_misses) > n // 100
    assert len(hits) == len(want_hits) and len(misses) == len(want_misses)
    assert np.array_equal(np.sort(hits), want_hits)      # same set, and no index twice (the lengths agree)
    assert np.array_equal(np.sort(misses), want_misses)
    empty = sc.intersect_compact(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert len(empty[0]) == 0 and len(empty[1]) == 0


def test_exact_fast_paths(binding):
    """The hand-scheduled division / square root of the intersection kernel equal the IEEE operators bit for bit on
    2^28 crafted operands each (random and all-ones / all-zeros / single-bit mantissas over the claimed exponent ranges)."""
    sc = binding.Scene(scenes.cornell_box())
    assert sc.selftest_arith(0, 1 << 28, seed=7) == 0
    assert sc.selftest_arith(1, 1 << 28, seed=9) == 0
    # normalize's sqrt + reciprocal from one rsqrt seed: exhaustive over all 2^32 float bit patterns
    assert sc.selftest_arith(2, 1 << 32) == 0


def test_general_planes_and_odd_rays(binding, oracle):
    """Tilted planes take the general path, axis-aligned ones the exact fast path; rays with zero components, on-plane
    origins, huge / tiny / non-finite components must all agree with the oracle bit for bit."""
    flat = scenes.cornell_box()
    n1 = np.array([1, 1, 0], np.float32) / np.sqrt(np.float32(2))
    n2 = np.array([0.3, -0.5, 0.81], np.float32)
    n2 = n2 / np.float32(np.linalg.norm(n2))
    flat["planes"] = np.concatenate([flat["planes"], [[*n1, 50, 200, 0, 300, 300, 0], [*n2, -100, 300, 100, 250, 400, 0]]]
                                    ).astype(np.float32)
    flat["plane_mat"] = np.concatenate([flat["plane_mat"], [1, 2]]).astype(np.int32)
    sc, osc = binding.Scene(flat), oracle.scene(flat)
    rng = np.random.default_rng(77)
    n = 200000
    o = (rng.random((n, 3), dtype=np.float32) * 560 - np.float32([280, 0, 280])).astype(np.float32)
    d = unit(rng, n)
    k = np.arange(n)
    d[k % 11 == 0, 0] = 0                      # axis-parallel directions (parallel to walls)
    d[k % 13 == 0, 1] = 0
    d[k % 17 == 0, 2] = 0
    o[k % 19 == 0, 1] = 0                      # origins exactly on the floor plane
    o[k % 23 == 0, 0] = np.float32(-275)       # ... on the left wall
    o[k % 29 == 0] = np.float32([0, 0, 0])     # the floor's own point: diff == 0
    d[k % 31 == 0] *= np.float32(1e-3)
    d[k % 37 == 0] *= np.float32(1e6)          # beyond the fast-path range
    o[k % 41 == 0] *= np.float32(1e8)
    d[k % 43 == 0, 0] = np.float32(np.inf)
    o[k % 47 == 0, 2] = np.float32(np.nan)
    d[k % 53 == 0] = np.float32(3e-5)          # degenerate direction
    o[k % 59 == 0] *= np.float32(1e-30)
    a, b = sc.intersect(o, d), osc.intersect(o, d)
    assert np.array_equal(a["prim"], b["prim"])
    assert ulp_distance(a["t"], b["t"]).max() == 0 and np.array_equal(np.isnan(a["t"]), np.isnan(b["t"]))
    assert (a["prim"] >= 9).any() and (a["prim"] == 3).any()


def test_plane_tables_odd_cases(binding, oracle):
    """The compact axis-plane table compares |e| with HALF the extents, which is exact only when the halving is: planes
    with tiny, zero, infinite or NaN extents (the first is classified general at scene creation), scenes without planes
    and scenes without spheres must all agree with the oracle bit for bit."""
    rng = np.random.default_rng(91)
    n = 100000
    o = (rng.random((n, 3), dtype=np.float32) * 400 - 200).astype(np.float32)
    d = unit(rng, n)
    o[::7, 1] = -100.0  # origins on the first plane
    base = scenes.cornell_box()
    for extents in ([1e-38, 50.0], [0.0, 300.0], [np.inf, 120.0], [np.nan, 80.0], [3e-33, 3e-33], [250.0, np.inf]):
        flat = dict(base)
        flat["planes"] = np.array([[0, 1, 0, 0, -100, 0, extents[0], extents[1], 0],
                                   [1, 0, 0, -150, 0, 0, 260, 260, 0],
                                   [0, 0, -1, 0, 0, 150, extents[1], extents[0], 0]], np.float32)
        flat["plane_mat"] = np.array([1, 2, 3], np.int32)
        a, b = binding.Scene(flat).intersect(o, d), oracle.scene(flat).intersect(o, d)
        assert np.array_equal(a["prim"], b["prim"]), extents
        assert bit_equal(a["t"], b["t"]), extents
    no_planes = dict(base)
    no_planes["planes"], no_planes["plane_mat"] = np.zeros((0, 9), np.float32), np.zeros(0, np.int32)
    no_spheres = dict(base)
    no_spheres["spheres"], no_spheres["sphere_mat"] = np.zeros((0, 4), np.float32), np.zeros(0, np.int32)
    for flat in (no_planes, no_spheres):
        a, b = binding.Scene(flat).intersect(o, d), oracle.scene(flat).intersect(o, d)
        assert np.array_equal(a["prim"], b["prim"]) and bit_equal(a["t"], b["t"])
        img, st = binding.Scene(flat).render(32, 24, 8)
        assert st["pixel_samples"] == 32 * 24 * 8 and img.shape == (24, 32, 3)


def test_rays_starting_on_planes_keep_the_sign_of_zero(binding, oracle):
    """A bounce off a plane starts ON it about one time in four (o_k == p0_k after rounding): t against that plane is
    a signed zero whose sign the reference derives from -(diff . N) / (d . N).  The fast path must return the same
    bits — t = -0.0 included — for normals of either sign and every sign pattern of the other components."""
    flat = scenes.cornell_box()
    sc, osc = binding.Scene(flat), oracle.scene(flat)
    rng = np.random.default_rng(5)
    n = 120000
    o = (rng.random((n, 3), dtype=np.float32) * 550 - np.float32([275, 0, 275])).astype(np.float32)
    d = unit(rng, n)
    k = np.arange(n) % 6
    o[k == 0, 0] = np.float32(-275.0)   # left wall   N = +x
    o[k == 1, 0] = np.float32(275.0)    # right wall  N = -x
    o[k == 2, 1] = np.float32(555.0)    # roof        N = -y
    o[k == 3, 1] = np.float32(0.0)      # floor       N = +y   (+0 and, below, -0 origins)
    o[k == 4, 2] = np.float32(275.0)    # back wall   N = -z
    o[(np.arange(n) % 12) == 9, 1] = np.float32(-0.0)
    o[(np.arange(n) % 24) == 5] = np.float32([0, 555, 0])   # the roof's own point, exactly
    a, b = sc.intersect(o, d, surface=False), osc.intersect(o, d)
    assert np.array_equal(a["prim"], b["prim"])
    assert bit_equal(a["t"], b["t"])
    zero = b["t"] == 0
    assert zero.sum() > n // 4 and np.signbit(b["t"][zero]).any() and (~np.signbit(b["t"][zero])).any()


def test_packed_fp32_scan_is_bit_identical(binding, oracle, monkeypatch):
    """cornelis_cuda_intersect scans the spheres two at a time on packed FP32 (FFMA2, csrc/packed_f32.cuh): every packed
    operation must be the scalar IEEE operation on the same operands.  Checked against the scalar scan of the same
    library (CORNELIS_BATCH_PACKED=0 at scene creation) on the microbench scene and against the oracle for odd and
    tiny sphere counts (the last sphere of an odd table has no partner; groups of 16 leave a remainder)."""
    org, dirs = scenes.microbench_rays(1 << 19)
    flat = scenes.microbench_scene(1024)
    packed = binding.Scene(flat)
    packed.set_acceleration(binding.ACCEL_NONE)
    a = packed.intersect(org, dirs, surface=False)
    monkeypatch.setenv("CORNELIS_BATCH_PACKED", "0")
    scalar = binding.Scene(flat)
    scalar.set_acceleration(binding.ACCEL_NONE)
    b = scalar.intersect(org, dirs, surface=False)
    monkeypatch.delenv("CORNELIS_BATCH_PACKED")
    assert np.array_equal(a["prim"], b["prim"]) and bit_equal(a["t"], b["t"])
    assert (a["prim"] >= 0).all() and (a["prim"] < 1024).mean() > 0.2
    for count in (1, 2, 3, 17, 33, 1023):
        flat = scenes.microbench_scene(count)
        sc = binding.Scene(flat)
        sc.set_acceleration(binding.ACCEL_NONE)
        got, ref = sc.intersect(org[:20000], dirs[:20000], surface=False), oracle.scene(flat).intersect(org[:20000], dirs[:20000])
        assert np.array_equal(got["prim"], ref["prim"]) and bit_equal(got["t"], ref["t"]), count


def test_two_rays_per_lane_scan_is_bit_identical(binding, oracle, monkeypatch):
    """closestHit2 (csrc/geometry2.cuh): two rays per lane, every FP32 operation of the fast paths issued once for the
    pair as an FFMA2.  Through CORNELIS_BATCH_PAIRS=1 the intersect entry point runs it; hit ids and t must equal the
    oracle's bit for bit — on the microbench scene, on the Cornell box (camera rays, rays that start ON walls and the
    floor, degenerate directions) and on batches of odd length (the last ray has no partner)."""
    monkeypatch.setenv("CORNELIS_BATCH_PAIRS", "1")
    rng = np.random.default_rng(23)
    flat = scenes.microbench_scene(1024)
    sc, osc = binding.Scene(flat), oracle.scene(flat)
    sc.set_acceleration(binding.ACCEL_NONE)
    for n in ((1 << 18) + 1, 1, 2, 33):
        org, dirs = scenes.microbench_rays(n)
        a, b = sc.intersect(org, dirs, surface=False), osc.intersect(org, dirs)
        assert np.array_equal(a["prim"], b["prim"]) and bit_equal(a["t"], b["t"]), n
    flat = scenes.cornell_box()
    sc, osc = binding.Scene(flat), oracle.scene(flat)
    n = 1 << 18
    o1, d1 = osc.camera_rays(rng.random(n // 2, dtype=np.float32), rng.random(n // 2, dtype=np.float32))
    o2 = (rng.random((n // 2, 3), dtype=np.float32) * np.float32(500) + np.float32([-250, 20, -250])).astype(np.float32)
    o2[::4, 1] = 0.0       # on the floor
    o2[1::4, 0] = -275.0   # on the left wall
    o2[2::8, 2] = 275.0    # on the back wall
    d2 = unit(rng, n // 2)
    d2[3::4099] = 0
    d2[5::4099, 0] = np.float32(1e-6)  # a component below RayEpsilon: the warp leaves the fast path
    org, dirs = np.concatenate([o1, o2]).astype(np.float32), np.concatenate([d1, d2]).astype(np.float32)
    a, b = sc.intersect(org, dirs, surface=False), osc.intersect(org, dirs)
    assert np.array_equal(a["prim"], b["prim"]) and bit_equal(a["t"], b["t"])
    assert (a["t"] == 0).sum() > 1000  # zero-distance re-hits of the plane a ray starts on: the sign of zero included


def test_rays_on_and_tangent_to_spheres(binding, oracle):
    """Zero numerators of the sphere test (Geometry.cpp:77-84): an origin exactly on a sphere (C == r^2), a direction
    perpendicular to the centre offset (B == 0), both at once (t = 0 from a zero discriminant) and exact tangents.
    The fast path returns zeros whose sign may differ from the operators' inside the test, which must never reach t;
    spheres with tiny radii (r^2 < 2^-50) switch the scene to the operator scan.  All against the oracle, bit for
    bit, through the exhaustive scan and through the grid."""
    flat = scenes.microbench_scene(64)
    sph = flat["spheres"].copy()
    sph[0] = [0, 0, 0, 1]            # unit sphere at the origin
    sph[1] = [8, -4, 2, 0.5]         # power-of-two data: exact zeros are reachable
    sph[2] = [-16, 32, 64, 4]
    flat["spheres"] = sph.astype(np.float32)
    rng = np.random.default_rng(11)
    n = 60000
    o = np.zeros((n, 3), np.float32)
    d = unit(rng, n)
    k = np.arange(n) % 12
    which = sph[(np.arange(n) // 12) % 3]
    c, r = which[:, :3], which[:, 3:4]
    axis = np.eye(3, dtype=np.float32)[(np.arange(n) // 36) % 3]
    other = np.roll(axis, 1, axis=1)
    sign = np.where((np.arange(n) // 108) % 2 == 0, np.float32(1), np.float32(-1))[:, None]
    o[:] = c + sign * r * axis                         # exactly on the surface: C == r^2
    d[k == 0] = (-sign * axis)[k == 0]                 # straight through the centre
    d[k == 1] = (sign * axis)[k == 1]                  # straight away from it
    d[k == 2] = other[k == 2]                          # tangent from the surface: B == 0 and C == r^2
    d[k == 3] = (-other)[k == 3]
    o[k == 4] = (c + np.float32(0.5) * r * axis)[k == 4]   # inside, perpendicular: B == 0
    d[k == 4] = other[k == 4]
    o[k == 5] = (c + sign * r * axis - np.float32(4) * other)[k == 5]   # exact tangent from outside
    d[k == 5] = other[k == 5]
    d[k == 6] = (other * np.float32(3))[k == 6]        # non-unit directions (test_Geometry.cpp:46)
    d[k == 7] = (-sign * axis * np.float32(0.25))[k == 7]
    # k >= 8: random directions from the surface
    for tiny in (False, True):
        if tiny:
            flat["spheres"][3] = [100, 100, 100, 1e-9]   # r^2 = 1e-18 < 2^-50
        sc, osc = binding.Scene(flat), oracle.scene(flat)
        ref = osc.intersect(o, d)
        for accel in (binding.ACCEL_NONE, binding.ACCEL_GRID):
            sc.set_acceleration(accel)
            got = sc.intersect(o, d, surface=False)
            assert np.array_equal(got["prim"], ref["prim"]), (tiny, accel)
            assert bit_equal(got["t"], ref["t"]), (tiny, accel)
    assert (ref["t"] == 0).sum() > n // 8 and (ref["prim"] < 3).sum() > n // 4


# ---------------------------------------------------------------------------------------------------- materials --

def _bsdf_inputs(rng, n, n_mat):
    N = unit(rng, n)
    N[: n // 16] = np.float32([0, 1, 0])
    N[n // 16: n // 8] = np.float32([0, 0, -1])
    wo = unit(rng, n)
    flip = (wo * N).sum(1) < 0
    wo[flip] = -wo[flip]
    x = rng.random((n, 3), dtype=np.float32)
    mat = rng.integers(0, n_mat, n).astype(np.int32)
    return mat, wo, N, x


def test_bsdf_sample_and_eval(binding, oracle, golden):
    flat = scenes.cornell_box()
    sc, osc = binding.Scene(flat), oracle.scene(flat)
    g = golden("bsdf_cornell.npz")
    s = sc.bsdf_sample(g["mat"], g["wo"], g["N"], g["x"])
    assert rel_err(s["wi"], g["wi"], 1.0).max() <= REL
    assert rel_err(s["pdf"], g["pdf"]).max() <= REL and rel_err(s["f"], g["f"]).max() <= REL
    e = sc.bsdf_eval(g["mat"], g["eval_wi"], g["wo"], g["N"])
    assert rel_err(e["f"], g["eval_f"]).max() <= REL and rel_err(e["pdf"], g["eval_pdf"]).max() <= REL

    rng = np.random.default_rng(23)
    for flat in (scenes.cornell_box(), scenes.many_spheres(16, 64)):
        sc, osc = binding.Scene(flat), oracle.scene(flat)
        n = 1 << 18
        mat, wo, N, x = _bsdf_inputs(rng, n, len(flat["materials"]) + 1)
        a, b = sc.bsdf_sample(mat, wo, N, x), osc.bsdf_sample(mat, wo, N, x)
        # lobe choice / early-outs are decided by identical comparisons on identical inputs
        assert np.array_equal((a["wi"] == 0).all(1), (b["wi"] == 0).all(1))
        err = np.maximum.reduce([rel_err(a["wi"], b["wi"], 1.0), rel_err(a["pdf"], b["pdf"]), rel_err(a["f"], b["f"])])
        print(f"bsdf_sample max relative error {err.max():.2e}, wi bit-identical: "
              f"{bool(np.array_equal(a['wi'].view(np.uint32), b['wi'].view(np.uint32)))}")
        assert err.max() <= REL, (err.max(), int(err.argmax()))
        wi = unit(rng, n)
        a, b = sc.bsdf_eval(mat, wi, wo, N), osc.bsdf_eval(mat, wi, wo, N)
        err = np.maximum(rel_err(a["f"], b["f"]), rel_err(a["pdf"], b["pdf"]))
        print(f"bsdf_eval max relative error {err.max():.2e} (budget {REL:.0e})")
        assert err.max() <= REL, (err.max(), int(err.argmax()))
        assert not np.isnan(a["f"]).any()


def test_shade_step(binding, oracle, golden):
    """One accumulateAndBounce pass: RR decision identical, new ray / throughput / radiance within 1e-5."""
    flat = scenes.cornell_box()
    sc, osc = binding.Scene(flat), oracle.scene(flat)
    g = golden("shade_cornell.npz")
    order = g["order"]
    for depth in (0, 5):
        u = g[f"d{depth}_u"]
        ux = np.stack([u[:, 0], u[:, 1 + order[0]], u[:, 1 + order[1]], u[:, 1 + order[2]]], axis=1)
        r = sc.shade(depth, ux, g["P"], g["N"], g["mat"], g["org"], g["dir"], g["thr"], g["rad"])
        alive = g[f"d{depth}_alive"]
        assert np.array_equal(r["alive"], alive)
        assert rel_err(r["rad"], g[f"d{depth}_rad"]).max() <= REL
        for k in ("org", "dir", "thr"):
            assert rel_err(r[k][alive], g[f"d{depth}_{k}"][alive], 1e-3).max() <= REL, (depth, k)
    rng = np.random.default_rng(31)
    n = 1 << 17
    mat, wo, N, _ = _bsdf_inputs(rng, n, 6)
    P = (rng.standard_normal((n, 3)) * 200).astype(np.float32)
    thr = (rng.random((n, 3), dtype=np.float32) * 1.5).astype(np.float32)
    rad = rng.random((n, 3), dtype=np.float32)
    order = oracle.sample_draw_order()
    for depth in (0, 2, 3, 7):
        b = osc.shade(depth, 777, P, N, mat, P, -wo, thr, rad)
        u = b["u"]
        ux = np.stack([u[:, 0], u[:, 1 + order[0]], u[:, 1 + order[1]], u[:, 1 + order[2]]], axis=1)
        a = sc.shade(depth, ux, P, N, mat, P, -wo, thr, rad)
        assert np.array_equal(a["alive"], b["alive"])
        live = b["alive"]
        assert rel_err(a["rad"], b["rad"]).max() <= REL
        assert rel_err(a["dir"][live], b["dir"][live], 1.0).max() <= REL
        assert rel_err(a["org"][live], b["org"][live], 1.0).max() <= REL
        assert rel_err(a["thr"][live], b["thr"][live]).max() <= REL


# ---------------------------------------------------------------------------------------------------------- rng --

def _philox_numpy(c0, c1, c2, k0, k1, rounds=10, c3=0):
    """Philox4x32-`rounds` (Salmon et al., SC'11) on uint32 arrays: counter (c0, c1, c2, c3), key (k0, k1)."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    mask, sh = np.uint64(0xFFFFFFFF), np.uint64(32)
    c = [np.asarray(x, np.uint64) for x in (c0, c1, c2, np.full(np.shape(c0), c3, np.uint64))]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(rounds):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        c = [(p1 >> sh) ^ c[1] ^ k0, p1 & mask, (p0 >> sh) ^ c[3] ^ k1, p0 & mask]
        k0, k1 = (k0 + np.uint64(W0)) & mask, (k1 + np.uint64(W1)) & mask
    return np.stack(c, axis=1).astype(np.uint32)


def test_counter_rng(binding):
    """The render loop's generator is Philox4x32 with rng_rounds() = 7 rounds.  (1) The numpy restatement reproduces the
    Random123 known-answer vectors of philox4x32-10; (2) so does the library's round function on the device
    (cornelis_cuda_rng_bits at 10 rounds); (3) the device's 7-round words, and the uniforms the render kernels' own code
    path produces, equal the restatement's at 7 rounds, with the reference's 24-bit float mapping."""
    sc = binding.Scene(scenes.cornell_box())
    rounds = binding.rng_rounds()
    assert rounds == 7
    # Random123 kat_vectors, philox4x32 10: (counter, key) -> output
    z, ones = np.zeros(1, np.uint32), np.full(1, 0xFFFFFFFF, np.uint32)
    assert [hex(v) for v in _philox_numpy(z, z, z, 0, 0)[0]] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(v) for v in _philox_numpy(ones, ones, ones, 0xFFFFFFFF, 0xFFFFFFFF, c3=0xFFFFFFFF)[0]] == \
        ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    pi = [np.array([v], np.uint32) for v in (0x243f6a88, 0x85a308d3, 0x13198a2e)]  # digits of pi: counter and key
    assert [hex(v) for v in _philox_numpy(*pi, 0xa4093822, 0x299f31d0, c3=0x03707344)[0]] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]
    assert [hex(v) for v in sc.rng_bits(10, 0, z, z, z)[0]] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    rng = np.random.default_rng(3)
    n = 1 << 16
    pixel = rng.integers(0, 1 << 23, n).astype(np.uint32)
    sample = rng.integers(0, 1 << 14, n).astype(np.uint32)
    block = rng.integers(0, 16, n).astype(np.uint32)
    seed = 19791102 | (7 << 32)
    for r in (7, 10):
        assert np.array_equal(sc.rng_bits(r, seed, pixel, sample, block),
                              _philox_numpy(pixel, sample, block, seed & 0xFFFFFFFF, seed >> 32, rounds=r))
    u = sc.rng_uniforms(seed, pixel, sample, block)
    bits = _philox_numpy(pixel, sample, block, seed & 0xFFFFFFFF, seed >> 32, rounds=rounds)
    expect = (bits >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)  # XoshiroCpp.hpp:651-655 mapping
    assert bit_equal(u, expect)
    assert (u >= 0).all() and (u < 1).all()
    assert abs(u.mean() - 0.5) < 0.005 and abs(u.var() - 1 / 12) < 0.002
    # neighbouring counters (consecutive pixels, samples and bounces are what a render draws) are uncorrelated
    grid = sc.rng_uniforms(seed, np.arange(n, dtype=np.uint32), np.zeros(n, np.uint32), np.ones(n, np.uint32))
    for lag in (1, 2, 1920):
        for a in range(4):
            for b in range(4):
                assert abs(np.corrcoef(grid[:-lag, a], grid[lag:, b])[0, 1]) < 0.02
    assert abs(np.corrcoef(grid[:, 0], grid[:, 1])[0, 1]) < 0.02


# ------------------------------------------------------------------------------------------------------- images --

def _three_sigma(mu_g, var_g, n_g, mu_r, var_r, n_r):
    sigma = np.sqrt(var_r.astype(np.float64) / n_r + var_g.astype(np.float64) / n_g)
    diff = np.abs(mu_g.astype(np.float64) - mu_r.astype(np.float64))
    ok = diff <= 3.0 * sigma + 1e-4 * (1.0 + np.abs(mu_r))
    return ok, diff, sigma


PIPELINES = [("wavefront", 1), ("persistent", 2)]


@pytest.mark.parametrize("pipeline", [p[1] for p in PIPELINES], ids=[p[0] for p in PIPELINES])
def test_converged_image_within_three_sigma(binding, golden, pipeline):
    """Cornell box 128x128: GPU 4096 spp against the reference's own 4096 spp render (mean + per-sample variance from
    the compiled reference, tests/golden/make_golden.py)."""
    g = golden("render_cornell_128x128_4096spp.npz")
    sc = binding.Scene(scenes.cornell_box())
    spp = 4096
    st = sc.render_accumulate(128, 128, spp, variance=True, pipeline=pipeline)
    mean, var = sc.resolve(spp, variance=True)
    ok, diff, sigma = _three_sigma(mean, var, spp, g["mean"], g["variance"], int(g["spp"]))
    # The reference's Oren-Nayar term turns NaN when |w.z| rounds above 1 (about once per 1e8 samples, see
    # include/cornelis_cuda.h CORNELIS_RENDER_DROP_NONFINITE); such pixels count as failures of the 3-sigma test.
    bad = ~np.isfinite(mean).all(axis=2)
    assert bad.sum() <= 3, bad.sum()
    frac = ok.mean()
    good = ~bad
    rmse = float(np.sqrt(np.mean(diff[good] ** 2)))
    rel_rmse = rmse / float(np.sqrt(np.mean(g["mean"].astype(np.float64) ** 2)))
    energy = float(mean[good].mean() / g["mean"][good].mean())
    print(f"3-sigma fraction {frac:.5f}  RMSE {rmse:.5f}  relRMSE {rel_rmse:.5f}  energy ratio {energy:.5f} "
          f"non-finite pixels {int(bad.sum())}")
    assert frac >= 0.99, frac                # 0.9973 for a normal estimator; the heavy tail costs a little
    assert abs(energy - 1.0) < 0.01
    assert rel_rmse < 0.05
    z = (mean.astype(np.float64) - g["mean"]) / np.maximum(sigma, 1e-9)
    zz = z[(sigma > 1e-6) & np.isfinite(z)]
    assert abs(zz.mean()) < 0.05 and 0.8 < zz.std() < 1.2  # unbiased, correctly scaled
    rays_per_sample = st["rays"] / st["pixel_samples"]
    ref_rays_per_sample = float(g["rays"]) / (128 * 128 * int(g["spp"]))
    assert abs(rays_per_sample - ref_rays_per_sample) < 0.01, (rays_per_sample, ref_rays_per_sample)
    assert st["pixel_samples"] == 128 * 128 * spp and 8 <= st["max_depth"] <= 40


def test_headline_frame_within_three_sigma(binding, golden):
    """The frame bench.py times — Cornell 1920x1080, camera aspect 0.5625, 4096 spp (BASELINE.json configs[1]) —
    against the reference's own 4096-spp estimate of every 8th pixel in each dimension (240 x 135 pixels, mean and
    per-sample variance from the compiled reference: tests/golden/make_golden.py headline_strided)."""
    g = golden("render_cornell_1080p_stride8_4096spp.npz")
    W, H, spp, stride = int(g["W"]), int(g["H"]), int(g["spp"]), int(g["stride"])
    sc = binding.Scene(scenes.cornell_box(aspect=H / W))
    st = sc.render_accumulate(W, H, spp, variance=True)
    mean, var = sc.resolve(spp, variance=True)
    assert st["pixel_samples"] == W * H * spp
    mean_s, var_s = mean[::stride, ::stride], var[::stride, ::stride]
    assert mean_s.shape == g["mean"].shape
    ok, diff, sigma = _three_sigma(mean_s, var_s, spp, g["mean"], g["variance"], spp)
    good = np.isfinite(mean_s).all(axis=2) & np.isfinite(g["mean"]).all(axis=2)
    frac = float(ok[good].mean())
    rmse = float(np.sqrt(np.mean(diff[good] ** 2)))
    rel_rmse = rmse / float(np.sqrt(np.mean(g["mean"][good].astype(np.float64) ** 2)))
    energy = float(mean_s[good].mean() / g["mean"][good].mean())
    z = (mean_s.astype(np.float64) - g["mean"]) / np.maximum(sigma, 1e-9)
    zz = z[(sigma > 1e-6) & np.isfinite(z)]
    print(f"headline frame, {good.sum()} strided pixels: 3-sigma fraction {frac:.5f}  RMSE {rmse:.5f}  relRMSE "
          f"{rel_rmse:.5f}  energy ratio {energy:.5f}  z mean {zz.mean():.4f} std {zz.std():.4f}  non-finite pixels in "
          f"the whole frame {int((~np.isfinite(mean).all(axis=2)).sum())}")
    assert (~good).sum() <= 8  # the reference's NaN quirk, on either side: about one pixel-sample in 1e8
    assert frac >= 0.99, frac
    assert abs(energy - 1.0) < 0.01 and rel_rmse < 0.12
    assert abs(zz.mean()) < 0.05 and 0.8 < zz.std() < 1.2
    rays_ref = float(g["rays"]) / float(g["pixel_samples"])
    assert abs(st["rays"] / st["pixel_samples"] - rays_ref) < 0.01
    assert int((~np.isfinite(mean).all(axis=2)).sum()) < 1000  # the reference's own NaN quirk: ~1 per 1e8 samples


def test_config1_cli_default_512x512_64spp(binding, oracle):
    """BASELINE.json configs[0], the reference CLI's own case: the Cornell box at 512x512, 64 spp, 32x32 tiles, seed
    19791102, rendered by the oracle on the host cores HERE and by the GPU at the same 64 spp (and at 1024 spp for a
    tighter comparison).  64 samples of a heavy-tailed estimator give poor variance estimates, so the per-pixel
    3-sigma fraction is checked loosely; image energy and rays per sample are the sharper statements."""
    W = H = 512
    ref = oracle.scene(scenes.cornell_box()).render(W, H, 64, tile=(32, 32), variance=True, stats=True)
    sc = binding.Scene(scenes.cornell_box())
    rays_ref = ref["stats"]["rays"] / ref["stats"]["pixel_samples"]
    for spp, floor in ((64, 0.95), (1024, 0.95)):  # 64-sample means of a heavy-tailed estimator are not Gaussian
        st = sc.render_accumulate(W, H, spp, variance=True)
        mean, var = sc.resolve(spp, variance=True)
        # the per-sample variance is a property of the estimator, the same on both sides; at 1024 spp the GPU's
        # estimate of it is the reliable one (64 samples rarely contain the bright paths that dominate it)
        ok, diff, sigma = _three_sigma(mean, var, spp, ref["mean"], ref["variance"] if spp == 64 else var, 64)
        good = np.isfinite(mean).all(axis=2)
        energy = float(mean[good].mean() / ref["mean"][good].mean())
        print(f"config 1 at {spp} spp: 3-sigma fraction {ok.mean():.5f}  energy ratio {energy:.4f}  rays/sample "
              f"{st['rays'] / st['pixel_samples']:.4f} (reference {rays_ref:.4f})")
        assert st["pixel_samples"] == W * H * spp
        assert ok.mean() >= floor, (spp, ok.mean())
        assert abs(energy - 1.0) < 0.02, energy
        assert abs(st["rays"] / st["pixel_samples"] - rays_ref) < 0.01


@pytest.mark.parametrize("pipeline", [p[1] for p in PIPELINES], ids=[p[0] for p in PIPELINES])
def test_render_against_oracle_other_scene(binding, oracle, pipeline):
    """A second scene (mixed materials, many-sphere style) at a frame that is not a multiple of anything."""
    flat = scenes.many_spheres(60, 12, aspect=0.6)
    W, H, spp_ref, spp = 50, 30, 1024, 2048
    ref = oracle.scene(flat).render(W, H, spp_ref, tile=(10, 10), variance=True, stats=True)
    sc = binding.Scene(flat)
    st = sc.render_accumulate(W, H, spp, variance=True, pipeline=pipeline)
    mean, var = sc.resolve(spp, variance=True)
    ok, diff, sigma = _three_sigma(mean, var, spp, ref["mean"], ref["variance"], spp_ref)
    assert ok.mean() >= 0.985, ok.mean()
    sc.render_accumulate(W, H, 256, drop_nonfinite=True, pipeline=pipeline)
    assert np.isfinite(sc.resolve(256)).all()
    assert abs(st["rays"] / st["pixel_samples"] - ref["stats"]["rays"] / ref["stats"]["pixel_samples"]) < 0.03


def test_sample_sharding_and_determinism(binding):
    """Sample ranges add up: [0,32) + [32,64) accumulated == [0,64) (same sample set; fp32 sums differ only by
    atomic ordering); the image depends neither on the pool size nor on the pipeline: both pipelines trace the SAME
    paths (identity and random numbers are keyed by pixel, sample, depth)."""
    sc = binding.Scene(scenes.cornell_box())
    W = H = 96
    st_w = sc.render_accumulate(W, H, 64, pipeline=1)
    whole = sc.resolve(64).copy()
    for pipeline in (1, 2):
        sc.render_accumulate(W, H, 64, first_sample=0, sample_count=32, pipeline=pipeline)
        sc.render_accumulate(W, H, 64, first_sample=32, sample_count=32, keep=True, pipeline=pipeline)
        parts = sc.resolve(64)
        assert np.allclose(whole, parts, rtol=1e-5, atol=1e-6)
    sc.render_accumulate(W, H, 64, pool_paths=4096, pipeline=1)
    small_pool = sc.resolve(64)
    assert np.allclose(whole, small_pool, rtol=1e-5, atol=1e-6)
    st_p = sc.render_accumulate(W, H, 64, pipeline=2)
    assert np.allclose(whole, sc.resolve(64), rtol=1e-5, atol=1e-6)
    for key in ("pixel_samples", "rays", "shaded_hits", "max_depth", "contributions"):
        assert st_w[key] == st_p[key], key
    other_seed, _ = sc.render(W, H, 64, seed=12345)
    assert not np.allclose(whole, other_seed, rtol=1e-3, atol=1e-4)


def test_render_end_to_end_and_display_transform(binding, oracle):
    sc = binding.Scene(scenes.cornell_box())
    img, st = sc.render(64, 48, 16)
    assert img.shape == (48, 64, 3) and np.isfinite(img).all() and img.max() > 1.0 and st["kernel_launches"] > 0
    again = sc.resolve(16)
    assert bit_equal(img, again)
    srgb = sc.resolve_srgb8(16)
    expect = oracle.to_srgb8(img.reshape(-1, 3)).reshape(48, 64, 3)
    assert np.array_equal(srgb, expect)  # byte work: exact (Color.cpp:64-80, FrameBuffer.hpp:91-95)
    # top row is j = 0: the light (bright) is in the upper half, the floor in the lower
    assert img[:24].mean() > 0


def test_display_transform_exhaustive(binding, oracle):
    """toSRGB + quantizeTo8bit on the device against the oracle for EVERY float in [0, 1] (bit patterns 0 .. 0x3f800000,
    2^30 + 1 values), plus values above 1, negatives, signed zeros, denormals, infinities and NaN.  Byte for byte."""
    from concurrent.futures import ThreadPoolExecutor
    import os
    sc = binding.Scene(scenes.cornell_box())
    workers = os.cpu_count() or 1

    def reference(bits):  # oracle on 3-float "pixels", split over the host threads (ctypes releases the GIL)
        values = bits.view(np.float32)
        pad = (-len(values)) % 3
        v = np.concatenate([values, np.zeros(pad, np.float32)]).reshape(-1, 3)
        parts = np.array_split(v, max(1, min(workers * 2, len(v) // 4096 or 1)))
        with ThreadPoolExecutor(workers) as pool:
            out = np.concatenate(list(pool.map(oracle.to_srgb8, parts)))
        return out.reshape(-1)[:len(values)]

    chunk = 1 << 26
    last = 0x3f800000
    histogram = np.zeros(256, np.int64)
    for first in range(0, last + 1, chunk):
        n = min(chunk, last + 1 - first)
        got = sc.selftest_srgb8(first, n)
        want = reference(np.arange(first, first + n, dtype=np.uint32))
        assert np.array_equal(got, want), (hex(first), int((got != want).sum()))
        histogram += np.bincount(got, minlength=256)
    assert histogram[0] > 0 and histogram[255] > 0 and (histogram > 0).all()  # every output level is reached
    # outside [0, 1]: (1, 4], the negative floats down to -4, tiny and special values
    for first, n in ((0x3f800000, 1 << 24), (0x80000000, 1 << 16), (0xbf000000, 1 << 16), (0x7f7fff00, 0x200),
                     (0xff7fff00, 0x200), (0x7fc00000, 16), (0x00000000, 1 << 16)):
        got = sc.selftest_srgb8(first, n)
        want = reference(np.arange(first, first + n, dtype=np.uint64).astype(np.uint32))
        assert np.array_equal(got, want), hex(first)


def test_keep_needs_an_image_on_this_handle(binding):
    """CORNELIS_RENDER_KEEP adds to the accumulators, so it is only accepted when they hold a render of this frame
    size made through this handle: a fresh handle gets its buffers from the library's device-memory cache, where they
    hold whatever the previous owner left."""
    flat = scenes.cornell_box()
    first = binding.Scene(flat)
    first.render_accumulate(64, 48, 8)
    expect = first.resolve(8).copy()
    first.close()  # its 64x48 accumulators go to the cache ...
    fresh = binding.Scene(flat)  # ... and this handle would be handed the same block
    with pytest.raises(binding.CornelisError) as e:
        fresh.render_accumulate(64, 48, 8, keep=True)
    assert e.value.code == binding.ERR_INVALID_ARGUMENT and "KEEP" in str(e.value)
    fresh.render_accumulate(64, 48, 8, first_sample=0, sample_count=4)
    fresh.render_accumulate(64, 48, 8, first_sample=4, sample_count=4, keep=True)
    assert np.allclose(fresh.resolve(8), expect, rtol=1e-5, atol=1e-6)
    with pytest.raises(binding.CornelisError):  # another frame size: new accumulators, nothing to keep
        fresh.render_accumulate(32, 32, 8, keep=True)
    with pytest.raises(binding.CornelisError):  # and the failed call has not made one
        fresh.resolve(8)
    fresh.render_accumulate(32, 32, 8)
    assert np.isfinite(fresh.resolve(8)).all()


def _closed_box(albedo):
    """Six inward-facing 200x200 faces around the origin, camera inside, one material on every face."""
    ext = [200.0, 200.0, 0.0]
    planes = [[1, 0, 0, -100, 0, 0, *ext], [-1, 0, 0, 100, 0, 0, *ext], [0, 1, 0, 0, -100, 0, *ext],
              [0, -1, 0, 0, 100, 0, *ext], [0, 0, 1, 0, 0, -100, *ext], [0, 0, -1, 0, 0, 100, *ext]]
    mats = [scenes.material(albedo=albedo, emissive=(0.5, 0.5, 0.5))]
    return scenes._flat([0, 0, -50, 0, 0, 0, 1.0, 0.7], np.zeros((0, 4)), [], planes, [1] * 6, mats)


@pytest.mark.parametrize("pipeline", [1, 2], ids=["wavefront", "persistent"])
def test_paths_end_in_a_closed_scene(binding, pipeline):
    """No ray leaves a closed box, so a path ends only by Russian roulette — which a NaN throughput never loses
    (`prob < u` is false for NaN, Render.cpp:189): the reference's loop would not terminate.  Here every path ends
    after 255 bounces at the latest.  With a finite albedo nothing comes near the limit."""
    nan_box = binding.Scene(_closed_box((float("nan"), 0.5, 0.5)))
    st = nan_box.render_accumulate(16, 16, 4, pipeline=pipeline)
    assert st["pixel_samples"] == 16 * 16 * 4 and st["max_depth"] == 255
    # every ray hits a wall; a path ends before the limit only through a failed glossy sample (w_in left at zero,
    # Materials.hpp:169-170, is ignored by every primitive)
    assert 32 * st["pixel_samples"] < st["rays"] <= 255 * st["pixel_samples"]
    st = binding.Scene(_closed_box((0.5, 0.5, 0.5))).render_accumulate(64, 64, 64, pipeline=pipeline)
    assert st["pixel_samples"] == 64 * 64 * 64 and 8 <= st["max_depth"] < 64
    with pytest.raises(binding.CornelisError) as e:
        nan_box.render_accumulate(16, 16, 4, max_depth=256)
    assert e.value.code == binding.ERR_INVALID_ARGUMENT


def test_handle_churn_reuses_device_memory(binding):
    """Handles created and destroyed in a loop (bench.py's end-to-end leg) take their buffers from the library's
    device-memory cache: a block that held another handle's framebuffer, or a smaller / larger frame, must not leak
    into the image.  cornelis_cuda_trim_memory hands the cache back and the next handle still works."""
    images = []
    for k, (w, h) in enumerate([(96, 64), (96, 64), (64, 32), (128, 96), (96, 64)]):
        sc = binding.Scene(scenes.cornell_box())
        img, st = sc.render(w, h, 16)
        assert st["pixel_samples"] == w * h * 16 and np.isfinite(img).all()
        if (w, h) == (96, 64):
            images.append(img.copy())
        sc.close()
        if k == 2:
            binding.trim_memory()
    for img in images[1:]:
        assert np.allclose(images[0], img, rtol=1e-5, atol=1e-6)  # same paths; atomics reorder the fp32 sums


def test_pipelines_trace_the_same_paths(binding):
    """The wavefront stages and the persistent kernel share path identity, random numbers and arithmetic, with one
    deliberate difference: a plane hit at t == 0 (a bounce that starts ON a wall: one ray in eight here) carries the
    reference's sign of zero in the wavefront's hit records and whatever zero the fast division returned in the
    persistent kernel, where t only enters P = o + d * t.  No path may notice: identical ray, hit, depth and
    contribution counts, images equal up to fp32 summation order."""
    sc = binding.Scene(scenes.cornell_box(aspect=0.5625))
    for (w, h, spp, depth) in [(192, 108, 64, 0), (61, 37, 128, 0), (128, 72, 32, 5)]:
        a = sc.render_accumulate(w, h, spp, pipeline=1, max_depth=depth)
        img_a = sc.resolve(spp).copy()
        b = sc.render_accumulate(w, h, spp, pipeline=2, max_depth=depth)
        img_b = sc.resolve(spp)
        for key in ("pixel_samples", "rays", "shaded_hits", "max_depth", "contributions"):
            assert a[key] == b[key], (w, h, spp, key)
        finite = np.isfinite(img_a) & np.isfinite(img_b)
        assert finite.mean() > 0.999 and np.allclose(img_a[finite], img_b[finite], rtol=1e-4, atol=1e-5)


def test_lane_refill_variant_traces_the_same_paths(binding, monkeypatch):
    """The persistent kernel without the queues (kept for scenes whose tables leave no shared memory for them;
    CORNELIS_PERSISTENT_QUEUE=0 selects it at scene creation) accounts for the same paths as the queued one."""
    queued = binding.Scene(scenes.cornell_box())
    monkeypatch.setenv("CORNELIS_PERSISTENT_QUEUE", "0")
    refill = binding.Scene(scenes.cornell_box())
    monkeypatch.delenv("CORNELIS_PERSISTENT_QUEUE")
    for (w, h, spp, depth) in [(96, 64, 32, 0), (7, 5, 9, 0), (64, 64, 16, 3)]:
        a = queued.render_accumulate(w, h, spp, pipeline=2, max_depth=depth)
        img_a = queued.resolve(spp).copy()
        b = refill.render_accumulate(w, h, spp, pipeline=2, max_depth=depth)
        for key in ("pixel_samples", "rays", "shaded_hits", "max_depth", "contributions"):
            assert a[key] == b[key], (w, h, spp, key)
        assert np.allclose(img_a, refill.resolve(spp), rtol=1e-5, atol=1e-6)


def test_persistent_batches_of_any_size(binding):
    """The persistent kernel parks Russian-roulette survivors per warp and scatters them 32 at a time; frames with
    fewer pixels than a warp, sample ranges that end inside a warp's claim, and the final partial batches must
    account for every path exactly once (same counters as the wavefront pipeline, which has no such batching)."""
    sc = binding.Scene(scenes.cornell_box())
    for (w, h, spp) in [(5, 3, 7), (1, 1, 1), (33, 1, 3), (64, 64, 5), (257, 3, 2)]:
        st_w = sc.render_accumulate(w, h, spp, pipeline=1)
        img_w = sc.resolve(spp).copy()
        st_p = sc.render_accumulate(w, h, spp, pipeline=2)
        img_p = sc.resolve(spp)
        for key in ("pixel_samples", "rays", "shaded_hits", "max_depth", "contributions"):
            assert st_w[key] == st_p[key], (w, h, spp, key, st_w[key], st_p[key])
        assert st_p["pixel_samples"] == w * h * spp
        assert np.allclose(img_w, img_p, rtol=1e-5, atol=1e-6)


def test_depth_cap_and_argument_errors(binding):
    sc = binding.Scene(scenes.cornell_box())
    st = sc.render_accumulate(64, 64, 8, max_depth=2)
    assert st["max_depth"] <= 2
    st = sc.render_accumulate(64, 64, 8)
    assert st["max_depth"] > 2
    for bad in (dict(width=0, height=16, samples=4), dict(width=16, height=16, samples=0),
                dict(width=16, height=-1, samples=4)):
        with pytest.raises(binding.CornelisError) as e:
            sc.render_accumulate(bad["width"], bad["height"], bad["samples"])
        assert e.value.code == binding.ERR_INVALID_ARGUMENT
    with pytest.raises(binding.CornelisError):
        sc.bsdf_eval([99], [[0, 0, 1]], [[0, 0, 1]], [[0, 0, 1]])
    calls = []
    with pytest.raises(binding.CornelisError) as e:
        sc.render_accumulate(256, 256, 64, pool_paths=65536, progress=lambda done, total: calls.append(done) or 1)
    assert e.value.code == binding.ERR_ABORTED and len(calls) == 1


def test_library_communicators_on_one_gpu(binding):
    """The exchange step through the library's own NCCL binding on whatever this box has.  One GPU: a communicator of
    one rank (ncclCommInitAll over [0], and ncclCommInitRank with a fresh unique id) leaves the image as it is — the
    call still goes through ncclAllReduce — and two scenes on the SAME GPU are summed locally by
    cornelis_cuda_reduce_framebuffers; either way the union of the sample ranges equals the one-shot render."""
    flat = scenes.cornell_box()
    W, H, spp = 96, 64, 32
    one = binding.Scene(flat)
    one.render_accumulate(W, H, spp, variance=True)
    expect, expect_var = [a.copy() for a in one.resolve(spp, variance=True)]
    for make in (lambda: binding.Comm.init_all([0]),
                 lambda: binding.Comm.init_rank(binding.comm_unique_id(), 0, 1, 0)):
        comm = make()
        info = comm.info()
        assert info["n_ranks"] == 1 and info["n_local"] == 1 and info["nccl_version"] >= 21800
        comm.allreduce_framebuffers([one])
        got, var = one.resolve(spp, variance=True)
        assert bit_equal(got, expect) and bit_equal(var, expect_var)
        with pytest.raises(binding.CornelisError):
            comm.allreduce_framebuffers([one, one])  # one scene per local rank
        comm.close()
    with pytest.raises(binding.CornelisError):
        binding.Comm.init_all([0, 0])
    halves = [binding.Scene(flat), binding.Scene(flat)]
    for rank, sc in enumerate(halves):
        sc.render_accumulate(W, H, spp, first_sample=rank * spp // 2, sample_count=spp // 2, variance=True)
    binding.reduce_framebuffers(halves)
    got, var = halves[0].resolve(spp, variance=True)
    assert np.allclose(got, expect, rtol=1e-5, atol=1e-6) and np.allclose(var, expect_var, rtol=1e-3, atol=1e-5)
    fresh = binding.Scene(flat)
    with pytest.raises(binding.CornelisError):
        binding.reduce_framebuffers([halves[0], fresh])  # nothing rendered on the second


def test_multi_gpu_sample_sharding_single_process(binding):
    """Two or more GPUs, one process: each renders its share of the sample indices; cornelis_cuda_reduce_framebuffers
    (one grouped ncclReduce onto the first scene's GPU) and cornelis_cuda_allreduce_framebuffers (ncclCommInitAll + one
    grouped ncclAllReduce: every GPU ends with the whole image) both equal the one-GPU render of all samples."""
    if binding.device_count() < 2:
        pytest.skip("needs two GPUs")
    flat = scenes.cornell_box()
    W = H = 128
    spp = 64
    one = binding.Scene(flat, device=0)
    one.render_accumulate(W, H, spp)
    expect = one.resolve(spp).copy()
    parts = [binding.Scene(flat, device=d) for d in (0, 1)]
    for rank, sc in enumerate(parts):
        sc.render_accumulate(W, H, spp, first_sample=rank * spp // 2, sample_count=spp // 2, variance=True)
    binding.reduce_framebuffers(parts)
    got, var = parts[0].resolve(spp, variance=True)
    assert np.allclose(got, expect, rtol=1e-5, atol=1e-6)
    assert np.isfinite(var).all() and var.max() > 0
    n = binding.device_count()
    devices = list(range(n))
    shards = [binding.Scene(flat, device=d) for d in devices]
    for rank, sc in enumerate(shards):
        sc.render_accumulate(W, H, spp, first_sample=spp * rank // n, sample_count=spp * (rank + 1) // n - spp * rank // n)
    comm = binding.Comm.init_all(devices)
    assert comm.info()["n_ranks"] == n
    comm.allreduce_framebuffers(shards)
    for sc in shards:  # every GPU holds the whole image
        assert np.allclose(sc.resolve(spp), expect, rtol=1e-5, atol=1e-6)
    comm.close()


# ------------------------------------------------------------------------- uniform grid (BASELINE config 4) --

def test_grid_bit_exact_against_exhaustive_and_oracle(binding, oracle):
    """10 000 spheres (BASELINE.json configs[3]): the grid returns the exhaustive scan's primitive id and t bit for
    bit — against the oracle on 2^16 rays, against the exhaustive shared-memory kernel on 2^21 rays of three kinds
    (camera rays, random rays through the cloud, bounce-like rays 1e-4 off a surface)."""
    flat = scenes.many_spheres(10000)
    sc = binding.Scene(flat)
    assert sc.acceleration()["grid"]                       # AUTO picks the grid for this scene
    info = sc.acceleration()
    assert info["references"] < 40 * 10000 and min(info["dims"]) >= 8
    ref = oracle.scene(flat)
    rng = np.random.default_rng(21)
    n = 1 << 16
    pi, pj = rng.integers(0, 1920, n), rng.integers(0, 1080, n)
    p1, p2 = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
    o1, d1 = sc.pixel_rays(1920, 1080, pi, pj, p1, p2)
    want = ref.intersect(o1, d1)
    got = sc.intersect(o1, d1)
    assert np.array_equal(got["prim"], want["prim"]) and bit_equal(got["t"], want["t"])
    assert bit_equal(got["P"], want["P"]) and bit_equal(got["N"], want["N"]) and np.array_equal(got["mat"], want["mat"])
    # bounce-like rays from the oracle's hit points
    g = unit(rng, n)
    hit = want["prim"] >= 0
    o3 = (want["P"][hit] + g[hit] * np.float32(1e-4)).astype(np.float32)
    d3 = g[hit]
    want3 = ref.intersect(o3, d3)
    got3 = sc.intersect(o3, d3, surface=False)
    assert np.array_equal(got3["prim"], want3["prim"]) and bit_equal(got3["t"], want3["t"])
    # large batch: grid vs exhaustive kernel
    m = 1 << 21
    o2, d2 = scenes.microbench_rays(m, seed=77)
    o2[: m // 2, 1] = np.abs(o2[: m // 2, 1])
    pi, pj = rng.integers(0, 1920, m // 4), rng.integers(0, 1080, m // 4)
    oc, dc = sc.pixel_rays(1920, 1080, pi, pj, rng.random(m // 4, dtype=np.float32), rng.random(m // 4, dtype=np.float32))
    org, dirs = np.concatenate([o2, oc]), np.concatenate([d2, dc])
    org[:64] *= np.float32(100.0)                          # outside the trusted region: exhaustive fallback path
    dirs[64:96, 0] = 0.0
    dirs[96:100] = np.float32(np.nan)
    dirs[100:104] = 0.0
    with_grid = sc.intersect(org, dirs, surface=False)
    sc.set_acceleration(binding.ACCEL_NONE)
    assert not sc.acceleration()["grid"]
    exhaustive = sc.intersect(org, dirs, surface=False)
    assert np.array_equal(with_grid["prim"], exhaustive["prim"])
    assert bit_equal(with_grid["t"], exhaustive["t"])
    assert (exhaustive["prim"] >= 0).mean() > 0.3 and ((exhaustive["prim"] >= 0) & (exhaustive["prim"] < 10000)).mean() > 0.2
    sc.set_acceleration(binding.ACCEL_AUTO)
    assert sc.acceleration()["grid"]


@pytest.mark.parametrize("pipeline", [p[1] for p in PIPELINES], ids=[p[0] for p in PIPELINES])
def test_grid_render_traces_the_same_paths(binding, pipeline):
    """Rendering through the grid and through the exhaustive scan traces identical paths: identical ray, hit, depth
    and contribution counts; images equal up to fp32 summation order."""
    flat = scenes.many_spheres(3000, 16)
    sc = binding.Scene(flat)
    W, H, spp = 160, 90, 32
    sc.set_acceleration(binding.ACCEL_NONE)
    st_a = sc.render_accumulate(W, H, spp, pipeline=pipeline, max_depth=64)
    img_a = sc.resolve(spp).copy()
    sc.set_acceleration(binding.ACCEL_GRID)
    st_b = sc.render_accumulate(W, H, spp, pipeline=pipeline, max_depth=64)
    img_b = sc.resolve(spp)
    for key in ("pixel_samples", "rays", "shaded_hits", "max_depth", "contributions"):
        assert st_a[key] == st_b[key], key
    finite = np.isfinite(img_a) & np.isfinite(img_b)
    assert finite.mean() > 0.999 and np.allclose(img_a[finite], img_b[finite], rtol=1e-4, atol=1e-5)
    # the Cornell box can be forced through the grid as well (4 spheres)
    box = binding.Scene(scenes.cornell_box())
    s1 = box.render_accumulate(64, 64, 32, pipeline=pipeline)
    a = box.resolve(32).copy()
    box.set_acceleration(binding.ACCEL_GRID)
    assert box.acceleration()["grid"]
    s2 = box.render_accumulate(64, 64, 32, pipeline=pipeline)
    assert s1["rays"] == s2["rays"] and s1["contributions"] == s2["contributions"]
    assert np.allclose(a, box.resolve(32), rtol=1e-4, atol=1e-5)


def test_config4_image_within_three_sigma(binding, golden):
    """BASELINE.json configs[3] (10 000 spheres, 64 mixed materials, max depth 64) at a reduced frame: the GPU render
    through the grid against the reference's brute-force render (tests/golden/make_golden.py config4: 48x27 at
    2048 spp), 3-sigma per pixel.  The estimator is heavy-tailed here — 2 % of the materials are lights — so both
    sides need thousands of samples per pixel for their variance estimates to mean anything."""
    g = golden("render_config4_48x27_2048spp.npz")
    flat = scenes.many_spheres(10000)
    W, H, spp_ref, spp = 48, 27, int(g["spp"]), 4096
    sc = binding.Scene(flat)
    st = sc.render_accumulate(W, H, spp, variance=True, max_depth=64)
    mean, var = sc.resolve(spp, variance=True)
    ok, diff, sigma = _three_sigma(mean, var, spp, g["mean"], g["variance"], spp_ref)
    good = np.isfinite(mean).all(axis=2)
    rmse = float(np.sqrt(np.mean(diff[good] ** 2)))
    rays_ref = float(g["rays"]) / (W * H * spp_ref)
    print(f"config 4: 3-sigma fraction {ok.mean():.5f}  RMSE {rmse:.5f}  rays/sample {st['rays'] / st['pixel_samples']:.3f}"
          f" (reference {rays_ref:.3f})")
    assert ok.mean() >= 0.99, ok.mean()
    assert abs(mean[good].mean() / g["mean"][good].mean() - 1.0) < 0.03
    assert abs(st["rays"] / st["pixel_samples"] - rays_ref) < 0.02
