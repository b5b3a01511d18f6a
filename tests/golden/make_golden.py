"""Generates the committed golden vectors under tests/golden/ from the REFERENCE ITSELF (oracle/_ref, i.e. the
unmodified /root/reference sources compiled by oracle/build_ref.sh).  Run in the build container, where
/root/reference exists:

    oracle/build_ref.sh && python tests/golden/make_golden.py

The fixtures let the GPU box (which has no /root/reference) check both the C restatement and the CUDA path against
outputs of the real reference.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from cornelis_b200 import scenes  # noqa: E402
from oracle import loader  # noqa: E402

OUT = Path(__file__).resolve().parent


def unit(rng, n):
    v = rng.standard_normal((n, 3)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True).astype(np.float32)
    return v.astype(np.float32)


def main():
    ref = loader.load("reference")
    rng = np.random.default_rng(19791102)

    # ---- intersection: microbench scene (config 3, reduced) and Cornell camera/bounce-like rays ----------------
    flat3 = scenes.microbench_scene(1024)
    s3 = ref.scene(flat3)
    org, dirs = scenes.microbench_rays(8192)
    h = s3.intersect(org, dirs)
    np.savez_compressed(OUT / "intersect_microbench.npz", org=org, dir=dirs, t=h["t"], prim=h["prim"], P=h["P"],
                        N=h["N"], mat=h["mat"])

    cornell = scenes.cornell_box()
    sc = ref.scene(cornell)
    n = 8192
    o1, d1 = sc.camera_rays(rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32))
    o2 = (rng.random((n, 3), dtype=np.float32) * np.float32(500) + np.float32([-250, 20, -250])).astype(np.float32)
    d2 = unit(rng, n)
    d2[:16] = 0  # degenerate rays are ignored by every primitive (Geometry.cpp:67-70)
    d2[16:32] *= np.float32(3.5)  # directions need not be unit
    org = np.concatenate([o1, o2]).astype(np.float32)
    dirs = np.concatenate([d1, d2]).astype(np.float32)
    h = sc.intersect(org, dirs)
    np.savez_compressed(OUT / "intersect_cornell.npz", org=org, dir=dirs, t=h["t"], prim=h["prim"], P=h["P"],
                        N=h["N"], mat=h["mat"])

    # ---- camera rays for pixels + jitter ------------------------------------------------------------------------
    W, H = 1920, 1080
    sc16 = ref.scene(scenes.cornell_box(aspect=0.5625))
    pi = rng.integers(0, W, 4096).astype(np.int32)
    pj = rng.integers(0, H, 4096).astype(np.int32)
    p1, p2 = rng.random(4096, dtype=np.float32), rng.random(4096, dtype=np.float32)
    o, d = sc16.pixel_rays(W, H, pi, pj, p1, p2)
    np.savez_compressed(OUT / "pixel_rays_1080p.npz", W=W, H=H, pi=pi, pj=pj, phi1=p1, phi2=p2, org=o, dir=d)

    # ---- BSDF sample / eval on fixed random inputs --------------------------------------------------------------
    n = 8192
    N = unit(rng, n)
    N[:64] = np.float32([0, 1, 0])
    N[64:128] = np.float32([0, 0, -1])
    wo = unit(rng, n)
    flip = (wo * N).sum(1) < 0
    wo[flip] = -wo[flip]
    x = rng.random((n, 3), dtype=np.float32)
    mat = rng.integers(0, 6, n).astype(np.int32)
    s = sc.bsdf_sample(mat, wo, N, x)
    wi = unit(rng, n)
    e = sc.bsdf_eval(mat, wi, wo, N)
    np.savez_compressed(OUT / "bsdf_cornell.npz", mat=mat, wo=wo, N=N, x=x, wi=s["wi"], pdf=s["pdf"], f=s["f"],
                        eval_wi=wi, eval_f=e["f"], eval_pdf=e["pdf"])

    # ---- one shading pass (accumulateAndBounce) at two depths ---------------------------------------------------
    n = 4096
    N, wo, mat = N[:n], wo[:n], mat[:n]
    P = (rng.standard_normal((n, 3)) * 100).astype(np.float32)
    thr = rng.random((n, 3), dtype=np.float32)
    rad = rng.random((n, 3), dtype=np.float32)
    out = {}
    for depth in (0, 5):
        r = sc.shade(depth, 4242, P, N, mat, P, -wo, thr, rad)
        for k, v in r.items():
            out[f"d{depth}_{k}"] = v
    np.savez_compressed(OUT / "shade_cornell.npz", P=P, N=N, mat=mat, org=P, dir=-wo, thr=thr, rad=rad,
                        order=ref.sample_draw_order(), **out)

    # ---- PRNG known answers, helper scalars ---------------------------------------------------------------------
    np.savez_compressed(OUT / "prng.npz", seed=19791102, tile0=ref.prng_floats(19791102, 0, 64),
                        tile3=ref.prng_floats(19791102, 3, 64))

    # ---- whole renders ------------------------------------------------------------------------------------------
    small = sc.render(32, 32, 4, tile=(32, 32), stats=True)
    np.savez_compressed(OUT / "render_cornell_32x32_4spp.npz", mean=small["mean"], rays=small["stats"]["rays"])
    big = sc.render(128, 128, 4096, tile=(32, 32), variance=True, stats=True)
    np.savez_compressed(OUT / "render_cornell_128x128_4096spp.npz", mean=big["mean"], variance=big["variance"],
                        spp=4096, rays=big["stats"]["rays"], max_depth=big["stats"]["max_depth"])
    print("rays/sample", big["stats"]["rays"] / big["stats"]["pixel_samples"], "seconds", big["stats"]["seconds"])
    config4(ref)


def config4(ref):
    """BASELINE.json configs[3] — 10 000 spheres, 64 mixed materials — at 48x27 and 2048 spp (about a minute of the
    reference's brute-force loop on 8 threads).  The estimator is heavy-tailed here (2 % of the materials are lights):
    a few hundred samples per pixel leave the reference's own variance estimate too poor for a 3-sigma test."""
    sc = ref.scene(scenes.many_spheres(10000))
    r = sc.render(48, 27, 2048, tile=(16, 9), variance=True, stats=True)
    np.savez_compressed(OUT / "render_config4_48x27_2048spp.npz", mean=r["mean"], variance=r["variance"], spp=2048,
                        rays=r["stats"]["rays"], max_depth=r["stats"]["max_depth"])
    print("config 4 rays/sample", r["stats"]["rays"] / r["stats"]["pixel_samples"], "seconds", r["stats"]["seconds"])


def headline_strided(ref):
    """The HEADLINE frame (BASELINE.json configs[1]: Cornell 1920x1080, aspect 0.5625) at its full 4096 spp on every
    8th pixel of each dimension: 240 x 135 pixels, each integrated exactly as in the full frame
    (ora_render_strided).  Mean and per-sample variance from the compiled reference: the 3-sigma test of the frame
    bench.py times."""
    sc = ref.scene(scenes.cornell_box(aspect=0.5625))
    r = sc.render_strided(1920, 1080, 4096, 8, variance=True, stats=True)
    np.savez_compressed(OUT / "render_cornell_1080p_stride8_4096spp.npz", mean=r["mean"], variance=r["variance"],
                        spp=4096, stride=8, W=1920, H=1080, rays=r["stats"]["rays"],
                        pixel_samples=r["stats"]["pixel_samples"], max_depth=r["stats"]["max_depth"])
    print("headline strided rays/sample", r["stats"]["rays"] / r["stats"]["pixel_samples"], "seconds",
          r["stats"]["seconds"])


if __name__ == "__main__":
    if "--config4-only" in sys.argv:
        config4(loader.load("reference"))
    elif "--headline-only" in sys.argv:
        headline_strided(loader.load("reference"))
    else:
        main()
        headline_strided(loader.load("reference"))
