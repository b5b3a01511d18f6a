"""Flat scene descriptions for the named BASELINE.json configurations.

A flat scene mirrors the reference's SceneDescription (reference include/cornelis/SceneDescription.hpp:14-92)
as plain float/int arrays — the form both the C-ABI (include/cornelis_cuda.h) and the test oracle accept:

  camera      (8,)    origin xyz, lookAt xyz, aspect, horizontalFov
  spheres     (n, 4)  center xyz, radius            sphere_mat (n,) scene material index, -1 = none (-> 0)
  planes      (n, 9)  normal xyz, point xyz, extents xyz (only [0], [1] are used: width, height)
  plane_mat   (n,)
  materials   (m, 11) albedo rgb, emissive rgb, roughness, reflectionTint rgb, ior — USER materials only:
                      scene material 0 is the implicit default material (SceneDescription.hpp:89), user
                      material k is scene material k + 1.
"""
from __future__ import annotations

import numpy as np

DEFAULT_SEED = 19791102  # reference include/cornelis/PRNG.hpp:12

# MaterialDescription defaults (SceneDescription.hpp:14-20)
_DEFAULT_MATERIAL = dict(albedo=(0.5, 0.5, 0.5), emissive=(0, 0, 0), roughness=0.2, tint=(0, 0, 0), ior=1.5)


def material(albedo=None, emissive=None, roughness=None, tint=None, ior=None):
    d = dict(_DEFAULT_MATERIAL)
    for k, v in dict(albedo=albedo, emissive=emissive, roughness=roughness, tint=tint, ior=ior).items():
        if v is not None:
            d[k] = v
    return [*d["albedo"], *d["emissive"], d["roughness"], *d["tint"], d["ior"]]


def _flat(camera, spheres, sphere_mat, planes, plane_mat, materials):
    return dict(
        camera=np.asarray(camera, np.float32).reshape(8),
        spheres=np.asarray(spheres, np.float32).reshape(-1, 4),
        sphere_mat=np.asarray(sphere_mat, np.int32).reshape(-1),
        planes=np.asarray(planes, np.float32).reshape(-1, 9),
        plane_mat=np.asarray(plane_mat, np.int32).reshape(-1),
        materials=np.asarray(materials, np.float32).reshape(-1, 11),
    )


def cornell_box(aspect: float = 1.0):
    """The CLI scene of the reference (src/cornelis.cpp:6-74): 5 finite planes, 4 spheres, 5 user materials.

    `aspect` multiplies the VERTICAL film vector (Camera.cpp:25): 1.0 for the 512x512 CLI render, H/W (0.5625
    at 1920x1080 / 3840x2160) for square pixels in the 16:9 configurations.
    """
    side = 555.0
    half = 550.0 / 2.0
    camera = [0, half, -1100, 0, half, 0, aspect, 0.7]
    mats = [
        material(albedo=(.65, .05, .05)),  # 1 red
        material(albedo=(.73, .73, .73)),  # 2 white
        material(albedo=(.12, .45, .15)),  # 3 green
        material(albedo=(0, 0, 0), emissive=(0, 0, 0), roughness=0.01, tint=(0.916, 0.61, 0.0), ior=0.470),  # 4 gold
        material(albedo=(0, 0, 0), emissive=(15, 15, 15)),  # 5 light
    ]
    red, white, green, gold, light = 1, 2, 3, 4, 5
    ext = [side, side, 0]
    planes = [
        [1, 0, 0, -half, half, 0, *ext],   # left wall
        [-1, 0, 0, half, half, 0, *ext],   # right wall
        [0, -1, 0, 0, side, 0, *ext],      # roof
        [0, 1, 0, 0, 0, 0, *ext],          # floor
        [0, 0, -1, 0, half, half, *ext],   # back wall
    ]
    plane_mat = [green, red, white, white, white]
    spheres = [
        [0, side - 60.0, 0, 60.0],
        [0, 50.0, 0, 50.0],
        [-160, 100.0, 0, 100.0],
        [160, 125.0, 200, 125.0],
    ]
    sphere_mat = [light, red, white, gold]
    return _flat(camera, spheres, sphere_mat, planes, plane_mat, mats)


def _uniform(rng: np.random.Generator, shape):
    """U[0,1) with the reference's 24-bit float mapping (XoshiroCpp.hpp:651-655): (u32 >> 8) * 2^-24."""
    bits = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64).astype(np.uint32)
    return (bits >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def six_materials():
    return [
        material(albedo=(.65, .05, .05)),
        material(albedo=(.73, .73, .73), roughness=0.4),
        material(albedo=(.12, .45, .15), tint=(0.6, 0.6, 0.6), roughness=0.15),
        material(albedo=(0, 0, 0), roughness=0.01, tint=(0.916, 0.61, 0.0), ior=0.470),
        material(albedo=(0.2, 0.3, 0.8), tint=(0.9, 0.9, 0.9), roughness=0.3, ior=1.8),
    ]


def microbench_scene(n_spheres: int = 1024, seed: int = DEFAULT_SEED):
    """Config 3: n_spheres random spheres (centres U[-800,800]^3, radii U[5,50]) inside
    six inward-facing 2000x2000 box faces at +-1000; scene material = primitive index mod 6."""
    rng = np.random.Generator(np.random.PCG64(seed))
    c = (_uniform(rng, (n_spheres, 3)) * np.float32(1600.0) - np.float32(800.0)).astype(np.float32)
    r = (_uniform(rng, (n_spheres, 1)) * np.float32(45.0) + np.float32(5.0)).astype(np.float32)
    spheres = np.concatenate([c, r], axis=1)
    ext = [2000.0, 2000.0, 0.0]
    planes = [
        [1, 0, 0, -1000, 0, 0, *ext], [-1, 0, 0, 1000, 0, 0, *ext],
        [0, 1, 0, 0, -1000, 0, *ext], [0, -1, 0, 0, 1000, 0, *ext],
        [0, 0, 1, 0, 0, -1000, *ext], [0, 0, -1, 0, 0, 1000, *ext],
    ]
    prim = np.arange(n_spheres + 6)
    camera = [0, 0, -900, 0, 0, 0, 1.0, 0.7]
    return _flat(camera, spheres, prim[:n_spheres] % 6, planes, prim[n_spheres:] % 6, six_materials())


def microbench_rays(n: int, seed: int = DEFAULT_SEED + 1):
    """Config 3 ray batch: origins U[-1000,1000]^3, directions uniform on the sphere, normalised in fp32."""
    rng = np.random.Generator(np.random.PCG64(seed))
    org = (_uniform(rng, (n, 3)) * np.float32(2000.0) - np.float32(1000.0)).astype(np.float32)
    g = rng.standard_normal((n, 3), dtype=np.float32)
    norm = np.sqrt((g * g).sum(axis=1, keepdims=True, dtype=np.float32)).astype(np.float32)
    norm[norm == 0] = 1
    dirs = (g / norm).astype(np.float32)
    return org, dirs


def many_spheres(n_spheres: int = 10000, n_materials: int = 64, seed: int = DEFAULT_SEED, aspect: float = 0.5625):
    """Config 4: random spheres over one 4000x4000 floor, mixed Oren-Nayar / glossy materials."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cx = _uniform(rng, (n_spheres, 1)) * np.float32(2000.0) - np.float32(1000.0)
    cz = _uniform(rng, (n_spheres, 1)) * np.float32(2000.0) - np.float32(1000.0)
    cy = _uniform(rng, (n_spheres, 1)) * np.float32(2000.0)
    r = _uniform(rng, (n_spheres, 1)) * np.float32(20.0) + np.float32(5.0)
    spheres = np.concatenate([cx, cy, cz, r], axis=1).astype(np.float32)
    mats = []
    for _ in range(n_materials):
        albedo = _uniform(rng, 3) * np.float32(0.8) + np.float32(0.1)
        glossy = _uniform(rng, 1)[0] < 0.5
        tint = (_uniform(rng, 3) * np.float32(0.5) + np.float32(0.5)) if glossy else np.zeros(3, np.float32)
        rough = float(_uniform(rng, 1)[0] * np.float32(0.59) + np.float32(0.01))
        ior = float(_uniform(rng, 1)[0] * np.float32(0.8) + np.float32(1.2))
        emissive = (15, 15, 15) if _uniform(rng, 1)[0] < 0.02 else (0, 0, 0)
        mats.append(material(albedo=tuple(albedo), emissive=emissive, roughness=rough, tint=tuple(tint), ior=ior))
    sphere_mat = 1 + (np.arange(n_spheres) % n_materials)
    planes = [[0, 1, 0, 0, 0, 0, 4000.0, 4000.0, 0.0]]
    camera = [0, 600, -2500, 0, 300, 0, aspect, 0.7]
    return _flat(camera, spheres, sphere_mat, planes, [2], mats)
