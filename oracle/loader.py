"""TEST INFRASTRUCTURE — ctypes front-end for the two oracle libraries (see oracle/oracle_api.h).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may import this module.
The product (cornelis_b200/) never does.

  load("reference")  oracle/_ref/libcornelis_ref.so   — the unmodified reference, compiled by build_ref.sh
  load("port")       oracle/libcornelis_oracle.so     — the plain-C restatement, compiled by `make -C oracle`
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PATHS = {
    "reference": HERE / "_ref" / "libcornelis_ref.so",
    "port": HERE / "libcornelis_oracle.so",
}

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def available(kind: str) -> bool:
    return PATHS[kind].exists()


class Oracle:
    """One loaded oracle library."""

    def __init__(self, kind: str):
        path = PATHS[kind]
        if not path.exists():
            raise FileNotFoundError(f"oracle library missing: {path} (run `make -C oracle` / oracle/build_ref.sh)")
        lib = C.CDLL(os.fspath(path))
        self.lib = lib
        self.path = path
        lib.ora_kind.restype = C.c_char_p
        self.kind = lib.ora_kind().decode()
        assert self.kind == kind, (self.kind, kind)

        lib.ora_scene_create.restype = C.c_void_p
        lib.ora_scene_create.argtypes = [_f32p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_int32, C.c_void_p, C.c_int32]
        lib.ora_scene_destroy.argtypes = [C.c_void_p]
        lib.ora_camera_rays.argtypes = [C.c_void_p, C.c_int64, _f32p, _f32p, _f32p, _f32p]
        lib.ora_pixel_rays.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, _i32p, _i32p, _f32p, _f32p,
                                       _f32p, _f32p]
        lib.ora_intersect.argtypes = [C.c_void_p, C.c_int64, _f32p, _f32p, C.c_void_p, _f32p, _i32p, _f32p, _f32p,
                                      _i32p]
        lib.ora_bsdf_sample.argtypes = [C.c_void_p, C.c_int64, _i32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p]
        lib.ora_bsdf_eval.argtypes = [C.c_void_p, C.c_int64, _i32p, _f32p, _f32p, _f32p, _f32p, _f32p]
        lib.ora_rr_factor.argtypes = [C.c_int64, _f32p, _i32p, _f32p]
        lib.ora_shade.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, _f32p, _f32p, _f32p, _i32p, _f32p,
                                  _f32p, _f32p, _f32p, _u8p]
        for name in ("ora_gtr2", "ora_lambda_tr"):
            getattr(lib, name).restype = C.c_float
            getattr(lib, name).argtypes = [C.c_float, C.c_float]
        for name in ("ora_shadow_masking_tr", "ora_schlick"):
            getattr(lib, name).restype = C.c_float
            getattr(lib, name).argtypes = [C.c_float, C.c_float, C.c_float]
        lib.ora_construct_basis.argtypes = [_f32p, _f32p]
        lib.ora_prng_floats.argtypes = [C.c_uint64, C.c_int64, C.c_int64, _f32p]
        lib.ora_render.restype = C.c_int
        lib.ora_render.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint64,
                                   C.c_int32, _f32p, C.c_void_p, C.c_void_p]
        lib.ora_render_strided.restype = C.c_int
        lib.ora_render_strided.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_uint64,
                                           C.c_int32, _f32p, C.c_void_p, C.c_void_p]
        lib.ora_to_srgb8.argtypes = [C.c_int64, _f32p, _u8p]
        lib.ora_frame_tiling.restype = C.c_int32
        lib.ora_frame_tiling.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        lib.ora_sample_draw_order.argtypes = [_i32p]

    # ---- scene -------------------------------------------------------------------------------------------------
    def scene(self, flat) -> "OracleScene":
        return OracleScene(self, flat)

    # ---- scene-free helpers ------------------------------------------------------------------------------------
    def rr_factor(self, throughput, depth):
        thr = _f32(throughput, (-1, 3))
        depth = _i32(depth)
        out = np.empty(len(thr), np.float32)
        self.lib.ora_rr_factor(len(thr), thr, depth, out)
        return out

    def gtr2(self, c, alpha):
        return float(self.lib.ora_gtr2(c, alpha))

    def lambda_tr(self, t, alpha):
        return float(self.lib.ora_lambda_tr(t, alpha))

    def shadow_masking_tr(self, ti, to, alpha):
        return float(self.lib.ora_shadow_masking_tr(ti, to, alpha))

    def schlick(self, c, n1, n2):
        return float(self.lib.ora_schlick(c, n1, n2))

    def construct_basis(self, N):
        out = np.empty(9, np.float32)
        self.lib.ora_construct_basis(_f32(N, (3,)), out)
        return out.reshape(3, 3)  # rows T, B, N

    def prng_floats(self, seed, jumps, n):
        out = np.empty(n, np.float32)
        self.lib.ora_prng_floats(seed, jumps, n, out)
        return out

    def to_srgb8(self, rgb):
        rgb = _f32(rgb, (-1, 3))
        out = np.empty((len(rgb), 3), np.uint8)
        self.lib.ora_to_srgb8(len(rgb), rgb, out)
        return out

    def frame_tiling(self, W, H, tw, th):
        n = self.lib.ora_frame_tiling(W, H, tw, th, None)
        rects = np.empty((n, 4), np.int32)
        self.lib.ora_frame_tiling(W, H, tw, th, rects.ctypes.data_as(C.c_void_p))
        return rects

    def sample_draw_order(self):
        order = np.empty(3, np.int32)
        self.lib.ora_sample_draw_order(order)
        return order


class OracleScene:
    """A scene instantiated inside an oracle library from the flat description of cornelis_b200.scenes."""

    def __init__(self, oracle: Oracle, flat):
        self.oracle = oracle
        self.flat = flat
        cam = _f32(flat["camera"], (8,))
        sph = _f32(flat["spheres"], (-1, 4))
        smat = _i32(flat["sphere_mat"])
        pl = _f32(flat["planes"], (-1, 9))
        pmat = _i32(flat["plane_mat"])
        mats = _f32(flat["materials"], (-1, 11))
        self.n_spheres, self.n_planes = len(sph), len(pl)
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a.size else None
        self.handle = oracle.lib.ora_scene_create(cam, vp(sph), vp(smat), len(sph), vp(pl), vp(pmat), len(pl),
                                                  vp(mats), len(mats))

    def close(self):
        if self.handle:
            self.oracle.lib.ora_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def camera_rays(self, x, y):
        x, y = _f32(x), _f32(y)
        org = np.empty((len(x), 3), np.float32)
        dirs = np.empty((len(x), 3), np.float32)
        self.oracle.lib.ora_camera_rays(self.handle, len(x), x, y, org, dirs)
        return org, dirs

    def pixel_rays(self, W, H, pi, pj, phi1, phi2):
        pi, pj, phi1, phi2 = _i32(pi), _i32(pj), _f32(phi1), _f32(phi2)
        org = np.empty((len(pi), 3), np.float32)
        dirs = np.empty((len(pi), 3), np.float32)
        self.oracle.lib.ora_pixel_rays(self.handle, W, H, len(pi), pi, pj, phi1, phi2, org, dirs)
        return org, dirs

    def intersect(self, org, dirs, t_init=None):
        org, dirs = _f32(org, (-1, 3)), _f32(dirs, (-1, 3))
        n = len(org)
        t = np.empty(n, np.float32)
        prim = np.empty(n, np.int32)
        P = np.empty((n, 3), np.float32)
        N = np.empty((n, 3), np.float32)
        mat = np.empty(n, np.int32)
        ti = None
        if t_init is not None:
            t_init = _f32(t_init)
            ti = t_init.ctypes.data_as(C.c_void_p)
        self.oracle.lib.ora_intersect(self.handle, n, org, dirs, ti, t, prim, P, N, mat)
        return dict(t=t, prim=prim, P=P, N=N, mat=mat)

    def bsdf_sample(self, mat, wo, N, x):
        mat, wo, N, x = _i32(mat), _f32(wo, (-1, 3)), _f32(N, (-1, 3)), _f32(x, (-1, 3))
        n = len(mat)
        wi = np.empty((n, 3), np.float32)
        pdf = np.empty(n, np.float32)
        f = np.empty((n, 3), np.float32)
        self.oracle.lib.ora_bsdf_sample(self.handle, n, mat, wo, N, x, wi, pdf, f)
        return dict(wi=wi, pdf=pdf, f=f)

    def bsdf_eval(self, mat, wi, wo, N):
        mat, wi, wo, N = _i32(mat), _f32(wi, (-1, 3)), _f32(wo, (-1, 3)), _f32(N, (-1, 3))
        n = len(mat)
        f = np.empty((n, 3), np.float32)
        pdf = np.empty(n, np.float32)
        self.oracle.lib.ora_bsdf_eval(self.handle, n, mat, wi, wo, N, f, pdf)
        return dict(f=f, pdf=pdf)

    def shade(self, depth, seed_base, P, N, mat, org, dirs, thr, rad):
        P, N, mat = _f32(P, (-1, 3)), _f32(N, (-1, 3)), _i32(mat)
        org, dirs = _f32(org, (-1, 3)).copy(), _f32(dirs, (-1, 3)).copy()
        thr, rad = _f32(thr, (-1, 3)).copy(), _f32(rad, (-1, 3)).copy()
        n = len(mat)
        u = np.empty((n, 4), np.float32)
        alive = np.empty(n, np.uint8)
        self.oracle.lib.ora_shade(self.handle, n, depth, seed_base, u, P, N, mat, org, dirs, thr, rad, alive)
        return dict(u=u, org=org, dir=dirs, thr=thr, rad=rad, alive=alive.astype(bool))

    def render(self, W, H, spp, tile=(32, 32), seed=19791102, threads=0, variance=False, stats=False):
        mean = np.zeros((H, W, 3), np.float32)
        var = np.zeros((H, W, 3), np.float32) if variance else None
        st = np.zeros(4, np.float64) if stats else None
        rc = self.oracle.lib.ora_render(self.handle, W, H, spp, tile[0], tile[1], seed, threads, mean,
                                        var.ctypes.data_as(C.c_void_p) if variance else None,
                                        st.ctypes.data_as(C.c_void_p) if stats else None)
        if rc != 0:
            raise RuntimeError(f"ora_render failed rc={rc}")
        out = dict(mean=mean)
        if variance:
            out["variance"] = var
        if stats:
            out["stats"] = dict(rays=st[0], seconds=st[1], pixel_samples=st[2], max_depth=int(st[3]))
        return out


    def render_strided(self, W, H, spp, stride, seed=19791102, threads=0, variance=False, stats=False):
        """Every stride-th pixel (both dimensions) of the W x H frame, integrated as in the full frame."""
        ow, oh = -(-W // stride), -(-H // stride)
        mean = np.zeros((oh, ow, 3), np.float32)
        var = np.zeros((oh, ow, 3), np.float32) if variance else None
        st = np.zeros(4, np.float64) if stats else None
        rc = self.oracle.lib.ora_render_strided(self.handle, W, H, spp, stride, seed, threads, mean,
                                                var.ctypes.data_as(C.c_void_p) if variance else None,
                                                st.ctypes.data_as(C.c_void_p) if stats else None)
        if rc != 0:
            raise RuntimeError(f"ora_render_strided failed rc={rc}")
        out = dict(mean=mean)
        if variance:
            out["variance"] = var
        if stats:
            out["stats"] = dict(rays=st[0], seconds=st[1], pixel_samples=st[2], max_depth=int(st[3]))
        return out


_cache: dict[str, Oracle] = {}


def load(kind: str = "port") -> Oracle:
    if kind not in _cache:
        _cache[kind] = Oracle(kind)
    return _cache[kind]


def best() -> Oracle:
    """The compiled reference when present, else the port."""
    return load("reference") if available("reference") else load("port")
