"""ctypes binding of the C-ABI library (include/cornelis_cuda.h) — what tests/ and bench.py drive.

There is no fallback: if cornelis_b200/lib/libcornelis_cuda.so is missing, importing a symbol from here raises, and
every compute call fails with CORNELIS_ERR_NO_DEVICE on a machine without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

# CORNELIS_CUDA_LIB: an alternative build of the same library (tools/variants.py A/B runs), never a fallback
LIB_PATH = Path(os.environ.get("CORNELIS_CUDA_LIB") or Path(__file__).resolve().parent / "lib" / "libcornelis_cuda.so")

# every symbol include/cornelis_cuda.h declares
EXPORTS = [
    "cornelis_cuda_abi_version", "cornelis_cuda_last_error", "cornelis_cuda_device_count",
    "cornelis_cuda_trim_memory",
    "cornelis_cuda_scene_create", "cornelis_cuda_scene_destroy", "cornelis_cuda_scene_set_stream",
    "cornelis_cuda_scene_set_acceleration", "cornelis_cuda_scene_acceleration", "cornelis_cuda_render_accumulate",
    "cornelis_cuda_framebuffer_device", "cornelis_cuda_reduce_framebuffers", "cornelis_cuda_resolve", "cornelis_cuda_resolve_device",
    "cornelis_cuda_comm_unique_id", "cornelis_cuda_comm_init_rank", "cornelis_cuda_comm_init_all",
    "cornelis_cuda_comm_destroy", "cornelis_cuda_comm_info", "cornelis_cuda_allreduce_framebuffers",
    "cornelis_cuda_resolve_srgb8",
    "cornelis_cuda_render", "cornelis_cuda_pixel_rays", "cornelis_cuda_intersect",
    "cornelis_cuda_intersect_device", "cornelis_cuda_bsdf_sample", "cornelis_cuda_bsdf_eval",
    "cornelis_cuda_shade", "cornelis_cuda_rng_uniforms", "cornelis_cuda_selftest_arith",
    "cornelis_cuda_intersect_compact", "cornelis_cuda_selftest_srgb8", "cornelis_cuda_rng_rounds",
    "cornelis_cuda_rng_bits",
]

OK, ERR_INVALID_ARGUMENT, ERR_NO_DEVICE, ERR_CUDA, ERR_OUT_OF_MEMORY, ERR_ABORTED, ERR_NCCL = range(7)
COMM_ID_BYTES = 128
RENDER_VARIANCE, RENDER_KEEP, RENDER_STAGE_TIMING, RENDER_DROP_NONFINITE = 1, 2, 4, 8
PIPELINE_DEFAULT, PIPELINE_WAVEFRONT, PIPELINE_PERSISTENT = 0, 1, 2
DEFAULT_SEED = 19791102


class CameraDesc(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("look_at", C.c_float * 3), ("aspect", C.c_float),
                ("horizontal_fov", C.c_float)]


class MaterialDesc(C.Structure):
    _fields_ = [("albedo", C.c_float * 3), ("emissive", C.c_float * 3), ("roughness", C.c_float),
                ("reflection_tint", C.c_float * 3), ("ior", C.c_float)]


class SphereDesc(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("radius", C.c_float), ("material", C.c_int32)]


class PlaneDesc(C.Structure):
    _fields_ = [("normal", C.c_float * 3), ("point", C.c_float * 3), ("extents", C.c_float * 3),
                ("material", C.c_int32)]


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("samples", C.c_int32), ("first_sample", C.c_int32),
                ("sample_count", C.c_int32), ("max_depth", C.c_int32), ("seed", C.c_uint64), ("flags", C.c_uint32),
                ("pipeline", C.c_int32), ("pool_paths", C.c_int32), ("reserved", C.c_int32)]


class RenderStats(C.Structure):
    _fields_ = [("pixel_samples", C.c_uint64), ("rays", C.c_uint64), ("shaded_hits", C.c_uint64),
                ("iterations", C.c_uint64), ("kernel_launches", C.c_uint64), ("contributions", C.c_uint64),
                ("max_depth", C.c_uint32),
                ("reserved", C.c_uint32), ("gpu_ms", C.c_float), ("intersect_ms", C.c_float),
                ("shade_ms", C.c_float), ("raygen_ms", C.c_float), ("accumulate_ms", C.c_float)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_ if name != "reserved"}


PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_uint64)


class CornelisError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"cornelis_cuda error {code}: {message}")
        self.code = code


_lib = None


def lib():
    """Loads the C-ABI library (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(f"{LIB_PATH} is missing — run `python cornelis_b200/build.py` "
                                    "(there is no CPU fallback)")
        L = C.CDLL(os.fspath(LIB_PATH))
        L.cornelis_cuda_last_error.restype = C.c_char_p
        vp, sz, i32, f = C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p
        L.cornelis_cuda_device_count.argtypes = [C.POINTER(C.c_int)]
        L.cornelis_cuda_scene_create.argtypes = [C.c_int, C.POINTER(CameraDesc), vp, sz, vp, sz, vp, sz, C.POINTER(vp)]
        L.cornelis_cuda_scene_destroy.argtypes = [vp]
        L.cornelis_cuda_scene_set_stream.argtypes = [vp, vp]
        L.cornelis_cuda_scene_set_acceleration.argtypes = [vp, C.c_int]
        L.cornelis_cuda_scene_acceleration.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_uint32 * 3),
                                                       C.POINTER(C.c_uint64)]
        L.cornelis_cuda_render_accumulate.argtypes = [vp, C.POINTER(RenderParams), vp, vp, C.POINTER(RenderStats)]
        L.cornelis_cuda_framebuffer_device.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
        L.cornelis_cuda_resolve.argtypes = [vp, i32, vp, vp]
        L.cornelis_cuda_reduce_framebuffers.argtypes = [C.POINTER(vp), C.c_int]
        L.cornelis_cuda_comm_unique_id.argtypes = [vp]
        L.cornelis_cuda_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.cornelis_cuda_comm_init_all.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
        L.cornelis_cuda_comm_destroy.argtypes = [vp]
        L.cornelis_cuda_comm_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.cornelis_cuda_allreduce_framebuffers.argtypes = [vp, C.POINTER(vp), C.c_int]
        L.cornelis_cuda_resolve_srgb8.argtypes = [vp, i32, vp]
        L.cornelis_cuda_resolve_device.argtypes = [vp, i32, C.POINTER(vp)]
        L.cornelis_cuda_render.argtypes = [vp, C.POINTER(RenderParams), vp, C.POINTER(RenderStats)]
        L.cornelis_cuda_pixel_rays.argtypes = [vp, i32, i32, sz, f, f, f, f, f, f]
        L.cornelis_cuda_intersect.argtypes = [vp, sz, f, f, f, f, f, f, f]
        L.cornelis_cuda_intersect_device.argtypes = [vp, sz, vp, vp, vp, C.c_int, C.POINTER(C.c_float)]
        L.cornelis_cuda_bsdf_sample.argtypes = [vp, sz, f, f, f, f, f, f, f]
        L.cornelis_cuda_bsdf_eval.argtypes = [vp, sz, f, f, f, f, f, f]
        L.cornelis_cuda_shade.argtypes = [vp, sz, i32, f, f, f, f, f, f, f, f, f]
        L.cornelis_cuda_rng_uniforms.argtypes = [vp, C.c_uint64, sz, f, f, f, f]
        L.cornelis_cuda_selftest_arith.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]
        L.cornelis_cuda_intersect_compact.argtypes = [vp, sz, f, f, f, C.POINTER(C.c_uint32), f, C.POINTER(C.c_uint32)]
        L.cornelis_cuda_selftest_srgb8.argtypes = [vp, C.c_uint32, sz, vp]
        L.cornelis_cuda_rng_bits.argtypes = [vp, C.c_int, C.c_uint64, sz, f, f, f, f]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise CornelisError(rc, lib().cornelis_cuda_last_error().decode())


def rng_rounds() -> int:
    """Rounds of the render loop's Philox4x32 generator."""
    return int(lib().cornelis_cuda_rng_rounds())


def trim_memory() -> None:
    """Returns the device memory cached from destroyed handles to the driver."""
    _check(lib().cornelis_cuda_trim_memory())


def device_count() -> int:
    n = C.c_int(0)
    _check(lib().cornelis_cuda_device_count(C.byref(n)))
    return n.value


def reduce_framebuffers(scene_list):
    """Sum the accumulators of several scenes (single process) into the first: one grouped ncclReduce over the
    scenes' GPUs (scenes sharing a GPU are added locally first)."""
    arr = (C.c_void_p * len(scene_list))(*[s.handle for s in scene_list])
    _check(lib().cornelis_cuda_reduce_framebuffers(arr, len(scene_list)))


def comm_unique_id() -> bytes:
    """ncclGetUniqueId: 128 bytes rank 0 hands to the other ranks of a one-process-per-GPU job."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    _check(lib().cornelis_cuda_comm_unique_id(buf))
    return bytes(buf)


class Comm:
    """An NCCL communicator owned by the library: the framebuffer sum is the render path's only exchange step."""

    def __init__(self, handle):
        self.handle = handle

    @classmethod
    def init_rank(cls, unique_id: bytes, rank: int, n_ranks: int, device: int):
        """One rank of a one-process-per-GPU job (ncclCommInitRank)."""
        if len(unique_id) != COMM_ID_BYTES:
            raise ValueError("unique id must be 128 bytes")
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(unique_id)
        h = C.c_void_p()
        _check(lib().cornelis_cuda_comm_init_rank(buf, rank, n_ranks, device, C.byref(h)))
        return cls(h)

    @classmethod
    def init_all(cls, devices):
        """All the given GPUs from this one process (ncclCommInitAll)."""
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _check(lib().cornelis_cuda_comm_init_all(arr, len(devices), C.byref(h)))
        return cls(h)

    def info(self):
        n, local, version = C.c_int(0), C.c_int(0), C.c_int(0)
        _check(lib().cornelis_cuda_comm_info(self.handle, C.byref(n), C.byref(local), C.byref(version)))
        return dict(n_ranks=n.value, n_local=local.value, nccl_version=version.value)

    def allreduce_framebuffers(self, scene_list):
        """In-place sum over all ranks of the accumulation images of this process's scenes (enqueued on their streams)."""
        arr = (C.c_void_p * len(scene_list))(*[s.handle for s in scene_list])
        _check(lib().cornelis_cuda_allreduce_framebuffers(self.handle, arr, len(scene_list)))

    def close(self):
        if self.handle:
            lib().cornelis_cuda_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


# MaterialDescription{} (reference SceneDescription.hpp:14-20): scene material 0
DEFAULT_MATERIAL = [0.5, 0.5, 0.5, 0, 0, 0, 0.2, 0, 0, 0, 1.5]


def descriptors(flat):
    """The C-ABI description structs of a flat scene (cornelis_b200.scenes): camera, sphere, plane and material arrays
    (scene material 0, the default material, prepended) plus the numpy views they were filled from."""
    cam = CameraDesc()
    c = _f32(flat["camera"], (8,))
    cam.origin[:] = c[0:3]
    cam.look_at[:] = c[3:6]
    cam.aspect, cam.horizontal_fov = float(c[6]), float(c[7])
    sph = _f32(flat["spheres"], (-1, 4))
    smat = np.asarray(flat["sphere_mat"], np.int32)
    pl = _f32(flat["planes"], (-1, 9))
    pmat = np.asarray(flat["plane_mat"], np.int32)
    mats = np.concatenate([np.asarray([DEFAULT_MATERIAL], np.float32), _f32(flat["materials"], (-1, 11))])
    S = (SphereDesc * max(len(sph), 1))()
    if len(sph):
        rec = np.zeros(len(sph), dtype=[("c", np.float32, 3), ("r", np.float32), ("m", np.int32)])
        rec["c"], rec["r"], rec["m"] = sph[:, 0:3], sph[:, 3], smat
        C.memmove(S, rec.ctypes.data, rec.nbytes)
    P = (PlaneDesc * max(len(pl), 1))()
    if len(pl):
        rec = np.zeros(len(pl), dtype=[("n", np.float32, 3), ("p", np.float32, 3), ("e", np.float32, 3), ("m", np.int32)])
        rec["n"], rec["p"], rec["e"], rec["m"] = pl[:, 0:3], pl[:, 3:6], pl[:, 6:9], pmat
        C.memmove(P, rec.ctypes.data, rec.nbytes)
    M = (MaterialDesc * len(mats))()
    C.memmove(M, np.ascontiguousarray(mats, np.float32).ctypes.data, mats.nbytes)
    return cam, S, P, M, (sph, pl, mats)


ACCEL_AUTO, ACCEL_NONE, ACCEL_GRID = 0, 1, 2


class Scene:
    """A scene resident on one GPU: SceneData (reference Scene.cpp:40-53) uploaded behind the C-ABI."""

    def __init__(self, flat, device: int = 0):
        self.handle = None
        L = lib()
        cam, S, P, M, (sph, pl, mats) = descriptors(flat)
        self.n_spheres, self.n_planes, self.n_materials = len(sph), len(pl), len(mats)
        self.scene_bytes = C.sizeof(cam) + C.sizeof(SphereDesc) * len(sph) + C.sizeof(PlaneDesc) * len(pl) + \
            C.sizeof(MaterialDesc) * len(mats)
        h = C.c_void_p()
        _check(L.cornelis_cuda_scene_create(device, C.byref(cam), S, len(sph), P, len(pl), M, len(mats), C.byref(h)))
        self.handle = h
        self.device = device
        self.frame = None

    def close(self):
        if self.handle:
            lib().cornelis_cuda_scene_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        """Run on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); None = the scene's own."""
        _check(lib().cornelis_cuda_scene_set_stream(self.handle, C.c_void_p(cuda_stream) if cuda_stream else None))

    # ---- hot path --------------------------------------------------------------------------------------------------
    def set_acceleration(self, mode: int):
        """ACCEL_AUTO / ACCEL_NONE (exhaustive scan from shared memory) / ACCEL_GRID (uniform grid over the spheres)."""
        _check(lib().cornelis_cuda_scene_set_acceleration(self.handle, int(mode)))

    def acceleration(self):
        on, dims, refs = C.c_int(0), (C.c_uint32 * 3)(), C.c_uint64(0)
        _check(lib().cornelis_cuda_scene_acceleration(self.handle, C.byref(on), C.byref(dims), C.byref(refs)))
        return dict(grid=bool(on.value), dims=tuple(dims), references=refs.value)

    def _params(self, width, height, samples, first_sample=0, sample_count=0, max_depth=0, seed=DEFAULT_SEED,
                variance=False, keep=False, stage_timing=False, drop_nonfinite=False, pipeline=PIPELINE_DEFAULT,
                pool_paths=0):
        p = RenderParams()
        p.width, p.height, p.samples = width, height, samples
        p.first_sample, p.sample_count, p.max_depth = first_sample, sample_count, max_depth
        p.seed = seed
        p.flags = (RENDER_VARIANCE if variance else 0) | (RENDER_KEEP if keep else 0) | \
            (RENDER_STAGE_TIMING if stage_timing else 0) | (RENDER_DROP_NONFINITE if drop_nonfinite else 0)
        p.pipeline, p.pool_paths = pipeline, pool_paths
        return p

    def render_accumulate(self, width, height, samples, progress=None, **kw):
        """Renders into the device accumulators; returns the stats dict."""
        p = self._params(width, height, samples, **kw)
        st = RenderStats()
        cb = PROGRESS_FN(lambda user, done, total: int(progress(done, total) or 0)) if progress else None
        _check(lib().cornelis_cuda_render_accumulate(self.handle, C.byref(p), C.cast(cb, C.c_void_p) if cb else None,
                                                     None, C.byref(st)))
        self.frame = (width, height)
        return st.as_dict()

    def framebuffer_device(self):
        """(device pointer, float count) of the float4 accumulation image."""
        ptr, n = C.c_void_p(), C.c_size_t()
        _check(lib().cornelis_cuda_framebuffer_device(self.handle, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def resolve(self, samples, variance=False, out=None):
        W, H = self.frame
        rgb = out if out is not None else np.empty((H, W, 3), np.float32)
        var = np.empty((H, W, 3), np.float32) if variance else None
        _check(lib().cornelis_cuda_resolve(self.handle, samples, _ptr(rgb), _ptr(var)))
        return (rgb, var) if variance else rgb

    def resolve_device(self, samples):
        """Resolve on the device only; returns the device pointer of the packed RGB image."""
        ptr = C.c_void_p()
        _check(lib().cornelis_cuda_resolve_device(self.handle, samples, C.byref(ptr)))
        return ptr.value

    def resolve_srgb8(self, samples):
        W, H = self.frame
        out = np.empty((H, W, 3), np.uint8)
        _check(lib().cornelis_cuda_resolve_srgb8(self.handle, samples, _ptr(out)))
        return out

    def render(self, width, height, samples, out=None, out_ptr=None, **kw):
        """The end-to-end call: render and download the framebuffer into host memory."""
        p = self._params(width, height, samples, **kw)
        st = RenderStats()
        if out_ptr is None:
            out = out if out is not None else np.empty((height, width, 3), np.float32)
            out_ptr = _ptr(out)
        _check(lib().cornelis_cuda_render(self.handle, C.byref(p), out_ptr, C.byref(st)))
        self.frame = (width, height)
        return out, st.as_dict()

    # ---- stages ------------------------------------------------------------------------------------------------------
    def pixel_rays(self, W, H, pi, pj, phi1, phi2):
        pi, pj = np.ascontiguousarray(pi, np.int32), np.ascontiguousarray(pj, np.int32)
        phi1, phi2 = _f32(phi1), _f32(phi2)
        n = len(pi)
        org, dirs = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        _check(lib().cornelis_cuda_pixel_rays(self.handle, W, H, n, _ptr(pi), _ptr(pj), _ptr(phi1), _ptr(phi2),
                                              _ptr(org), _ptr(dirs)))
        return org, dirs

    def intersect(self, org, dirs, surface=True):
        org, dirs = _f32(org, (-1, 3)), _f32(dirs, (-1, 3))
        n = len(org)
        t, prim = np.empty(n, np.float32), np.empty(n, np.int32)
        P = np.empty((n, 3), np.float32) if surface else None
        N = np.empty((n, 3), np.float32) if surface else None
        mat = np.empty(n, np.int32) if surface else None
        _check(lib().cornelis_cuda_intersect(self.handle, n, _ptr(org), _ptr(dirs), _ptr(t), _ptr(prim), _ptr(P), _ptr(N),
                                             _ptr(mat)))
        return dict(t=t, prim=prim, P=P, N=N, mat=mat)

    def intersect_compact(self, org, dirs):
        """Intersect stage of one wavefront pass incl. compaction: (indices of the rays that hit, of those that missed),
        each in the order the device queues hold them."""
        org, dirs = _f32(org, (-1, 3)), _f32(dirs, (-1, 3))
        n = len(org)
        hits, misses = np.empty(n, np.uint32), np.empty(n, np.uint32)
        nh, nm = C.c_uint32(0), C.c_uint32(0)
        _check(lib().cornelis_cuda_intersect_compact(self.handle, n, _ptr(org), _ptr(dirs), _ptr(hits), C.byref(nh),
                                                     _ptr(misses), C.byref(nm)))
        return hits[:nh.value], misses[:nm.value]

    def selftest_srgb8(self, first_bits, n):
        """Device display transform + quantisation of the floats with bit patterns first_bits .. first_bits + n - 1."""
        out = np.empty(n, np.uint8)
        _check(lib().cornelis_cuda_selftest_srgb8(self.handle, first_bits, n, _ptr(out)))
        return out

    def intersect_device(self, n, d_org4, d_dir4, d_hit2, repeats=1):
        ms = C.c_float(0)
        _check(lib().cornelis_cuda_intersect_device(self.handle, n, d_org4, d_dir4, d_hit2, repeats, C.byref(ms)))
        return ms.value

    def bsdf_sample(self, mat, wo, N, x):
        mat = np.ascontiguousarray(mat, np.int32)
        wo, N, x = _f32(wo, (-1, 3)), _f32(N, (-1, 3)), _f32(x, (-1, 3))
        n = len(mat)
        wi, pdf, f = np.empty((n, 3), np.float32), np.empty(n, np.float32), np.empty((n, 3), np.float32)
        _check(lib().cornelis_cuda_bsdf_sample(self.handle, n, _ptr(mat), _ptr(wo), _ptr(N), _ptr(x), _ptr(wi), _ptr(pdf),
                                               _ptr(f)))
        return dict(wi=wi, pdf=pdf, f=f)

    def bsdf_eval(self, mat, wi, wo, N):
        mat = np.ascontiguousarray(mat, np.int32)
        wi, wo, N = _f32(wi, (-1, 3)), _f32(wo, (-1, 3)), _f32(N, (-1, 3))
        n = len(mat)
        f, pdf = np.empty((n, 3), np.float32), np.empty(n, np.float32)
        _check(lib().cornelis_cuda_bsdf_eval(self.handle, n, _ptr(mat), _ptr(wi), _ptr(wo), _ptr(N), _ptr(f), _ptr(pdf)))
        return dict(f=f, pdf=pdf)

    def shade(self, depth, u, P, N, mat, org, dirs, thr, rad):
        """u[n,4] = (RR draw, x0, x1, x2)."""
        mat = np.ascontiguousarray(mat, np.int32)
        u, P, N = _f32(u, (-1, 4)), _f32(P, (-1, 3)), _f32(N, (-1, 3))
        org, dirs = _f32(org, (-1, 3)).copy(), _f32(dirs, (-1, 3)).copy()
        thr, rad = _f32(thr, (-1, 3)).copy(), _f32(rad, (-1, 3)).copy()
        n = len(mat)
        alive = np.empty(n, np.uint8)
        _check(lib().cornelis_cuda_shade(self.handle, n, depth, _ptr(u), _ptr(P), _ptr(N), _ptr(mat), _ptr(org),
                                         _ptr(dirs), _ptr(thr), _ptr(rad), _ptr(alive)))
        return dict(org=org, dir=dirs, thr=thr, rad=rad, alive=alive.astype(bool))

    def selftest_arith(self, mode, n, seed=1):
        bad = C.c_uint64(0)
        _check(lib().cornelis_cuda_selftest_arith(self.handle, mode, n, seed, C.byref(bad)))
        return bad.value

    def rng_bits(self, rounds, seed, pixel, sample, block):
        """Raw Philox4x32-`rounds` words for the counters (pixel, sample, block, 0)."""
        pixel = np.ascontiguousarray(pixel, np.uint32)
        sample = np.ascontiguousarray(sample, np.uint32)
        block = np.ascontiguousarray(block, np.uint32)
        out = np.empty((len(pixel), 4), np.uint32)
        _check(lib().cornelis_cuda_rng_bits(self.handle, rounds, seed, len(pixel), _ptr(pixel), _ptr(sample),
                                            _ptr(block), _ptr(out)))
        return out

    def rng_uniforms(self, seed, pixel, sample, block):
        pixel = np.ascontiguousarray(pixel, np.uint32)
        sample = np.ascontiguousarray(sample, np.uint32)
        block = np.ascontiguousarray(block, np.uint32)
        out = np.empty((len(pixel), 4), np.float32)
        _check(lib().cornelis_cuda_rng_uniforms(self.handle, seed, len(pixel), _ptr(pixel), _ptr(sample), _ptr(block),
                                                _ptr(out)))
        return out
