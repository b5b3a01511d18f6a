/* cornelis_cuda.h — C-ABI of the B200 render path.
 *
 * This is the drop-in boundary for the reference's batched render loop.  The reference has no FFI: the seam it
 * replaces is the call `integrateTile(tileInfo, options, scene, fb)` that RenderSession::render makes for every
 * tile from its TBB workers (reference src/Render.cpp:343, body :220-255), with read-only SceneData in and disjoint
 * pixel writes into one RGBFrameBuffer out.  One cornelis_cuda_render() call produces what the whole tile loop
 * (Render.cpp:327-354) produces: the W x H RGB framebuffer, row-major j*W+i, three packed floats per pixel
 * (reference include/cornelis/FrameBuffer.hpp:66, Color.hpp:56).
 *
 * Plain C: POD structs, pointers and sizes only; no exceptions, no STL, no torch types.  Every function returns
 * 0 on success and a non-zero cornelis_status otherwise; cornelis_cuda_last_error() gives the message for the
 * calling thread.  There is no CPU fallback: without a CUDA device every compute entry point fails with
 * CORNELIS_ERR_NO_DEVICE.
 *
 * Ownership: a scene handle owns all device memory (scene tables, path pool, queues, framebuffer accumulators).
 * Pointers passed in are borrowed for the duration of the call; output buffers are caller-owned.
 * Threading: a handle may be used by one host thread at a time; different handles are independent.
 */
#ifndef CORNELIS_CUDA_H
#define CORNELIS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CORNELIS_CUDA_ABI_VERSION 2

typedef enum cornelis_status {
    CORNELIS_OK = 0,
    CORNELIS_ERR_INVALID_ARGUMENT = 1, /* null pointer, zero-area frame, samples <= 0 (Render.cpp:310-313), bad index */
    CORNELIS_ERR_NO_DEVICE = 2,        /* no CUDA device / driver: the product has no CPU path */
    CORNELIS_ERR_CUDA = 3,             /* a CUDA runtime call failed; see cornelis_cuda_last_error() */
    CORNELIS_ERR_OUT_OF_MEMORY = 4,
    CORNELIS_ERR_ABORTED = 5,          /* the progress callback asked to stop (RenderCommand::Abort, Render.hpp:10-14) */
    CORNELIS_ERR_NCCL = 6              /* libnccl.so.2 could not be loaded, or an NCCL call failed */
} cornelis_status;

/* ---- scene description PODs: field-for-field the reference's SceneDescription.hpp:14-53 ---------------------- */

typedef struct cornelis_camera_desc { /* PerspectiveCameraDescription, SceneDescription.hpp:45-53 */
    float origin[3];
    float look_at[3];
    float aspect;         /* multiplies the VERTICAL film vector (Camera.cpp:25) */
    float horizontal_fov; /* radians */
} cornelis_camera_desc;

typedef struct cornelis_material_desc { /* MaterialDescription, SceneDescription.hpp:14-22 */
    float albedo[3];
    float emissive[3];
    float roughness;
    float reflection_tint[3];
    float ior;
} cornelis_material_desc;

typedef struct cornelis_sphere_desc { /* SphereDescription, SceneDescription.hpp:30-35 */
    float center[3];
    float radius;
    int32_t material; /* index into the material list; < 0 = unset -> material 0 (Scene.cpp:16) */
} cornelis_sphere_desc;

typedef struct cornelis_plane_desc { /* PlaneDescription, SceneDescription.hpp:37-43 */
    float normal[3];
    float point[3];
    float extents[3]; /* [0] = width along T, [1] = height along B (Scene.cpp:33-34); [2] unused */
    int32_t material;
} cornelis_plane_desc;

/* ---- render parameters ---------------------------------------------------------------------------------------- */

enum {
    CORNELIS_RENDER_VARIANCE = 1u << 0, /* also accumulate the per-pixel second moment (for the 3-sigma image test) */
    CORNELIS_RENDER_KEEP = 1u << 1,     /* add to the accumulators instead of clearing them first (progressive) */
    CORNELIS_RENDER_STAGE_TIMING = 1u << 2, /* time every stage kernel of one pass in 32 with CUDA events (stats *_ms) */
    CORNELIS_RENDER_DROP_NONFINITE = 1u << 3 /* skip finished paths whose radiance is NaN/inf.  Off by default: the
                                                reference lets them through (|w.z| rounding above 1 makes its
                                                Oren-Nayar term NaN about once per 1e8 samples) */
};

enum {
    CORNELIS_PIPELINE_DEFAULT = 0,   /* persistent for scenes scanned from shared memory, wavefront for grid scenes */
    CORNELIS_PIPELINE_WAVEFRONT = 1, /* raygen -> intersect(+compact) -> shade(+compact) -> accumulate kernels over
                                        a path pool in HBM; with the grid the intersect stage is a kernel of warps
                                        that pull rays from the pool plus a streaming compaction pass */
    CORNELIS_PIPELINE_PERSISTENT = 2 /* the same stage functions in ONE persistent kernel: a path lives in registers,
                                        Russian-roulette survivors wait in a warp-private shared-memory queue and are
                                        scattered 32 at a time, free warps start 32 camera paths (no pool in HBM) */
};

typedef struct cornelis_render_params {
    int32_t width, height;   /* the reference hard-codes 512 x 512 (Render.cpp:307) */
    int32_t samples;         /* RenderOptions::samplesAA of the whole image: the resolve divides by it */
    int32_t first_sample;    /* this call renders global sample indices [first_sample, first_sample + sample_count) */
    int32_t sample_count;    /*   of every pixel — the unit of multi-GPU sharding; 0 = all `samples` */
    int32_t max_depth;       /* <= 0: no cap, as the reference (Render.cpp:237); else paths stop after this many bounces
                                (<= 255).  Without a cap a path still ends at 255 bounces: only paths whose throughput
                                became NaN get there, and in a closed scene they would otherwise never end */
    uint64_t seed;           /* PRNG::DefaultSeed = 19791102 (PRNG.hpp:12) */
    uint32_t flags;          /* CORNELIS_RENDER_* */
    int32_t pipeline;        /* CORNELIS_PIPELINE_* */
    int32_t pool_paths;      /* paths in flight (wavefront width); 0 = default: 2^26, 2^25 or 2^24, the largest whose
                                pools and queues (172 bytes per path) fit a quarter of the free device memory */
    int32_t reserved;
} cornelis_render_params;

typedef struct cornelis_render_stats {
    uint64_t pixel_samples; /* camera paths started */
    uint64_t rays;          /* rays intersected = sum of active-list sizes in the reference's loop (Render.cpp:237-238) */
    uint64_t shaded_hits;   /* hits that reached accumulateAndBounce */
    uint64_t iterations;    /* wavefront passes */
    uint64_t kernel_launches;
    uint64_t contributions; /* finished paths whose (non-zero) radiance was added to a pixel */
    uint32_t max_depth;     /* deepest path, in bounces */
    uint32_t reserved;
    float gpu_ms;           /* device time of the render, CUDA events on the scene's stream */
    float intersect_ms, shade_ms, raygen_ms, accumulate_ms; /* mean device ms per launch of each stage kernel over the
                                                               sampled passes (CORNELIS_RENDER_STAGE_TIMING) */
} cornelis_render_stats;

typedef struct cornelis_cuda_scene cornelis_cuda_scene;

/* Progress callback: called on the calling thread between wavefront batches; return non-zero to abort
 * (RenderSession::ProgressCallback, Render.hpp:18-19). */
typedef int (*cornelis_progress_fn)(void *user, uint64_t samples_done, uint64_t samples_total);

/* ---- library ---------------------------------------------------------------------------------------------------- */

int cornelis_cuda_abi_version(void);
const char *cornelis_cuda_last_error(void);
int cornelis_cuda_device_count(int *count);
/* Device memory released by destroyed (or resized) handles is kept for the next handle of this process instead of
 * going back to the driver (the reference allocates its FrameBuffer per render, Render.cpp:307; on the GPU that
 * allocation is the expensive part of a short render).  This call returns the cached blocks to the driver. */
int cornelis_cuda_trim_memory(void);

/* ---- scene: replaces SceneData construction (Scene.cpp:40-53) + upload ------------------------------------------ */

/* `materials` is the complete list as SceneDescription::materials() returns it (index 0 = the default material). */
int cornelis_cuda_scene_create(int device, const cornelis_camera_desc *camera,
                               const cornelis_sphere_desc *spheres, size_t n_spheres,
                               const cornelis_plane_desc *planes, size_t n_planes,
                               const cornelis_material_desc *materials, size_t n_materials,
                               cornelis_cuda_scene **out_scene);
int cornelis_cuda_scene_destroy(cornelis_cuda_scene *scene);

/* Run this scene's kernels and copies on the caller's stream (a cudaStream_t passed as void*; NULL restores the
 * scene's own non-blocking stream).  Lets a host framework bracket the work with its own events. */
int cornelis_cuda_scene_set_stream(cornelis_cuda_scene *scene, void *cuda_stream);

/* How the intersect stage finds the closest sphere.  The reference scans every sphere for every ray
 * (Render.cpp:115-123); the grid evaluates the same per-sphere test on the spheres registered in the cells a ray
 * crosses and returns the same hit primitive and t bit for bit (cornelis_b200/csrc/geometry.cuh).
 * AUTO (the state after scene_create): grid for scenes with many spheres, exhaustive scan from shared memory otherwise. */
enum { CORNELIS_ACCEL_AUTO = 0, CORNELIS_ACCEL_NONE = 1, CORNELIS_ACCEL_GRID = 2 };
int cornelis_cuda_scene_set_acceleration(cornelis_cuda_scene *scene, int mode);
/* What is in use: grid on/off, its resolution, and the number of (cell, sphere) references.  Any output may be NULL. */
int cornelis_cuda_scene_acceleration(cornelis_cuda_scene *scene, int *grid_enabled, uint32_t dims[3],
                                     uint64_t *references);

/* ---- the hot path: replaces the tile loop + integrateTile (Render.cpp:220-255, 327-354) ------------------------- */

/* Render the sample range into the device accumulators (sum of per-sample radiance per pixel). */
int cornelis_cuda_render_accumulate(cornelis_cuda_scene *scene, const cornelis_render_params *params,
                                    cornelis_progress_fn progress, void *progress_user, cornelis_render_stats *stats);

/* Device pointer of the accumulation buffer: width*height float4 (sum r, g, b, and in .w the number of finished paths
 * that CONTRIBUTED, i.e. ended with non-zero radiance — not the sample count: the resolve divides by `samples`).
 * Valid until the next render with a different frame size or scene destroy. */
int cornelis_cuda_framebuffer_device(cornelis_cuda_scene *scene, void **device_ptr, size_t *n_floats);

/* ---- multi-GPU: the one exchange step of the path ----------------------------------------------------------------
 *
 * The estimator is a plain mean over samples (Render.cpp:245-250) and the device random numbers are keyed by the
 * GLOBAL sample index, so GPU g of G renders sample indices [g*spp/G, (g+1)*spp/G) of every pixel into its own
 * accumulators and the accumulators are summed ONCE: one ncclAllReduce (or ncclReduce) of width*height*4 floats per
 * GPU over NVLink; cornelis_cuda_resolve then applies 1/samples.  NCCL (libnccl.so.2) is loaded on first use —
 * the copy already in the process if there is one (e.g. torch's), else the system's — and its absence is an error
 * (CORNELIS_ERR_NCCL), not a fallback.  A communicator is either
 *   - all GPUs of one process: cornelis_cuda_comm_init_all (ncclCommInitAll), one scene handle per device, or
 *   - one rank of a one-process-per-GPU job: rank 0 calls cornelis_cuda_comm_unique_id, hands the 128 bytes to the
 *     other ranks by whatever means the job has (bench.py: torch.distributed), every rank calls
 *     cornelis_cuda_comm_init_rank (ncclCommInitRank).  */
typedef struct cornelis_cuda_comm cornelis_cuda_comm;
#define CORNELIS_COMM_ID_BYTES 128

int cornelis_cuda_comm_unique_id(uint8_t id[CORNELIS_COMM_ID_BYTES]);
int cornelis_cuda_comm_init_rank(const uint8_t id[CORNELIS_COMM_ID_BYTES], int rank, int n_ranks, int device,
                                 cornelis_cuda_comm **out_comm);
int cornelis_cuda_comm_init_all(const int *devices, int n_devices, cornelis_cuda_comm **out_comm);
int cornelis_cuda_comm_destroy(cornelis_cuda_comm *comm);
/* n_ranks of the communicator, how many of them live in this process, and NCCL's version code.  Outputs may be NULL. */
int cornelis_cuda_comm_info(cornelis_cuda_comm *comm, int *n_ranks, int *n_local, int *nccl_version);

/* In-place sum of the accumulation images (and of the second-moment images after CORNELIS_RENDER_VARIANCE renders) of
 * all ranks: afterwards every rank holds the image of all samples.  `scenes` are this process's handles in the
 * communicator's device order (n_local = 1 for a per-rank communicator); all must hold renders of the same frame.
 * One grouped ncclAllReduce, enqueued on each scene's stream: later calls on a scene are ordered behind it. */
int cornelis_cuda_allreduce_framebuffers(cornelis_cuda_comm *comm, cornelis_cuda_scene *const *scenes, int n_local);

/* One process, several GPUs: sum the accumulation images of n scenes into scenes[0] and wait for it.  Scenes on
 * distinct GPUs are summed by one grouped ncclReduce (root = scenes[0]'s device) over a process-wide communicator
 * that is created for that device set on first use and kept; scenes that share a GPU with an earlier one are first
 * added to it locally (no exchange to make).  RenderSession with RenderOptions::devices > 1 calls this. */
int cornelis_cuda_reduce_framebuffers(cornelis_cuda_scene *const *scenes, int n);

/* color = sum * (1.0f / samples) (Render.cpp:250) and download: host_rgb[3*W*H]; host_variance[3*W*H] optional
 * (unbiased per-sample variance, only after a CORNELIS_RENDER_VARIANCE render). */
int cornelis_cuda_resolve(cornelis_cuda_scene *scene, int32_t samples, float *host_rgb, float *host_variance);

/* The resolve alone: packed RGB (3*W*H floats) left in device memory owned by the scene; no host copy. */
int cornelis_cuda_resolve_device(cornelis_cuda_scene *scene, int32_t samples, void **device_rgb);

/* Same as cornelis_cuda_resolve, followed by the display transform and 8-bit quantisation on the device (Color.cpp:64-80,
 * FrameBuffer.hpp:91-95, i.e. what saveImage does before stbi_write_png, Render.cpp:257-265): host_rgb8[3*W*H]. */
int cornelis_cuda_resolve_srgb8(cornelis_cuda_scene *scene, int32_t samples, uint8_t *host_rgb8);

/* render_accumulate + resolve: scene in HBM, framebuffer back on the host — the end-to-end call. */
int cornelis_cuda_render(cornelis_cuda_scene *scene, const cornelis_render_params *params, float *host_rgb,
                         cornelis_render_stats *stats);

/* ---- stage entry points (parity tests and the intersection microbench); host buffers, packed xyz --------------- */

/* generateCameraRays arithmetic (Render.cpp:29-37, 85-100; Camera.cpp:11-13) for explicit jitter. */
int cornelis_cuda_pixel_rays(cornelis_cuda_scene *scene, int32_t width, int32_t height, size_t n, const int32_t *pi,
                             const int32_t *pj, const float *phi1, const float *phi2, float *org, float *dir);

/* Closest hit over all spheres then all planes (Render.cpp:110-140, Geometry.cpp:34-178).  prim = sphere index, or
 * n_spheres + plane index, or -1 (t = +inf).  P, N, mat may be NULL. */
int cornelis_cuda_intersect(cornelis_cuda_scene *scene, size_t n, const float *org, const float *dir, float *t,
                            int32_t *prim, float *P, float *N, int32_t *mat);

/* Same kernel on device-resident float4 ray arrays (org.xyz|-, dir.xyz|-) writing float2 hits (t, prim as int bits);
 * repeats the launch `repeats` times and returns the mean device milliseconds per launch. */
int cornelis_cuda_intersect_device(cornelis_cuda_scene *scene, size_t n, const void *d_org4, const void *d_dir4,
                                   void *d_hit2, int repeats, float *ms_per_launch);

/* The intersect stage of ONE wavefront pass on an explicit ray batch, compaction included (Render.cpp:110-150): the
 * rays are placed in the path pool as the render loop would hold them, the pass's own kernels run (k_intersect, or
 * k_walk + k_compact_hits for grid scenes), and the queues come back: hit_queue[0 .. *n_hits) are the indices of the
 * rays with t < INF — the reference's rebuilt activeList (Render.cpp:142-149), in no particular order — and
 * miss_queue[0 .. *n_misses) the indices of the rays that left the list (every ray of the batch carries non-zero
 * radiance, so each miss reaches the finished queue).  Both arrays must hold n entries. */
int cornelis_cuda_intersect_compact(cornelis_cuda_scene *scene, size_t n, const float *org, const float *dir,
                                    uint32_t *hit_queue, uint32_t *n_hits, uint32_t *miss_queue, uint32_t *n_misses);

/* LayeredBRDF::generateDirection / operator() / pdf (Materials.hpp:255-293) on explicit inputs. */
int cornelis_cuda_bsdf_sample(cornelis_cuda_scene *scene, size_t n, const int32_t *mat, const float *wo,
                              const float *N, const float *x, float *wi, float *pdf, float *f);
int cornelis_cuda_bsdf_eval(cornelis_cuda_scene *scene, size_t n, const int32_t *mat, const float *wi,
                            const float *wo, const float *N, float *f, float *pdf);

/* One accumulateAndBounce pass (Render.cpp:167-218) with explicit random numbers u[4*n] = (RR draw, x0, x1, x2).
 * In/out org, dir, thr, rad (3*n); in P, N (3*n), mat (n); out alive (n). */
int cornelis_cuda_shade(cornelis_cuda_scene *scene, size_t n, int32_t depth, const float *u, const float *P,
                        const float *N, const int32_t *mat, float *org, float *dir, float *thr, float *rad,
                        uint8_t *alive);

/* The counter-based generator that replaces the per-tile xoshiro stream (PRNG.hpp:11-37): Philox4x32 with
 * cornelis_cuda_rng_rounds() rounds (7: Crush-resistant per Salmon et al., SC'11; see cornelis_b200/csrc/rng.cuh), keyed
 * by `seed`, counter (pixel, sample, dimension block, 0).  out[4*n] = the four U[0,1) floats of each counter, with the
 * reference's 24-bit mapping (u >> 8) * 2^-24 (XoshiroCpp.hpp:651-655), produced by the render kernels' own code path.
 * Block 0 feeds the camera jitter, block d+1 the bounce at depth d. */
int cornelis_cuda_rng_uniforms(cornelis_cuda_scene *scene, uint64_t seed, size_t n, const uint32_t *pixel,
                               const uint32_t *sample, const uint32_t *block, float *out);
int cornelis_cuda_rng_rounds(void);
/* The raw 4x32-bit words of Philox4x32 with an explicit number of rounds (1..16), same key and counter layout: lets a
 * test put the published 10-round known-answer vectors through the library's round function. */
int cornelis_cuda_rng_bits(cornelis_cuda_scene *scene, int rounds, uint64_t seed, size_t n, const uint32_t *pixel,
                           const uint32_t *sample, const uint32_t *block, uint32_t *out);

/* Device self-test of the hand-scheduled exact arithmetic (cornelis_b200/csrc/exact_arith.cuh) against the IEEE
 * operators: mode 0 division, mode 1 square root (n crafted operand pairs, zeros and out-of-range operands included),
 * mode 2 the (sqrt, reciprocal) pair of normalize over float bit patterns 0..n-1 — n = 2^32 checks EVERY float.
 * *mismatches receives the number of results whose bits differ. */
int cornelis_cuda_selftest_arith(cornelis_cuda_scene *scene, int mode, uint64_t n, uint32_t seed, uint64_t *mismatches);

/* Device self-test of the display transform + 8-bit quantisation of cornelis_cuda_resolve_srgb8 (Color.cpp:64-80,
 * FrameBuffer.hpp:91-95) over consecutive float bit patterns: host_out[i] = srgb8(float whose bits are first_bits + i). */
int cornelis_cuda_selftest_srgb8(cornelis_cuda_scene *scene, uint32_t first_bits, size_t n, uint8_t *host_out);

#ifdef __cplusplus
}
#endif
#endif /* CORNELIS_CUDA_H */
