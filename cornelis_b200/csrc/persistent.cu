// Persistent-thread variant of the render loop: the same stage functions as the wavefront kernels (raygen,
// closestHit, shadeBounce, per-pixel accumulation), but a path lives in its thread's registers from its camera ray
// to its end and the lane that held it is refilled at once.
//
// Why: ncu on the wavefront kernels (profiles/) shows both compute stages bound by instruction latency at ~0.65 issued
// instructions per cycle per scheduler — a third of every warp's stall time is waiting for the pool's global loads
// and for the block barriers of the compaction, and ~20 % of the issued instructions move path state and queue
// entries.  For scenes whose primitives fit in shared memory none of that traffic is needed.
//
// What replaces the wavefront's machinery:
//   * pool + compaction (reference Render.cpp:142-149, :215-217): a lane whose path ended (miss, Russian roulette,
//     depth cap) claims the next camera path.  The claim is a warp operation: ballot of the lanes that need a path,
//     popc prefix for each lane's offset, ONE atomicAdd on the global cursor per 1024 camera paths per warp (a
//     warp-private stash of indices), so the lanes of a warp stay full without any queue.
//   * accumulate kernel (Render.cpp:245-248): the path's radiance is added to its pixel when the path ends, with
//     the same 128-bit vector reduction, skipped when it is exactly zero.
// Path identity, random numbers (Philox keyed by pixel, sample, depth) and all arithmetic are shared with the
// wavefront pipeline, so both produce the same per-path results; only fp32 summation order in the pixel differs.
#include <cstdlib>

#include "kernels.cuh"
#include "wavefront.h"

namespace cornelis_b200 {

constexpr unsigned long long kClaim = 1024; // camera paths a warp claims per atomic

#ifndef CORNELIS_PERSISTENT_MIN_BLOCKS
#define CORNELIS_PERSISTENT_MIN_BLOCKS 4
#endif

template <bool kGrid>
__global__ void __launch_bounds__(kBlockThreads, CORNELIS_PERSISTENT_MIN_BLOCKS)
    k_persistent(RenderConfig cfg, SceneView scene, unsigned long long *__restrict__ cursor, unsigned long long limit,
                 float4 *__restrict__ accum, float4 *__restrict__ accum2, bool dropNonFinite, Control *__restrict__ ctl) {
    extern __shared__ __align__(16) unsigned char smem[];
    SharedScene const sh = stageScene<kGrid>(scene, smem, true);
    constexpr unsigned kFull = 0xffffffffu;
    unsigned const lane = threadIdx.x & 31u;
    unsigned const below = (1u << lane) - 1u;

    // path state (registers)
    bool alive = false, exhausted = false;
    V3 org{0.f, 0.f, 0.f}, dir{0.f, 0.f, 0.f};
    RGBf thr{0.f, 0.f, 0.f}, rad{0.f, 0.f, 0.f};
    uint32_t pixel = 0, sample = 0, depth = 0;
    // warp-private stash of claimed camera-path indices [stashNext, stashEnd)
    unsigned long long stashNext = 0, stashEnd = 0;
    uint32_t stashPix = 0, stashSmp = 0; // (pixel, local sample) of stashNext
    // statistics
    uint32_t rays = 0, shaded = 0, started = 0, deepest = 0, contributed = 0;

    for (;;) {
        // ---- regeneration: lanes without a path claim camera paths (generateCameraRays, Render.cpp:85-100) ----
        bool const need = !alive && !exhausted;
        unsigned const needMask = __ballot_sync(kFull, need);
        if (needMask) {
            unsigned const count = __popc(needMask);
            unsigned const avail = static_cast<unsigned>(stashEnd - stashNext);
            unsigned long long fresh = 0;
            uint32_t freshPix = 0, freshSmp = 0;
            if (count > avail) { // refill: one atomic for the next kClaim paths; one 64-bit division per refill
                if (lane == 0) {
                    fresh = atomicAdd(cursor, kClaim);
                    unsigned long long const q = fresh / cfg.npixels;
                    freshSmp = static_cast<uint32_t>(q);
                    freshPix = static_cast<uint32_t>(fresh - q * cfg.npixels);
                }
                fresh = __shfl_sync(kFull, fresh, 0);
                freshPix = __shfl_sync(kFull, freshPix, 0);
                freshSmp = __shfl_sync(kFull, freshSmp, 0);
            }
            unsigned const rank = __popc(needMask & below);
            bool const fromStash = rank < avail;
            unsigned long long const index = fromStash ? stashNext + rank : fresh + (rank - avail);
            // path p = sampleLocal * npixels + pixel, from the stash's (pixel, sample) position: 32-bit arithmetic
            // (at most 32 past the last pixel of a frame: the wrap loops below run zero or one time unless the frame
            // has fewer than 32 pixels)
            uint32_t newPixel = fromStash ? stashPix + rank : freshPix + (rank - avail);
            uint32_t newSample = fromStash ? stashSmp : freshSmp;
            while (newPixel >= cfg.npixels) {
                newPixel -= cfg.npixels;
                newSample++;
            }
            if (count > avail) {
                stashNext = fresh + (count - avail);
                stashEnd = fresh + kClaim;
                stashPix = freshPix + (count - avail);
                stashSmp = freshSmp;
            } else {
                stashNext += count;
                stashPix += count;
            }
            while (stashPix >= cfg.npixels) { // keep the stash position normalised
                stashPix -= cfg.npixels;
                stashSmp++;
            }
            if (need) {
                if (index >= limit) {
                    exhausted = true;
                } else {
                    pixel = newPixel;
                    sample = cfg.firstSample + newSample;
                    uint32_t const j = fastDivide(pixel, cfg.byWidth), i = pixel - j * cfg.width;
                    Philox4 const r = philox4x32_10(pixel, sample, 0u, 0u, cfg.keys);
                    dir = pixelRayDirection(scene.camera, i, j, cfg.dx, cfg.dy, uniformFromBits(r.v[0]),
                                            uniformFromBits(r.v[1]));
                    org = V3{scene.camera.ex, scene.camera.ey, scene.camera.ez};
                    thr = RGBf{1.0f, 1.0f, 1.0f}; // Render.cpp:58
                    rad = RGBf{0.0f, 0.0f, 0.0f}; // Render.cpp:60
                    depth = 0;
                    alive = true;
                    started++;
                }
            }
        }
        if (!__any_sync(kFull, alive))
            break; // every lane is out of paths and the budget is exhausted

        // ---- intersect (Render.cpp:110-150) ----
        float t = INFINITY; // IntersectionData::reset, Geometry.cpp:7-12
        int32_t prim = -1;
        closestHitScene<kGrid>(alive, org, dir, sh, scene, t, prim);

        // ---- accumulateAndBounce (Render.cpp:167-218) ----
        bool finished = false;
        if (alive) {
            rays++;
            if (!(t < INFINITY)) { // Render.cpp:146: misses leave the active list
                finished = true;
            } else {
                V3 P, N;
                uint32_t material;
                hitSurface(org, dir, t, prim, sh.spheres, sh.sphereMaterial, scene.nSpheres, sh.planes, P, N, material);
                Philox4 const r = philox4x32_10(pixel, sample, depth + 1u, 0u, cfg.keys);
                bool const survives =
                    shadeBounce(sh.materials[material], P, N, depth, uniformFromBits(r.v[0]), uniformFromBits(r.v[1]),
                                uniformFromBits(r.v[2]), uniformFromBits(r.v[3]), org, dir, thr, rad);
                shaded++;
                depth = depth < 255u ? depth + 1u : 255u;
                deepest = depth > deepest ? depth : deepest;
                finished = !survives || (cfg.maxDepth && depth >= cfg.maxDepth);
            }
        }
        // ---- per-pixel accumulation (Render.cpp:245-248) ----
        if (finished) {
            alive = false;
            bool const nonZero = rad.r != 0.0f || rad.g != 0.0f || rad.b != 0.0f;
            bool const finite = isfinite(rad.r) && isfinite(rad.g) && isfinite(rad.b);
            contributed += nonZero ? 1u : 0u;
            if (nonZero && (finite || !dropNonFinite)) {
                atomicAdd(&accum[pixel], make_float4(rad.r, rad.g, rad.b, 1.0f));
                if (accum2)
                    atomicAdd(&accum2[pixel], make_float4(rad.r * rad.r, rad.g * rad.g, rad.b * rad.b, 0.0f));
            }
        }
    }

    // statistics: one set of atomics per warp
    rays = __reduce_add_sync(kFull, rays);
    shaded = __reduce_add_sync(kFull, shaded);
    started = __reduce_add_sync(kFull, started);
    contributed = __reduce_add_sync(kFull, contributed);
    deepest = __reduce_max_sync(kFull, deepest);
    if (lane == 0) {
        atomicAdd(&ctl->rays, static_cast<unsigned long long>(rays));
        atomicAdd(&ctl->shaded, static_cast<unsigned long long>(shaded));
        atomicAdd(&ctl->cursor, static_cast<unsigned long long>(started)); // camera paths actually started
        atomicAdd(&ctl->contributions, static_cast<unsigned long long>(contributed));
        atomicMax(&ctl->maxDepth, deepest);
    }
}

template <bool kGrid>
static cudaError_t configureOne(const LaunchShape &shape, int &grid) {
    cudaError_t e;
    if (shape.sceneSmemBytes > 48 * 1024)
        if ((e = cudaFuncSetAttribute(k_persistent<kGrid>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(shape.sceneSmemBytes))) != cudaSuccess)
            return e;
    int blocks = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_persistent<kGrid>, kBlockThreads,
                                                           shape.sceneSmemBytes)) != cudaSuccess)
        return e;
    if (const char *env = std::getenv("CORNELIS_PERSISTENT_BLOCKS_PER_SM"))
        if (std::atoi(env) > 0)
            blocks = std::atoi(env);
    grid = shape.numSMs * (blocks > 0 ? blocks : 1);
    return cudaSuccess;
}

cudaError_t configurePersistent(LaunchShape &shape, bool gridScene, int &grid) {
    return gridScene ? configureOne<true>(shape, grid) : configureOne<false>(shape, grid);
}

void launchPersistent(cudaStream_t s, const LaunchShape &shape, int grid, const RenderConfig &cfg, const SceneView &scene,
                      unsigned long long *cursor, unsigned long long limit, float4 *accum, float4 *accum2,
                      bool dropNonFinite, Control *ctl) {
    if (scene.grid.enabled)
        k_persistent<true><<<grid, kBlockThreads, shape.sceneSmemBytes, s>>>(cfg, scene, cursor, limit, accum, accum2,
                                                                             dropNonFinite, ctl);
    else
        k_persistent<false><<<grid, kBlockThreads, shape.sceneSmemBytes, s>>>(cfg, scene, cursor, limit, accum, accum2,
                                                                              dropNonFinite, ctl);
}

} // namespace cornelis_b200
