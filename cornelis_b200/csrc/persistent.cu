// Persistent-thread pipeline of the render loop: the same stage functions as the wavefront kernels (camera rays,
// closestHit, the two halves of accumulateAndBounce, per-pixel accumulation) in ONE kernel whose paths never leave the
// SM.  Two variants:
//   k_persistent_queued (default)  Russian-roulette survivors wait in a warp-private shared-memory queue; the BSDF half
//                                  runs on full batches of 32, free warps start 32 camera paths (see below)
//   k_persistent                   a path stays in its lane from camera ray to its end and the lane is refilled at once;
//                                  kept for scenes whose tables leave no shared memory for the queues
//
// Why: ncu on the wavefront kernels (profiles/) shows both compute stages bound by instruction latency at ~0.65 issued
// instructions per cycle per scheduler — a third of every warp's stall time is waiting for the pool's global loads
// and for the block barriers of the compaction, and ~20 % of the issued instructions move path state and queue
// entries.  For scenes whose primitives fit in shared memory none of that traffic is needed.
//
// What replaces the wavefront's machinery in k_persistent:
//   * pool + compaction (reference Render.cpp:142-149, :215-217): a lane whose path ended (miss, Russian roulette,
//     depth cap) claims the next camera path.  The claim is a warp operation: ballot of the lanes that need a path,
//     popc prefix for each lane's offset, ONE atomicAdd on the global cursor per claim of up to 1024 camera paths per
//     warp (a warp-private stash of indices), so the lanes of a warp stay full without any queue.
//   * accumulate kernel (Render.cpp:245-248): the path's radiance is added to its pixel when the path ends, with
//     the same 128-bit vector reduction, skipped when it is exactly zero.
// Path identity, random numbers (Philox keyed by pixel, sample, depth) and all arithmetic are shared with the
// wavefront pipeline, so all of them produce the same per-path results; only fp32 summation order in the pixel differs.
#include <cstdlib>

#include "kernels.cuh"
#include "wavefront.h"

namespace cornelis_b200 {

constexpr unsigned long long kMaxClaim = 1024; // camera paths a warp claims per atomic (RenderConfig::claim <= this)

// CTA shape of the persistent kernels, measured on B200 with the queued variant (Cornell 1080p, Msamples/s).
// Round 1 (hot loop ~28 KB against a 32 KB instruction cache, 72 registers wanted):
//   256 threads x 4 CTAs/SM (64 registers, 32 warps/SM) 6332     256 x 3 (80 registers, 24 warps) 7103
//   256 x 2 (128 registers, 16 warps) 7260    128 x 6 (80, 24 warps) 7396    128 x 7 (72 registers, 28 warps) 7509
// Round 2 (branch-free plane tests, 7 Philox rounds: a smaller loop that fits 64 registers without spills or
// rematerialisation — the same 2952 instructions at either budget; profiles/r2_persistent/variants.log):
//   128 x 6 (78 registers, 24 warps) 7947     128 x 7 (64 registers) 8097     128 x 8 (64 registers, 32 warps/SM) 8240
// A minimum of 7 CTAs per SM makes ptxas settle on 64 registers, and the occupancy calculator then finds room for 8.
// Later in round 2 (a shorter loop after the SASS-level pass over the shading step; gpurun_out/r3b in the same log):
//   __launch_bounds__(128, 7): 66 registers, 9404     (128, 6): 67 registers — still 7 CTAs per SM — and a schedule
//   that runs at 9593
#ifndef CORNELIS_PERSISTENT_MIN_BLOCKS
#define CORNELIS_PERSISTENT_MIN_BLOCKS 6
#endif
// Spheres whose discriminants are formed back to back before one vote decides whether any lane has a root among them
// (geometry.cuh scanSpheres); 1 = one sphere per loop trip.
#ifndef CORNELIS_RENDER_SPHERE_GROUP
#define CORNELIS_RENDER_SPHERE_GROUP 4
#endif
// The queued kernel's sphere scan: scalar, or two spheres per FFMA2 (packed FP32) from a paired copy of the table.
#ifndef CORNELIS_RENDER_PACKED
#define CORNELIS_RENDER_PACKED 0
#endif
#ifndef CORNELIS_PERSISTENT_THREADS
#define CORNELIS_PERSISTENT_THREADS 128
#endif
constexpr int kRenderScan = (CORNELIS_RENDER_PACKED && CORNELIS_RENDER_SPHERE_GROUP > 1) ? kScanPacked : kScanScalar;
constexpr int kPersistentThreads = CORNELIS_PERSISTENT_THREADS; // CTA size of the persistent kernels
constexpr int kPersistentWarps = kPersistentThreads / 32;
// Register budget: either a minimum number of resident CTAs (ptxas derives the cap) or an explicit cap
// (-DCORNELIS_PERSISTENT_MAXNREG=n; the two qualifiers exclude each other).
#ifdef CORNELIS_PERSISTENT_MAXNREG
#define CB_PERSISTENT_BOUNDS __maxnreg__(CORNELIS_PERSISTENT_MAXNREG)
#else
#define CB_PERSISTENT_BOUNDS __launch_bounds__(kPersistentThreads, CORNELIS_PERSISTENT_MIN_BLOCKS)
#endif

template <bool kGrid>
__global__ void CB_PERSISTENT_BOUNDS
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
            if (count > avail) { // refill: one atomic for the next cfg.claim paths; one 64-bit division per refill
                if (lane == 0) {
                    fresh = atomicAdd(cursor, static_cast<unsigned long long>(cfg.claim));
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
                stashEnd = fresh + cfg.claim;
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
                    Philox4 const r = philoxRender(pixel, sample, 0u, 0u, cfg.keys);
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
        closestHitScene<kGrid, CORNELIS_RENDER_SPHERE_GROUP, kScanEither, false>(alive, org, dir, sh, scene, t, prim);

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
                Philox4 const r = philoxRender(pixel, sample, depth + 1u, 0u, cfg.keys);
                bool const survives =
                    shadeBounce(sh.materials[material], P, N, depth, uniformFromBits(r.v[0]), uniformFromBits(r.v[1]),
                                uniformFromBits(r.v[2]), uniformFromBits(r.v[3]), org, dir, thr, rad);
                shaded++;
                depth += 1u; // <= kDepthLimit: a path that reaches it ends below
                deepest = depth > deepest ? depth : deepest;
                finished = !survives || (cfg.maxDepth && depth >= cfg.maxDepth) || depth >= kDepthLimit;
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

// ---------------------------------------------------------------------------------------- queued variant --
//
// ncu on k_persistent (profiles/r1_diet) shows the issue slots 86 % busy but only 23 of 32 lanes active per
// instruction: the BSDF half of a bounce runs with the 71 % of the lanes whose ray hit something AND survived Russian
// roulette, and the camera-ray code with the 29 % that need a new path.  k_persistent_queued parks the roulette
// survivors in a WARP-PRIVATE shared-memory queue (compaction #2 of the reference, Render.cpp:215-217, as a stack of
// 80-byte records) and runs the BSDF half only on full batches of 32; when fewer than 32 survivors are parked every
// lane is free, so the warp generates 32 camera rays (32 consecutive pixels of one sample index: a coherent packet).
// Every stage — camera rays, intersect, BSDF — therefore runs with all 32 lanes; only the short roulette half runs at
// the hit rate.  A record is five float4 in SoA order (slot-major within a chunk: a warp's 128-bit accesses are
// consecutive and conflict-free):
//     org.xyz | t      dir.xyz | prim      thr.rgb | pixel      rad.rgb | sample << 8 | depth      x0 x1 x2 | prob
// P, N and the material are re-derived from (org, dir, t, prim) when the record is popped, as in the wavefront shade
// kernel.  No block barrier anywhere: the queue belongs to one warp (__syncwarp orders its lanes' accesses).
constexpr uint32_t kQueueSlots = 64;  // a pop leaves fewer than 32 records, a push adds at most 32 (scatterParkedSlow too)
constexpr uint32_t kQueueChunks = 5;
constexpr size_t kQueueBytesPerWarp = kQueueSlots * kQueueChunks * sizeof(float4);

// The scatter half once more for a parked record whose fast normalisations left their range (math.cuh normalize with
// `odd`): hit point, normal and BSDF step with the per-call range checks, read from and written back to the record's
// queue slot (chunks 0..2: org, dir, thr), so nothing but the slot address crosses the call.  Rare and out of line.
static __device__ __noinline__ void scatterParkedSlow(float4 *queue, uint32_t slot, const DevSphere *spheres,
                                                      const uint32_t *sphereMaterial, uint32_t nSpheres,
                                                      const DevPlane *planes, const DevMaterial *materials) {
    constexpr uint32_t kSlots = kQueueSlots;
    float4 const q0 = queue[slot], q1 = queue[kSlots + slot], q2 = queue[2 * kSlots + slot], q4 = queue[4 * kSlots + slot];
    V3 org{q0.x, q0.y, q0.z}, dir{q1.x, q1.y, q1.z};
    RGBf thr{q2.x, q2.y, q2.z};
    V3 P, N;
    uint32_t material;
    hitSurface(org, dir, q0.w, static_cast<int32_t>(__float_as_uint(q1.w)), spheres, sphereMaterial, nSpheres, planes, P, N,
               material);
    shadeScatter(materials[material], P, N, q4.w, q4.x, q4.y, q4.z, org, dir, thr);
    queue[slot] = make_float4(org.x, org.y, org.z, q0.w);
    queue[kSlots + slot] = make_float4(dir.x, dir.y, dir.z, q1.w);
    queue[2 * kSlots + slot] = make_float4(thr.r, thr.g, thr.b, q2.w);
}

template <bool kGrid>
__global__ void CB_PERSISTENT_BOUNDS
    k_persistent_queued(RenderConfig cfg, SceneView scene, unsigned long long *__restrict__ cursor,
                        unsigned long long limit, float4 *__restrict__ accum, float4 *__restrict__ accum2,
                        bool dropNonFinite, Control *__restrict__ ctl, uint32_t queueOffset, uint32_t pairsOffset,
                        PackedConstants neutral) {
    extern __shared__ __align__(16) unsigned char smem[];
    SharedScene const sh = stageScene<kGrid>(scene, smem, true);
    // pairsOffset != 0: the sphere table once more as pairs behind the queues, for the packed-FP32 scan (geometry.cuh)
    float4 *const pairs = (!kGrid && pairsOffset) ? reinterpret_cast<float4 *>(smem + pairsOffset) : nullptr;
    if (pairs)
        stageSpherePairs(sh, scene.nSpheres, pairs);
    constexpr unsigned kFull = 0xffffffffu;
    unsigned const lane = threadIdx.x & 31u;
    unsigned const below = (1u << lane) - 1u;
    float4 *const queue = reinterpret_cast<float4 *>(smem + queueOffset) + (threadIdx.x >> 5) * (kQueueSlots * kQueueChunks);

    uint32_t parked = 0;        // records in the queue (warp-uniform)
    bool camerasLeft = true;    // warp-uniform: the camera-path range is not exhausted yet
    // warp-private stash of claimed camera-path indices [stashNext, stashEnd), always a multiple of 32 long
    unsigned long long stashNext = 0, stashEnd = 0;
    uint32_t stashPix = 0, stashSmp = 0; // (pixel, local sample) of stashNext
    uint32_t rays = 0, shaded = 0, started = 0, deepest = 0, contributed = 0;

    for (;;) {
        bool alive = false;
        V3 org{0.f, 0.f, 0.f}, dir{0.f, 0.f, 0.f};
        RGBf thr{0.f, 0.f, 0.f}, rad{0.f, 0.f, 0.f};
        uint32_t pixel = 0, sample = 0, depth = 0;
        if (parked >= 32u || (!camerasLeft && parked != 0u)) {
            // ---- second half of accumulateAndBounce (Render.cpp:194-213) for a batch of parked survivors ----
            // Every lane runs the step: a lane beyond the batch (only while the queue drains at the end of the run)
            // repeats the batch's last record and is not alive afterwards.  No lane-dependent branch around the
            // longest stretch of the loop, and no default values to set up for the lanes that would skip it.
            uint32_t const n = parked < 32u ? parked : 32u, base = parked - n;
            {
                uint32_t const slot = base + (lane < n ? lane : n - 1u);
                float4 const q0 = queue[slot], q1 = queue[kQueueSlots + slot], q2 = queue[2 * kQueueSlots + slot],
                             q3 = queue[3 * kQueueSlots + slot], q4 = queue[4 * kQueueSlots + slot];
                org = V3{q0.x, q0.y, q0.z};
                dir = V3{q1.x, q1.y, q1.z};
                thr = RGBf{q2.x, q2.y, q2.z};
                rad = RGBf{q3.x, q3.y, q3.z};
                pixel = __float_as_uint(q2.w);
                uint32_t const sd = __float_as_uint(q3.w);
                sample = sd >> 8;
                depth = sd & 255u;
                V3 P, N;
                uint32_t material;
                OddWatch odd = 0; // a normalisation outside the fast sequence's range: redo this record out of line
                hitSurface(org, dir, q0.w, static_cast<int32_t>(__float_as_uint(q1.w)), sh.spheres, sh.sphereMaterial,
                           scene.nSpheres, sh.planes, P, N, material, &odd);
                shadeScatter(sh.materials[material], P, N, q4.w, q4.x, q4.y, q4.z, org, dir, thr, &odd);
                if (oddRaised(odd)) {
                    scatterParkedSlow(queue, slot, sh.spheres, sh.sphereMaterial, scene.nSpheres, sh.planes, sh.materials);
                    float4 const r0 = queue[slot], r1 = queue[kQueueSlots + slot], r2 = queue[2 * kQueueSlots + slot];
                    org = V3{r0.x, r0.y, r0.z};
                    dir = V3{r1.x, r1.y, r1.z};
                    thr = RGBf{r2.x, r2.y, r2.z};
                }
                alive = lane < n;
            }
            parked = base;
            __syncwarp(); // the slots read here are overwritten by this iteration's pushes
        } else if (camerasLeft) {
            // ---- generateCameraRays (Render.cpp:85-100) for 32 consecutive camera paths ----
            if (stashNext == stashEnd) { // one atomic per cfg.claim paths per warp; one 64-bit division per refill
                unsigned long long fresh = 0;
                uint32_t freshPix = 0, freshSmp = 0;
                if (lane == 0) {
                    fresh = atomicAdd(cursor, static_cast<unsigned long long>(cfg.claim));
                    unsigned long long const q = fresh / cfg.npixels;
                    freshSmp = static_cast<uint32_t>(q);
                    freshPix = static_cast<uint32_t>(fresh - q * cfg.npixels);
                }
                stashNext = __shfl_sync(kFull, fresh, 0);
                stashEnd = stashNext + cfg.claim;
                stashPix = __shfl_sync(kFull, freshPix, 0);
                stashSmp = __shfl_sync(kFull, freshSmp, 0);
            }
            unsigned long long const index = stashNext + lane;
            uint32_t newPixel = stashPix + lane, newSample = stashSmp;
            while (newPixel >= cfg.npixels) { // runs at most once unless the frame has fewer than 32 pixels
                newPixel -= cfg.npixels;
                newSample++;
            }
            stashNext += 32u;
            stashPix += 32u;
            while (stashPix >= cfg.npixels) {
                stashPix -= cfg.npixels;
                stashSmp++;
            }
            camerasLeft = stashNext < limit; // the cursor only grows: once past the limit, always past it
            if (index < limit) {
                pixel = newPixel;
                sample = cfg.firstSample + newSample;
                uint32_t const j = fastDivide(pixel, cfg.byWidth), i = pixel - j * cfg.width;
                Philox4 const r = philoxRender(pixel, sample, 0u, 0u, cfg.keys);
                dir = pixelRayDirection(scene.camera, i, j, cfg.dx, cfg.dy, uniformFromBits(r.v[0]),
                                        uniformFromBits(r.v[1]));
                org = V3{scene.camera.ex, scene.camera.ey, scene.camera.ez};
                thr = RGBf{1.0f, 1.0f, 1.0f}; // Render.cpp:58
                rad = RGBf{0.0f, 0.0f, 0.0f}; // Render.cpp:60
                alive = true;
                started++;
            }
            if (!__any_sync(kFull, alive))
                continue; // this claim lay past the limit: drain the queue, or stop
        } else {
            break; // no camera paths left and nothing parked
        }

        // ---- intersect (Render.cpp:110-150) ----
        float t = INFINITY; // IntersectionData::reset, Geometry.cpp:7-12
        int32_t prim = -1;
        closestHitScene<kGrid, CORNELIS_RENDER_SPHERE_GROUP, kRenderScan, false>(alive, org, dir, sh, scene, t, prim, pairs,
                                                                                 neutral);

        // ---- first half of accumulateAndBounce (Render.cpp:174-192): emission, Russian roulette ----
        // Straight-line for every lane, no branch to reconverge before the ballot below: a lane without a hit — a miss
        // (Render.cpp:146: misses leave the active list) or a dead lane, which closestHit returns as one — draws its
        // numbers against the scene's first material and drops the results.
        bool const hit = t < INFINITY;
        uint32_t material = 0;
        if (hit)
            material = static_cast<uint32_t>(prim) < scene.nSpheres
                           ? sh.sphereMaterial[prim]
                           : sh.planes[prim - static_cast<int32_t>(scene.nSpheres)].material;
        Philox4 const r = philoxRender(pixel, sample, depth + 1u, 0u, cfg.keys);
        float prob;
        RGBf radNext = rad;
        bool const survives = shadeRoulette(sh.materials[material], depth, uniformFromBits(r.v[0]), thr, radNext, prob);
        float const x0 = uniformFromBits(r.v[1]), x1 = uniformFromBits(r.v[2]), x2 = uniformFromBits(r.v[3]);
        rad = hit ? radNext : rad;
        rays += alive ? 1u : 0u;
        shaded += hit ? 1u : 0u;
        depth += hit ? 1u : 0u; // <= kDepthLimit: a path that reaches it ends below
        deepest = depth > deepest ? depth : deepest;
        bool const park = hit && survives && !(cfg.maxDepth && depth >= cfg.maxDepth) && depth < kDepthLimit;
        bool const finished = alive && !park;
        // ---- compaction #2 (Render.cpp:215-217): survivors are pushed onto the warp's queue ----
        unsigned const parkMask = __ballot_sync(kFull, park);
        if (park) {
            uint32_t const slot = parked + __popc(parkMask & below);
            queue[slot] = make_float4(org.x, org.y, org.z, t);
            queue[kQueueSlots + slot] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(static_cast<uint32_t>(prim)));
            queue[2 * kQueueSlots + slot] = make_float4(thr.r, thr.g, thr.b, __uint_as_float(pixel));
            queue[3 * kQueueSlots + slot] = make_float4(rad.r, rad.g, rad.b, __uint_as_float((sample << 8) | depth));
            queue[4 * kQueueSlots + slot] = make_float4(x0, x1, x2, prob);
        }
        parked += __popc(parkMask);
        __syncwarp();
        // ---- per-pixel accumulation (Render.cpp:245-248) ----
        // Most paths end with nothing to add (only those that met the light carry radiance): one test for both.
        bool const nonZero = (rad.r != 0.0f) | (rad.g != 0.0f) | (rad.b != 0.0f);
        if (finished & nonZero) {
            contributed++;
            bool const finite = isfinite(rad.r) && isfinite(rad.g) && isfinite(rad.b);
            if (finite || !dropNonFinite) {
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

static bool packedRenderScan() { return kRenderScan == kScanPacked; }

static bool queuedVariant() {
    if (const char *env = std::getenv("CORNELIS_PERSISTENT_QUEUE"))
        return std::atoi(env) != 0;
    return true;
}

template <bool kGrid>
static cudaError_t configureOne(LaunchShape &shape, uint32_t nSpheres, int &grid) {
    cudaError_t e;
    // the queued variant appends one queue per warp to the staged scene tables; scenes whose tables leave no room for
    // the queues (close to the 227 KB carve-out) keep the lane-refill variant
    size_t const queueOffset = (shape.sceneSmemBytes + 15u) & ~static_cast<size_t>(15u);
    size_t queuedBytes = queueOffset + kPersistentWarps * kQueueBytesPerWarp;
    shape.persistentQueued = queuedVariant() && queuedBytes <= shape.smemOptin;
    shape.persistentPairsOffset = 0;
    if (shape.persistentQueued && !kGrid && packedRenderScan() && nSpheres >= 2u) {
        // the packed scan reads the paired table unconditionally: without room for it the lane-refill kernel (scalar
        // scan) renders the scene
        size_t const withPairs = queuedBytes + sizeof(float4) * (nSpheres & ~1u);
        if (withPairs <= shape.smemOptin) {
            shape.persistentPairsOffset = static_cast<uint32_t>(queuedBytes);
            queuedBytes = withPairs;
        } else {
            shape.persistentQueued = false;
        }
    }
    shape.persistentSmemBytes = shape.persistentQueued ? queuedBytes : shape.sceneSmemBytes;
    shape.persistentQueueOffset = static_cast<uint32_t>(queueOffset);
    auto kernel = shape.persistentQueued ? reinterpret_cast<const void *>(k_persistent_queued<kGrid>)
                                         : reinterpret_cast<const void *>(k_persistent<kGrid>);
    if (shape.smemOptin > 48 * 1024) { // per function and device, not per scene: always the device's limit
        cudaFuncAttributes attr{};
        if ((e = cudaFuncGetAttributes(&attr, kernel)) != cudaSuccess)
            return e;
        if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(shape.smemOptin - attr.sharedSizeBytes))) != cudaSuccess)
            return e;
    }
    if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared)) != cudaSuccess)
        return e;
    int blocks = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, kPersistentThreads,
                                                           shape.persistentSmemBytes)) != cudaSuccess)
        return e;
    if (const char *env = std::getenv("CORNELIS_PERSISTENT_BLOCKS_PER_SM"))
        if (std::atoi(env) > 0)
            blocks = std::atoi(env);
    grid = shape.numSMs * (blocks > 0 ? blocks : 1);
    return cudaSuccess;
}

// Camera paths a warp claims per atomic: 1024 for long renders (one 64-bit division per claim), fewer for short ones
// so that every warp of the grid gets at least ~8 claims and the last claim of the slowest warp is a small part of the
// launch.  Always a multiple of 32: the queued variant hands out whole batches.
uint32_t persistentClaim(unsigned long long paths, int grid) {
    unsigned long long const warps = static_cast<unsigned long long>(grid) * kPersistentWarps;
    unsigned long long claim = paths / (warps * 8ull);
    claim = claim / 32ull * 32ull;
    if (claim < 32ull)
        claim = 32ull;
    if (claim > kMaxClaim)
        claim = kMaxClaim;
    return static_cast<uint32_t>(claim);
}

cudaError_t configurePersistent(LaunchShape &shape, bool gridScene, uint32_t nSpheres, int &grid) {
    return gridScene ? configureOne<true>(shape, nSpheres, grid) : configureOne<false>(shape, nSpheres, grid);
}

void launchPersistent(cudaStream_t s, const LaunchShape &shape, int grid, const RenderConfig &cfg, const SceneView &scene,
                      unsigned long long *cursor, unsigned long long limit, float4 *accum, float4 *accum2,
                      bool dropNonFinite, Control *ctl) {
    size_t const smem = shape.persistentSmemBytes;
    if (shape.persistentQueued) {
        if (scene.grid.enabled)
            k_persistent_queued<true><<<grid, kPersistentThreads, smem, s>>>(cfg, scene, cursor, limit, accum, accum2,
                                                                        dropNonFinite, ctl, shape.persistentQueueOffset,
                                                                        shape.persistentPairsOffset, hostPackedConstants());
        else
            k_persistent_queued<false><<<grid, kPersistentThreads, smem, s>>>(cfg, scene, cursor, limit, accum, accum2,
                                                                         dropNonFinite, ctl, shape.persistentQueueOffset,
                                                                         shape.persistentPairsOffset, hostPackedConstants());
    } else if (scene.grid.enabled) {
        k_persistent<true><<<grid, kPersistentThreads, smem, s>>>(cfg, scene, cursor, limit, accum, accum2, dropNonFinite, ctl);
    } else {
        k_persistent<false><<<grid, kPersistentThreads, smem, s>>>(cfg, scene, cursor, limit, accum, accum2, dropNonFinite, ctl);
    }
}

} // namespace cornelis_b200
