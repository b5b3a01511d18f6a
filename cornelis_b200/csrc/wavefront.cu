// Wavefront kernels (see kernels.cuh for the pipeline overview) and their launchers.
#include "kernels.cuh"
#include "wavefront.h"

#include <cstdlib>

namespace cornelis_b200 {

// ---------------------------------------------------------------------------------------------------- plan --

// One thread.  Turns the survivors of the previous pass into the head of the new pool and decides how many camera
// paths to regenerate behind them.
__global__ void k_plan(Control *ctl, RenderConfig cfg) {
    uint32_t const survivors = ctl->nSurvive;
    unsigned long long const remaining = ctl->total - ctl->cursor;
    uint32_t const room = cfg.poolPaths - survivors;
    // (tiled generation: whole tiles of 32 paths, so that every warp of k_raygen and k_walk holds one tile)
    uint32_t const gen = remaining < room ? static_cast<uint32_t>(remaining) : cfg.tilesPerRow ? room & ~31u : room;
    ctl->genBase = survivors;
    ctl->genCount = gen;
    ctl->genFirst = ctl->cursor;
    ctl->cursor += gen;
    ctl->nIn = survivors + gen;
    ctl->nSurvive = 0;
    ctl->nHit = 0;
    ctl->nFinished = 0;
    ctl->walkCursor = 0;
    ctl->walkCursorCamera = survivors;
    ctl->rays += survivors + gen;
    ctl->iterations += (survivors + gen) ? 1 : 0;
}

// -------------------------------------------------------------------------------------------------- raygen --

// generateCameraRays (Render.cpp:85-100) for camera paths [genFirst, genFirst + genCount), written behind the
// survivors.  Path p = sampleLocal * npixels + pixel: consecutive threads take consecutive pixels of one sample
// index, so a warp's rays are coherent and its later framebuffer atomics hit distinct pixels.
__global__ void __launch_bounds__(kBlockThreads) k_raygen(const Control *ctl, RenderConfig cfg, DevCamera cam,
                                                          PathPool pool) {
    uint32_t const count = ctl->genCount;
    uint32_t const base = ctl->genBase;
    unsigned long long const first = ctl->genFirst;
    uint32_t const sample0 = static_cast<uint32_t>(first / cfg.npixels);
    uint32_t const pixel0 = static_cast<uint32_t>(first % cfg.npixels);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
        uint32_t pixel = pixel0 + k; // < npixels + poolPaths, no overflow
        uint32_t const wraps = pixel / cfg.npixels;
        pixel -= wraps * cfg.npixels;
        uint32_t const sample = cfg.firstSample + sample0 + wraps;
        uint32_t j, i;
        pixel = pixelOfPosition(pixel, cfg, i, j); // (so far the position within the frame)
        Philox4 const r = philoxRender(pixel, sample, 0u, 0u, cfg.keys);
        float const phi1 = uniformFromBits(r.v[0]), phi2 = uniformFromBits(r.v[1]); // Render.cpp:94-95
        V3 const d = pixelRayDirection(cam, i, j, cfg.dx, cfg.dy, phi1, phi2);
        uint32_t const slot = base + k;
        pool.org[slot] = make_float4(cam.ex, cam.ey, cam.ez, 0.0f);
        pool.dir[slot] = make_float4(d.x, d.y, d.z, 0.0f);
        pool.thr[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(packSampleDepth(sample, 0))); // Render.cpp:58
        pool.rad[slot] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(pixel));                     // Render.cpp:60
    }
}

// ------------------------------------------------------------------------------------------------ intersect --

// intersect (Render.cpp:110-150): closest hit of every pooled ray against all spheres then all planes held in
// shared memory, then compaction #1: hits are appended to the hit queue; misses end the path — if it carries
// radiance it goes to the finished queue, otherwise it simply disappears.
#ifndef CORNELIS_INTERSECT_MIN_BLOCKS
#define CORNELIS_INTERSECT_MIN_BLOCKS 5
#endif
template <bool kGrid>
__global__ void __launch_bounds__(kBlockThreads, CORNELIS_INTERSECT_MIN_BLOCKS) k_intersect(Control *ctl, SceneView scene, PathPool pool,
                                                             HitRecord *__restrict__ hits,
                                                             uint32_t *__restrict__ hitQueue,
                                                             FinishedPath *__restrict__ finished) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t scratch[2][kWarpsPerBlock + 1];
    SharedScene const sh = stageScene<kGrid>(scene, smem, false);
    uint32_t const n = ctl->nIn;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        uint32_t const i = base + threadIdx.x;
        bool const valid = i < n;
        bool hit = false, finish = false;
        float4 radiance = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = o4;
        if (valid) {
            o4 = pool.org[i];
            d4 = pool.dir[i];
        }
        float t = INFINITY; // IntersectionData::reset, Geometry.cpp:7-12
        int32_t prim = -1;
        closestHitScene<kGrid>(valid, V3{o4.x, o4.y, o4.z}, V3{d4.x, d4.y, d4.z}, sh, scene, t, prim);
        if (valid) {
            hit = t < INFINITY; // Render.cpp:146
            hits[i] = HitRecord{t, prim};
            if (!hit) {
                radiance = pool.rad[i];
                finish = radiance.x != 0.0f || radiance.y != 0.0f || radiance.z != 0.0f;
            }
        }
        AppendSlots const slot = blockAppend2(hit, finish, &ctl->nHit, &ctl->nFinished, scratch);
        if (hit)
            hitQueue[slot.a] = i;
        if (finish)
            finished[slot.b] = FinishedPath{radiance.x, radiance.y, radiance.z, __float_as_uint(radiance.w)};
    }
}

// ----------------------------------------------------------------------------------- intersect, grid scenes --

// Walking the grid costs anything from a handful to a hundred steps per ray, and with one ray per thread a warp waits
// for its longest walk: ncu on the one-ray-per-lane form shows 7.7 of 32 lanes active per instruction on config 4
// (profiles/r1_diet).  k_walk is a persistent kernel of warps that PULL rays: a lane whose walk is over takes the
// next pooled ray (a warp claims indices 64 at a time with one atomic, like the persistent pipeline's camera paths),
// so all lanes keep stepping until the pool is drained.  Refills are batched — they run when a quarter of the warp is
// idle — because setting up a walk (plane tests, clipping, DDA) is itself ~150 instructions.  Only the hit records are
// written here; k_compact_hits builds the queues (compaction #1) in a streaming pass.
#ifndef CORNELIS_WALK_CLAIM
#define CORNELIS_WALK_CLAIM 64
#endif
// rays a warp claims per atomic (a multiple of 32).  Small claims keep the warps of an SM on neighbouring stretches of the
// pool — whose paths, generated tile by tile, are neighbours in the scene — and shorten the tail of a pass: config 4 with
// 512 / 256 / 128 / 64 / 32: 1725 / 1800 / 1834 / 1852 / 1852 Msamples/s (profiles/r2_walk/variants.log).
constexpr unsigned kWalkClaim = CORNELIS_WALK_CLAIM;
static_assert(kWalkClaim % 32u == 0u && kWalkClaim != 0u, "the prepared walks are set up 32 at a time");
#ifndef CORNELIS_WALK_CAMERA_PHASE
#define CORNELIS_WALK_CAMERA_PHASE 1
#endif
#ifndef CORNELIS_WALK_CAMERA_CLAIM
#define CORNELIS_WALK_CAMERA_CLAIM 128
#endif
constexpr unsigned kWalkCameraClaim = CORNELIS_WALK_CAMERA_CLAIM; // camera rays a warp claims per atomic in phase 1: four coherent packets
// (8 while a refill meant setting rays up with the idle lanes; with prepared walks a refill is five shared-memory loads
// per lane and 4-6 measure best, profiles/r2_walk/variants.log)
#ifndef CORNELIS_WALK_REFILL
#define CORNELIS_WALK_REFILL 4
#endif
#ifndef CORNELIS_WALK_TEST_COST
#define CORNELIS_WALK_TEST_COST 11
#endif
#ifndef CORNELIS_WALK_ADVANCE_COST
#define CORNELIS_WALK_ADVANCE_COST 5
#endif
constexpr unsigned kWalkRefill = CORNELIS_WALK_REFILL; // idle lanes that trigger a refill
// relative instruction counts of the two kinds of step
constexpr unsigned kWalkTestCost = CORNELIS_WALK_TEST_COST, kWalkAdvanceCost = CORNELIS_WALK_ADVANCE_COST;

// 1: the walker reads its rays and writes its hit records with streaming (evict-first) accesses, leaving the L1 to
// the grid's cell ranges and references, which are what its lanes wait for (config 4: 1533 against 1506 Msamples/s,
// profiles/r2_walk/variants.log).
#ifndef CORNELIS_WALK_STREAMING
#define CORNELIS_WALK_STREAMING 1
#endif
#if CORNELIS_WALK_STREAMING
#define CB_WALK_LOAD4(p) __ldcs(p)
#define CB_WALK_STORE_HIT(p, t, prim) \
    __stcs(reinterpret_cast<float2 *>(p), make_float2(t, __int_as_float(prim)))
#else
#define CB_WALK_LOAD4(p) (*(p))
#define CB_WALK_STORE_HIT(p, t, prim) (*(p) = HitRecord{t, prim})
#endif
// 1: pooled rays are set up 32 at a time with every lane busy and parked, walk state included, in a queue of the warp
// in shared memory, from which the lanes whose walks are over take them; 0: a lane sets its next ray up itself.
#ifndef CORNELIS_WALK_PREPARED
#define CORNELIS_WALK_PREPARED 1
#endif
constexpr uint32_t kWalkQueueChunks = 5; // float4 per parked walk (walkBody)
constexpr size_t kWalkQueueBytesPerWarp = kWalkQueueChunks * 32u * sizeof(float4);
#ifndef CORNELIS_WALK_LEAN_ROUNDS
#define CORNELIS_WALK_LEAN_ROUNDS 1
#endif
#ifndef CORNELIS_WALK_MIN_BLOCKS
#define CORNELIS_WALK_MIN_BLOCKS 4
#endif
// kRangesInShared: the CTA is one of 1024 threads per SM and carries the cells' reference ranges in shared memory
// (nCells + 1 prefix offsets, up to ~200 KB) behind the staged scene tables.  ncu on the 256-thread form showed the
// walker waiting on memory, not issuing (long scoreboard 3.4 of 10.8 cycles per issue, issue slots 69 % busy, L1 hit
// rate 39 %): every cell crossing is a dependent load of the cell's range, then of its references.  The first of the two
// becomes a shared-memory access.  Same warps per SM either way (32: 62 registers).  Measured: no gain (1439 against
// 1488 Msamples/s on config 4) — the references and the spheres behind them are the loads that matter, and they stay
// where they were.  Off by default (LaunchShape::walkRangesInShared).
constexpr int kWalkSharedThreads = 1024;
template <bool kRangesInShared>
__device__ __forceinline__ void walkBody(Control *ctl, const SceneView &scene, const PathPool &pool,
                                         HitRecord *__restrict__ hits, uint32_t rangesOffset, uint32_t queueOffset) {
    extern __shared__ __align__(16) unsigned char smem[];
    SharedScene const sh = stageScene<true>(scene, smem, false);
    const uint32_t *cellStart = nullptr;
    if (kRangesInShared) {
        uint32_t *const dst = reinterpret_cast<uint32_t *>(smem + rangesOffset);
        uint32_t const nCells = scene.grid.nx * scene.grid.ny * scene.grid.nz;
        for (uint32_t c = threadIdx.x; c < nCells; c += blockDim.x)
            dst[c] = scene.grid.cellRange[c].x;
        if (threadIdx.x == 0)
            dst[nCells] = scene.grid.cellRange[nCells - 1u].y; // ranges are consecutive: the prefix sums of the builder
        __syncthreads();
        cellStart = dst;
    }
    constexpr unsigned kFull = 0xffffffffu;
    unsigned const lane = threadIdx.x & 31u;
    unsigned const below = (1u << lane) - 1u;
    // Phase 1: this pass's NEW camera rays, [genBase, nIn) of the pool (k_plan puts them behind the survivors).  A warp's
    // 32 consecutive camera paths are 32 consecutive pixels of one sample index: they cross nearly the same cells and
    // test the same spheres, so one ray per lane walked to completion keeps the lanes together WITHOUT the per-round
    // votes and refills of the pull model below (28 % of its instructions) and runs the set-up at full width
    // (9.5 lanes there).  CORNELIS_WALK_CAMERA_PHASE=0 at build time sends every ray through the pull model.
#if CORNELIS_WALK_CAMERA_PHASE
    {
        unsigned long long const cameraEnd = ctl->nIn;
        for (;;) {
            unsigned long long first = 0;
            if (lane == 0)
                first = atomicAdd(&ctl->walkCursorCamera, static_cast<unsigned long long>(kWalkCameraClaim));
            first = __shfl_sync(kFull, first, 0);
            if (first >= cameraEnd)
                break;
#pragma unroll 1
            for (unsigned k = 0; k < kWalkCameraClaim; k += 32u) {
                unsigned long long const mine = first + k + lane;
                if (mine < cameraEnd) {
                    uint32_t const at = static_cast<uint32_t>(mine);
                    float4 const o4 = CB_WALK_LOAD4(pool.org + at), d4 = CB_WALK_LOAD4(pool.dir + at);
                    float tc = INFINITY; // IntersectionData::reset, Geometry.cpp:7-12
                    int32_t pc = -1;
                    closestHitGrid(true, V3{o4.x, o4.y, o4.z}, V3{d4.x, d4.y, d4.z}, scene, sh.planes, tc, pc, nullptr,
                                   cellStart);
                    CB_WALK_STORE_HIT(hits + at, tc, pc);
                }
            }
        }
    }
    unsigned long long const n = ctl->genBase; // phase 2: the survivors of the last pass, [0, genBase)
#else
    unsigned long long const n = ctl->nIn;
#endif
    unsigned long long stashNext = 0, stashEnd = 0;
    bool walking = false, exhausted = false;
    uint32_t index = 0;
    V3 o{0.f, 0.f, 0.f}, d{0.f, 0.f, 0.f};
    float t = INFINITY;
    int32_t prim = -1;
    GridWalk w{};
#if CORNELIS_WALK_PREPARED
    // The warp's queue of prepared walks: up to 32 records of five float4, chunk c of record e at [32 c + e]:
    //   (o, t so far) (d, primitive so far) (tx, ty, tz, cell) (dtx, dty, dtz, ray index)
    //   (tMargin, steps left per axis in 9 bits each | the three direction signs, k, last)
    // In the pull model a lane used to set its next ray up when its walk was over — 9 lanes of 32 at a time on config 4,
    // and 38 % of the kernel's instructions (profiles/r2_walk).  Now the warp sets 32 rays up together whenever its
    // queue is empty and enough lanes wait, every lane busy, next to the walks the other lanes still hold in their
    // registers; a ray that needs no walk gets its hit record at once.
    float4 *const queue = reinterpret_cast<float4 *>(smem + queueOffset) + (threadIdx.x >> 5) * (kWalkQueueChunks * 32u);
    unsigned qHead = 0, qCount = 0; // warp-uniform
    bool raysLeft = n != 0;         // warp-uniform: the claims have not run past the pool yet
    // The lanes that hold a walk, warp-uniform and kept up to date where it changes (a refill, a round of advances): the
    // rounds below then cost ONE vote each (CORNELIS_WALK_LEAN_ROUNDS; 0: the round-1 loop that votes on both kinds of
    // step every round and leaves the burst every 32 rounds).
    unsigned walkMask = 0u;
    for (;;) {
#if !CORNELIS_WALK_LEAN_ROUNDS
        walkMask = __ballot_sync(kFull, walking);
#endif
        unsigned const idleMask = ~walkMask;
        if (walkMask == 0u && qCount == 0u && !raysLeft)
            break; // nothing walking, nothing parked, nothing left to claim
        unsigned const nIdle = __popc(idleMask);
        bool const refill = nIdle >= kWalkRefill || walkMask == 0u;
        if (refill && qCount == 0u && raysLeft) {
            if (stashNext == stashEnd) {
                unsigned long long fresh = 0;
                if (lane == 0)
                    fresh = atomicAdd(&ctl->walkCursor, static_cast<unsigned long long>(kWalkClaim));
                stashNext = __shfl_sync(kFull, fresh, 0);
                stashEnd = stashNext + kWalkClaim;
            }
            unsigned long long const mine = stashNext + lane;
            stashNext += 32u;
            raysLeft = stashNext < n; // the cursor only grows: once past the pool, always past it
            bool started = false;
            V3 po{0.f, 0.f, 0.f}, pd{0.f, 0.f, 0.f};
            float pt = INFINITY; // IntersectionData::reset, Geometry.cpp:7-12
            int32_t pprim = -1;
            GridWalk pw{};
            uint32_t const pindex = static_cast<uint32_t>(mine);
            if (mine < n) {
                float4 const o4 = CB_WALK_LOAD4(pool.org + pindex), d4 = CB_WALK_LOAD4(pool.dir + pindex);
                po = V3{o4.x, o4.y, o4.z};
                pd = V3{d4.x, d4.y, d4.z};
                started = gridWalkBegin(pw, po, pd, scene, sh.planes, pt, pprim, nullptr, cellStart);
                if (!started)
                    CB_WALK_STORE_HIT(hits + pindex, pt, pprim); // decided without a walk (degenerate, outside the grid, ...)
            }
            unsigned const startedMask = __ballot_sync(kFull, started);
            if (started) {
                unsigned const slot = __popc(startedMask & below);
                uint32_t const packed = static_cast<uint32_t>(pw.leftX) | static_cast<uint32_t>(pw.leftY) << 9 |
                                        static_cast<uint32_t>(pw.leftZ) << 18 | (pw.strideX > 0 ? 1u << 27 : 0u) |
                                        (pw.strideY > 0 ? 1u << 28 : 0u) | (pw.strideZ > 0 ? 1u << 29 : 0u);
                queue[slot] = make_float4(po.x, po.y, po.z, pt);
                queue[32u + slot] = make_float4(pd.x, pd.y, pd.z, __int_as_float(pprim));
                queue[64u + slot] = make_float4(pw.tx, pw.ty, pw.tz, __int_as_float(pw.cell));
                queue[96u + slot] = make_float4(pw.dtx, pw.dty, pw.dtz, __uint_as_float(pindex));
                queue[128u + slot] = make_float4(pw.tMargin, __uint_as_float(packed), __uint_as_float(pw.k),
                                                 __uint_as_float(pw.last));
            }
            qHead = 0;
            qCount = __popc(startedMask);
            __syncwarp();
        }
        if (refill && qCount != 0u) {
            unsigned const take = nIdle < qCount ? nIdle : qCount;
            unsigned const rank = __popc(idleMask & below);
            if (!walking && rank < take) {
                unsigned const e = qHead + rank;
                float4 const q0 = queue[e], q1 = queue[32u + e], q2 = queue[64u + e], q3 = queue[96u + e],
                             q4 = queue[128u + e];
                o = V3{q0.x, q0.y, q0.z}, t = q0.w;
                d = V3{q1.x, q1.y, q1.z}, prim = __float_as_int(q1.w);
                w.tx = q2.x, w.ty = q2.y, w.tz = q2.z, w.cell = __float_as_int(q2.w);
                w.dtx = q3.x, w.dty = q3.y, w.dtz = q3.z, index = __float_as_uint(q3.w);
                uint32_t const packed = __float_as_uint(q4.y);
                int32_t const nx = static_cast<int32_t>(scene.grid.nx), nxy = static_cast<int32_t>(scene.grid.nx * scene.grid.ny);
                w.tMargin = q4.x, w.k = __float_as_uint(q4.z), w.last = __float_as_uint(q4.w);
                w.leftX = static_cast<int32_t>(packed & 511u), w.leftY = static_cast<int32_t>((packed >> 9) & 511u),
                w.leftZ = static_cast<int32_t>((packed >> 18) & 511u);
                w.strideX = (packed & (1u << 27)) ? 1 : -1, w.strideY = (packed & (1u << 28)) ? nx : -nx,
                w.strideZ = (packed & (1u << 29)) ? nxy : -nxy;
                w.A = dot(d, d);
                w.rA = rcpSeedRefined(w.A);
                w.lastTested = 0xffffffffu;
#if CORNELIS_GRID_MAILBOX2
                w.prevTested = 0xffffffffu;
#endif
                walking = true;
            }
            qHead += take;
            qCount -= take;
            __syncwarp(); // the records read here are overwritten by the next refill
            walkMask = __ballot_sync(kFull, walking);
        }
#if CORNELIS_WALK_LEAN_ROUNDS
        // A burst of rounds (see below for the two kinds of step and why a round runs only one of them).  Lanes finish
        // only in a round of advances and start only in a refill, so the set of walking lanes is re-voted there and
        // nowhere else: a round of tests costs one vote, one population count and one compare on top of the test.
        // nAdvance * kWalkTestCost >= nTest * kWalkAdvanceCost with nTest = nWalk - nAdvance.
        {
            unsigned nWalk = __popc(walkMask);
#pragma unroll 1
            while (nWalk != 0u) {
                bool const wantAdvance = walking && w.k >= w.last;
                unsigned const nAdvance = __popc(__ballot_sync(kFull, wantAdvance));
                if (nAdvance * (kWalkTestCost + kWalkAdvanceCost) >= nWalk * kWalkAdvanceCost) {
                    if (wantAdvance) {
                        walking = gridWalkAdvance(w, scene.grid, t, nullptr, cellStart);
                        if (!walking)
                            CB_WALK_STORE_HIT(hits + index, t, prim);
                    }
                    walkMask = __ballot_sync(kFull, walking);
                    nWalk = __popc(walkMask);
                    // (leave the burst for a refill only while there is something to refill from)
                    if ((qCount != 0u || raysLeft) && 32u - nWalk >= kWalkRefill)
                        break;
                } else if (walking && !wantAdvance) {
                    gridWalkTest(w, o, d, scene.grid, t, prim);
                }
            }
        }
        continue;
#endif
#else
    for (;;) {
        unsigned const idleMask = __ballot_sync(kFull, !walking && !exhausted);
        unsigned const walkMask = __ballot_sync(kFull, walking);
        if (idleMask && (__popc(idleMask) >= kWalkRefill || walkMask == 0u)) {
            unsigned const count = __popc(idleMask);
            unsigned const avail = static_cast<unsigned>(stashEnd - stashNext);
            unsigned long long fresh = 0;
            if (count > avail) {
                if (lane == 0)
                    fresh = atomicAdd(&ctl->walkCursor, static_cast<unsigned long long>(kWalkClaim));
                fresh = __shfl_sync(kFull, fresh, 0);
            }
            unsigned const rank = __popc(idleMask & below);
            unsigned long long const mine = rank < avail ? stashNext + rank : fresh + (rank - avail);
            if (count > avail) {
                stashNext = fresh + (count - avail);
                stashEnd = fresh + kWalkClaim;
            } else {
                stashNext += count;
            }
            if (!walking && !exhausted) {
                if (mine >= n) {
                    exhausted = true;
                } else {
                    index = static_cast<uint32_t>(mine);
                    float4 const o4 = CB_WALK_LOAD4(pool.org + index), d4 = CB_WALK_LOAD4(pool.dir + index);
                    o = V3{o4.x, o4.y, o4.z};
                    d = V3{d4.x, d4.y, d4.z};
                    t = INFINITY; // IntersectionData::reset, Geometry.cpp:7-12
                    prim = -1;
                    walking = gridWalkBegin(w, o, d, scene, sh.planes, t, prim, nullptr, cellStart);
                    if (!walking)
                        CB_WALK_STORE_HIT(hits + index, t, prim); // decided without a walk (degenerate, outside the grid, ...)
                }
            }
        } else if (walkMask == 0u) {
            break; // nothing walking, nothing left to claim
        }
#endif
        // A burst of rounds.  A walking lane wants one of two things: to TEST the next sphere of its cell (~55
        // instructions) or, when the cell has none left — empty cells are the common case — to ADVANCE to the next
        // cell (~25).  Executing both per round (one mixed step per lane) left 11.9 of 32 lanes active per instruction
        // (profiles/r1_queue: 18.1 lanes after the change); instead every round runs only the kind of step that moves more lanes per instruction
        // issued, and the others wait for a round of their kind.  The burst stops early when enough lanes have finished
        // to make a refill worthwhile.
#pragma unroll 1
        for (int burst = 0; burst < 32; burst++) {
            bool const wantAdvance = walking && w.k >= w.last;
            unsigned const nAdvance = __popc(__ballot_sync(kFull, wantAdvance));
            unsigned const nTest = __popc(__ballot_sync(kFull, walking && !wantAdvance));
            if (nAdvance + nTest == 0u)
                break;
            if (nAdvance * kWalkTestCost >= nTest * kWalkAdvanceCost) {
                if (wantAdvance) {
                    walking = gridWalkAdvance(w, scene.grid, t, nullptr, cellStart);
                    if (!walking)
                        CB_WALK_STORE_HIT(hits + index, t, prim);
                }
#if CORNELIS_WALK_PREPARED
                // (leave the burst for a refill only while there is something to refill from)
                if ((qCount != 0u || raysLeft) && __popc(__ballot_sync(kFull, !walking)) >= kWalkRefill)
                    break;
#else
                if (__popc(__ballot_sync(kFull, !walking && !exhausted)) >= kWalkRefill)
                    break;
#endif
            } else if (walking && !wantAdvance) {
                gridWalkTest(w, o, d, scene.grid, t, prim);
            }
        }
    }
}

__global__ void __launch_bounds__(kBlockThreads, CORNELIS_WALK_MIN_BLOCKS)
    k_walk(Control *ctl, SceneView scene, PathPool pool, HitRecord *__restrict__ hits, uint32_t queueOffset) {
    walkBody<false>(ctl, scene, pool, hits, 0u, queueOffset);
}

__global__ void __launch_bounds__(kWalkSharedThreads, 1)
    k_walk_shared(Control *ctl, SceneView scene, PathPool pool, HitRecord *__restrict__ hits, uint32_t rangesOffset,
                  uint32_t queueOffset) {
    walkBody<true>(ctl, scene, pool, hits, rangesOffset, queueOffset);
}

// Compaction #1 for grid scenes (Render.cpp:142-149): hits to the hit queue, misses that carry radiance to the finished
// queue — the tail of k_intersect as a streaming pass over the hit records.
__global__ void __launch_bounds__(kBlockThreads) k_compact_hits(Control *ctl, PathPool pool,
                                                                const HitRecord *__restrict__ hits,
                                                                uint32_t *__restrict__ hitQueue,
                                                                FinishedPath *__restrict__ finished) {
    __shared__ uint32_t scratch[2][kWarpsPerBlock + 1];
    uint32_t const n = ctl->nIn;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        uint32_t const i = base + threadIdx.x;
        bool hit = false, finish = false;
        float4 radiance = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n) {
            hit = hits[i].t < INFINITY; // Render.cpp:146
            if (!hit) {
                radiance = pool.rad[i];
                finish = radiance.x != 0.0f || radiance.y != 0.0f || radiance.z != 0.0f;
            }
        }
        AppendSlots const slot = blockAppend2(hit, finish, &ctl->nHit, &ctl->nFinished, scratch);
        if (hit)
            hitQueue[slot.a] = i;
        if (finish)
            finished[slot.b] = FinishedPath{radiance.x, radiance.y, radiance.z, __float_as_uint(radiance.w)};
    }
}

// ---------------------------------------------------------------------------------------------------- shade --

// accumulateAndBounce (Render.cpp:167-218) over the hit queue, then compaction #2: survivors are written
// CONTIGUOUSLY into the next pool (so the next pass reads coalesced float4 streams); paths killed by Russian
// roulette or the depth cap go to the finished queue if they carry radiance.
#ifndef CORNELIS_SHADE_MIN_BLOCKS
#define CORNELIS_SHADE_MIN_BLOCKS 5
#endif
template <bool kGrid>
__global__ void __launch_bounds__(kBlockThreads, CORNELIS_SHADE_MIN_BLOCKS) k_shade(Control *ctl, RenderConfig cfg, SceneView scene,
                                                         PathPool in, PathPool out,
                                                         const HitRecord *__restrict__ hits,
                                                         const uint32_t *__restrict__ hitQueue,
                                                         FinishedPath *__restrict__ finished) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t scratch[2][kWarpsPerBlock + 1];
    SharedScene const sh = stageScene<kGrid>(scene, smem, true);
    uint32_t const n = ctl->nHit;
    uint32_t deepest = 0;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        uint32_t const q = base + threadIdx.x;
        bool const valid = q < n;
        bool alive = false, finish = false;
        V3 org{}, dir{};
        RGBf thr{}, rad{};
        uint32_t pixel = 0, sample = 0, depth = 0;
        if (valid) {
            uint32_t const i = hitQueue[q];
            float4 const o4 = in.org[i], d4 = in.dir[i], t4 = in.thr[i], r4 = in.rad[i];
            HitRecord const h = hits[i];
            org = V3{o4.x, o4.y, o4.z};
            dir = V3{d4.x, d4.y, d4.z};
            thr = RGBf{t4.x, t4.y, t4.z};
            rad = RGBf{r4.x, r4.y, r4.z};
            uint32_t const sd = __float_as_uint(t4.w);
            sample = sd >> 8;
            depth = sd & 255u;
            pixel = __float_as_uint(r4.w);
            V3 P, N;
            uint32_t material;
            hitSurface(org, dir, h.t, h.prim, sh.spheres, sh.sphereMaterial, scene.nSpheres, sh.planes, P, N, material);
            Philox4 const r = philoxRender(pixel, sample, depth + 1u, 0u, cfg.keys);
            alive = shadeBounce(sh.materials[material], P, N, depth, uniformFromBits(r.v[0]), uniformFromBits(r.v[1]),
                                uniformFromBits(r.v[2]), uniformFromBits(r.v[3]), org, dir, thr, rad);
            depth += 1;
            deepest = depth > deepest ? depth : deepest;
            if ((cfg.maxDepth && depth >= cfg.maxDepth) || depth >= kDepthLimit)
                alive = false;
            finish = !alive && (rad.r != 0.0f || rad.g != 0.0f || rad.b != 0.0f);
        }
        AppendSlots const slot = blockAppend2(alive, finish, &ctl->nSurvive, &ctl->nFinished, scratch);
        if (alive) {
            out.org[slot.a] = make_float4(org.x, org.y, org.z, 0.0f);
            out.dir[slot.a] = make_float4(dir.x, dir.y, dir.z, 0.0f);
            out.thr[slot.a] = make_float4(thr.r, thr.g, thr.b, __uint_as_float(packSampleDepth(sample, depth)));
            out.rad[slot.a] = make_float4(rad.r, rad.g, rad.b, __uint_as_float(pixel));
        }
        if (finish)
            finished[slot.b] = FinishedPath{rad.r, rad.g, rad.b, pixel};
    }
    // statistics: one atomic per block
    deepest = __reduce_max_sync(0xffffffffu, deepest);
    if ((threadIdx.x & 31u) == 0 && deepest)
        atomicMax(&ctl->maxDepth, deepest);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&ctl->shaded, static_cast<unsigned long long>(n));
}

// ----------------------------------------------------------------------------------------------- accumulate --

// Per-pixel accumulation (Render.cpp:245-248): every finished path adds its radiance to its pixel's running sum
// with one 128-bit vector reduction (red.global.add.v4.f32); .w counts contributing paths.  With the variance
// option the squares go to a second float4 image.
__global__ void __launch_bounds__(kBlockThreads) k_accumulate(Control *ctl,
                                                              const FinishedPath *__restrict__ finished,
                                                              float4 *__restrict__ accum, float4 *__restrict__ accum2,
                                                              bool dropNonFinite) {
    uint32_t const n = ctl->nFinished;
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&ctl->contributions, static_cast<unsigned long long>(n));
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        FinishedPath const f = finished[q];
        if (dropNonFinite && !(isfinite(f.r) && isfinite(f.g) && isfinite(f.b)))
            continue;
        atomicAdd(&accum[f.pixel], make_float4(f.r, f.g, f.b, 1.0f));
        if (accum2)
            atomicAdd(&accum2[f.pixel], make_float4(f.r * f.r, f.g * f.g, f.b * f.b, 0.0f));
    }
}

// -------------------------------------------------------------------------------------------------- resolve --

// color = sum * (1.0f / samplesAA) (Render.cpp:250) into packed RGB (FrameBuffer.hpp:66, Color.hpp:56);
// optionally the unbiased per-sample variance from the second moments.
__global__ void __launch_bounds__(kBlockThreads) k_resolve(uint32_t npixels, float invSamples, uint32_t samples,
                                                           const float4 *__restrict__ accum,
                                                           const float4 *__restrict__ accum2, float *__restrict__ rgb,
                                                           float *__restrict__ variance) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < npixels; p += gridDim.x * blockDim.x) {
        float4 const s = accum[p];
        rgb[3 * p + 0] = s.x * invSamples;
        rgb[3 * p + 1] = s.y * invSamples;
        rgb[3 * p + 2] = s.z * invSamples;
        if (variance) {
            float4 const q = accum2[p];
            double const n = samples;
            double const sx[3] = {s.x, s.y, s.z}, sq[3] = {q.x, q.y, q.z};
            for (int c = 0; c < 3; c++) {
                double v = samples > 1 ? (sq[c] - sx[c] * sx[c] / n) / (n - 1.0) : 0.0;
                variance[3 * p + c] = static_cast<float>(v > 0.0 ? v : 0.0);
            }
        }
    }
}

// toSRGB (Color.cpp:64-80: 12.95 slope, double pow) + quantizeTo8bit (FrameBuffer.hpp:91-95) fused with the resolve
// — what saveImage (Render.cpp:257-261) does on the CPU before PNG encoding.
__device__ __forceinline__ uint8_t srgb8(float x) {
    float s;
    if (static_cast<double>(x) <= 0.0031308)
        s = x * 12.95f;
    else
        s = static_cast<float>((1 + 0.055f) * pow(static_cast<double>(x), static_cast<double>(1.0f / 2.4f)) - 0.055f);
    double v = round(255.0 * static_cast<double>(s));
    v = (v < 0.0) ? 0.0 : (255.0 < v) ? 255.0 : v;
    return static_cast<uint8_t>(v);
}

__global__ void __launch_bounds__(kBlockThreads) k_resolve_srgb8(uint32_t npixels, float invSamples,
                                                                 const float4 *__restrict__ accum,
                                                                 uint8_t *__restrict__ rgb8) {
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < npixels; p += gridDim.x * blockDim.x) {
        float4 const s = accum[p];
        rgb8[3 * p + 0] = srgb8(s.x * invSamples);
        rgb8[3 * p + 1] = srgb8(s.y * invSamples);
        rgb8[3 * p + 2] = srgb8(s.z * invSamples);
    }
}

// The display transform of k_resolve_srgb8 over consecutive float bit patterns: out[i] = srgb8(float with bits first + i).
// 2^30 + 1 patterns cover every float in [0, 1] (tests/test_gpu_parity.py compares all of them with the oracle).
__global__ void __launch_bounds__(kBlockThreads) k_srgb8_sweep(uint32_t first, size_t n, uint8_t *__restrict__ out) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = srgb8(__uint_as_float(first + static_cast<uint32_t>(i)));
}

// dst += src over float4 images; src may live on a peer GPU (NVLink peer access): the loads go over the link, the
// adds and stores are local.
__global__ void __launch_bounds__(kBlockThreads) k_add_images(size_t n4, float4 *__restrict__ dst,
                                                              const float4 *__restrict__ src) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float4 a = dst[i];
        float4 const b = src[i];
        a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
        dst[i] = a;
    }
}

// ------------------------------------------------------------------------------- stage kernels on plain arrays --

__global__ void __launch_bounds__(kBlockThreads) k_pixel_rays(DevCamera cam, uint32_t n, float dx, float dy,
                                                              const int32_t *pi, const int32_t *pj, const float *phi1,
                                                              const float *phi2, float *org, float *dir) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        V3 const d = pixelRayDirection(cam, static_cast<uint32_t>(pi[k]), static_cast<uint32_t>(pj[k]), dx, dy,
                                       phi1[k], phi2[k]);
        org[3 * k] = cam.ex, org[3 * k + 1] = cam.ey, org[3 * k + 2] = cam.ez;
        dir[3 * k] = d.x, dir[3 * k + 1] = d.y, dir[3 * k + 2] = d.z;
    }
}

// The intersect stage on a plain float4 ray batch — the same closestHit as k_intersect without the queues.
// Used by cornelis_cuda_intersect(_device): parity tests and the intersection microbench (config 3).
#ifndef CORNELIS_BATCH_SPHERE_GROUP
#define CORNELIS_BATCH_SPHERE_GROUP 16 // spheres per discriminant vote (geometry.cuh scanSpheres)
#endif
template <bool kGrid>
__global__ void __launch_bounds__(kBlockThreads) k_intersect_batch(SceneView scene, size_t n,
                                                                   const float4 *__restrict__ org,
                                                                   const float4 *__restrict__ dir,
                                                                   HitRecord *__restrict__ hits, uint32_t pairOffset,
                                                                   PackedConstants neutral) {
    extern __shared__ __align__(16) unsigned char smem[];
    SharedScene const sh = stageScene<kGrid>(scene, smem, false);
    // pairOffset != 0: room for the paired sphere table behind the staged scene (packed FP32 scan)
    float4 *pairs = (!kGrid && pairOffset) ? reinterpret_cast<float4 *>(smem + pairOffset) : nullptr;
    if (pairs)
        stageSpherePairs(sh, scene.nSpheres, pairs);
    for (size_t base = static_cast<size_t>(blockIdx.x) * blockDim.x; base < n;
         base += static_cast<size_t>(gridDim.x) * blockDim.x) {
        size_t const i = base + threadIdx.x;
        bool const valid = i < n;
        float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = o4;
        if (valid) {
            o4 = org[i];
            d4 = dir[i];
        }
        float t = INFINITY;
        int32_t prim = -1;
        closestHitScene<kGrid, CORNELIS_BATCH_SPHERE_GROUP>(valid, V3{o4.x, o4.y, o4.z}, V3{d4.x, d4.y, d4.z}, sh, scene, t,
                                                            prim, pairs, neutral);
        if (valid)
            hits[i] = HitRecord{t, prim};
    }
}

// The same on closestHit2 (geometry2.cuh): two rays per thread — ray i and ray i + ceil(n / 2) — on packed FP32.  What
// the render kernel's pair mode runs, exposed here so that the parity tests of the intersect stage cover it
// (CORNELIS_BATCH_PAIRS=1 routes cornelis_cuda_intersect(_device) through it; scenes scanned from shared memory only).
__global__ void __launch_bounds__(kBlockThreads) k_intersect_batch2(SceneView scene, size_t n,
                                                                    const float4 *__restrict__ org,
                                                                    const float4 *__restrict__ dir,
                                                                    HitRecord *__restrict__ hits, uint32_t splatOffset,
                                                                    PackedConstants neutral) {
    extern __shared__ __align__(16) unsigned char smem[];
    SharedScene const sh = stageScene<false>(scene, smem, false);
    SplattedScene const splat = stageSplatted(sh, scene, reinterpret_cast<float4 *>(smem + splatOffset));
    size_t const half = (n + 1) / 2;
    for (size_t base = static_cast<size_t>(blockIdx.x) * blockDim.x; base < half;
         base += static_cast<size_t>(gridDim.x) * blockDim.x) {
        size_t const a = base + threadIdx.x, b = a + half;
        bool const liveA = a < half, liveB = liveA && b < n;
        float4 const zero = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 const oa = liveA ? org[a] : zero, da = liveA ? dir[a] : zero;
        float4 const ob = liveB ? org[b] : zero, db = liveB ? dir[b] : zero;
        float tA, tB;
        int32_t primA, primB;
        closestHit2(liveA, liveB, V3{oa.x, oa.y, oa.z}, V3{da.x, da.y, da.z}, V3{ob.x, ob.y, ob.z}, V3{db.x, db.y, db.z}, sh,
                    splat, scene, neutral, tA, primA, tB, primB);
        if (liveA)
            hits[a] = HitRecord{tA, primA};
        if (liveB)
            hits[b] = HitRecord{tB, primB};
    }
}

// Expands hit records into the reference's IntersectionData fields (P, N, MaterialId; Geometry.hpp:7-15).
template <bool kGrid>
__global__ void __launch_bounds__(kBlockThreads) k_hit_surface(SceneView scene, size_t n,
                                                               const float4 *__restrict__ org,
                                                               const float4 *__restrict__ dir,
                                                               const HitRecord *__restrict__ hits, float *P, float *N,
                                                               int32_t *mat) {
    extern __shared__ __align__(16) unsigned char smem[];
    SharedScene const sh = stageScene<kGrid>(scene, smem, true);
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        HitRecord const h = hits[i];
        V3 p{0, 0, 0}, nrm{0, 0, 0};
        uint32_t m = 0xffffffffu;
        if (h.prim >= 0) {
            float4 const o4 = org[i], d4 = dir[i];
            hitSurface(V3{o4.x, o4.y, o4.z}, V3{d4.x, d4.y, d4.z}, h.t, h.prim, sh.spheres, sh.sphereMaterial,
                       scene.nSpheres, sh.planes, p, nrm, m);
        }
        if (P)
            P[3 * i] = p.x, P[3 * i + 1] = p.y, P[3 * i + 2] = p.z;
        if (N)
            N[3 * i] = nrm.x, N[3 * i + 1] = nrm.y, N[3 * i + 2] = nrm.z;
        if (mat)
            mat[i] = static_cast<int32_t>(m);
    }
}

__global__ void __launch_bounds__(kBlockThreads) k_bsdf_sample(const DevMaterial *materials, uint32_t n,
                                                               const int32_t *mat, const float *wo, const float *N,
                                                               const float *x, float *wi, float *pdf, float *f) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        Basis const b = constructBasis(V3{N[3 * k], N[3 * k + 1], N[3 * k + 2]});
        V3 w;
        float p;
        RGBf const v = layeredSample(materials[mat[k]], V3{wo[3 * k], wo[3 * k + 1], wo[3 * k + 2]}, x[3 * k],
                                     x[3 * k + 1], x[3 * k + 2], b, w, p);
        wi[3 * k] = w.x, wi[3 * k + 1] = w.y, wi[3 * k + 2] = w.z;
        f[3 * k] = v.r, f[3 * k + 1] = v.g, f[3 * k + 2] = v.b;
        pdf[k] = p;
    }
}

__global__ void __launch_bounds__(kBlockThreads) k_bsdf_eval(const DevMaterial *materials, uint32_t n,
                                                             const int32_t *mat, const float *wi, const float *wo,
                                                             const float *N, float *f, float *pdf) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        V3 const a{wi[3 * k], wi[3 * k + 1], wi[3 * k + 2]}, o{wo[3 * k], wo[3 * k + 1], wo[3 * k + 2]};
        V3 const nrm{N[3 * k], N[3 * k + 1], N[3 * k + 2]};
        DevMaterial const &m = materials[mat[k]];
        float p;
        RGBf const v = layeredEvalPdf(m, a, o, nrm, p);
        f[3 * k] = v.r, f[3 * k + 1] = v.g, f[3 * k + 2] = v.b;
        pdf[k] = p;
    }
}

__global__ void __launch_bounds__(kBlockThreads) k_shade_explicit(const DevMaterial *materials, uint32_t n,
                                                                  uint32_t depth, const float *u, const float *P,
                                                                  const float *N, const int32_t *mat, float *org,
                                                                  float *dir, float *thr, float *rad, uint8_t *alive) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        V3 o{org[3 * k], org[3 * k + 1], org[3 * k + 2]}, d{dir[3 * k], dir[3 * k + 1], dir[3 * k + 2]};
        RGBf T{thr[3 * k], thr[3 * k + 1], thr[3 * k + 2]}, L{rad[3 * k], rad[3 * k + 1], rad[3 * k + 2]};
        bool const a = shadeBounce(materials[mat[k]], V3{P[3 * k], P[3 * k + 1], P[3 * k + 2]},
                                   V3{N[3 * k], N[3 * k + 1], N[3 * k + 2]}, depth, u[4 * k], u[4 * k + 1],
                                   u[4 * k + 2], u[4 * k + 3], o, d, T, L);
        alive[k] = a ? 1 : 0;
        org[3 * k] = o.x, org[3 * k + 1] = o.y, org[3 * k + 2] = o.z;
        dir[3 * k] = d.x, dir[3 * k + 1] = d.y, dir[3 * k + 2] = d.z;
        thr[3 * k] = T.r, thr[3 * k + 1] = T.g, thr[3 * k + 2] = T.b;
        rad[3 * k] = L.r, rad[3 * k + 1] = L.g, rad[3 * k + 2] = L.b;
    }
}

// rounds == 0: the render loop's generator (philoxRender over the expanded keys), as the kernels call it; otherwise
// Philox4x32 with that many rounds.  Writes the four uniforms and / or the four raw words of every counter.
__global__ void __launch_bounds__(kBlockThreads) k_rng(uint32_t n, int rounds, uint32_t key0, uint32_t key1,
                                                       PhiloxKeys keys, const uint32_t *pixel, const uint32_t *sample,
                                                       const uint32_t *block, float *uniforms, uint32_t *bits) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        Philox4 const r = rounds == 0 ? philoxRender(pixel[k], sample[k], block[k], 0u, keys)
                                      : philox4x32(rounds, pixel[k], sample[k], block[k], 0u, key0, key1);
        for (int c = 0; c < 4; c++) {
            if (uniforms)
                uniforms[4 * k + c] = uniformFromBits(r.v[c]);
            if (bits)
                bits[4 * k + c] = r.v[c];
        }
    }
}

// Self-test of the exact fast paths of geometry.cuh against the IEEE operators: operands with random sign and
// mantissa (one in eight with an all-ones / all-zeros / single-bit mantissa) and exponents spanning the ranges the fast
// paths claim.  mode 0: division, mode 1: square root, mode 2: the normalize pair (sqrt, reciprocal) over EVERY float
// bit pattern i < 2^32 in its range (n = 2^32 makes it exhaustive).  Counts results whose bits differ.
__global__ void __launch_bounds__(kBlockThreads) k_selftest_arith(int mode, unsigned long long n, uint32_t seed,
                                                                  unsigned long long *mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
        Philox4 const r = philox4x32(10, static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32), 0x5e1f7e57u, 0u, seed, 0u);
        auto craft = [](uint32_t bits, uint32_t style, int eLo, int eHi) {
            uint32_t mant = bits & 0x7fffffu;
            switch (style & 7u) {
            case 0: mant = 0x7fffffu; break;
            case 1: mant = 0u; break;
            case 2: mant = 1u << ((bits >> 3) % 23u); break;
            case 3: mant = 0x7fffffu ^ (1u << ((bits >> 3) % 23u)); break;
            default: break;
            }
            int const e = eLo + static_cast<int>((bits >> 23) % static_cast<uint32_t>(eHi - eLo + 1));
            return __uint_as_float((bits & 0x80000000u) | (static_cast<uint32_t>(e + 127) << 23) | mant);
        };
        if (mode == 2) { // exhaustive: i enumerates float bit patterns; every float in the fast normalize range
            float const x = __uint_as_float(static_cast<uint32_t>(i));
            if (i < (1ull << 32) && inFastNormalizeRange(x)) {
                float len, s;
                sqrtAndReciprocalExactFast(x, len, s);
                float const lenRef = sqrtf(x);
                bad += __float_as_uint(len) != __float_as_uint(lenRef);
                bad += (__float_as_uint(s) != __float_as_uint(1.0f / lenRef)) ? (1ull << 32) : 0ull; // high word: s
            }
            continue;
        }
        if (mode == 0) {
            float a = craft(r.v[0], r.v[2], -80, 79);
            float const b = craft(r.v[1], r.v[2] >> 3, -40, 39);
            float const rb = rcpSeedRefined(b);
            bad += __float_as_uint(divideExactFast(a, b, rb)) != __float_as_uint(a / b);
            if ((r.v[3] & 15u) == 0u)
                a = __uint_as_float(r.v[3] & 0x80000000u); // +-0: the zero-numerator variant must keep the sign
            bad += __float_as_uint(divideExactFast0(a, b, rb)) != __float_as_uint(a / b);
            // the per-lane wrapper, with operands inside and outside the fast ranges
            float const a2 = craft(r.v[0], r.v[2], -126, 126), b2 = craft(r.v[1], r.v[2] >> 3, -126, 126);
            bad += __float_as_uint(divideExact(a2, b2)) != __float_as_uint(a2 / b2);
            bad += __float_as_uint(divideExact(a, b2)) != __float_as_uint(a / b2);
        } else {
            float const x = fabsf(craft(r.v[0], r.v[2], -100, 125));
            bad += __float_as_uint(sqrtExactFast(x)) != __float_as_uint(sqrtf(x));
            float y = fabsf(craft(r.v[1], r.v[2] >> 3, -126, 127));
            if ((r.v[3] & 15u) == 0u)
                y = __uint_as_float(r.v[3] & 0x80000000u); // +-0
            if ((r.v[3] & 15u) == 1u)
                y = __uint_as_float(r.v[3] >> 9);          // denormals
            bad += __float_as_uint(sqrtExact(y)) != __float_as_uint(sqrtf(y));
        }
    }
    if (bad)
        atomicAdd(mismatches, bad);
}

// Packed xyz -> float4 staging for the host-buffer stage entry points.
__global__ void __launch_bounds__(kBlockThreads) k_pack4(size_t n, const float *xyz, float4 *out) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.0f);
}

__global__ void __launch_bounds__(kBlockThreads) k_unpack_hits(size_t n, const HitRecord *hits, float *t,
                                                               int32_t *prim) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        t[i] = hits[i].t;
        prim[i] = hits[i].prim;
    }
}

// ---------------------------------------------------------------------------------------------- launchers --

static inline int gridFor(size_t n, int numSMs, int blocksPerSM) {
    size_t const want = (n + kBlockThreads - 1) / kBlockThreads;
    size_t const cap = static_cast<size_t>(numSMs) * blocksPerSM;
    size_t g = want < cap ? want : cap;
    return static_cast<int>(g ? g : 1);
}

void launchPlan(cudaStream_t s, Control *ctl, const RenderConfig &cfg) { k_plan<<<1, 1, 0, s>>>(ctl, cfg); }

void launchRaygen(cudaStream_t s, const LaunchShape &shape, const Control *ctl, const RenderConfig &cfg,
                  const DevCamera &cam, const PathPool &pool) {
    k_raygen<<<shape.gridRaygen, kBlockThreads, 0, s>>>(ctl, cfg, cam, pool);
}

// shared memory of the warps' queues of prepared walks behind a walker CTA's other tables
static size_t walkQueueBytes(int threads) {
    return CORNELIS_WALK_PREPARED ? static_cast<size_t>(threads / 32) * kWalkQueueBytesPerWarp : 0u;
}

void launchIntersect(cudaStream_t s, const LaunchShape &shape, Control *ctl, const SceneView &scene,
                     const PathPool &pool, HitRecord *hits, uint32_t *hitQueue, FinishedPath *finished) {
    if (scene.grid.enabled) {
        // (a scene whose tables leave no shared memory for the walker's queues takes the one-ray-per-thread kernel)
        bool const roomForQueues = ((shape.sceneSmemBytes + 15u) & ~static_cast<size_t>(15u)) +
                                       walkQueueBytes(kBlockThreads) + 1024u <= shape.smemOptin;
        if (shape.walkPull && roomForQueues) {
            // the cells' ranges ride in shared memory when they fit behind the scene tables (one 1024-thread CTA per SM);
            // the warps' queues of prepared walks come last
            size_t const rangesOffset = (shape.sceneSmemBytes + 15u) & ~static_cast<size_t>(15u);
            size_t const nCells = static_cast<size_t>(scene.grid.nx) * scene.grid.ny * scene.grid.nz;
            size_t const withRanges = (rangesOffset + sizeof(uint32_t) * (nCells + 1u) + 15u) & ~static_cast<size_t>(15u);
            size_t const sharedQueues = walkQueueBytes(kWalkSharedThreads);
            if (shape.walkRangesInShared && withRanges + sharedQueues + 1024u <= shape.smemOptin)
                k_walk_shared<<<shape.numSMs, kWalkSharedThreads, withRanges + sharedQueues, s>>>(
                    ctl, scene, pool, hits, static_cast<uint32_t>(rangesOffset), static_cast<uint32_t>(withRanges));
            else
                k_walk<<<shape.gridWalk, kBlockThreads, rangesOffset + walkQueueBytes(kBlockThreads), s>>>(
                    ctl, scene, pool, hits, static_cast<uint32_t>(rangesOffset));
            k_compact_hits<<<shape.gridAccumulate, kBlockThreads, 0, s>>>(ctl, pool, hits, hitQueue, finished);
        } else {
            k_intersect<true><<<shape.gridIntersectGrid, kBlockThreads, shape.sceneSmemBytes, s>>>(ctl, scene, pool, hits,
                                                                                                   hitQueue, finished);
        }
    } else
        k_intersect<false><<<shape.gridIntersect, kBlockThreads, shape.sceneSmemBytes, s>>>(ctl, scene, pool, hits,
                                                                                            hitQueue, finished);
}

void launchShade(cudaStream_t s, const LaunchShape &shape, Control *ctl, const RenderConfig &cfg,
                 const SceneView &scene, const PathPool &in, const PathPool &out, const HitRecord *hits,
                 const uint32_t *hitQueue, FinishedPath *finished) {
    if (scene.grid.enabled)
        k_shade<true><<<shape.gridShade, kBlockThreads, shape.sceneSmemBytes, s>>>(ctl, cfg, scene, in, out, hits,
                                                                                   hitQueue, finished);
    else
        k_shade<false><<<shape.gridShade, kBlockThreads, shape.sceneSmemBytes, s>>>(ctl, cfg, scene, in, out, hits,
                                                                                    hitQueue, finished);
}

void launchAccumulate(cudaStream_t s, const LaunchShape &shape, Control *ctl, const FinishedPath *finished,
                      float4 *accum, float4 *accum2, bool dropNonFinite) {
    k_accumulate<<<shape.gridAccumulate, kBlockThreads, 0, s>>>(ctl, finished, accum, accum2, dropNonFinite);
}

void launchResolve(cudaStream_t s, const LaunchShape &shape, uint32_t npixels, uint32_t samples, const float4 *accum,
                   const float4 *accum2, float *rgb, float *variance) {
    float const inv = 1.0f / static_cast<float>(static_cast<int32_t>(samples)); // 1.0f / options.samplesAA
    k_resolve<<<gridFor(npixels, shape.numSMs, 8), kBlockThreads, 0, s>>>(npixels, inv, samples, accum, accum2, rgb,
                                                                          variance);
}

void launchResolveSrgb8(cudaStream_t s, const LaunchShape &shape, uint32_t npixels, uint32_t samples,
                        const float4 *accum, uint8_t *rgb8) {
    float const inv = 1.0f / static_cast<float>(static_cast<int32_t>(samples));
    k_resolve_srgb8<<<gridFor(npixels, shape.numSMs, 8), kBlockThreads, 0, s>>>(npixels, inv, accum, rgb8);
}

void launchSrgb8Sweep(cudaStream_t s, const LaunchShape &shape, uint32_t first, size_t n, uint8_t *out) {
    k_srgb8_sweep<<<gridFor(n, shape.numSMs, 8), kBlockThreads, 0, s>>>(first, n, out);
}

void launchAddImages(cudaStream_t s, const LaunchShape &shape, size_t n4, float4 *dst, const float4 *src) {
    k_add_images<<<gridFor(n4, shape.numSMs, 8), kBlockThreads, 0, s>>>(n4, dst, src);
}

void launchPixelRays(cudaStream_t s, const LaunchShape &shape, const DevCamera &cam, uint32_t n, float dx, float dy,
                     const int32_t *pi, const int32_t *pj, const float *phi1, const float *phi2, float *org,
                     float *dir) {
    k_pixel_rays<<<gridFor(n, shape.numSMs, 8), kBlockThreads, 0, s>>>(cam, n, dx, dy, pi, pj, phi1, phi2, org, dir);
}

void launchIntersectBatch(cudaStream_t s, const LaunchShape &shape, const SceneView &scene, size_t n,
                          const float4 *org, const float4 *dir, HitRecord *hits) {
    if (scene.grid.enabled) {
        k_intersect_batch<true><<<gridFor(n, shape.numSMs, shape.blocksPerSM), kBlockThreads, shape.sceneSmemBytes, s>>>(
            scene, n, org, dir, hits, 0u, hostPackedConstants());
        return;
    }
    size_t const pairOffset = (shape.sceneSmemBytes + 15u) & ~static_cast<size_t>(15u);
    if (const char *env = std::getenv("CORNELIS_BATCH_PAIRS")) {
        size_t const withSplat = pairOffset + splattedBytes(scene.nSpheres, scene.planeEnd[2]);
        if (std::atoi(env) != 0 && withSplat <= shape.smemOptin) {
            static bool optedIn = false;
            if (!optedIn) {
                cudaFuncSetAttribute(k_intersect_batch2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(shape.smemOptin));
                optedIn = true;
            }
            k_intersect_batch2<<<gridFor((n + 1) / 2, shape.numSMs, shape.blocksPerSM), kBlockThreads, withSplat, s>>>(
                scene, n, org, dir, hits, static_cast<uint32_t>(pairOffset), hostPackedConstants());
            return;
        }
    }
    // the exhaustive scan runs on packed FP32 when the paired copy of the sphere table fits behind the staged scene
    size_t const withPairs = pairOffset + sizeof(float4) * (scene.nSpheres & ~1u);
    bool const packed = shape.batchPacked && scene.nSpheres >= 2u && withPairs <= shape.smemOptin;
    k_intersect_batch<false><<<gridFor(n, shape.numSMs, shape.blocksPerSM), kBlockThreads,
                               packed ? withPairs : shape.sceneSmemBytes, s>>>(
        scene, n, org, dir, hits, packed ? static_cast<uint32_t>(pairOffset) : 0u, hostPackedConstants());
}

void launchHitSurface(cudaStream_t s, const LaunchShape &shape, const SceneView &scene, size_t n, const float4 *org,
                      const float4 *dir, const HitRecord *hits, float *P, float *N, int32_t *mat) {
    if (scene.grid.enabled)
        k_hit_surface<true><<<gridFor(n, shape.numSMs, shape.blocksPerSM), kBlockThreads, shape.sceneSmemBytes, s>>>(
            scene, n, org, dir, hits, P, N, mat);
    else
        k_hit_surface<false><<<gridFor(n, shape.numSMs, shape.blocksPerSM), kBlockThreads, shape.sceneSmemBytes, s>>>(
            scene, n, org, dir, hits, P, N, mat);
}

void launchBsdfSample(cudaStream_t s, const LaunchShape &shape, const DevMaterial *materials, uint32_t n,
                      const int32_t *mat, const float *wo, const float *N, const float *x, float *wi, float *pdf,
                      float *f) {
    k_bsdf_sample<<<gridFor(n, shape.numSMs, 8), kBlockThreads, 0, s>>>(materials, n, mat, wo, N, x, wi, pdf, f);
}

void launchBsdfEval(cudaStream_t s, const LaunchShape &shape, const DevMaterial *materials, uint32_t n,
                    const int32_t *mat, const float *wi, const float *wo, const float *N, float *f, float *pdf) {
    k_bsdf_eval<<<gridFor(n, shape.numSMs, 8), kBlockThreads, 0, s>>>(materials, n, mat, wi, wo, N, f, pdf);
}

void launchShadeExplicit(cudaStream_t s, const LaunchShape &shape, const DevMaterial *materials, uint32_t n,
                         uint32_t depth, const float *u, const float *P, const float *N, const int32_t *mat,
                         float *org, float *dir, float *thr, float *rad, uint8_t *alive) {
    k_shade_explicit<<<gridFor(n, shape.numSMs, 8), kBlockThreads, 0, s>>>(materials, n, depth, u, P, N, mat, org, dir,
                                                                           thr, rad, alive);
}

void launchRng(cudaStream_t s, const LaunchShape &shape, uint32_t n, int rounds, uint32_t key0, uint32_t key1,
               const uint32_t *pixel, const uint32_t *sample, const uint32_t *block, float *uniforms, uint32_t *bits) {
    k_rng<<<gridFor(n, shape.numSMs, 8), kBlockThreads, 0, s>>>(n, rounds, key0, key1, makePhiloxKeys(key0, key1), pixel,
                                                                sample, block, uniforms, bits);
}

void launchSelftestArith(cudaStream_t s, const LaunchShape &shape, int mode, unsigned long long n, uint32_t seed,
                          unsigned long long *mismatches) {
    k_selftest_arith<<<shape.numSMs * 8, kBlockThreads, 0, s>>>(mode, n, seed, mismatches);
}

void launchPack4(cudaStream_t s, const LaunchShape &shape, size_t n, const float *xyz, float4 *out) {
    k_pack4<<<gridFor(n, shape.numSMs, 8), kBlockThreads, 0, s>>>(n, xyz, out);
}

void launchUnpackHits(cudaStream_t s, const LaunchShape &shape, size_t n, const HitRecord *hits, float *t,
                      int32_t *prim) {
    k_unpack_hits<<<gridFor(n, shape.numSMs, 8), kBlockThreads, 0, s>>>(n, hits, t, prim);
}

cudaError_t configureKernels(LaunchShape &shape) {
    cudaError_t e;
    // Every kernel that stages the scene opts in to the device's largest dynamic shared-memory size once.  The attribute
    // is per function and per device, not per scene: sizing it from one scene's tables would lower the limit under
    // another live scene whose tables are larger.
    if (shape.smemOptin > 48 * 1024) {
        int const most = static_cast<int>(shape.smemOptin);
        const void *staging[] = {reinterpret_cast<const void *>(k_intersect<false>),
                                 reinterpret_cast<const void *>(k_intersect<true>),
                                 reinterpret_cast<const void *>(k_walk),
                                 reinterpret_cast<const void *>(k_walk_shared),
                                 reinterpret_cast<const void *>(k_shade<false>),
                                 reinterpret_cast<const void *>(k_shade<true>),
                                 reinterpret_cast<const void *>(k_intersect_batch<false>),
                                 reinterpret_cast<const void *>(k_intersect_batch<true>),
                                 reinterpret_cast<const void *>(k_hit_surface<false>),
                                 reinterpret_cast<const void *>(k_hit_surface<true>)};
        for (const void *kernel : staging) { // the opt-in limit covers static + dynamic shared memory
            cudaFuncAttributes attr{};
            if ((e = cudaFuncGetAttributes(&attr, kernel)) != cudaSuccess)
                return e;
            if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          most - static_cast<int>(attr.sharedSizeBytes))) != cudaSuccess)
                return e;
        }
    }
    if (const char *env = std::getenv("CORNELIS_BATCH_PACKED"))
        shape.batchPacked = std::atoi(env) != 0;
    auto resident = [&](auto kernel, size_t smem, int &grid) -> cudaError_t {
        int blocks = 0;
        cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, kBlockThreads, smem);
        if (err != cudaSuccess)
            return err;
        if (const char *env = std::getenv("CORNELIS_BLOCKS_PER_SM"))
            if (std::atoi(env) > 0)
                blocks = std::atoi(env);
        grid = shape.numSMs * (blocks > 0 ? blocks : 1);
        return cudaSuccess;
    };
    if ((e = resident(k_raygen, 0, shape.gridRaygen)) != cudaSuccess)
        return e;
    if ((e = resident(k_intersect<false>, shape.sceneSmemBytes, shape.gridIntersect)) != cudaSuccess)
        return e;
    if ((e = resident(k_intersect<true>, shape.sceneSmemBytes, shape.gridIntersectGrid)) != cudaSuccess)
        return e;
    if ((e = resident(k_walk, ((shape.sceneSmemBytes + 15u) & ~static_cast<size_t>(15u)) + walkQueueBytes(kBlockThreads),
                      shape.gridWalk)) != cudaSuccess)
        return e;
    if (const char *env = std::getenv("CORNELIS_WALK_PULL"))
        shape.walkPull = std::atoi(env) != 0;
    if (const char *env = std::getenv("CORNELIS_WALK_SHARED_RANGES"))
        shape.walkRangesInShared = std::atoi(env) != 0;
    if ((e = resident(k_shade<false>, shape.sceneSmemBytes, shape.gridShade)) != cudaSuccess)
        return e;
    return resident(k_accumulate, 0, shape.gridAccumulate);
}

} // namespace cornelis_b200
