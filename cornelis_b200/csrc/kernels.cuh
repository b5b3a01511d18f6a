// sm_100a kernels of the wavefront render loop.
//
//   plan        1 thread: sizes the pass (how many camera paths to regenerate), resets the queue tails
//   raygen      generateCameraRays            (reference src/Render.cpp:85-100)            HBM-write bound
//   intersect   intersect + compaction #1     (Render.cpp:110-150, Geometry.cpp:34-178)    FP32-pipe bound
//   shade       accumulateAndBounce + compaction #2 (Render.cpp:167-218, Materials.*)      FP32/SFU-pipe bound
//   accumulate  per-pixel sum of finished paths (Render.cpp:245-248)                       HBM (atomics) bound
//   resolve     color * (1/spp), optional display transform (Render.cpp:250, 257-261)      HBM bound
//
// A "pass" runs all live paths one bounce forward.  Paths that end (miss, Russian roulette, depth cap) are replaced
// by fresh camera paths at the start of the next pass, so every pass works on a full pool until the sample budget
// is exhausted; the reference instead drains one pixel's samples at a time (Render.cpp:232-243).
//
// All kernels are persistent grid-stride loops (grid = a multiple of the SM count) that read their trip counts from
// the device-resident Control block, so the host never synchronises inside a batch of passes.
#pragma once

#include "device_types.h"
#include "geometry.cuh"
#include "geometry2.cuh"
#include "materials.cuh"
#include "math.cuh"
#include "rng.cuh"

namespace cornelis_b200 {

constexpr int kBlockThreads = 256;
constexpr int kWarpsPerBlock = kBlockThreads / 32;

// --------------------------------------------------------------------------------------------- shared scene --

// Cooperative 16-byte copies global -> shared; every table is a multiple of 16 bytes except the two index lists.
// kGrid is a compile-time parameter so that, without the grid, every pointer provably addresses shared memory
// (LDS.128 instead of generic loads).
template <bool kGrid>
__device__ __forceinline__ SharedScene stageScene(const SceneView &scene, unsigned char *smem, bool wantMaterials) {
    float4 *dst = reinterpret_cast<float4 *>(smem);
    uint32_t const nS4 = kGrid ? 0u : scene.nSpheres;             // 1 float4 per sphere
    uint32_t const nP4 = scene.nPlanes * 4;                       // 4 float4 per plane
    uint32_t const nM4 = scene.nMaterials * 4;                    // 4 float4 per material
    uint32_t const nA4 = scene.planeEnd[2] * 2;                   // 2 float4 per axis-aligned plane
    uint32_t const nO4 = (scene.nPlanes + 3u) / 4u;               // plane order, padded to 16 bytes
    const float4 *srcS = reinterpret_cast<const float4 *>(scene.spheres);
    const float4 *srcP = reinterpret_cast<const float4 *>(scene.planes);
    const float4 *srcM = reinterpret_cast<const float4 *>(scene.materials);
    const float4 *srcA = reinterpret_cast<const float4 *>(scene.axisPlanes);
    for (uint32_t k = threadIdx.x; k < nS4; k += blockDim.x)
        dst[k] = srcS[k];
    for (uint32_t k = threadIdx.x; k < nP4; k += blockDim.x)
        dst[nS4 + k] = srcP[k];
    if (wantMaterials)
        for (uint32_t k = threadIdx.x; k < nM4; k += blockDim.x)
            dst[nS4 + nP4 + k] = srcM[k];
    for (uint32_t k = threadIdx.x; k < nA4; k += blockDim.x)
        dst[nS4 + nP4 + nM4 + k] = srcA[k];
    uint32_t *order = reinterpret_cast<uint32_t *>(dst + nS4 + nP4 + nM4 + nA4);
    for (uint32_t k = threadIdx.x; k < scene.nPlanes; k += blockDim.x)
        order[k] = scene.planeOrder[k];
    uint32_t *ids = reinterpret_cast<uint32_t *>(dst + nS4 + nP4 + nM4 + nA4 + nO4);
    if (wantMaterials && !kGrid)
        for (uint32_t k = threadIdx.x; k < scene.nSpheres; k += blockDim.x)
            ids[k] = scene.sphereMaterial[k];
    __syncthreads();
    SharedScene s;
    s.spheres = kGrid ? scene.spheres : reinterpret_cast<const DevSphere *>(dst);
    s.planes = reinterpret_cast<const DevPlane *>(dst + nS4);
    s.materials = reinterpret_cast<const DevMaterial *>(dst + nS4 + nP4);
    s.axisPlanes = dst + nS4 + nP4 + nM4;
    s.planeOrder = order;
    s.sphereMaterial = kGrid ? scene.sphereMaterial : ids;
    return s;
}

// The intersect stage for either scene representation.  kGrid is a template parameter of the kernels (the host picks
// the instantiation from SceneView::grid.enabled) so the few-primitive kernels keep their register budget.
template <bool kGrid, int kSphereUnroll = 1, int kPacked = kScanEither, bool kSettleZero = true>
__device__ __forceinline__ void closestHitScene(bool live, V3 o, V3 d, const SharedScene &sh, const SceneView &scene,
                                                float &tBest, int32_t &primBest, const float4 *pairs = nullptr,
                                                PackedConstants neutral = PackedConstants{1.0f, -0.0f}) {
    if (kGrid)
        closestHitGrid(live, o, d, scene, sh.planes, tBest, primBest);
    else
        closestHit<kSphereUnroll, kPacked, kSettleZero>(live, o, d, sh, scene, tBest, primBest, pairs, neutral);
}

// The sphere table of a staged scene once more, as pairs for scanSpheresPacked (geometry.cuh):
//     (cx_2p, cx_2p+1, cy_2p, cy_2p+1), (cz_2p, cz_2p+1, r2_2p, r2_2p+1)      for p < nSpheres / 2
__device__ __forceinline__ void stageSpherePairs(const SharedScene &sh, uint32_t nSpheres, float4 *pairs) {
    const float4 *spheres4 = reinterpret_cast<const float4 *>(sh.spheres);
    for (uint32_t p = threadIdx.x; p < nSpheres / 2u; p += blockDim.x) {
        float4 const a = spheres4[2u * p], b = spheres4[2u * p + 1u];
        pairs[2u * p] = make_float4(a.x, b.x, a.y, b.y);
        pairs[2u * p + 1u] = make_float4(a.z, b.z, a.w, b.w);
    }
    __syncthreads();
}

// The tables of closestHit2 (geometry2.cuh SplattedScene) built from a staged scene: every value twice.
__device__ __forceinline__ SplattedScene stageSplatted(const SharedScene &sh, const SceneView &scene, float4 *dst) {
    const float4 *spheres4 = reinterpret_cast<const float4 *>(sh.spheres);
    for (uint32_t k = threadIdx.x; k < scene.nSpheres; k += blockDim.x) {
        float4 const s = spheres4[k];
        dst[2u * k] = make_float4(s.x, s.x, s.y, s.y);
        dst[2u * k + 1u] = make_float4(s.z, s.z, s.w, s.w);
    }
    float4 *planes = dst + 2u * scene.nSpheres;
    for (uint32_t k = threadIdx.x; k < scene.planeEnd[2]; k += blockDim.x) {
        float4 const a = sh.axisPlanes[2u * k], b = sh.axisPlanes[2u * k + 1u]; // (p0k, p0T, p0B, w/2) (h/2, id, -, -)
        planes[3u * k] = make_float4(a.x, a.x, a.y, a.y);
        planes[3u * k + 1u] = make_float4(a.z, a.z, a.w, a.w);
        planes[3u * k + 2u] = make_float4(b.x, b.x, b.y, 0.0f);
    }
    __syncthreads();
    return SplattedScene{dst, planes};
}

// ------------------------------------------------------------------------------------------------ compaction --

// Active-path compaction (the reference's order-preserving list rebuilds, Render.cpp:142-149 and :215-217) as an
// unordered stream append: warp ballot + popc prefix inside each warp, a 8-entry shared scan across the block's
// warps, ONE atomicAdd per block per queue on the device-side tail.  Two queues are served per call so that the
// two __syncthreads are shared.  Order need not be preserved: a path's random numbers are keyed by
// (pixel, sample, depth), not by its position in a list.
struct AppendSlots {
    uint32_t a, b; // destination index in queue A / queue B (valid where the flag was set)
};

__device__ __forceinline__ AppendSlots blockAppend2(bool flagA, bool flagB, uint32_t *tailA, uint32_t *tailB,
                                                    uint32_t (*scratch)[kWarpsPerBlock + 1]) {
    unsigned const lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned const maskA = __ballot_sync(0xffffffffu, flagA);
    unsigned const maskB = __ballot_sync(0xffffffffu, flagB);
    unsigned const below = (1u << lane) - 1u;
    if (lane == 0) {
        scratch[0][warp] = __popc(maskA);
        scratch[1][warp] = __popc(maskB);
    }
    __syncthreads();
    if (threadIdx.x < 2) { // thread 0 scans queue A, thread 1 queue B
        uint32_t running = 0;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; w++) {
            uint32_t c = scratch[threadIdx.x][w];
            scratch[threadIdx.x][w] = running;
            running += c;
        }
        uint32_t base = 0;
        if (running)
            base = atomicAdd(threadIdx.x == 0 ? tailA : tailB, running);
        scratch[threadIdx.x][kWarpsPerBlock] = base;
    }
    __syncthreads();
    AppendSlots s;
    s.a = scratch[0][kWarpsPerBlock] + scratch[0][warp] + __popc(maskA & below);
    s.b = scratch[1][kWarpsPerBlock] + scratch[1][warp] + __popc(maskB & below);
    __syncthreads(); // scratch is reused by the next grid-stride step
    return s;
}

// ----------------------------------------------------------------------------------------------- path packing --

// Path length limit.  The depth travels in 8 bits next to the sample index and selects the Philox block of a bounce, so
// a path is ended when it reaches 255 bounces.  Finite paths never get there (Russian roulette lets a path survive a
// bounce with probability <= 0.5445 from depth 3 on, Render.cpp:153-165: 255 bounces have probability < 1e-66); the
// paths that do are those whose throughput became NaN (the reference's Oren-Nayar quirk, about one in 1e8 samples):
// `prob < u` is then false for ever (Render.cpp:189) and in a closed scene the reference's loop would never end.
constexpr uint32_t kDepthLimit = 255u;

__device__ __forceinline__ uint32_t packSampleDepth(uint32_t sample, uint32_t depth) {
    return (sample << 8) | (depth > 255u ? 255u : depth);
}

// -------------------------------------------------------------------------------------------------- raygen --

// PerspectiveCamera::operator() (Camera.cpp:11-13): (corner + x*u + y*v).normalize(), where nanovdb's normalize
// multiplies by the rounded reciprocal of the length without any small-length guard (NanoVDB.h:919-953).
CB_HD V3 cameraDirection(const DevCamera &c, float x, float y) {
    V3 xu{x * c.ux, x * c.uy, x * c.uz};
    V3 yv{y * c.vx, y * c.vy, y * c.vz};
    V3 d = (V3{c.cx, c.cy, c.cz} + xu) + yv;
    float const len2 = d.x * d.x + d.y * d.y + d.z * d.z;
    float s;
#ifdef __CUDA_ARCH__
    if (inFastNormalizeRange(len2)) {
        float len;
        sqrtAndReciprocalExactFast(len2, len, s); // == 1.0f / sqrtf(len2), bit for bit (exact_arith.cuh)
    } else
#endif
        s = 1.0f / sqrtf(len2);
    return V3{d.x * s, d.y * s, d.z * s};
}

// NormalizedFrameBufferCoord + the jittered film position of Render.cpp:29-37, 96.
// Camera paths are numbered sample by sample and, within a sample, by POSITION: position p of a frame is pixel p in row
// order, or — when the frame divides into tiles of 32 pixels (RenderConfig::tilesPerRow != 0) — pixel (p & 31) of tile
// p / 32, tiles in row order.  32 consecutive paths, one warp's worth, then cover a compact tile instead of a strip of a
// row: their rays cross the same cells of the grid and meet the same primitives.  A bijection of the frame either way,
// so every (pixel, sample) pair is still generated exactly once, with the random numbers of its (pixel, sample) key.
__device__ __forceinline__ uint32_t pixelOfPosition(uint32_t position, const RenderConfig &cfg, uint32_t &i, uint32_t &j) {
    if (cfg.tilesPerRow) {
        uint32_t const tile = position >> 5, within = position & 31u;
        uint32_t const tileRow = fastDivide(tile, cfg.byTilesPerRow), tileCol = tile - tileRow * cfg.tilesPerRow;
        i = tileCol * kTileWidth + (within & (kTileWidth - 1u));
        j = tileRow * kTileHeight + (within >> CORNELIS_RAYGEN_TILE_SHIFT);
        return j * cfg.width + i;
    }
    j = fastDivide(position, cfg.byWidth), i = position - j * cfg.width;
    return position;
}

CB_HD V3 pixelRayDirection(const DevCamera &c, uint32_t i, uint32_t j, float dx, float dy, float phi1, float phi2) {
    float x = static_cast<float>(i) * dx;
    float y = static_cast<float>(j) * dy;
    return cameraDirection(c, x + phi1 * dx, y + phi2 * dy);
}

// ----------------------------------------------------------------------------------------------- shade body --

// accumulateAndBounce for one ray (Render.cpp:173-216), in two halves so that the persistent pipeline can park the
// survivors of Russian roulette between them (persistent.cu).
//
// First half, Render.cpp:174-192: emission is added before the roulette test (Render.cpp:187), then the path survives
// with probability `prob`.  u0 = the RR draw.
__device__ __forceinline__ bool shadeRoulette(const DevMaterial &mat, uint32_t depth, float u0, RGBf thr, RGBf &rad,
                                              float &prob) {
    prob = russianRouletteFactor(thr.r, thr.g, thr.b, depth);              // Render.cpp:182
    rad.r += thr.r * mat.er;                                                // Render.cpp:187, :67-69
    rad.g += thr.g * mat.eg;
    rad.b += thr.b * mat.eb;
    return !(prob < u0);                                                    // Render.cpp:189
}

// Second half, Render.cpp:194-213: sample the layered BSDF at the hit, build the next ray and update the throughput.
// dir is the incoming ray's direction on entry and the sampled direction on exit.
// `odd`: see normalize (math.cuh) — the caller redoes the step without it when it comes back raised.
__device__ __forceinline__ void shadeScatter(const DevMaterial &mat, V3 P, V3 N, float prob, float x0, float x1, float x2,
                                             V3 &org, V3 &dir, RGBf &thr, OddWatch *odd = nullptr) {
    V3 const wOut = -dir;                                                   // Render.cpp:174
    Basis const basis = constructBasis(N, odd);                             // Render.cpp:194
    V3 wIn;
    float pdf;
    RGBf const f = layeredSample(mat, wOut, x0, x1, x2, basis, wIn, pdf, odd); // Render.cpp:200
    org = P + wIn * 0.0001f;                                                // Render.cpp:207
    dir = wIn;                                                              // Render.cpp:208
    float const c = fabsf(dot(wIn, N));
    float const denom = pdf * prob;
    // thr *= f * c / denom, Render.cpp:210-213 (RGB / float = three divisions, Color.cpp:11-17).  The throughput update
    // belongs to the tolerant tail: one MUFU reciprocal (<= 1 ulp) serves the three quotients.  denom =
    // pdf * prob >= 0.5 / (2 Pi) * 0.05 > 0 for finite inputs, zero numerators (black albedo: the light, gold's
    // diffuse part) give zero, NaN stays NaN.
    float const scale = c * approxRcp(denom);
    thr.r *= f.r * scale;
    thr.g *= f.g * scale;
    thr.b *= f.b * scale;
}

// Both halves back to back.  u = (RR draw, x0, x1, x2).
// Returns true if the path survives; org/dir/thr are then the next ray and the updated throughput.
__device__ __forceinline__ bool shadeBounce(const DevMaterial &mat, V3 P, V3 N, uint32_t depth, float u0, float x0,
                                            float x1, float x2, V3 &org, V3 &dir, RGBf &thr, RGBf &rad) {
    float prob;
    if (!shadeRoulette(mat, depth, u0, thr, rad, prob))
        return false;
    shadeScatter(mat, P, N, prob, x0, x1, x2, org, dir, thr);
    return true;
}

} // namespace cornelis_b200
