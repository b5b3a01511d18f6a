// Host-callable launchers for the kernels in wavefront.cu.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "device_types.h"

namespace cornelis_b200 {

struct LaunchShape {
    int numSMs = 148;         // B200: 148 SMs; queried at scene creation
    int blocksPerSM = 4;      // resident 256-thread CTAs per SM for the stage entry points
    // Persistent grids: numSMs x (CTAs of that kernel resident per SM, from the occupancy calculator), so every CTA
    // is resident for the whole launch and the grid-stride loops split the pool evenly.
    int gridRaygen = 592, gridIntersect = 592, gridIntersectGrid = 592, gridWalk = 592, gridShade = 592, gridAccumulate = 592;
    bool walkPull = true;     // grid scenes: k_walk (warps pull rays) + k_compact_hits instead of one ray per thread
    // k_walk_shared (1024-thread CTAs, cell ranges in shared memory): measured on config 4 and left off — 1439 against
    // 1488 Msamples/s (profiles/r2_walk); CORNELIS_WALK_SHARED_RANGES=1 switches it on
    bool walkRangesInShared = false;
    bool batchPacked = true;  // k_intersect_batch scans the spheres two at a time on packed FP32 (FFMA2)
    size_t sceneSmemBytes = 0;
    size_t smemOptin = 227 * 1024;        // largest dynamic shared memory a CTA may opt in to; queried at scene creation
    // persistent pipeline (persistent.cu): which variant runs, its dynamic shared memory and where its queues start
    bool persistentQueued = true;
    size_t persistentSmemBytes = 0;
    uint32_t persistentQueueOffset = 0;
    uint32_t persistentPairsOffset = 0; // 0: no paired sphere table (scalar scan)
};

// Opts the kernels in to the scene's shared-memory size and fills the persistent grid sizes.
cudaError_t configureKernels(LaunchShape &shape);

// persistent-thread pipeline (persistent.cu)
cudaError_t configurePersistent(LaunchShape &shape, bool gridScene, uint32_t nSpheres, int &grid);
uint32_t persistentClaim(unsigned long long paths, int grid); // RenderConfig::claim for a launch over `paths` camera paths
void launchPersistent(cudaStream_t s, const LaunchShape &shape, int grid, const RenderConfig &cfg, const SceneView &scene,
                      unsigned long long *cursor, unsigned long long limit, float4 *accum, float4 *accum2,
                      bool dropNonFinite, Control *ctl);

void launchPlan(cudaStream_t s, Control *ctl, const RenderConfig &cfg);
void launchRaygen(cudaStream_t s, const LaunchShape &shape, const Control *ctl, const RenderConfig &cfg,
                  const DevCamera &cam, const PathPool &pool);
void launchIntersect(cudaStream_t s, const LaunchShape &shape, Control *ctl, const SceneView &scene,
                     const PathPool &pool, HitRecord *hits, uint32_t *hitQueue, FinishedPath *finished);
void launchShade(cudaStream_t s, const LaunchShape &shape, Control *ctl, const RenderConfig &cfg,
                 const SceneView &scene, const PathPool &in, const PathPool &out, const HitRecord *hits,
                 const uint32_t *hitQueue, FinishedPath *finished);
void launchAccumulate(cudaStream_t s, const LaunchShape &shape, Control *ctl, const FinishedPath *finished,
                      float4 *accum, float4 *accum2, bool dropNonFinite);
void launchResolve(cudaStream_t s, const LaunchShape &shape, uint32_t npixels, uint32_t samples, const float4 *accum,
                   const float4 *accum2, float *rgb, float *variance);
void launchResolveSrgb8(cudaStream_t s, const LaunchShape &shape, uint32_t npixels, uint32_t samples,
                        const float4 *accum, uint8_t *rgb8);

void launchSrgb8Sweep(cudaStream_t s, const LaunchShape &shape, uint32_t first, size_t n, uint8_t *out);
void launchAddImages(cudaStream_t s, const LaunchShape &shape, size_t n4, float4 *dst, const float4 *src);

void launchPixelRays(cudaStream_t s, const LaunchShape &shape, const DevCamera &cam, uint32_t n, float dx, float dy,
                     const int32_t *pi, const int32_t *pj, const float *phi1, const float *phi2, float *org,
                     float *dir);
void launchIntersectBatch(cudaStream_t s, const LaunchShape &shape, const SceneView &scene, size_t n,
                          const float4 *org, const float4 *dir, HitRecord *hits);
void launchHitSurface(cudaStream_t s, const LaunchShape &shape, const SceneView &scene, size_t n, const float4 *org,
                      const float4 *dir, const HitRecord *hits, float *P, float *N, int32_t *mat);
void launchBsdfSample(cudaStream_t s, const LaunchShape &shape, const DevMaterial *materials, uint32_t n,
                      const int32_t *mat, const float *wo, const float *N, const float *x, float *wi, float *pdf,
                      float *f);
void launchBsdfEval(cudaStream_t s, const LaunchShape &shape, const DevMaterial *materials, uint32_t n,
                    const int32_t *mat, const float *wi, const float *wo, const float *N, float *f, float *pdf);
void launchShadeExplicit(cudaStream_t s, const LaunchShape &shape, const DevMaterial *materials, uint32_t n,
                         uint32_t depth, const float *u, const float *P, const float *N, const int32_t *mat,
                         float *org, float *dir, float *thr, float *rad, uint8_t *alive);
void launchRng(cudaStream_t s, const LaunchShape &shape, uint32_t n, int rounds, uint32_t key0, uint32_t key1,
               const uint32_t *pixel, const uint32_t *sample, const uint32_t *block, float *uniforms, uint32_t *bits);
void launchSelftestArith(cudaStream_t s, const LaunchShape &shape, int mode, unsigned long long n, uint32_t seed,
                          unsigned long long *mismatches);
void launchPack4(cudaStream_t s, const LaunchShape &shape, size_t n, const float *xyz, float4 *out);
void launchUnpackHits(cudaStream_t s, const LaunchShape &shape, size_t n, const HitRecord *hits, float *t,
                      int32_t *prim);

} // namespace cornelis_b200
