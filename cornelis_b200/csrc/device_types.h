// Device-side data layout of the render path (shared by host upload code and kernels).
//
// Everything a kernel reads per ray lives in float4 arrays so that a warp's accesses are 16-byte vector loads of
// consecutive addresses (the reference's SoA vectors, include/cornelis/SoA.hpp:144-176 and src/Render.cpp:47-61,
// regrouped four floats at a time).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "fastdiv.h"
#include "rng.cuh"

namespace cornelis_b200 {

// Scene tables.  Built on the host with the reference's own arithmetic (api.cu) and staged into shared memory by
// every kernel that intersects or shades.
struct __align__(16) DevSphere {   // float4
    float cx, cy, cz; // center (Scene.cpp:12-14)
    float r2;         // sphereRadius * sphereRadius, the only form Geometry.cpp:81 uses
};

struct __align__(16) DevPlane {    // 4 x float4
    float px, py, pz, width;   // point on the plane; extents[0] (Scene.cpp:28-33)
    float nx, ny, nz, height;  // normal;             extents[1]
    float tx, ty, tz; uint32_t material; // constructBasis(normal).T (Math.hpp:424-434), hoisted out of Geometry.cpp:165
    float bx, by, bz; uint32_t pad;      // constructBasis(normal).B; pad = axis class: 0/1/2 normal = +-x/y/z, 3 general
};

// Axis-aligned planes once more, as closestHit's fast path reads them (geometry.cuh axisPlaneTest): per plane two float4,
// class by class (normal along x, then y, then z; by index within a class):
//     (p0_k, p0_T, p0_B, width / 2)   (height / 2, primitive id as int bits, 0, 0)
// with k the normal's axis and T, B the in-plane axes constructBasis gives that class.
struct __align__(16) DevAxisPlane {
    float p0k, p0T, p0B, halfWidth;
    float halfHeight;
    int32_t id;
    uint32_t pad0, pad1;
};

struct __align__(16) DevMaterial { // 4 x float4 — StandardMaterial (Materials.hpp:325-338) with its constants folded
    float er, eg, eb, on_a;        // emission; OrenNayarBRDF::a_ (Materials.hpp:208)
    float dr, dg, db, on_b;        // albedo / Pi (Materials.hpp:226, Color.cpp:11-17); OrenNayarBRDF::b_
    float tr, tg, tb, alpha;       // glossy tint; GlossyBRDF::alpha_ = roughness^2 (Materials.hpp:296-299)
    float ior, r0, gtr_a, alpha2;  // refidx; Schlick R0 for (1, ior) (Materials.cpp:39-40); alpha2/(2 Pi); alpha^2
};

struct DevCamera {   // PerspectiveCamera (Camera.hpp:57-60), produced by lookAt (Camera.cpp:15-34)
    float ex, ey, ez, pad0;
    float cx, cy, cz, pad1; // corner_
    float ux, uy, uz, pad2;
    float vx, vy, vz, pad3;
};

// 1: the wavefront pipeline generates camera paths in 8x4-pixel tiles (RenderConfig::tilesPerRow); 0: in row order.
#ifndef CORNELIS_RAYGEN_TILES
#define CORNELIS_RAYGEN_TILES 1
#endif
#ifndef CORNELIS_RAYGEN_TILE_SHIFT
#define CORNELIS_RAYGEN_TILE_SHIFT 3 // log2 of the tile width: 8 x 4 pixels
#endif
constexpr unsigned kTileWidth = 1u << CORNELIS_RAYGEN_TILE_SHIFT, kTileHeight = 32u / kTileWidth;

// 1: the walk remembers the last TWO spheres it tested (0: the last one): a sphere is listed in ~4 cells, and with two
// spheres sharing consecutive cells a one-entry mailbox forgets the first while testing the second.  Config 4: 14 % fewer
// sphere tests (14.1 -> 12.0 per camera ray, tests/grid_walk_stats.py); three entries: 22 % fewer, but no faster.
#ifndef CORNELIS_GRID_MAILBOX2
#define CORNELIS_GRID_MAILBOX2 1
#endif

// 1: the walk's termination slack follows the ray's best hit so far (geometry.cuh); 0: one worst-case slack for the
// whole trusted region (round 1).
#ifndef CORNELIS_GRID_RAY_MARGIN
#define CORNELIS_GRID_RAY_MARGIN 1
#endif

// Uniform grid over the spheres.  Every sphere is listed, by ascending index, in every cell
// its padded bounding box overlaps; closestHitGrid walks the cells a ray crosses front to back and runs the SAME
// per-sphere test on the listed spheres, so hit ids and t are those of the exhaustive scan (see geometry.cuh for the
// argument and api.cu buildGrid for the padding).
struct DevGrid {
    float minx, miny, minz;        // lower corner of the grid
    float maxx, maxy, maxz;        // upper corner
    float cellx, celly, cellz;     // cell edge lengths
    float invx, invy, invz;        // their reciprocals
    float rminx, rminy, rminz;     // "trusted region": ray origins inside it are covered by the error bound the
    float rmaxx, rmaxy, rmaxz;     //   padding was derived from; other rays take the exhaustive scan
    float margin;                  // slack (a distance) on the front-to-back termination test: the part that does not
                                   //   depend on the hit so far (geometry.cuh: per-ray margin)
    float marginScale;             // 1 + slope: the walk stops when t_best * marginScale + margin / |d| < t(cell exit)
    uint32_t nx, ny, nz;           // resolution
    uint32_t enabled;              // 0: scan all spheres from shared memory
    const uint2 *cellRange;        // per cell: [first, last) into the two reference arrays
    const float4 *cellSpheres;     // per reference: a copy of the sphere (c.xyz, r^2), so a test costs one load
    const uint32_t *cellIds;       // per reference: the sphere index, ascending within a cell
};

struct SceneView {
    const DevSphere *spheres;
    const uint32_t *sphereMaterial;
    const DevPlane *planes;
    const DevMaterial *materials;
    uint32_t nSpheres, nPlanes, nMaterials;
    uint32_t radiiSafe; // every r^2 >= 2^-50: C - r^2 is then 0 or at least 2^-75 in magnitude (geometry.cuh scanSpheres)
    // Plane indices sorted by axis class (then by index): [0, planeEnd[0]) have normals along x, [planeEnd[0],
    // planeEnd[1]) along y, [planeEnd[1], planeEnd[2]) along z, the rest are general.  closestHit runs one tight loop
    // per class instead of dispatching on the class of every plane.
    const uint32_t *planeOrder;
    const DevAxisPlane *axisPlanes; // planeEnd[2] records
    uint32_t planeEnd[3];
    uint32_t pad2;
    DevCamera camera;
    DevGrid grid;
};

// Scene tables staged in dynamic shared memory: [spheres][planes][materials][axis planes][plane order][sphere
// material ids] (kernels.cuh stageScene).  With the grid the sphere tables stay in global memory and the pointers say so.
struct SharedScene {
    const DevSphere *spheres;
    const DevPlane *planes;
    const DevMaterial *materials;
    const float4 *axisPlanes; // DevAxisPlane records as float4 pairs
    const uint32_t *planeOrder;
    const uint32_t *sphereMaterial;
};

// With the grid enabled the spheres (and their material ids) stay in global memory — a ray touches a few dozen of
// them through the read-only cache — and only planes, materials and the plane order are staged.
__host__ __device__ inline size_t sharedSceneBytes(uint32_t nSpheres, uint32_t nPlanes, uint32_t nAxisPlanes,
                                                   uint32_t nMaterials, bool spheresInShared) {
    size_t const s = spheresInShared ? nSpheres : 0u;
    size_t const order = (sizeof(uint32_t) * nPlanes + 15u) & ~static_cast<size_t>(15u);
    return sizeof(DevSphere) * s + sizeof(DevPlane) * nPlanes + sizeof(DevMaterial) * nMaterials +
           sizeof(DevAxisPlane) * nAxisPlanes + order + sizeof(uint32_t) * s;
}

// Path pool: four float4 arrays — 64 B per path.
//   org  = ray origin xyz | unused
//   dir  = ray direction xyz | unused
//   thr  = path throughput rgb | pixel index (uint bits)         (PathThroughputTag, Render.cpp:39-41)
//   rad  = radiance collected so far rgb | sample << 8 | depth   (LightInTag, Render.cpp:43-45)
struct PathPool {
    float4 *org, *dir, *thr, *rad;
};

// Hit record: t and the primitive that produced it (sphere index, or nSpheres + plane index, or -1).
// P, N and the material id are re-derived in the shade kernel with the reference's expressions.
struct __align__(8) HitRecord {
    float t;
    int32_t prim;
};

// Terminated path with non-zero radiance, waiting for the accumulate kernel.
struct __align__(16) FinishedPath {
    float r, g, b;
    uint32_t pixel;
};

// Device-resident bookkeeping of the wavefront; one instance per scene handle.
struct Control {
    uint32_t nIn;       // rays in the current pool (survivors of the last pass + regenerated camera paths)
    uint32_t nSurvive;  // survivors appended to the next pool by shade
    uint32_t nHit;      // entries in the hit queue (compaction #1)
    uint32_t nFinished; // entries in the finished queue
    uint32_t genBase;   // raygen plan: first free slot
    uint32_t genCount;  //              number of camera paths to start this pass
    uint32_t maxDepth;  // deepest bounce seen
    uint32_t pad;
    unsigned long long cursor;     // next camera path (0 .. total)
    unsigned long long total;      // npixels * sample_count
    unsigned long long genFirst;   // raygen plan: first camera path index of this pass
    unsigned long long rays;       // rays intersected
    unsigned long long shaded;     // hits shaded
    unsigned long long iterations; // passes
    unsigned long long contributions; // finished paths added to the framebuffer
    unsigned long long walkCursor;    // grid scenes: next pooled ray to be claimed by k_walk (reset by k_plan)
    unsigned long long walkCursorCamera; // ... and the next of this pass's NEW camera rays [genBase, nIn) (reset by k_plan)
};

struct RenderConfig {
    uint32_t width, height, npixels;
    uint32_t firstSample;
    uint32_t maxDepth;  // 0 = unlimited
    uint32_t poolPaths;
    uint32_t variance;  // accumulate second moments
    uint32_t claim;     // persistent pipeline: camera paths a warp claims per atomic (a multiple of 32)
    uint32_t key0, key1; // Philox key = seed
    float dx, dy;       // 1.0f / width, 1.0f / height (Render.cpp:31)
    FastDiv byWidth;    // pixel -> row without a software divide
    // Wavefront pipeline: camera paths are generated in 8x4-pixel tiles (32 consecutive paths = one tile) when the frame
    // divides into them, so that a warp's camera rays cross the same cells of the grid (k_raygen); 0: row order.
    uint32_t tilesPerRow;
    FastDiv byTilesPerRow;
    PhiloxKeys keys;    // the ten Philox round keys of (key0, key1)
};

} // namespace cornelis_b200
