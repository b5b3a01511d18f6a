// TEST INFRASTRUCTURE — not part of the product.  Compiles the grid builder (csrc/scene_tables.h) and the grid
// traversal (csrc/geometry.cuh closestHitGrid, a __host__ __device__ function) for the HOST so that the CPU test
// suite (`-m "not gpu"`) can check, without a GPU, that walking the grid returns exactly what the exhaustive
// reference scan returns.  Built by tests/conftest.py into tests/native/_build/ (git-ignored); nothing under
// cornelis_b200/ or include/ references it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../cornelis_b200/csrc/geometry.cuh"
#include "../../cornelis_b200/csrc/scene_tables.h"

using namespace cornelis_b200;

extern "C" {

// Returns 0 on success, 1 if the grid could not be built.  info = {nx, ny, nz, references, margin bits, cells}.
// walk (may be null) receives per ray {cells visited, sphere tests}.
int grid_host_intersect(const cornelis_camera_desc *camera, const cornelis_sphere_desc *spheres, size_t nSpheres,
                        const cornelis_plane_desc *planes, size_t nPlanes, size_t nRays, const float *org,
                        const float *dir, float *t, int32_t *prim, uint32_t *walk, uint64_t *info) {
    std::vector<DevSphere> hs(nSpheres);
    for (size_t i = 0; i < nSpheres; i++)
        hs[i] = DevSphere{spheres[i].center[0], spheres[i].center[1], spheres[i].center[2],
                          spheres[i].radius * spheres[i].radius};
    std::vector<DevPlane> hp(nPlanes);
    for (size_t i = 0; i < nPlanes; i++)
        hp[i] = makeDevPlane(planes[i]);
    double lo[3], hi[3];
    sceneOriginBox(*camera, spheres, nSpheres, hp.data(), nPlanes, lo, hi);
    HostGrid grid;
    if (!buildGrid(spheres, nSpheres, lo, hi, grid))
        return 1;
    SceneView view{};
    view.spheres = hs.data();
    view.planes = hp.data();
    view.nSpheres = static_cast<uint32_t>(nSpheres);
    view.nPlanes = static_cast<uint32_t>(nPlanes);
    view.grid = grid.g;
    view.grid.cellStart = grid.cellStart.data();
    view.grid.cellItems = grid.cellItems.data();
    if (info) {
        info[0] = grid.g.nx, info[1] = grid.g.ny, info[2] = grid.g.nz;
        info[3] = grid.cellStart.back();
        uint32_t bits;
        std::memcpy(&bits, &grid.g.margin, 4);
        info[4] = bits;
        info[5] = grid.cellStart.size() - 1;
    }
    for (size_t k = 0; k < nRays; k++) {
        float tb = INFINITY;
        int32_t pb = -1;
        uint32_t w[2] = {0, 0};
        closestHitGrid(true, V3{org[3 * k], org[3 * k + 1], org[3 * k + 2]}, V3{dir[3 * k], dir[3 * k + 1], dir[3 * k + 2]},
                       view, hp.data(), tb, pb, w);
        t[k] = tb;
        prim[k] = pb;
        if (walk)
            walk[2 * k] = w[0], walk[2 * k + 1] = w[1];
    }
    return 0;
}

// Structural check of the grid: every sphere must be listed, in ascending order, in every cell its (unpadded)
// bounding box touches.  Returns the number of violations.
uint64_t grid_host_check_structure(const cornelis_camera_desc *camera, const cornelis_sphere_desc *spheres,
                                   size_t nSpheres, const cornelis_plane_desc *planes, size_t nPlanes) {
    std::vector<DevPlane> hp(nPlanes);
    for (size_t i = 0; i < nPlanes; i++)
        hp[i] = makeDevPlane(planes[i]);
    double lo[3], hi[3];
    sceneOriginBox(*camera, spheres, nSpheres, hp.data(), nPlanes, lo, hi);
    HostGrid grid;
    if (!buildGrid(spheres, nSpheres, lo, hi, grid))
        return ~0ull;
    DevGrid const &g = grid.g;
    uint64_t bad = 0;
    size_t const ncell = static_cast<size_t>(g.nx) * g.ny * g.nz;
    for (size_t c = 0; c < ncell; c++)
        for (uint32_t k = grid.cellStart[c] + 1; k < grid.cellStart[c + 1]; k++)
            bad += grid.cellItems[k - 1] >= grid.cellItems[k];
    double const gmin[3] = {g.minx, g.miny, g.minz}, cell[3] = {g.cellx, g.celly, g.cellz};
    uint32_t const dim[3] = {g.nx, g.ny, g.nz};
    for (size_t i = 0; i < nSpheres; i++) {
        int64_t f[3], l[3];
        for (int a = 0; a < 3; a++) {
            double const r = std::fabs(static_cast<double>(spheres[i].radius));
            f[a] = static_cast<int64_t>(std::floor((spheres[i].center[a] - r - gmin[a]) / cell[a]));
            l[a] = static_cast<int64_t>(std::floor((spheres[i].center[a] + r - gmin[a]) / cell[a]));
            f[a] = std::max<int64_t>(0, std::min<int64_t>(dim[a] - 1, f[a]));
            l[a] = std::max<int64_t>(0, std::min<int64_t>(dim[a] - 1, l[a]));
        }
        for (int64_t z = f[2]; z <= l[2]; z++)
            for (int64_t y = f[1]; y <= l[1]; y++)
                for (int64_t x = f[0]; x <= l[0]; x++) {
                    size_t const c = (static_cast<size_t>(z) * g.ny + y) * g.nx + x;
                    bool found = false;
                    for (uint32_t k = grid.cellStart[c]; k < grid.cellStart[c + 1] && !found; k++)
                        found = grid.cellItems[k] == i;
                    bad += !found;
                }
    }
    return bad;
}

} // extern "C"
