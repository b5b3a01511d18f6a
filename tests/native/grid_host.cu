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
    view.grid.cellRange = grid.cellRange.data();
    view.grid.cellSpheres = grid.cellSpheres.data();
    view.grid.cellIds = grid.cellIds.data();
    if (info) {
        info[0] = grid.g.nx, info[1] = grid.g.ny, info[2] = grid.g.nz;
        info[3] = grid.cellIds.size();
        uint32_t bits;
        std::memcpy(&bits, &grid.g.margin, 4);
        info[4] = bits;
        info[5] = grid.cellRange.size();
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

// Structural check of the grid: every sphere must be listed, in ascending order and with a faithful copy, in every
// cell that the sphere itself touches.  Returns the number of violations.
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
    bad += grid.cellRange.size() != ncell;
    for (size_t c = 0; c < ncell; c++) {
        bad += grid.cellRange[c].x > grid.cellRange[c].y || grid.cellRange[c].y > grid.cellIds.size();
        bad += c > 0 && grid.cellRange[c].x != grid.cellRange[c - 1].y;
        for (uint32_t k = grid.cellRange[c].x; k < grid.cellRange[c].y; k++) {
            uint32_t const i = grid.cellIds[k];
            bad += k > grid.cellRange[c].x && grid.cellIds[k - 1] >= i;
            float4 const s = grid.cellSpheres[k];
            bad += !(s.x == spheres[i].center[0] && s.y == spheres[i].center[1] && s.z == spheres[i].center[2] &&
                     s.w == spheres[i].radius * spheres[i].radius);
        }
    }
    double const gmin[3] = {g.minx, g.miny, g.minz}, cell[3] = {g.cellx, g.celly, g.cellz};
    uint32_t const dim[3] = {g.nx, g.ny, g.nz};
    for (size_t i = 0; i < nSpheres; i++) {
        int64_t f[3], l[3];
        double const r = std::fabs(static_cast<double>(spheres[i].radius));
        for (int a = 0; a < 3; a++) {
            f[a] = static_cast<int64_t>(std::floor((spheres[i].center[a] - r - gmin[a]) / cell[a]));
            l[a] = static_cast<int64_t>(std::floor((spheres[i].center[a] + r - gmin[a]) / cell[a]));
            f[a] = std::max<int64_t>(0, std::min<int64_t>(dim[a] - 1, f[a]));
            l[a] = std::max<int64_t>(0, std::min<int64_t>(dim[a] - 1, l[a]));
        }
        for (int64_t z = f[2]; z <= l[2]; z++)
            for (int64_t y = f[1]; y <= l[1]; y++)
                for (int64_t x = f[0]; x <= l[0]; x++) {
                    int64_t const c3[3] = {x, y, z};
                    double dist2 = 0;
                    for (int a = 0; a < 3; a++) {
                        double const lo_ = gmin[a] + c3[a] * cell[a], hi_ = lo_ + cell[a], p = spheres[i].center[a];
                        double const dd = p < lo_ ? lo_ - p : p > hi_ ? p - hi_ : 0.0;
                        dist2 += dd * dd;
                    }
                    if (dist2 > r * r)
                        continue; // the sphere does not reach this cell of its bounding box
                    size_t const c = (static_cast<size_t>(z) * g.ny + y) * g.nx + x;
                    bool found = false;
                    for (uint32_t k = grid.cellRange[c].x; k < grid.cellRange[c].y && !found; k++)
                        found = grid.cellIds[k] == i;
                    bad += !found;
                }
    }
    return bad;
}

} // extern "C"
