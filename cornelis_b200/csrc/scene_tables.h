// Host-side construction of the device scene tables from the C-ABI descriptions: plane records (with the reference's
// constructBasis hoisted out of Geometry.cpp:165) and the uniform grid over the spheres.  Header-only so that the
// C-ABI library (api.cu) and the CPU test helper (tests/native/grid_host.cu) build the tables with the same code.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../include/cornelis_cuda.h"
#include "device_types.h"
#include "math.cuh"

namespace cornelis_b200 {

inline DevPlane makeDevPlane(const cornelis_plane_desc &d) {
    Basis const b = constructBasis(V3{d.normal[0], d.normal[1], d.normal[2]}); // Geometry.cpp:165, hoisted
    DevPlane p{};
    p.px = d.point[0], p.py = d.point[1], p.pz = d.point[2], p.width = d.extents[0];
    p.nx = d.normal[0], p.ny = d.normal[1], p.nz = d.normal[2], p.height = d.extents[1];
    p.tx = b.T.x, p.ty = b.T.y, p.tz = b.T.z;
    p.material = d.material >= 0 ? static_cast<uint32_t>(d.material) : 0u;
    p.bx = b.B.x, p.by = b.B.y, p.bz = b.B.z;
    // Axis class for the exact fast path of closestHit: the normal is +-e_k and constructBasis produced the
    // in-plane axes closestHit assumes (x: T=z, B=y; y: T=x, B=z; z: T=x, B=y), all with unit magnitude.
    auto unitAxis = [](float x, float y, float z) -> int {
        if (fabsf(x) == 1.0f && y == 0.0f && z == 0.0f) return 0;
        if (x == 0.0f && fabsf(y) == 1.0f && z == 0.0f) return 1;
        if (x == 0.0f && y == 0.0f && fabsf(z) == 1.0f) return 2;
        return 3;
    };
    int const kN = unitAxis(p.nx, p.ny, p.nz), kT = unitAxis(p.tx, p.ty, p.tz), kB = unitAxis(p.bx, p.by, p.bz);
    bool const expected = (kN == 0 && kT == 2 && kB == 1) || (kN == 1 && kT == 0 && kB == 2) ||
                          (kN == 2 && kT == 0 && kB == 1);
    // (the fast path also wants p0's components to be 0 or >= 2^-56: geometry.cuh differenceSafe)
    auto safe = [](float x) { return !(fabsf(x) < 0x1.0p-56f) || x == 0.0f; };
    bool const finite = std::isfinite(p.px) && std::isfinite(p.py) && std::isfinite(p.pz) &&
                        fabsf(p.px) <= 0x1.0p30f && fabsf(p.py) <= 0x1.0p30f && fabsf(p.pz) <= 0x1.0p30f;
    // ... and compares |e| against HALF the extents (axisPlaneTest): the halving must be exact, i.e. an extent is 0, NaN,
    // infinite or at least 2^-100 in magnitude
    auto halves = [](float w) { return w == 0.0f || !(fabsf(w) < 0x1.0p-100f); };
    p.pad = expected && finite && safe(p.px) && safe(p.py) && safe(p.pz) && halves(p.width) && halves(p.height)
                ? static_cast<uint32_t>(kN)
                : 3u;
    return p;
}

// The compact record of an axis-aligned plane (device_types.h DevAxisPlane); `index` is the plane's position in the
// scene's plane list, its primitive id is nSpheres + index.
inline DevAxisPlane makeDevAxisPlane(const DevPlane &p, uint32_t nSpheres, uint32_t index) {
    float const point[3] = {p.px, p.py, p.pz};
    int const k = static_cast<int>(p.pad);          // 0, 1, 2
    int const axisT = k == 0 ? 2 : 0, axisB = k == 1 ? 2 : 1; // constructBasis: x -> (T z, B y); y -> (T x, B z); z -> (T x, B y)
    DevAxisPlane a{};
    a.p0k = point[k];
    a.p0T = point[axisT];
    a.p0B = point[axisB];
    a.halfWidth = p.width * 0.5f;
    a.halfHeight = p.height * 0.5f;
    a.id = static_cast<int32_t>(nSpheres + index);
    return a;
}

// Bounding box of everything a ray of the render loop can start from: the camera eye, the sphere boxes and the
// plane rectangles (bounce origins lie 1e-4 off a surface, Render.cpp:207).
inline void sceneOriginBox(const cornelis_camera_desc &camera, const cornelis_sphere_desc *spheres, size_t nSpheres,
                           const DevPlane *planes, size_t nPlanes, double boxMin[3], double boxMax[3]) {
    double lo[3] = {camera.origin[0], camera.origin[1], camera.origin[2]}, hi[3] = {lo[0], lo[1], lo[2]};
    auto grow = [&](double x, double y, double z) {
        double const p[3] = {x, y, z};
        for (int a = 0; a < 3; a++)
            lo[a] = std::min(lo[a], p[a]), hi[a] = std::max(hi[a], p[a]);
    };
    for (size_t i = 0; i < nSpheres; i++) {
        double const r = std::fabs(static_cast<double>(spheres[i].radius));
        grow(spheres[i].center[0] - r, spheres[i].center[1] - r, spheres[i].center[2] - r);
        grow(spheres[i].center[0] + r, spheres[i].center[1] + r, spheres[i].center[2] + r);
    }
    for (size_t i = 0; i < nPlanes; i++)
        for (int c = 0; c < 4; c++) {
            double const a = (c & 1 ? 0.5 : -0.5) * planes[i].width, b = (c & 2 ? 0.5 : -0.5) * planes[i].height;
            grow(planes[i].px + a * planes[i].tx + b * planes[i].bx, planes[i].py + a * planes[i].ty + b * planes[i].by,
                 planes[i].pz + a * planes[i].tz + b * planes[i].bz);
        }
    for (int a = 0; a < 3; a++)
        boxMin[a] = lo[a], boxMax[a] = hi[a];
}

// Uniform grid over the spheres (device_types.h DevGrid; the exactness argument is in geometry.cuh).  All sizing is
// done in double.  `box` is the bounding box of everything a ray of the render loop can start from: the sphere
// boxes, the plane rectangles and the camera eye.
struct HostGrid {
    DevGrid g{};
    std::vector<uint2> cellRange;      // per cell
    std::vector<float4> cellSpheres;   // per reference
    std::vector<uint32_t> cellIds;     // per reference
};

inline bool buildGrid(const cornelis_sphere_desc *spheres, size_t n, const double boxMin[3], const double boxMax[3],
               HostGrid &out) {
    if (n == 0)
        return false;
    // trusted region: the box inflated by 1e-3 of its diagonal on every side (bounce origins sit 1e-4 off a surface)
    double diag = 0;
    for (int a = 0; a < 3; a++)
        diag += (boxMax[a] - boxMin[a]) * (boxMax[a] - boxMin[a]);
    diag = std::sqrt(diag);
    if (!(diag > 0) || !std::isfinite(diag))
        return false;
    double rmin[3], rmax[3];
    for (int a = 0; a < 3; a++) {
        rmin[a] = boxMin[a] - 1e-3 * diag;
        rmax[a] = boxMax[a] + 1e-3 * diag;
    }
    double const D = diag * 1.0035; // >= diagonal of the trusted region: (1 + 2e-3 sqrt 3) diag
    double const eps = std::ldexp(D * D, -19);  // bound on the error of the computed discriminant (geometry.cuh)
    double const delta = 1e-4 * D;              // slack for the DDA's own rounding
    // padded sphere boxes and the grid bounds
    std::vector<double> lo(3 * n), hi(3 * n);
    double gmin[3] = {INFINITY, INFINITY, INFINITY}, gmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = 0; i < n; i++) {
        double const r = std::fabs(static_cast<double>(spheres[i].radius));
        double const reach = std::sqrt(r * r + eps) + delta;
        for (int a = 0; a < 3; a++) {
            lo[3 * i + a] = spheres[i].center[a] - reach;
            hi[3 * i + a] = spheres[i].center[a] + reach;
            gmin[a] = std::min(gmin[a], lo[3 * i + a]);
            gmax[a] = std::max(gmax[a], hi[3 * i + a]);
        }
    }
    double ext[3], volume = 1;
    for (int a = 0; a < 3; a++) {
        gmin[a] -= delta, gmax[a] += delta;
        if (!(std::fabs(gmin[a]) <= 0x1.0p30 && std::fabs(gmax[a]) <= 0x1.0p30 && std::fabs(rmin[a]) <= 0x1.0p30 &&
              std::fabs(rmax[a]) <= 0x1.0p30))
            return false; // outside the range the hoisted sphere test is exact for (geometry.cuh)
        ext[a] = gmax[a] - gmin[a];
        if (!(ext[a] > 0) || !std::isfinite(ext[a]))
            return false;
        volume *= ext[a];
    }
    double density = 4.0; // target cells per sphere
    if (const char *env = std::getenv("CORNELIS_GRID_DENSITY"))
        if (std::atof(env) > 0)
            density = std::atof(env);
    for (int attempt = 0; attempt < 8; attempt++, density *= 0.25) {
        double const edge = std::cbrt(volume / std::max(1.0, density * static_cast<double>(n)));
        uint32_t dim[3];
        for (int a = 0; a < 3; a++)
            dim[a] = static_cast<uint32_t>(std::min(256.0, std::max(1.0, std::ceil(ext[a] / edge))));
        double cell[3];
        for (int a = 0; a < 3; a++)
            cell[a] = ext[a] / dim[a];
        size_t const ncell = static_cast<size_t>(dim[0]) * dim[1] * dim[2];
        auto range = [&](size_t i, int a, uint32_t &first, uint32_t &last) {
            double const f = std::floor((lo[3 * i + a] - gmin[a]) / cell[a]), l = std::floor((hi[3 * i + a] - gmin[a]) / cell[a]);
            first = static_cast<uint32_t>(std::min<double>(dim[a] - 1, std::max(0.0, f)));
            last = static_cast<uint32_t>(std::min<double>(dim[a] - 1, std::max(0.0, l)));
        };
        // a sphere is listed in the cells of its padded box that come within `reach` of its centre
        std::vector<double> reach2(n);
        for (size_t i = 0; i < n; i++) {
            double const r = std::fabs(static_cast<double>(spheres[i].radius));
            double const reach = std::sqrt(r * r + eps) + delta;
            reach2[i] = reach * reach;
        }
        auto touches = [&](size_t i, uint32_t x, uint32_t y, uint32_t z) {
            uint32_t const c[3] = {x, y, z};
            double dist2 = 0;
            for (int a = 0; a < 3; a++) {
                double const lo_ = gmin[a] + c[a] * cell[a], hi_ = lo_ + cell[a], p = spheres[i].center[a];
                double const dd = p < lo_ ? lo_ - p : p > hi_ ? p - hi_ : 0.0;
                dist2 += dd * dd;
            }
            return dist2 <= reach2[i] * (1.0 + 1e-9);
        };
        std::vector<uint32_t> start(ncell + 1, 0);
        uint64_t total = 0;
        for (size_t i = 0; i < n; i++) {
            uint32_t f[3], l[3];
            for (int a = 0; a < 3; a++)
                range(i, a, f[a], l[a]);
            total += static_cast<uint64_t>(l[0] - f[0] + 1) * (l[1] - f[1] + 1) * (l[2] - f[2] + 1);
            if (total > (1ull << 28))
                break;
            for (uint32_t z = f[2]; z <= l[2]; z++)
                for (uint32_t y = f[1]; y <= l[1]; y++)
                    for (uint32_t x = f[0]; x <= l[0]; x++)
                        if (touches(i, x, y, z))
                            start[(static_cast<size_t>(z) * dim[1] + y) * dim[0] + x + 1]++;
        }
        if (total > (1ull << 28))
            continue; // too many references at this resolution: coarsen
        for (size_t c = 0; c < ncell; c++)
            start[c + 1] += start[c];
        size_t const nrefs = start[ncell];
        std::vector<uint32_t> ids(nrefs ? nrefs : 1), fill(start.begin(), start.end() - 1);
        std::vector<float4> copies(nrefs ? nrefs : 1);
        for (size_t i = 0; i < n; i++) { // ascending sphere index within every cell
            uint32_t f[3], l[3];
            for (int a = 0; a < 3; a++)
                range(i, a, f[a], l[a]);
            float4 const copy = make_float4(spheres[i].center[0], spheres[i].center[1], spheres[i].center[2],
                                            spheres[i].radius * spheres[i].radius); // as DevSphere (api.cu)
            for (uint32_t z = f[2]; z <= l[2]; z++)
                for (uint32_t y = f[1]; y <= l[1]; y++)
                    for (uint32_t x = f[0]; x <= l[0]; x++)
                        if (touches(i, x, y, z)) {
                            uint32_t const slot = fill[(static_cast<size_t>(z) * dim[1] + y) * dim[0] + x]++;
                            ids[slot] = static_cast<uint32_t>(i);
                            copies[slot] = copy;
                        }
        }
        std::vector<uint2> ranges(ncell);
        for (size_t c = 0; c < ncell; c++)
            ranges[c] = make_uint2(start[c], start[c + 1]);
        DevGrid &g = out.g;
        // round the grid bounds outwards, the trusted region inwards
        g.minx = std::nextafterf(static_cast<float>(gmin[0]), -INFINITY), g.maxx = std::nextafterf(static_cast<float>(gmax[0]), INFINITY);
        g.miny = std::nextafterf(static_cast<float>(gmin[1]), -INFINITY), g.maxy = std::nextafterf(static_cast<float>(gmax[1]), INFINITY);
        g.minz = std::nextafterf(static_cast<float>(gmin[2]), -INFINITY), g.maxz = std::nextafterf(static_cast<float>(gmax[2]), INFINITY);
        g.cellx = static_cast<float>(cell[0]), g.celly = static_cast<float>(cell[1]), g.cellz = static_cast<float>(cell[2]);
        g.invx = static_cast<float>(1.0 / cell[0]), g.invy = static_cast<float>(1.0 / cell[1]), g.invz = static_cast<float>(1.0 / cell[2]);
        g.rminx = std::nextafterf(static_cast<float>(rmin[0]), INFINITY), g.rmaxx = std::nextafterf(static_cast<float>(rmax[0]), -INFINITY);
        g.rminy = std::nextafterf(static_cast<float>(rmin[1]), INFINITY), g.rmaxy = std::nextafterf(static_cast<float>(rmax[1]), -INFINITY);
        g.rminz = std::nextafterf(static_cast<float>(rmin[2]), INFINITY), g.rmaxz = std::nextafterf(static_cast<float>(rmax[2]), -INFINITY);
#if CORNELIS_GRID_RAY_MARGIN
        // Per-ray termination slack (geometry.cuh): 2 sqrt(kappa) (t_best |d| + r_max) with kappa = 2^-19, inflated by
        // 1.003 (|o - c| against t |d| + r, see there) and 1 % for the roundings of the slack itself.
        double rmaxSphere = 0;
        for (size_t i = 0; i < n; i++)
            rmaxSphere = std::max(rmaxSphere, std::fabs(static_cast<double>(spheres[i].radius)));
        double const slope = 2.0 * std::sqrt(std::ldexp(1.0, -19)) * 1.003 * 1.01;
        g.margin = static_cast<float>(slope * rmaxSphere + 1e-6 * D);
        g.marginScale = std::nextafterf(static_cast<float>(1.0 + slope), INFINITY);
#else
        g.margin = static_cast<float>(2.0 * std::sqrt(eps) * 1.001);
        g.marginScale = 1.0f;
#endif
        g.nx = dim[0], g.ny = dim[1], g.nz = dim[2];
        g.enabled = 1;
        out.cellRange.swap(ranges);
        out.cellSpheres.swap(copies);
        out.cellIds.swap(ids);
        out.cellIds.resize(nrefs);
        out.cellSpheres.resize(nrefs);
        return true;
    }
    return false;
}

} // namespace cornelis_b200
