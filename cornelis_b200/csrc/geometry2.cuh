// closestHit for TWO rays per lane on packed FP32 (PTX *.f32x2 -> SASS FFMA2): the arithmetic of geometry.cuh's fast
// paths — one sphere head, one square-root refinement, one axis-aligned plane test — issued once for a pair of rays.
//
// Why: ncu on the render kernel (profiles/r2_ncu) shows it bound by issue slots (85 % busy) with the FMA pipe at 35 %:
// the FP32 instructions are already at the algorithmic count, so the only way to issue fewer of them is to make each
// one do two IEEE operations.  Packing two SPHERES for one ray (scanSpheresPacked) pays for a 1024-sphere scan and does
// nothing for four; packing two RAYS pays for every primitive, and the loop control, table loads and votes around the
// arithmetic are shared by the pair as well.
//
// Exactness: every packed operation is an fma.rn.f32x2 with an opaque neutral operand (packed_f32.cuh), i.e. the same
// IEEE operation on the same operands as the scalar code, so each half of the pair carries the bits closestHit computes
// for that ray.  The subtractions are written fma(b, -1, a) (b * -1 is exact) and negations fma(x, -1, -0) (exact, and
// the sign of a zero flips as it does for unary minus).  Everything the scalar code settles per ray — range checks,
// root selection, the strict compares, the sign of a zero-distance plane hit — stays scalar, once per half.
// A warp in which any live ray fails the fast-path preconditions, and any scene with planes outside the axis classes,
// takes the scalar closestHit for both halves (out of line).
#pragma once

#include "geometry.cuh"

#ifdef __CUDACC__
namespace cornelis_b200 {

// The tables closestHit2 reads, staged behind the scene tables by stageSplatted (kernels.cuh): every value twice, so
// that one 128-bit load fills two register pairs.
//   sphere i:      (cx, cx, cy, cy) (cz, cz, r2, r2)
//   axis plane k:  (p0k, p0k, p0T, p0T) (p0B, p0B, w/2, w/2) (h/2, h/2, id bits, 0)        class by class (DevAxisPlane)
struct SplattedScene {
    const float4 *spheres; // 2 per sphere
    const float4 *planes;  // 3 per axis-aligned plane
};

__host__ __device__ inline size_t splattedBytes(uint32_t nSpheres, uint32_t nAxisPlanes) {
    return sizeof(float4) * (2u * static_cast<size_t>(nSpheres) + 3u * static_cast<size_t>(nAxisPlanes));
}

struct Pair {
    float a, b;
};
__device__ __forceinline__ Pair halves(F2 v) {
    Pair p;
    unpack2(v, p.a, p.b);
    return p;
}

struct PackedOps {
    F2 one, negOne, negZero;
    __device__ __forceinline__ F2 mul(F2 x, F2 y) const { return fma2(x, y, negZero); }
    __device__ __forceinline__ F2 add(F2 x, F2 y) const { return fma2(x, one, y); }
    __device__ __forceinline__ F2 sub(F2 x, F2 y) const { return fma2(y, negOne, x); } // x - y
    __device__ __forceinline__ F2 neg(F2 x) const { return fma2(x, negOne, negZero); }
};

__device__ __forceinline__ float rcpApproxFtz(float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return r;
}
__device__ __forceinline__ float rsqrtApproxFtz(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// rcpSeedRefined (exact_arith.cuh) for both halves: the seeds come from MUFU per half, the Newton step is packed.
__device__ __forceinline__ F2 rcpSeedRefined2(const PackedOps &k, F2 b, F2 negB) {
    Pair const h = halves(b);
    F2 const r0 = pack2(rcpApproxFtz(h.a), rcpApproxFtz(h.b));
    F2 const e = fma2(negB, r0, k.one);
    return fma2(r0, e, r0);
}

// The scalar per-ray tail of one sphere test (geometry.cuh sphereTail, select-chain form) given the packed roots.
__device__ __forceinline__ void sphereRoots(bool live, float discriminant, float t0, float t1, uint32_t i, float &tBest,
                                            int32_t &primBest, bool &odd) {
    uint32_t const bits = __float_as_uint(discriminant);
    constexpr uint32_t kLo = 0x0d800000u, kHi = 0x7e800000u; // 2^-100, 2^126: inFastSqrtRange
    odd = odd | ((discriminant >= 0.0f) & (bits - kLo >= kHi - kLo));
    float const t = !(t0 < 0.0f) ? t0 : t1;
    if (live & (discriminant >= 0.0f) & !(t < 0.0f) & (tBest > t)) { // Geometry.cpp:97 — strict
        tBest = t;
        primBest = static_cast<int32_t>(i);
    }
}

// axisPlaneTest's compares for one half.
__device__ __forceinline__ void planeOffer(bool live, float t, float eT, float eB, float halfW, float halfH, int32_t id,
                                           float &tBest, int32_t &primBest) {
    bool const outside = (fabsf(eT) > halfW) | (fabsf(eB) > halfH);      // Geometry.cpp:166-167
    bool const closer = (tBest > t) | ((tBest == t) & (id < primBest));  // Geometry.cpp:169
    if (live & !(t < 0.0f) & !outside & closer) {                        // Geometry.cpp:161-163
        tBest = t;
        primBest = id;
    }
}

struct Rays2 {
    F2 ox, oy, oz, dx, dy, dz;
};

template <int AXIS>
__device__ __forceinline__ void axisPlaneTest2(const PackedOps &k, bool liveA, bool liveB, const Rays2 &r, F2 rk, F2 negDk,
                                               const float4 *__restrict__ p, float &tA, int32_t &primA, float &tB,
                                               int32_t &primB) {
    float4 const q0 = p[0], q1 = p[1], q2 = p[2];
    F2 const p0k = pack2(q0.x, q0.y), p0T = pack2(q0.z, q0.w), p0B = pack2(q1.x, q1.y);
    F2 const ok_ = AXIS == 0 ? r.ox : AXIS == 1 ? r.oy : r.oz;
    F2 const oT = AXIS == 0 ? r.oz : r.ox, dT = AXIS == 0 ? r.dz : r.dx;
    F2 const oB = AXIS == 1 ? r.oz : r.oy, dB = AXIS == 1 ? r.dz : r.dy;
    F2 const num = k.neg(k.sub(ok_, p0k));               // -(o_k - p0_k)
    F2 const q = k.mul(num, rk);                         // divideExactFast(num, d_k, r_k)
    F2 const t = fma2(rk, fma2(negDk, q, num), q);
    F2 const eT = k.sub(k.add(oT, k.mul(dT, t)), p0T);   // (o_T + d_T t) - p0_T
    F2 const eB = k.sub(k.add(oB, k.mul(dB, t)), p0B);
    Pair const tt = halves(t), et = halves(eT), eb = halves(eB);
    int32_t const id = __float_as_int(q2.z);
    planeOffer(liveA, tt.a, et.a, eb.a, q1.z, q2.x, id, tA, primA);
    planeOffer(liveB, tt.b, et.b, eb.b, q1.z, q2.x, id, tB, primB);
}

// The scalar closestHit for one of the halves, out of line (warp-cooperative like closestHit itself: every lane calls).
static __device__ __noinline__ HitPair closestHitOne(bool live, V3 o, V3 d, const SharedScene &sh, const SceneView &scene) {
    float t = INFINITY; // IntersectionData::reset, Geometry.cpp:7-12
    int32_t prim = -1;
    closestHit<1, kScanScalar>(live, o, d, sh, scene, t, prim);
    return HitPair{t, prim};
}

// Per-ray preconditions of the fast paths (closestHit; A = d.d): returns planeOk, clears `live` for degenerate directions.
__device__ __forceinline__ bool fastRay(bool &live, V3 o, V3 d, float A) {
    live = live && !isDegenerateDirection(d); // Geometry.cpp:67-70, :145-148
    bool const sane = fabsf(o.x) <= 0x1.0p30f && fabsf(o.y) <= 0x1.0p30f && fabsf(o.z) <= 0x1.0p30f &&
                      fabsf(d.x) <= 0x1.0p19f && fabsf(d.y) <= 0x1.0p19f && fabsf(d.z) <= 0x1.0p19f && A >= 0x1.0p-40f;
    return sane && !isAlmostZero(d.x) && !isAlmostZero(d.y) && !isAlmostZero(d.z) && differenceSafe(o.x) &&
           differenceSafe(o.y) && differenceSafe(o.z);
}

// Closest hits of rays A and B of every lane; t = +INF, prim = -1 on entry is implied (results are returned).
__device__ __forceinline__ void closestHit2(bool liveA, bool liveB, V3 oA, V3 dA, V3 oB, V3 dB, const SharedScene &sh,
                                            const SplattedScene &splat, const SceneView &scene, PackedConstants neutral,
                                            float &tA, int32_t &primA, float &tB, int32_t &primB) {
    constexpr unsigned kFull = 0xffffffffu;
    uint32_t const nSpheres = scene.nSpheres, nPlanes = scene.nPlanes;
    PackedOps const k{splat2(neutral.one), splat2(-neutral.one), splat2(neutral.negZero)};
    Rays2 const r{pack2(oA.x, oB.x), pack2(oA.y, oB.y), pack2(oA.z, oB.z), pack2(dA.x, dB.x), pack2(dA.y, dB.y),
                  pack2(dA.z, dB.z)};
    F2 const A = k.add(k.add(k.mul(r.dx, r.dx), k.mul(r.dy, r.dy)), k.mul(r.dz, r.dz)); // dot(d, d), both rays
    Pair const A1 = halves(A);
    bool const okA = fastRay(liveA, oA, dA, A1.a), okB = fastRay(liveB, oB, dB, A1.b);
    bool const fast = scene.radiiSafe && scene.planeEnd[2] == nPlanes &&
                      __all_sync(kFull, (okA || !liveA) && (okB || !liveB));
    if (!fast) {
        HitPair const a = closestHitOne(liveA, oA, dA, sh, scene), b = closestHitOne(liveB, oB, dB, sh, scene);
        tA = a.t, primA = a.prim, tB = b.t, primB = b.prim;
        return;
    }
    tA = INFINITY, tB = INFINITY, primA = -1, primB = -1;

    // ---- spheres (geometry.cuh sphereHead / sphereTail) ----
    F2 const negA = k.neg(A);
    F2 const rA = rcpSeedRefined2(k, A, negA);
    F2 const quarter = splat2(0.25f), half = splat2(0.5f), negHalf = splat2(-0.5f);
    float smallestA = INFINITY, smallestB = INFINITY;
    bool oddA = false, oddB = false;
    const float4 *__restrict__ s = splat.spheres;
#pragma unroll 1
    for (uint32_t i = 0; i < nSpheres; i++, s += 2) {
        float4 const c0 = s[0], c1 = s[1];
        F2 const cx = pack2(c0.x, c0.y), cy = pack2(c0.z, c0.w), cz = pack2(c1.x, c1.y), r2 = pack2(c1.z, c1.w);
        F2 const Px = k.sub(r.ox, cx), Py = k.sub(r.oy, cy), Pz = k.sub(r.oz, cz);                   // o - c
        F2 const B = k.add(k.add(k.mul(Px, r.dx), k.mul(Py, r.dy)), k.mul(Pz, r.dz));                // dot(P, d)
        F2 const C = k.add(k.add(k.mul(Px, Px), k.mul(Py, Py)), k.mul(Pz, Pz));                      // mag2(P)
        F2 const nu = k.add(B, B);                                                                  // 2 B
        F2 const nvNeg = k.sub(r2, C);                                                              // -(C - r^2)
        F2 const qu = k.mul(nu, rA);
        F2 const u = fma2(rA, fma2(negA, qu, nu), qu);                                               // 2 B / A
        F2 const qv = k.mul(nvNeg, rA);
        F2 const vNeg = fma2(rA, fma2(negA, qv, nvNeg), qv);                                         // -v
        F2 const disc = k.add(k.mul(k.mul(u, u), quarter), vNeg);                                    // -v + u^2 / 4
        Pair const n = halves(nu), dd = halves(disc);
        smallestA = fminf(smallestA, fabsf(n.a));
        smallestB = fminf(smallestB, fabsf(n.b));
        if (!__any_sync(kFull, (liveA & (dd.a >= 0.0f)) | (liveB & (dd.b >= 0.0f))))
            continue; // no lane has a root in this sphere for either ray (Geometry.cpp:85-86)
        // sqrtExactFast for both halves, then t0 = -u/2 - shift, t1 = -u/2 + shift
        F2 const rs = pack2(rsqrtApproxFtz(dd.a), rsqrtApproxFtz(dd.b));
        F2 const q = k.mul(disc, rs), h = k.mul(rs, half);
        F2 const shift = fma2(fma2(k.neg(q), q, disc), h, q);
        F2 const mid = k.mul(u, negHalf);
        Pair const t0 = halves(k.sub(mid, shift)), t1 = halves(k.add(mid, shift));
        sphereRoots(liveA, dd.a, t0.a, t1.a, i, tA, primA, oddA);
        sphereRoots(liveB, dd.b, t0.b, t1.b, i, tB, primB, oddB);
    }
    // a tiny numerator or a discriminant outside the fast square root's range: that half is scanned again, slowly
    bool const redoA = __any_sync(kFull, liveA && (oddA || smallestA < 0x1.0p-80f));
    bool const redoB = __any_sync(kFull, liveB && (oddB || smallestB < 0x1.0p-80f));
    if (redoA) {
        HitPair const hit = scanSpheresSlow(liveA, oA, dA, A1.a, sh.spheres, nSpheres, INFINITY, -1);
        tA = hit.t, primA = hit.prim;
    }
    if (redoB) {
        HitPair const hit = scanSpheresSlow(liveB, oB, dB, A1.b, sh.spheres, nSpheres, INFINITY, -1);
        tB = hit.t, primB = hit.prim;
    }

    // ---- axis-aligned planes (geometry.cuh axisPlaneTest), class by class ----
    F2 const negDx = k.neg(r.dx), negDy = k.neg(r.dy), negDz = k.neg(r.dz);
    F2 const rx = rcpSeedRefined2(k, r.dx, negDx), ry = rcpSeedRefined2(k, r.dy, negDy), rz = rcpSeedRefined2(k, r.dz, negDz);
    const float4 *__restrict__ p = splat.planes;
    const float4 *const endX = p + 3u * scene.planeEnd[0], *const endY = p + 3u * scene.planeEnd[1],
                       *const endZ = p + 3u * scene.planeEnd[2];
#pragma unroll 1
    for (; p != endX; p += 3)
        axisPlaneTest2<0>(k, liveA, liveB, r, rx, negDx, p, tA, primA, tB, primB);
#pragma unroll 1
    for (; p != endY; p += 3)
        axisPlaneTest2<1>(k, liveA, liveB, r, ry, negDy, p, tA, primA, tB, primB);
#pragma unroll 1
    for (; p != endZ; p += 3)
        axisPlaneTest2<2>(k, liveA, liveB, r, rz, negDz, p, tA, primA, tB, primB);
    // a plane hit at t == 0: the sign of the zero is the reference's (geometry.cuh settleZeroPlaneHit)
    if (nPlanes) {
        int32_t const hitA = primA - static_cast<int32_t>(nSpheres), hitB = primB - static_cast<int32_t>(nSpheres);
        float const settledA = settleZeroPlaneHit(oA, dA, sh.planes[hitA > 0 ? hitA : 0], tA);
        float const settledB = settleZeroPlaneHit(oB, dB, sh.planes[hitB > 0 ? hitB : 0], tB);
        tA = (tA == 0.0f) & (hitA >= 0) ? settledA : tA;
        tB = (tB == 0.0f) & (hitB >= 0) ? settledB : tB;
    }
}

} // namespace cornelis_b200
#endif // __CUDACC__
