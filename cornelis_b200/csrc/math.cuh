// FP32 recipes of the reference's Math.hpp, written once for host and device.
//
// Operation ORDER matters here: the intersection kernel has to reproduce the reference's t bit for bit, so every
// expression keeps the reference's association (dot = a0*b0 + a1*b1 + a2*b2 left to right, Math.hpp:278) and the
// translation unit is compiled with --fmad=false (no contraction), IEEE division and square root.
#pragma once

#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

#ifndef CB_HD
#define CB_HD __host__ __device__ __forceinline__
#endif

#include "exact_arith.cuh"

namespace cornelis_b200 {

struct V3 {
    float x, y, z;
};

constexpr float kRayEpsilon = 0.00005f;   // Math.hpp:20
constexpr float kPi = 3.14159265359f;     // Math.hpp:25

CB_HD bool isAlmostZero(float v) { return fabsf(v) < kRayEpsilon; } // Math.hpp:22

// std::max / std::min / std::clamp semantics — NOT fmaxf/fminf: with a NaN second argument std::max returns the
// first (Materials.hpp:223-227 relies on std::max(0.0f, NaN) == 0).
CB_HD float stdMax(float a, float b) { return (a < b) ? b : a; }
CB_HD float stdMin(float a, float b) { return (b < a) ? b : a; }
CB_HD float stdClamp(float v, float lo, float hi) { return (v < lo) ? lo : (hi < v) ? hi : v; }

CB_HD V3 v3(float x, float y, float z) { return V3{x, y, z}; }
CB_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }   // Math.hpp:63-70
CB_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }   // Math.hpp:75-82
CB_HD V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }                        // Math.hpp:86-93
CB_HD V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }   // Math.hpp:98-105
CB_HD V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }      // Math.hpp:110-117
CB_HD V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }      // Math.hpp:121-128
CB_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }        // Math.hpp:278
CB_HD float mag2(V3 a) { return dot(a, a); }                                     // Math.hpp:284
CB_HD V3 rayT(V3 o, V3 d, float t) { return o + d * V3{t, t, t}; }               // Math.hpp:290-292

CB_HD V3 cross(V3 a, V3 b) { // Math.hpp:380-384
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Math.hpp:392-398: length below RayEpsilon collapses to the zero vector; otherwise multiply by the ROUNDED
// reciprocal (not a division per component).
CB_HD V3 normalizeGeneric(V3 v, float x) {
    float len = sqrtf(x);
    if (isAlmostZero(len))
        return V3{0.0f, 0.0f, 0.0f};
    float s = 1.0f / len;
    return v * V3{s, s, s};
}
#ifdef __CUDACC__
// out of line on the device: lengths outside the fast range are never seen in a render, and six inlined copies of the
// operator sequences with their slow-path calls cost instruction-cache space in the hot loop
static __device__ __noinline__ V3 normalizeOutOfRange(V3 v, float x) { return normalizeGeneric(v, x); }
#endif
// `odd`, when given (the persistent kernel's scatter half), replaces the per-call branch to the out-of-range path: the
// fast sequence runs unconditionally and the caller redoes the whole step out of line for a lane whose squared length
// was outside its range.  *odd accumulates the LARGEST (bits(x) - bits(2^-28)) seen, as unsigned: in range means at
// most kNormalizeSpan, anything else (smaller, larger, negative, NaN) wraps or lands above it — one subtract and one
// unsigned maximum per call, one compare per step (oddRaised).  The watched range starts at 2^-28 rather than at the
// 2^-40 the fast sequence could take: from there up the length is at least 2^-14 > RayEpsilon, so the reference's
// collapse of a short vector to zero (Math.hpp:394) cannot apply and the watched call does not test for it.
constexpr uint32_t kNormalizeLo = 0x31800000u, kNormalizeSpan = 0x67800000u - 0x31800000u; // 2^-28 .. 2^80
typedef uint32_t OddWatch;
CB_HD bool oddRaised(OddWatch w) { return w > kNormalizeSpan; }
CB_HD V3 normalize(V3 v, OddWatch *odd = nullptr) {
    float const x = mag2(v);
#ifdef __CUDA_ARCH__
    if (odd) {
        uint32_t const excess = __float_as_uint(x) - kNormalizeLo; // inFastNormalizeRange on the bit pattern
        *odd = *odd > excess ? *odd : excess;
    } else if (!inFastNormalizeRange(x)) {
        return normalizeOutOfRange(v, x);
    }
    float len, s; // same values as normalizeGeneric, bit for bit (exact_arith.cuh)
    sqrtAndReciprocalExactFast(x, len, s);
    if (!odd && isAlmostZero(len))
        return V3{0.0f, 0.0f, 0.0f};
    return v * V3{s, s, s};
#else
    (void)odd;
    return normalizeGeneric(v, x);
#endif
}

struct Basis {
    V3 N, T, B;
};

// Math.hpp:424-434.  `abs(N(1)) > 0.95` compares a float against a double literal; the smallest float above
// 0.95 is also the smallest float above 0.95f, so the float comparison below decides identically.
CB_HD Basis constructBasis(V3 N, OddWatch *odd = nullptr) {
    V3 helper{0.0f, 1.0f, 0.0f};
    if (fabsf(N.y) > 0.95f)
        helper = V3{0.0f, 0.0f, 1.0f};
    Basis b;
    b.N = N;
    b.T = normalize(cross(helper, N), odd);
    b.B = cross(b.T, N);
    return b;
}

} // namespace cornelis_b200
