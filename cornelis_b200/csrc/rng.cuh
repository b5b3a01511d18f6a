// Counter-based random numbers: Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as
// 1, 2, 3", SC'11).  Replaces the reference's per-tile xoshiro128+ stream (include/cornelis/PRNG.hpp:11-37): a
// path's numbers depend only on (seed; pixel, global sample index, dimension block), never on which thread, which
// wavefront pass or which GPU processes it, so sample-sharded multi-GPU renders draw the same sample set as one GPU.
//
//   key     = (seed low 32 bits, seed high 32 bits)
//   counter = (pixel index, global sample index, dimension block, 0)
//   block 0     -> camera jitter (phi1, phi2) = outputs 0, 1            (Render.cpp:94-95)
//   block d + 1 -> bounce at depth d: RR draw = output 0, x0..x2 = outputs 1..3   (Render.cpp:189, 199)
//
// Floats keep the reference's mapping (XoshiroCpp.hpp:651-655): (u >> 8) * 2^-24, i.e. U[0,1) on a 24-bit grid.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace cornelis_b200 {

#ifndef CORNELIS_PHILOX_UNROLL
#define CORNELIS_PHILOX_UNROLL 10 // rounds unrolled per loop trip (10 = straight-line code)
#endif
#define CB_PHILOX_PRAGMA(x) _Pragma(#x)
#define CB_PHILOX_UNROLL_N(n) CB_PHILOX_PRAGMA(unroll n)
#define CB_PHILOX_UNROLL CB_PHILOX_UNROLL_N(CORNELIS_PHILOX_UNROLL)

struct Philox4 {
    uint32_t v[4];
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int round = 0; round < 10; round++) {
        unsigned long long p0 = static_cast<unsigned long long>(M0) * c0;
        unsigned long long p1 = static_cast<unsigned long long>(M1) * c2;
        uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
        c1 = static_cast<uint32_t>(p1);
        c3 = static_cast<uint32_t>(p0);
        c0 = n0;
        c2 = n2;
        k0 += W0;
        k1 += W1;
    }
    return Philox4{{c0, c1, c2, c3}};
}

// The ten round keys depend only on the seed: the host expands them once per render into the kernel parameters, so
// the per-round key bumps disappear from the device code (the keys become constant-bank operands of the xors).
struct PhiloxKeys {
    uint32_t k[20]; // k[2r] = key0 + r * W0, k[2r + 1] = key1 + r * W1
};

inline PhiloxKeys makePhiloxKeys(uint32_t key0, uint32_t key1) {
    PhiloxKeys keys{};
    for (uint32_t r = 0; r < 10; r++) {
        keys.k[2 * r] = key0 + r * 0x9E3779B9u;
        keys.k[2 * r + 1] = key1 + r * 0xBB67AE85u;
    }
    return keys;
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          const PhiloxKeys &keys) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    CB_PHILOX_UNROLL
    for (int round = 0; round < 10; round++) {
        unsigned long long p0 = static_cast<unsigned long long>(M0) * c0;
        unsigned long long p1 = static_cast<unsigned long long>(M1) * c2;
        uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ keys.k[2 * round];
        uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ keys.k[2 * round + 1];
        c1 = static_cast<uint32_t>(p1);
        c3 = static_cast<uint32_t>(p0);
        c0 = n0;
        c2 = n2;
    }
    return Philox4{{c0, c1, c2, c3}};
}

__host__ __device__ __forceinline__ float uniformFromBits(uint32_t u) {
    return static_cast<float>(u >> 8) * 0x1.0p-24f; // exact: 24-bit integer times a power of two
}

} // namespace cornelis_b200
