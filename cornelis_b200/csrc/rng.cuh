// Counter-based random numbers: Philox4x32 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as
// 1, 2, 3", SC'11) with kPhiloxRounds = 7 rounds.  The paper's table 2 reports Philox4x32 Crush-resistant (all of
// TestU01's SmallCrush, Crush and BigCrush) from 7 rounds on; 10 is its default only as a safety margin.  A path
// tracer needs streams that pass as i.i.d. uniform, not margin against cryptanalysis, and the generator runs once per
// camera ray and once per shaded hit: three rounds fewer are 2 % of the render kernel's instructions
// (-DCORNELIS_PHILOX_ROUNDS=10 restores the default).  The round function is the library's: tests check the 10-round
// Random123 known-answer vector through the same code (cornelis_cuda_rng_bits) and the 7-round stream against a numpy
// restatement validated by that vector.  Replaces the reference's per-tile xoshiro128+ stream (include/cornelis/PRNG.hpp:11-37): a
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

#ifndef CORNELIS_PHILOX_ROUNDS
#define CORNELIS_PHILOX_ROUNDS 7
#endif
constexpr int kPhiloxRounds = CORNELIS_PHILOX_ROUNDS; // of the render loop's generator
constexpr int kPhiloxMaxRounds = 10;
static_assert(kPhiloxRounds >= 7 && kPhiloxRounds <= kPhiloxMaxRounds, "Philox4x32 is Crush-resistant from 7 rounds on");

#ifndef CORNELIS_PHILOX_UNROLL
#define CORNELIS_PHILOX_UNROLL 10 // rounds unrolled per loop trip (>= the round count = straight-line code)
#endif
#define CB_PHILOX_PRAGMA(x) _Pragma(#x)
#define CB_PHILOX_UNROLL_N(n) CB_PHILOX_PRAGMA(unroll n)
#define CB_PHILOX_UNROLL CB_PHILOX_UNROLL_N(CORNELIS_PHILOX_UNROLL)

struct Philox4 {
    uint32_t v[4];
};

// Philox4x32-R with the key schedule computed on the fly (stage entry points and self-tests).
__host__ __device__ __forceinline__ Philox4 philox4x32(int rounds, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int round = 0; round < rounds; round++) {
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

// The round keys depend only on the seed: the host expands them once per render into the kernel parameters, so
// the per-round key bumps disappear from the device code (the keys become constant-bank operands of the xors).
struct PhiloxKeys {
    uint32_t k[2 * kPhiloxMaxRounds]; // k[2r] = key0 + r * W0, k[2r + 1] = key1 + r * W1
};

inline PhiloxKeys makePhiloxKeys(uint32_t key0, uint32_t key1) {
    PhiloxKeys keys{};
    for (uint32_t r = 0; r < static_cast<uint32_t>(kPhiloxMaxRounds); r++) {
        keys.k[2 * r] = key0 + r * 0x9E3779B9u;
        keys.k[2 * r + 1] = key1 + r * 0xBB67AE85u;
    }
    return keys;
}

// The render loop's generator: kPhiloxRounds rounds over the pre-expanded keys.
__host__ __device__ __forceinline__ Philox4 philoxRender(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         const PhiloxKeys &keys) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    CB_PHILOX_UNROLL
    for (int round = 0; round < kPhiloxRounds; round++) {
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
