// Division of a 32-bit unsigned integer by a run-time constant without the ~25-instruction software divide:
// q = floor(x / d) for every 32-bit x, d >= 1, as  t = umulhi(x, magic); q = (((x - t) >> 1) + t) >> shift
// (the "round-up, always add" form of Granlund-Montgomery division; d == 1 and powers of two use magic = 0 and a
// pre-shift).  The constants are computed on the host per render (frame width).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace cornelis_b200 {

struct FastDiv {
    uint32_t magic;
    uint32_t shift;   // applied after the averaging step
    uint32_t one;     // d == 1: the quotient is x itself
    uint32_t divisor;
};

inline FastDiv makeFastDiv(uint32_t d) {
    FastDiv f{};
    f.divisor = d;
    if (d <= 1) {
        f.one = 1;
        return f;
    }
    uint32_t L = 31;
    while (!(d >> L))
        L--; // floor(log2 d)
    if ((d & (d - 1)) == 0) { // power of two: t = 0, ((x - 0) >> 1) >> (L - 1) = x >> L
        f.magic = 0;
        f.shift = L - 1;
        return f;
    }
    // magic = floor(2^(33 + L) / d) - 2^32 + 1 (a 33-bit multiplier with its top bit implied by the add)
    uint64_t const num = 1ull << (32 + L);
    uint64_t m = num / d, rem = num % d;
    m += m;
    uint64_t const twice = rem + rem;
    if (twice >= d)
        m += 1;
    f.magic = static_cast<uint32_t>(m + 1);
    f.shift = L;
    return f;
}

__host__ __device__ __forceinline__ uint32_t fastDivide(uint32_t x, const FastDiv &f) {
#ifdef __CUDA_ARCH__
    uint32_t const t = __umulhi(x, f.magic);
#else
    uint32_t const t = static_cast<uint32_t>((static_cast<uint64_t>(x) * f.magic) >> 32);
#endif
    uint32_t const q = (((x - t) >> 1) + t) >> f.shift;
    return f.one ? x : q;
}

} // namespace cornelis_b200
