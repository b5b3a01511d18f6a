// Packed FP32 arithmetic (PTX *.f32x2 -> SASS FFMA2, new with sm_100): one instruction issues two IEEE round-to-nearest
// operations on a 64-bit register pair.  The FMA pipe needs two cycles for it (tools/ubench/f32x2_rate.cu: 1.7 packed
// against 3.5 scalar warp-instructions per clock per SM, the same 32 T lane-operations/s), so the gain is ISSUE SLOTS:
// a loop whose FP32 work is paired issues half as many arithmetic instructions and leaves the slots to the loads,
// compares and votes around it.
//
// Exactness.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under --fmad=false (CUDA 12.9), and it
// does the same after simplifying fma(a, b, -0) to a product and fma(a, 1, b) to a sum — which would change the
// reference's bits.  So every operation here is an fma.rn.f32x2 whose neutral operand ptxas cannot see through:
//     a * b == fma(a, b, NEGZERO)   (a -0 addend keeps the sign of a zero product: (+0) + (-0) = +0, (-0) + (-0) = -0)
//     a + b == fma(a, ONE, b)       (a * 1 is exact)
// for every input including NaN and the infinities, with ONE = 1.0f and NEGZERO = -0.0f arriving as KERNEL PARAMETERS
// (PackedConstants, filled in by the host).  An FFMA2 with three register / constant-bank operands of unknown value
// cannot be simplified or contracted.  tests/test_gpu_parity.py compares the packed scan with the oracle bit for bit.
#pragma once

#include <cuda_runtime.h>

namespace cornelis_b200 {

struct F2 {
    unsigned long long bits; // {lo, hi} floats in one aligned register pair
};

__device__ __forceinline__ F2 pack2(float lo, float hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.bits) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ F2 splat2(float x) { return pack2(x, x); }
__device__ __forceinline__ void unpack2(F2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v.bits));
}
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.bits) : "l"(a.bits), "l"(b.bits), "l"(c.bits));
    return r;
}

// The neutral operands, opaque to ptxas (see above): passed to the kernel by the host as (1.0f, -0.0f).
struct PackedConstants {
    float one, negZero;
};
inline PackedConstants hostPackedConstants() { return PackedConstants{1.0f, -0.0f}; }

struct PackedNeutral {
    F2 one, negZero;
};
__device__ __forceinline__ PackedNeutral packedNeutral(PackedConstants c) {
    return PackedNeutral{splat2(c.one), splat2(c.negZero)};
}
__device__ __forceinline__ F2 mul2(F2 a, F2 b, const PackedNeutral &k) { return fma2(a, b, k.negZero); }
__device__ __forceinline__ F2 add2(F2 a, F2 b, const PackedNeutral &k) { return fma2(a, k.one, b); }

} // namespace cornelis_b200
