// The reference's "standard material": emission + LayeredBRDF = Oren-Nayar diffuse under a GGX glossy coat
// (include/cornelis/Materials.hpp:59-338, src/Materials.cpp:16-42), as flat device functions over DevMaterial.
// The virtual BRDF hierarchy has one concrete leaf (Materials.hpp:331), so dispatch is static here.
//
// All of the reference's quirks are kept because they shape the converged image:
// world-space Oren-Nayar angles, GTR2 normalised by 2*Pi with the alpha^2 < 5e-5 -> 1 shortcut, the unweighted
// 0.5*(1/(2 Pi) + pdf_glossy) layered pdf, the uniform (not cosine) hemisphere sampling of the diffuse lobe, and
// w_in left at zero when a glossy half-vector falls below the surface.
//
// Double-precision leakage.  The reference's unqualified sin/cos bind to the C double functions, so the sampled
// direction is  float(cos_double(a) * b)  (PRNG.hpp:42-45, Materials.hpp:163-168).  The GGX lobe is so peaked for
// small roughness that a 1-ulp change of wi moves D(h.N) by ~1e-4 relative, so the 1e-5 material parity budget can
// only be met if wi itself matches to the bit: the device therefore evaluates sincos in double here too and rounds
// the double product once, like the reference.  The Oren-Nayar sines (Materials.hpp:227) only scale the small b_ term
// and stay in float.
#pragma once

#include "device_types.h"
#include "exact_arith.cuh"
#include "math.cuh"

namespace cornelis_b200 {

struct RGBf {
    float r, g, b;
};

CB_HD RGBf operator*(RGBf a, float s) { return RGBf{a.r * s, a.g * s, a.b * s}; }
CB_HD RGBf operator+(RGBf a, RGBf b) { return RGBf{a.r + b.r, a.g + b.g, a.b + b.b}; }

constexpr float kHemispherePdf = 1.0f / (2.0f * kPi); // randomHemispherePDF(), PRNG.hpp:62

// ---- arithmetic classes -----------------------------------------------------------------------------------------
// EXACT chain: everything between the random numbers and base = 1 + (alpha^2 - 1) (h.N)^2 — the sampled direction
// wi, the half vector h, its cosine — uses IEEE division / square root in the reference's operation order.  The GGX
// term amplifies a 1-ulp change of h.N into ~1e-4 of D for the reference's default roughness, so these values must
// match the reference to the bit (tests/test_gpu_parity.py checks wi bit for bit).
// TOLERANT tail: values that are only multiplied into the result (the reciprocal of base^2, Smith G, Fresnel,
// tan = sin/cos, the Oren-Nayar factor, the final quotients) may be off by a couple of ulp; they use the
// single-instruction MUFU approximations (<= 2 ulp each), keeping f and pdf within ~3e-6 of the reference against a
// budget of 1e-5.  The .ftz forms: MUFU itself flushes subnormals, and without .ftz the compiler wraps every use in a
// range test and two rescalings (seven instructions per reciprocal, four per square root — 4 % of the render kernel's
// instructions, profiles/r2_ncu).  No operand of the tail is subnormal or above 2^126: the reciprocals are taken of
// base^2 in [alpha^4, 1], of sums >= 1, of cosines the reference has already compared with its 5e-5 epsilon, of
// pdf * prob >= 4e-3; the square roots of 1 - c^2 (0 or >= 2^-24) and of 1 + k^2.
__device__ __forceinline__ float approxSqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float approxRcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float approxDiv(float a, float b) { return a * approxRcp(b); } // (__fdividef adds range scaling)

// Contractions in the TOLERANT tail are written out as explicit fused multiply-adds: the translation unit is compiled
// with --fmad=false for the exact chains, and one rounding instead of two is within the tail's budget.
CB_HD float fma1(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return a * b + c;
#endif
}

// x^5 for Schlick's (1 - cos)^5 (Materials.cpp:41 uses std::pow(x, 5.0f)): exact-ish in double, rounded once, which
// equals glibc's powf except for its ~7e-4 fraction of not-correctly-rounded results.
CB_HD float pow5(float x) {
    double d = static_cast<double>(x);
    double d2 = d * d;
    return static_cast<float>(d2 * d2 * d);
}

// models::schlick(cos_theta, 1.0f, ior), Materials.cpp:38-42, with R0 precomputed per material.  EXACT: the layered
// BRDF weights the diffuse lobe with 1 - F(N.wi) (Materials.hpp:261), which cancels for grazing wi — one ulp of F is
// 6e-8 / (1 - F) of the weight.
CB_HD float schlick(float cosTheta, float r0) { return r0 + (1.0f - r0) * pow5(1.0f - cosTheta); }

// The same for the Fresnel factor of the half vector, which only multiplies the glossy term: tolerant (three float
// products, one fused multiply-add; within 3 ulp).
CB_HD float schlickApprox(float cosTheta, float r0) {
    float const x = 1.0f - cosTheta, x2 = x * x;
    return fma1(1.0f - r0, x2 * x2 * x, r0);
}

// models::distributionGTR2, Materials.cpp:16-26 (std::pow(x, 2.0f) is folded to x*x by the reference's compiler).
// The argument chain up to `base` is exact; the reciprocal is tolerant.
__device__ __forceinline__ float distributionGTR2(float cosThetaH, const DevMaterial &m) {
    if (isAlmostZero(m.alpha2))
        return 1.0f;
    float c2 = cosThetaH * cosThetaH;
    float base = 1.0f + (m.alpha2 - 1.0f) * c2;
    return m.gtr_a * approxRcp(base * base);
}

// models::lambdaTR, Materials.cpp:28-32.
__device__ __forceinline__ float lambdaTR(float tanTheta, float alpha) {
    if (isinf(tanTheta))
        return 0.0f;
    float k = fabsf(tanTheta) * alpha;
    return fma1(approxSqrt(fma1(k, k, 1.0f)), 0.5f, -0.5f);
}

// models::shadowMaskingTR, Materials.cpp:34-36.
__device__ __forceinline__ float shadowMaskingTR(float tanI, float tanO, float alpha) {
    return approxRcp(1.0f + lambdaTR(tanI, alpha) + lambdaTR(tanO, alpha));
}

// sqrtf(x) for x = 1 - c^2 with |c| <= 1, i.e. x == 0 or 2^-24 <= x <= 1: the exact fast sequence, straight-line (the
// compiler's sqrtf carries a range test and a branch to an out-of-line call).  x < 0 or NaN (|c| > 1, NaN) gives NaN
// like sqrtf; so would x == 0 (0 * rsqrt(0)), hence the select.
__device__ __forceinline__ float sqrtUnitExact(float x) {
    float const r = sqrtExactFast(x);
    return x == 0.0f ? 0.0f : r;
}
// a / s for s = sqrtUnitExact(..) in [2^-12, 1] and |a| <= 1: RN(a / s) by the exact fast sequence for |a| >= 2^-80
// and for a == 0.  A smaller non-zero a (a direction component below 1e-24) may come out one ulp off: its only use is
// r * r and r * r' in cosD, where it vanishes against 1.  s == 0 or NaN gives NaN where the operator gives +-INF or
// NaN; both make cosD NaN (INF * 0 and 1 - INF^2 under the root), which std::max(0, NaN) turns into 0.
__device__ __forceinline__ float divideByUnitExact(float a, float s) { return divideExactFast0(a, s, rcpSeedRefined(s)); }

// OrenNayarBRDF::operator(), Materials.hpp:211-228, without its seven transcendental calls.
//
// The reference takes the angles from WORLD-space components (cos theta = w.z, cos phi = w.x / sin theta — not
// relative to N), then evaluates  a + b * max(0, cos(phiI - phiO)) * sin(alpha) * sin(beta)  with
// phi = acos(r), theta = acos(w.z), alpha/beta = max/min(thetaI, thetaO).  Since acos maps into [0, pi]:
//   cos(phiI - phiO)         = rI rO + sqrt(1 - rI^2) sqrt(1 - rO^2)
//   sin(alpha) sin(beta)     = sin(thetaI) sin(thetaO) = sqrt(1 - cI^2) sqrt(1 - cO^2)
// which differ from the acos/cos/sin route by ~1e-6 absolute, scaled by b (<= 0.33) against a (>= 0.79).
// NaN behaviour is kept: a NaN azimuth is swallowed by std::max(0, NaN) == 0 in both forms; |wi.z| > 1 makes the
// factor NaN in both; |wo.z| > 1 alone makes std::max/std::min pick thetaI twice (Materials.hpp:223-224 with a NaN
// second argument), i.e. sin^2(thetaI).
__device__ __forceinline__ RGBf orenNayarEval(const DevMaterial &m, V3 wi, V3 wo) {
    float cI = wi.z, cO = wo.z;
    // exact up to r: whether |r| exceeds 1 (-> NaN azimuth -> the b-term vanishes) is a discontinuity of size ~b
    float sI = sqrtUnitExact(1.0f - cI * cI);
    float sO = sqrtUnitExact(1.0f - cO * cO);
    float rI = divideByUnitExact(wi.x, sI);
    float rO = divideByUnitExact(wo.x, sO);
    // (the sign of 1 - r^2, hence the NaN that switches the b-term off, is the same fused or not: r^2 > 1 iff |r| > 1)
    float cosD = fma1(approxSqrt(fma1(-rI, rI, 1.0f)), approxSqrt(fma1(-rO, rO, 1.0f)), rI * rO);
    bool const thetaONaN = !(fabsf(cO) <= 1.0f);
    float sinProduct = sI * (thetaONaN ? sI : sO);
    float s = fma1(m.on_b * stdMax(0.0f, cosD), sinProduct, m.on_a);
    return RGBf{m.dr, m.dg, m.db} * s;
}

// GlossyBRDF::operator() (Materials.hpp:130-154) and GlossyBRDF::pdf (Materials.hpp:177-188) share the half vector
// h = normalize(wi + wo), its cosine and D; evaluated together.  Returns the scalar that multiplies the tint.
__device__ __forceinline__ float glossyEvalPdf(const DevMaterial &m, V3 wi, V3 wo, V3 N, float &pdf, OddWatch *odd = nullptr) {
    V3 const h = normalize(wi + wo, odd);                  // exact chain
    float const cosThetaH = stdMax(0.0f, dot(h, N));
    float D = 1.0f;
    if (isAlmostZero(cosThetaH)) {
        pdf = 1.0f;                                        // Materials.hpp:180-181
    } else {
        D = distributionGTR2(cosThetaH, m);
        float const pdfh = D * fabsf(cosThetaH);
        float const wiDotH = dot(wi, h);
        pdf = isAlmostZero(wiDotH) ? pdfh : approxDiv(pdfh, 4.0f * wiDotH);
    }
    float const cosThetaO = stdMax(0.0f, dot(wo, N));
    float const cosThetaI = stdMax(0.0f, dot(wi, N));
    if (isAlmostZero(cosThetaO) || isAlmostZero(cosThetaI))  // Materials.hpp:141-142
        return 0.0f;
    if (isAlmostZero(h.x) && isAlmostZero(h.y) && isAlmostZero(h.z))
        return 0.0f;
    if (isAlmostZero(cosThetaH))
        D = distributionGTR2(cosThetaH, m);                // eval has no shortcut for a grazing half vector
    float const sinThetaO = approxSqrt(fma1(-cosThetaO, cosThetaO, 1.0f));
    float const sinThetaI = approxSqrt(fma1(-cosThetaI, cosThetaI, 1.0f));
    float const G = shadowMaskingTR(approxDiv(sinThetaI, cosThetaI), approxDiv(sinThetaO, cosThetaO), m.alpha);
    float const F = schlickApprox(cosThetaH, m.r0);
    return approxDiv(F * D * G, 4.0f * cosThetaO * cosThetaI);
}

// LayeredBRDF::operator() and ::pdf (Materials.hpp:255-277): f = (1 - F(N.wi)) * diffuse + glossy;
// pdf = 0.5 * (1/(2 Pi) + pdf_glossy) — the unweighted average whatever lobe was sampled.
__device__ __forceinline__ RGBf layeredEvalPdf(const DevMaterial &m, V3 wi, V3 wo, V3 N, float &pdf, OddWatch *odd = nullptr) {
    float pdfGlossy;
    float const g = glossyEvalPdf(m, wi, wo, N, pdfGlossy, odd);
    pdf = fma1(0.5f, pdfGlossy, 0.5f * kHemispherePdf);
    RGBf const Df = orenNayarEval(m, wi, wo);
    RGBf const Gf = RGBf{m.tr, m.tg, m.tb} * g;
    float const k = 1.0f - schlick(stdMax(0.0f, dot(N, wi)), m.r0);
    return RGBf{fma1(Df.r, k, Gf.r), fma1(Df.g, k, Gf.g), fma1(Df.b, k, Gf.b)};
}

static __constant__ double kSinCos[16] = {
    6.36619772367581382433e-01,  // 0: 2 / Pi
    1.57079632679489655800e+00,  // 1: Pi / 2, high part
    6.12323399573676603587e-17,  // 2: Pi / 2, low part
    6755399441055744.0,          // 3: 1.5 * 2^52
    1.58969099521155010221e-10,  -2.50507602534068634195e-08, 2.75573137070700676789e-06,  // 4..9: k_sin.c S6..S1
    -1.98412698298579493134e-04, 8.33333333332248946124e-03,  -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09,  -2.75573143513906633035e-07, // 10..15: k_cos.c C6..C1
    2.48015872894767294178e-05,  -1.38888888888741095749e-03, 4.16666666666666019037e-02,
};

// sin and cos of a double in [0, 2 Pi] — the only arguments the direction sampling has (angle = float(2 Pi x), x in
// [0, 1)) — to within one ulp: Cody-Waite reduction by multiples of Pi / 2 with two fused steps (the quadrant is at most
// 4, so the products are exact), then the classic degree-13 / degree-14 minimax kernels on [-Pi/4, Pi/4] (the
// coefficients of fdlibm's k_sin.c / k_cos.c).  The CUDA library's sincos(double) computes the same thing behind a
// range check whose slow path (huge arguments) passes its result through local memory: a store and a load executed
// on every call, plus the call set-up.  The reference (glibc) and either device version agree to the last bit or
// differ by one ulp of a DOUBLE; what the render uses is float(cos * radial), which a restatement of this function in C
// reproduced for all 2^24 possible angles x 8 radial values without one mismatch against glibc (and
// tests/test_gpu_parity.py::test_bsdf_sample_and_eval requires the sampled direction bit for bit on 2^18 inputs).
__device__ __forceinline__ void sincosFirstTurn(double a, double &sn, double &cs) {
    // (the constants live in constant memory, kSinCos: a 64-bit immediate costs two UMOVs per use, a constant-bank
    // operand none)
    const double *const k = kSinCos;
    double const shifted = fma(a, k[0], k[3]); // the quadrant, rounded to nearest, sits in the low word
    int const quadrant = __double2loint(shifted);
    double const j = shifted - k[3];
    double r = fma(-j, k[1], a);
    r = fma(-j, k[2], r);
    double const z = r * r;
    double ps = fma(z, k[4], k[5]);
    ps = fma(z, ps, k[6]);
    ps = fma(z, ps, k[7]);
    ps = fma(z, ps, k[8]);
    ps = fma(z, ps, k[9]);
    double const s = fma(z * r, ps, r);
    double pc = fma(z, k[10], k[11]);
    pc = fma(z, pc, k[12]);
    pc = fma(z, pc, k[13]);
    pc = fma(z, pc, k[14]);
    pc = fma(z, pc, k[15]);
    double const c = fma(z * z, pc, fma(-0.5, z, 1.0));
    double const S = (quadrant & 1) ? c : s, C = (quadrant & 1) ? s : c;
    sn = (quadrant & 2) ? -S : S;
    cs = ((quadrant + 1) & 2) ? -C : C;
}

// LayeredBRDF::generateDirection, Materials.hpp:279-293: x2 < 0.5 samples the diffuse lobe UNIFORMLY over the
// hemisphere (BRDF::generateDirection -> randomHemisphere, PRNG.hpp:39-55), otherwise the GGX half vector
// (GlossyBRDF::generateDirection, Materials.hpp:156-175).  Both lobes build
//     float(cos_d(angle) * radial) * B + float(sin_d(angle) * radial) * T + axial * N
// with the sine/cosine evaluated in DOUBLE and the double product rounded once — as the reference's unqualified
// cos()/sin() do — so one sincos serves whichever lobe a lane picked:
//     diffuse: angle = 2 Pi x1, radial = sqrt(1 - x0^2),         axial = x0
//     glossy:  angle = 2 Pi x0, radial = sin(theta_h),           axial = cos(theta_h) = sqrt((1-x1)/(1+(a^2-1)x1))
// The lobe's own f/pdf are discarded by the reference (Materials.hpp:281-289); pdf and f come from the layered
// functions at the sampled wi.  If the half vector falls below the surface wi stays 0 (Materials.hpp:169-170).
// `odd`: see normalize (math.cuh).
__device__ __forceinline__ RGBf layeredSample(const DevMaterial &m, V3 wo, float x0, float x1, float x2,
                                              const Basis &b, V3 &wi, float &pdf, OddWatch *odd = nullptr) {
    bool const diffuse = x2 < 0.5f;
    float radial, axial;
    if (diffuse) {
        radial = sqrtf(1.0f - x0 * x0);
        axial = x0;
    } else {
        float const A = 1.0f - x1;
        float const B = 1.0f + (m.alpha2 - 1.0f) * x1;
        axial = sqrtf(A / B);
        radial = sqrtExact(1.0f - axial * axial); // == sqrtf; near-mirror lobes (gold: alpha^2 = 1e-8) give sqrt(0) here
                                                  // every time, which sqrtf answers through an out-of-line call
    }
    float const angle = 2.0f * kPi * (diffuse ? x1 : x0); // == float(2.0 * Pi * x): the 48-bit product is exact in double
    double sn, cs;
    sincosFirstTurn(static_cast<double>(angle), sn, cs);
    float const kB = static_cast<float>(cs * static_cast<double>(radial));
    float const kT = static_cast<float>(sn * static_cast<double>(radial));
    V3 const v = (kB * b.B + kT * b.T) + axial * b.N;
    wi = v;
    if (!diffuse) {
        wi = V3{0.0f, 0.0f, 0.0f};
        V3 const h = normalize(v, odd);
        if (!(dot(h, b.N) < 0.0f))
            wi = normalize((2.0f * dot(wo, h)) * h - wo, odd);
    }
    return layeredEvalPdf(m, wi, wo, b.N, pdf, odd);
}

// russianRouletteFactor, Render.cpp:153-165.
CB_HD float russianRouletteFactor(float tr, float tg, float tb, uint32_t depth) {
    constexpr float Base = 0.55f;
    // (one select instead of an early return: nearly every warp holds a path at depth >= 3, so a branch saves nothing)
    float const power = stdClamp(tr * tr + tg * tg + tb * tb, 0.05f / Base, 0.99f);
    float const deep = Base * power;
    return depth < 3 ? 0.99f : deep;
}

// Host-side construction of the per-material constants, with the reference's constructor arithmetic
// (Scene.cpp:46-52 -> Materials.hpp:327-329, 251-253, 206-209, 296-302; Materials.cpp:17-24, 39-40).
inline DevMaterial makeDevMaterial(const float albedo[3], const float emissive[3], float roughness,
                                   const float tint[3], float ior) {
    DevMaterial m{};
    m.er = emissive[0];
    m.eg = emissive[1];
    m.eb = emissive[2];
    float glossyRough = roughness * roughness;
    float sigma = fabsf(0.5f * glossyRough);
    float sigma2 = sigma * sigma;
    m.on_a = 1.0f - (sigma2 / (2.0f * (sigma2 + 0.333f)));
    m.on_b = 0.45f * sigma2 / (sigma2 + 0.09f);
    m.dr = albedo[0] / kPi;
    m.dg = albedo[1] / kPi;
    m.db = albedo[2] / kPi;
    m.tr = tint[0];
    m.tg = tint[1];
    m.tb = tint[2];
    m.alpha = glossyRough;
    m.alpha2 = glossyRough * glossyRough;
    m.ior = ior;
    float r0 = (1.0f - ior) / (1.0f + ior);
    m.r0 = r0 * r0;
    m.gtr_a = m.alpha2 / (2.0f * kPi);
    return m;
}

} // namespace cornelis_b200
