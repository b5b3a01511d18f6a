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
// the double product once, like the reference.  (1 - cos)^5 is formed in double and rounded once, which equals
// glibc's powf except for its ~7e-4 fraction of not-correctly-rounded results.  The Oren-Nayar sines
// (Materials.hpp:227) only scale the small b_ term and stay in float.
#pragma once

#include "device_types.h"
#include "math.cuh"

namespace cornelis_b200 {

struct RGBf {
    float r, g, b;
};

CB_HD RGBf operator*(RGBf a, float s) { return RGBf{a.r * s, a.g * s, a.b * s}; }
CB_HD RGBf operator+(RGBf a, RGBf b) { return RGBf{a.r + b.r, a.g + b.g, a.b + b.b}; }

constexpr float kHemispherePdf = 1.0f / (2.0f * kPi); // randomHemispherePDF(), PRNG.hpp:62

// x^5 for Schlick's (1 - cos)^5 (Materials.cpp:41 uses std::pow(x, 5.0f)): exact-ish in double, rounded once.
CB_HD float pow5(float x) {
    double d = static_cast<double>(x);
    double d2 = d * d;
    return static_cast<float>(d2 * d2 * d);
}

// models::schlick(cos_theta, 1.0f, ior), Materials.cpp:38-42, with R0 precomputed per material.
CB_HD float schlick(float cosTheta, float r0) { return r0 + (1.0f - r0) * pow5(1.0f - cosTheta); }

// models::distributionGTR2, Materials.cpp:16-26 (std::pow(x, 2.0f) is folded to x*x by the reference's compiler).
CB_HD float distributionGTR2(float cosThetaH, const DevMaterial &m) {
    if (isAlmostZero(m.alpha2))
        return 1.0f;
    float c2 = cosThetaH * cosThetaH;
    float base = 1.0f + (m.alpha2 - 1.0f) * c2;
    float B = 1.0f / (base * base);
    return m.gtr_a * B;
}

// models::lambdaTR, Materials.cpp:28-32.
CB_HD float lambdaTR(float tanTheta, float alpha) {
    if (isinf(tanTheta))
        return 0.0f;
    float k = fabsf(tanTheta) * alpha;
    return (-1.0f + sqrtf(1.0f + k * k)) * 0.5f;
}

// models::shadowMaskingTR, Materials.cpp:34-36.
CB_HD float shadowMaskingTR(float tanI, float tanO, float alpha) {
    return 1.0f / (1.0f + lambdaTR(tanI, alpha) + lambdaTR(tanO, alpha));
}

// OrenNayarBRDF::operator(), Materials.hpp:211-228.  Angles come from WORLD-space components (wi.z, wi.x), not
// relative to N; std::max(0, NaN) == 0 swallows the NaN azimuth of a vertical direction.
__device__ __forceinline__ RGBf orenNayarEval(const DevMaterial &m, V3 wi, V3 wo) {
    float cosThetaI = wi.z;
    float cosThetaO = wo.z;
    float sinThetaI = sqrtf(1.0f - cosThetaI * cosThetaI);
    float sinThetaO = sqrtf(1.0f - cosThetaO * cosThetaO);
    float phiI = acosf(wi.x / sinThetaI);
    float phiO = acosf(wo.x / sinThetaO);
    float thetaO = acosf(cosThetaO);
    float thetaI = acosf(cosThetaI);
    float alpha = stdMax(thetaI, thetaO);
    float beta = stdMin(thetaI, thetaO);
    float s = m.on_a + m.on_b * stdMax(0.0f, cosf(phiI - phiO)) * sinf(alpha) * sinf(beta);
    return RGBf{m.dr, m.dg, m.db} * s;
}

// GlossyBRDF::operator(), Materials.hpp:130-154.
__device__ __forceinline__ RGBf glossyEval(const DevMaterial &m, V3 wi, V3 wo, V3 N) {
    float cosThetaO = stdMax(0.0f, dot(wo, N));
    float sinThetaO = sqrtf(1.0f - cosThetaO * cosThetaO);
    float cosThetaI = stdMax(0.0f, dot(wi, N));
    float sinThetaI = sqrtf(1.0f - cosThetaI * cosThetaI);
    if (isAlmostZero(cosThetaO) || isAlmostZero(cosThetaI))
        return RGBf{0.0f, 0.0f, 0.0f};
    V3 h = normalize(wi + wo);
    if (isAlmostZero(h.x) && isAlmostZero(h.y) && isAlmostZero(h.z))
        return RGBf{0.0f, 0.0f, 0.0f};
    float cosThetaH = stdMax(0.0f, dot(h, N));
    float D = distributionGTR2(cosThetaH, m);
    float G = shadowMaskingTR(sinThetaI / cosThetaI, sinThetaO / cosThetaO, m.alpha);
    float F = schlick(cosThetaH, m.r0);
    return RGBf{m.tr, m.tg, m.tb} * (F * D * G / (4.0f * cosThetaO * cosThetaI));
}

// GlossyBRDF::pdf, Materials.hpp:177-188.
__device__ __forceinline__ float glossyPdf(const DevMaterial &m, V3 wi, V3 wo, V3 N) {
    V3 h = normalize(wi + wo);
    float cosThetaH = stdMax(0.0f, dot(h, N));
    if (isAlmostZero(cosThetaH))
        return 1.0f;
    float D = distributionGTR2(cosThetaH, m);
    float pdfh = D * fabsf(cosThetaH);
    float wiDotH = dot(wi, h);
    if (isAlmostZero(wiDotH))
        return pdfh;
    return pdfh / (4.0f * wiDotH);
}

// randomHemisphere(float2, Basis), PRNG.hpp:39-55 — UNIFORM over the hemisphere.
__device__ __forceinline__ V3 sampleHemisphere(float x1, float x2, const Basis &b) {
    float a = 2.0f * kPi * x2; // the reference's double product rounds to the same float (exact 48-bit product)
    float r = sqrtf(1.0f - x1 * x1);
    double sa, ca;
    sincos(static_cast<double>(a), &sa, &ca);
    V3 v{static_cast<float>(ca * static_cast<double>(r)), static_cast<float>(sa * static_cast<double>(r)), x1};
    return b.B * v.x + b.T * v.y + b.N * v.z;
}

// GlossyBRDF::generateDirection, Materials.hpp:156-175 — only wi matters to the caller (Materials.hpp:281-289);
// on the early-out wi keeps the zero it was initialised with (Render.cpp:198).
__device__ __forceinline__ void sampleGlossy(const DevMaterial &m, V3 wo, float x0, float x1, const Basis &b, V3 &wi) {
    float A = 1.0f - x1;
    float B = 1.0f + (m.alpha2 - 1.0f) * x1;
    float cosThetaH = sqrtf(A / B);
    float sinThetaH = sqrtf(1.0f - cosThetaH * cosThetaH);
    float phi = 2.0f * kPi * x0;
    double sp, cp;
    sincos(static_cast<double>(phi), &sp, &cp);
    float const kB = static_cast<float>(static_cast<double>(sinThetaH) * cp);
    float const kT = static_cast<float>(static_cast<double>(sinThetaH) * sp);
    V3 h = normalize(kB * b.B + kT * b.T + cosThetaH * b.N);
    if (dot(h, b.N) < 0.0f)
        return;
    wi = normalize((2.0f * dot(wo, h)) * h - wo);
}

// LayeredBRDF::operator(), Materials.hpp:255-263.
__device__ __forceinline__ RGBf layeredEval(const DevMaterial &m, V3 wi, V3 wo, V3 N) {
    RGBf Df = orenNayarEval(m, wi, wo);
    RGBf Gf = glossyEval(m, wi, wo, N);
    float k = 1.0f - schlick(stdMax(0.0f, dot(N, wi)), m.r0);
    return Df * k + Gf;
}

// LayeredBRDF::pdf, Materials.hpp:265-277 — the unweighted average whatever lobe was sampled.
__device__ __forceinline__ float layeredPdf(const DevMaterial &m, V3 wi, V3 wo, V3 N) {
    return 0.5f * (kHemispherePdf + glossyPdf(m, wi, wo, N));
}

// LayeredBRDF::generateDirection, Materials.hpp:279-293.  x2 picks the lobe (its rescaled value is unused).
__device__ __forceinline__ RGBf layeredSample(const DevMaterial &m, V3 wo, float x0, float x1, float x2,
                                              const Basis &b, V3 &wi, float &pdf) {
    wi = V3{0.0f, 0.0f, 0.0f};
    if (x2 < 0.5f)
        wi = sampleHemisphere(x0, x1, b);
    else
        sampleGlossy(m, wo, x0, x1, b, wi);
    pdf = layeredPdf(m, wi, wo, b.N);
    return layeredEval(m, wi, wo, b.N);
}

// russianRouletteFactor, Render.cpp:153-165.
CB_HD float russianRouletteFactor(float tr, float tg, float tb, uint32_t depth) {
    constexpr float Base = 0.55f;
    if (depth < 3)
        return 0.99f;
    float power = stdClamp(tr * tr + tg * tg + tb * tb, 0.05f / Base, 0.99f);
    return Base * power;
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
