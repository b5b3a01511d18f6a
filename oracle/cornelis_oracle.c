/* TEST INFRASTRUCTURE — not part of the product.
 *
 * Plain-C restatement of the reference's batched render loop (skurmedel/cornelis), implementing
 * oracle/oracle_api.h.  It exists so that the checker can be rebuilt anywhere gcc exists (the GPU box has no
 * /root/reference) and so that every arithmetic step the CUDA path must reproduce is written down once, with the
 * reference file:line it follows.  Parity is PINNED: tests/test_oracle.py checks this file bit-for-bit against
 * the reference compiled from its own sources (oracle/_ref, built by oracle/build_ref.sh) on intersections,
 * camera rays, BSDF sampling/evaluation, shading steps, PRNG streams and whole renders, and against the
 * reference's own known-answer tests (tests/test_Geometry.cpp, test_Camera.cpp, test_Math.cpp, test_Tiles.cpp,
 * test_FrameBuffer.cpp, test_Color.cpp) and the committed golden vectors under tests/golden/.
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no -march, no -ffast-math) — the same floating-point regime as
 * oracle/_ref: IEEE binary32 operations, no fused multiply-add, glibc libm.
 *
 * Double-precision leakage in the reference is reproduced where it exists: unqualified sqrt/sin/cos inside
 * namespace cornelis bind to the C double functions (checked with nm on the reference objects: sin, sincos,
 * acosf, cosf, powf, pow are the only libm symbols), `2.0 * Pi * x` is a double product.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "oracle_api.h"

/* ---------------------------------------------------------------------------------------------- Math.hpp */

typedef struct { float x, y, z; } f3;

static const float RayEpsilon = 0.00005f;      /* Math.hpp:20 */
static const float Pi = 3.14159265359f;        /* Math.hpp:25 */

static inline int isAlmostZero(float v) { return fabsf(v) < RayEpsilon; } /* Math.hpp:22 */

/* std::max / std::min / std::clamp semantics (they differ from fmaxf/fminf on NaN). */
static inline float std_max(float a, float b) { return (a < b) ? b : a; }
static inline float std_min(float a, float b) { return (b < a) ? b : a; }
static inline float std_clamp(float v, float lo, float hi) { return (v < lo) ? lo : (hi < v) ? hi : v; }

static inline f3 mk(float x, float y, float z) { f3 r = {x, y, z}; return r; }
static inline f3 add(f3 a, f3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }          /* Math.hpp:63-70 */
static inline f3 sub(f3 a, f3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }          /* Math.hpp:75-82 */
static inline f3 neg(f3 a) { return mk(-a.x, -a.y, -a.z); }                               /* Math.hpp:86-93 */
static inline f3 mul(f3 a, f3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }          /* Math.hpp:98-105 */
static inline f3 scale(f3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }           /* Math.hpp:110-128 */
static inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }         /* Math.hpp:278 */
static inline float mag2(f3 a) { return dot(a, a); }                                      /* Math.hpp:284 */
static inline f3 rayT(f3 o, f3 d, float t) { return add(o, mul(d, mk(t, t, t))); }        /* Math.hpp:290-292 */

static inline f3 cross(f3 a, f3 b) { /* Math.hpp:380-384 */
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

static inline f3 normalize(f3 v) { /* Math.hpp:392-398 */
    float len = sqrtf(mag2(v));
    if (isAlmostZero(len))
        return mk(0.0f, 0.0f, 0.0f);
    float s = 1.0f / len;
    return mul(v, mk(s, s, s));
}

typedef struct { f3 N, T, B; } Basis;

static inline Basis constructBasis(f3 N) { /* Math.hpp:424-434 */
    f3 helper = mk(0.0f, 1.0f, 0.0f);
    if ((double)fabsf(N.y) > 0.95) /* float abs compared against a double literal */
        helper = mk(0.0f, 0.0f, 1.0f);
    Basis b;
    b.N = N;
    b.T = normalize(cross(helper, N));
    b.B = cross(b.T, N);
    return b;
}

/* ---------------------------------------------------------------------------------------------- PRNG.hpp */

typedef struct { uint32_t s[4]; } Xoshiro128Plus;

static uint64_t splitmix64(uint64_t *state) { /* XoshiroCpp.hpp:684-690 */
    uint64_t z = (*state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

static void prng_seed(Xoshiro128Plus *g, uint64_t seed) { /* XoshiroCpp.hpp:1305-1314 */
    uint64_t sm = seed;
    for (int i = 0; i < 4; i++)
        g->s[i] = (uint32_t)splitmix64(&sm);
}

static inline uint32_t rotl32(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }

static inline uint32_t prng_bits(Xoshiro128Plus *g) { /* XoshiroCpp.hpp:1319-1331 */
    uint32_t const result = g->s[0] + g->s[3];
    uint32_t const t = g->s[1] << 9;
    g->s[2] ^= g->s[0];
    g->s[3] ^= g->s[1];
    g->s[1] ^= g->s[2];
    g->s[0] ^= g->s[3];
    g->s[2] ^= t;
    g->s[3] = rotl32(g->s[3], 11);
    return result;
}

static void prng_jump(Xoshiro128Plus *g) { /* XoshiroCpp.hpp:1333-1360 */
    static const uint32_t JUMP[] = {0x8764000b, 0xf542d2d3, 0x6fa035c3, 0x77f2db5b};
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int j = 0; j < 4; j++)
        for (int b = 0; b < 32; b++) {
            if (JUMP[j] & (UINT32_C(1) << b)) {
                s0 ^= g->s[0];
                s1 ^= g->s[1];
                s2 ^= g->s[2];
                s3 ^= g->s[3];
            }
            prng_bits(g);
        }
    g->s[0] = s0;
    g->s[1] = s1;
    g->s[2] = s2;
    g->s[3] = s3;
}

static inline float prng_next(Xoshiro128Plus *g) { /* PRNG.hpp:20, XoshiroCpp.hpp:651-655 */
    return (float)(prng_bits(g) >> 8) * 0x1.0p-24f;
}

static Xoshiro128Plus cloneForThread(const Xoshiro128Plus *g, size_t k) { /* PRNG.hpp:32-37 */
    Xoshiro128Plus c = *g;
    for (size_t i = 0; i < k; i++)
        prng_jump(&c);
    return c;
}

static inline float randomHemispherePDF(void) { return 1.0f / (2.0f * Pi); } /* PRNG.hpp:62 */

static inline f3 randomHemisphereLocal(float x1, float x2) { /* PRNG.hpp:39-46 */
    float a = (float)(2.0 * Pi * x2);          /* double product, rounded once */
    float b = (float)sqrt(1.0f - x1 * x1);     /* double sqrt of a float expression */
    double ca = cos(a), sa = sin(a);            /* double cos/sin of the float angle */
    return mk((float)(ca * b), (float)(sa * b), x1);
}

static inline f3 randomHemisphere(float x1, float x2, const Basis *base) { /* PRNG.hpp:52-55 */
    f3 v = randomHemisphereLocal(x1, x2);
    return add(add(scale(base->B, v.x), scale(base->T, v.y)), scale(base->N, v.z));
}

/* ------------------------------------------------------------------------------------------ Materials.cpp */

static float distributionGTR2(float cos_theta_H, float alpha) { /* Materials.cpp:16-26 */
    float alpha2 = alpha * alpha;
    float c2 = cos_theta_H * cos_theta_H;
    if (isAlmostZero(alpha2))
        return 1.0f;
    float A = alpha2 / (2.0f * Pi);
    float B = 1.0f / powf(1.0f + (alpha2 - 1.0f) * c2, 2.0f);
    return A * B;
}

static float lambdaTR(float tan_theta, float alpha) { /* Materials.cpp:28-32 */
    if (isinf(tan_theta))
        return 0.0f;
    return (-1.0f + sqrtf(1.0f + (fabsf(tan_theta) * alpha) * (fabsf(tan_theta) * alpha))) * 0.5f;
}

static float shadowMaskingTR(float tan_i, float tan_o, float alpha) { /* Materials.cpp:34-36 */
    return 1.0f / (1.0f + lambdaTR(tan_i, alpha) + lambdaTR(tan_o, alpha));
}

static float schlick(float cos_theta, float n1, float n2) { /* Materials.cpp:38-42 */
    float R0 = (n1 - n2) / (n1 + n2);
    R0 *= R0;
    return R0 + (1.0f - R0) * powf(1.0f - cos_theta, 5.0f);
}

/* ------------------------------------------------------------------------------------------ Materials.hpp */

typedef struct {
    f3 emission;       /* StandardMaterial::emission_  Materials.hpp:336 */
    f3 albedo;         /* OrenNayarBRDF::albedo_ */
    float on_a, on_b;  /* OrenNayarBRDF::a_, b_        Materials.hpp:206-209 */
    f3 tint;           /* GlossyBRDF::tint_ */
    float alpha;       /* GlossyBRDF::alpha_ = roughness^2   Materials.hpp:296-299 */
    float ior;         /* GlossyBRDF::refidx_ */
} Material;

static Material makeMaterial(const float *p) { /* Scene.cpp:46-52 -> Materials.hpp:327-329, 251-253, 206-209 */
    Material m;
    m.albedo = mk(p[0], p[1], p[2]);
    m.emission = mk(p[3], p[4], p[5]);
    float perceptual = p[6];
    m.tint = mk(p[7], p[8], p[9]);
    m.ior = p[10];
    float glossyRough = perceptual * perceptual;       /* Materials.hpp:296-299 */
    float sigma = fabsf(0.5f * glossyRough);           /* Materials.hpp:300-302 */
    float sigma2 = sigma * sigma;
    m.on_a = 1.0f - (sigma2 / (2.0f * (sigma2 + 0.333f)));
    m.on_b = 0.45f * sigma2 / (sigma2 + 0.09f);
    m.alpha = glossyRough;
    return m;
}

static f3 orenNayarEval(const Material *m, f3 wi, f3 wo) { /* Materials.hpp:211-228 (world-space angles!) */
    float cosThetaI = wi.z;
    float cosThetaO = wo.z;
    float sinThetaI = sqrtf(1.0f - cosThetaI * cosThetaI);
    float sinThetaO = sqrtf(1.0f - cosThetaO * cosThetaO);
    float phiI = acosf(wi.x / sinThetaI);
    float phiO = acosf(wo.x / sinThetaO);
    float thetaO = acosf(cosThetaO);
    float thetaI = acosf(cosThetaI);
    float alpha = std_max(thetaI, thetaO);
    float beta = std_min(thetaI, thetaO);
    /* float * double * double, + float: evaluated in double, converted to float when it multiplies the RGB */
    double s = m->on_a + m->on_b * std_max(0.0f, cosf(phiI - phiO)) * sin(alpha) * sin(beta);
    f3 base = mk(m->albedo.x / Pi, m->albedo.y / Pi, m->albedo.z / Pi); /* Color.cpp:11-17 */
    return scale(base, (float)s);
}

static f3 glossyEval(const Material *m, f3 wi, f3 wo, f3 N) { /* Materials.hpp:130-154 */
    float cos_thetaO = std_max(0.0f, dot(wo, N));
    float sin_thetaO = sqrtf(1.0f - cos_thetaO * cos_thetaO);
    float cos_thetaI = std_max(0.0f, dot(wi, N));
    float sin_thetaI = sqrtf(1.0f - cos_thetaI * cos_thetaI);
    if (isAlmostZero(cos_thetaO) || isAlmostZero(cos_thetaI))
        return mk(0.0f, 0.0f, 0.0f);
    f3 h = normalize(add(wi, wo));
    if (isAlmostZero(h.x) && isAlmostZero(h.y) && isAlmostZero(h.z))
        return mk(0.0f, 0.0f, 0.0f);
    float cos_theta_H = std_max(0.0f, dot(h, N));
    float D = distributionGTR2(cos_theta_H, m->alpha);
    float G = shadowMaskingTR(sin_thetaI / cos_thetaI, sin_thetaO / cos_thetaO, m->alpha);
    float F = schlick(cos_theta_H, 1.0f, m->ior);
    return scale(m->tint, F * D * G / (4.0f * cos_thetaO * cos_thetaI));
}

static float glossyPdf(const Material *m, f3 wi, f3 wo, const Basis *b) { /* Materials.hpp:177-188 */
    f3 h = normalize(add(wi, wo));
    float cos_theta_h = std_max(0.0f, dot(h, b->N));
    if (isAlmostZero(cos_theta_h))
        return 1.0f;
    float D = distributionGTR2(cos_theta_h, m->alpha);
    float pdfh = D * fabsf(cos_theta_h);
    float wi_dot_h = dot(wi, h);
    if (isAlmostZero(wi_dot_h))
        return pdfh;
    return pdfh / (4.0f * wi_dot_h);
}

/* GlossyBRDF::generateDirection, Materials.hpp:156-175.  Its own pdf/f results are discarded by the caller
 * (Materials.hpp:281-289); only wi matters.  On the early-out wi is left untouched. */
static void glossySample(const Material *m, f3 wo, f3 x, const Basis *b, f3 *wi) {
    float alpha2 = m->alpha * m->alpha;
    float A = 1.0f - x.y;
    float B = 1.0f + (alpha2 - 1.0f) * x.y;
    float cos_theta_H = sqrtf(A / B);
    float sin_theta_H = (float)sqrt(1.0f - cos_theta_H * cos_theta_H);
    float phih = 2.0f * Pi * x.x;
    double cp = cos(phih), sp = sin(phih);
    float kB = (float)(sin_theta_H * cp); /* double product converted to the float element type */
    float kT = (float)(sin_theta_H * sp);
    f3 h = normalize(add(add(scale(b->B, kB), scale(b->T, kT)), scale(b->N, cos_theta_H)));
    if (dot(h, b->N) < 0.0f)
        return;
    float k = (float)(2.0 * dot(wo, h));
    *wi = normalize(sub(scale(h, k), wo));
}

static f3 layeredEval(const Material *m, f3 wi, f3 wo, f3 N) { /* Materials.hpp:255-263 */
    f3 D_f = orenNayarEval(m, wi, wo);
    f3 G_f = glossyEval(m, wi, wo, N);
    float k = 1.0f - schlick(std_max(0.0f, dot(N, wi)), 1.0f, m->ior);
    return add(scale(D_f, k), G_f);
}

static float layeredPdf(const Material *m, f3 wi, f3 wo, const Basis *b) { /* Materials.hpp:265-277 */
    return 0.5f * (randomHemispherePDF() + glossyPdf(m, wi, wo, b));
}

static f3 layeredSample(const Material *m, f3 wo, f3 x, const Basis *b, f3 *wi, float *pdf) { /* Materials.hpp:279-293 */
    if (x.z < 0.5f) {
        *wi = randomHemisphere(x.x, x.y, b); /* BRDF::generateDirection, Materials.hpp:105-111 */
    } else {
        glossySample(m, wo, x, b, wi);
    }
    *pdf = layeredPdf(m, *wi, wo, b);
    return layeredEval(m, *wi, wo, b->N);
}

/* --------------------------------------------------------------------------------------------- Camera.cpp */

typedef struct { f3 eye, corner, u, v; } Camera;

static f3 vdbNormalize(f3 v) { /* NanoVDB.h:919-953: v *= T(1) / Sqrt(lengthSqr) — no small-length guard */
    float len = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    float s = 1.0f / len;
    return mk(v.x * s, v.y * s, v.z * s);
}

static Camera cameraLookAt(f3 from, f3 at, float aspectRatio, float hFov) { /* Camera.cpp:15-34 */
    f3 up = mk(0.0f, 1.0f, 0.0f);
    f3 dir = vdbNormalize(sub(at, from));
    f3 u = cross(up, dir);
    f3 v = cross(u, dir);
    float fovScale = (float)(2.0 * sin(hFov * 0.5));
    u = scale(u, fovScale);
    float vs = aspectRatio * fovScale;
    v = scale(v, vs);
    Camera cam;
    cam.eye = from;
    f3 hu = mk((float)(0.5 * u.x), (float)(0.5 * u.y), (float)(0.5 * u.z));
    f3 hv = mk((float)(0.5 * v.x), (float)(0.5 * v.y), (float)(0.5 * v.z));
    cam.corner = sub(sub(dir, hu), hv);
    cam.u = u;
    cam.v = v;
    return cam;
}

static void cameraRay(const Camera *c, float x, float y, f3 *org, f3 *dir) { /* Camera.cpp:11-13 */
    f3 xu = mk(x * c->u.x, x * c->u.y, x * c->u.z);
    f3 yv = mk(y * c->v.x, y * c->v.y, y * c->v.z);
    *org = c->eye;
    *dir = vdbNormalize(add(add(c->corner, xu), yv));
}

/* ------------------------------------------------------------------------------------------------- Scene */

typedef struct { f3 c; float r; int32_t mat; } Sphere;
typedef struct { f3 n, p; float w, h; int32_t mat; } Plane;

struct ora_scene {
    Camera camera;
    int32_t nSpheres, nPlanes, nMaterials;
    Sphere *spheres;
    Plane *planes;
    Material *materials;
};

const char *ora_kind(void) { return "port"; }

ora_scene *ora_scene_create(const float *camera,
                            const float *spheres, const int32_t *sphereMat, int32_t nSpheres,
                            const float *planes, const int32_t *planeMat, int32_t nPlanes,
                            const float *materials, int32_t nMaterials) {
    ora_scene *s = (ora_scene *)calloc(1, sizeof(ora_scene));
    s->camera = cameraLookAt(mk(camera[0], camera[1], camera[2]), mk(camera[3], camera[4], camera[5]), camera[6],
                             camera[7]); /* Scene.cpp:40-45 */
    s->nSpheres = nSpheres;
    s->nPlanes = nPlanes;
    s->nMaterials = nMaterials + 1;
    s->spheres = (Sphere *)calloc((size_t)(nSpheres > 0 ? nSpheres : 1), sizeof(Sphere));
    s->planes = (Plane *)calloc((size_t)(nPlanes > 0 ? nPlanes : 1), sizeof(Plane));
    s->materials = (Material *)calloc((size_t)s->nMaterials, sizeof(Material));
    /* SceneDescription.hpp:89: material 0 is MaterialDescription{} */
    static const float defaultMaterial[11] = {0.5f, 0.5f, 0.5f, 0, 0, 0, 0.2f, 0, 0, 0, 1.5f};
    s->materials[0] = makeMaterial(defaultMaterial);
    for (int32_t m = 0; m < nMaterials; m++)
        s->materials[m + 1] = makeMaterial(materials + 11 * m);
    for (int32_t i = 0; i < nSpheres; i++) { /* Scene.cpp:5-18 */
        s->spheres[i].c = mk(spheres[4 * i], spheres[4 * i + 1], spheres[4 * i + 2]);
        s->spheres[i].r = spheres[4 * i + 3];
        s->spheres[i].mat = (sphereMat && sphereMat[i] >= 0) ? sphereMat[i] : 0;
    }
    for (int32_t i = 0; i < nPlanes; i++) { /* Scene.cpp:20-38 */
        const float *p = planes + 9 * i;
        s->planes[i].n = mk(p[0], p[1], p[2]);
        s->planes[i].p = mk(p[3], p[4], p[5]);
        s->planes[i].w = p[6];
        s->planes[i].h = p[7];
        s->planes[i].mat = (planeMat && planeMat[i] >= 0) ? planeMat[i] : 0;
    }
    return s;
}

void ora_scene_destroy(ora_scene *s) {
    if (!s)
        return;
    free(s->spheres);
    free(s->planes);
    free(s->materials);
    free(s);
}

/* ------------------------------------------------------------------------------------------ Geometry.cpp */

typedef struct { float t; int32_t prim; f3 P, N; int32_t mat; } Hit;

/* intersectSphere body for one ray, Geometry.cpp:67-105. */
static inline void sphereTest(f3 o, f3 d, const Sphere *s, int32_t id, Hit *hit) {
    if (isAlmostZero(d.x) && isAlmostZero(d.y) && isAlmostZero(d.z))
        return;
    f3 P = sub(o, s->c);
    float A = dot(d, d);
    float B = dot(P, d);
    float C = mag2(P);
    float u = 2.0f * B / A;
    float v = (C - s->r * s->r) / A;
    float discriminant = -v + (u * u) / 4.0f;
    if (discriminant < 0.0f)
        return;
    float shift = (float)sqrt(discriminant); /* double sqrt; equals sqrtf for a float argument */
    float t0 = -u / 2.0f - shift;
    float t1 = -u / 2.0f + shift;
    if (t0 < 0.0f)
        t0 = INFINITY;
    if (t1 < 0.0f)
        t1 = INFINITY;
    float t = t0 < t1 ? t0 : t1;
    if (hit->t > t) {
        hit->t = t;
        hit->prim = id;
        f3 sP = rayT(o, d, t);
        hit->P = sP;
        hit->N = normalize(sub(sP, s->c));
        hit->mat = s->mat;
    }
}

/* intersectPlane body for one ray, Geometry.cpp:145-176. */
static inline void planeTest(f3 o, f3 d, const Plane *p, int32_t id, Hit *hit) {
    if (isAlmostZero(d.x) && isAlmostZero(d.y) && isAlmostZero(d.z))
        return;
    f3 diff = sub(o, p->p);
    float A = -dot(diff, p->n);
    float B = dot(d, p->n);
    int diffNonZero = !(diff.x == 0.0f && diff.y == 0.0f && diff.z == 0.0f);
    if (diffNonZero && isAlmostZero(B))
        return;
    float t = 0.0f;
    if (!isAlmostZero(B))
        t = A / B;
    if (t < 0.0f)
        return;
    f3 sP = rayT(o, d, t);
    Basis b = constructBasis(p->n);
    f3 e = sub(sP, p->p);
    if (fabsf(dot(e, b.T)) * 2.0f > p->w || fabsf(dot(e, b.B)) * 2.0f > p->h)
        return;
    if (hit->t > t) {
        hit->t = t;
        hit->prim = id;
        hit->P = sP;
        hit->N = p->n;
        hit->mat = p->mat;
    }
}

/* Render.cpp:110-140: every sphere, then every plane; strict closer-than update keeps the lowest index on ties.
 * The reference loops primitive-outer / ray-inner; per ray the sequence of tests is the same. */
static inline void closestHit(const ora_scene *s, f3 o, f3 d, Hit *hit) {
    for (int32_t i = 0; i < s->nSpheres; i++)
        sphereTest(o, d, &s->spheres[i], i, hit);
    for (int32_t i = 0; i < s->nPlanes; i++)
        planeTest(o, d, &s->planes[i], s->nSpheres + i, hit);
}

/* --------------------------------------------------------------------------------------------- threading */

typedef void (*chunk_fn)(void *ctx, int64_t lo, int64_t hi);
typedef struct { chunk_fn fn; void *ctx; int64_t n, chunk; atomic_llong cursor; } ParFor;

static void *parforWorker(void *arg) {
    ParFor *pf = (ParFor *)arg;
    for (;;) {
        int64_t c = atomic_fetch_add(&pf->cursor, 1);
        int64_t lo = c * pf->chunk;
        if (lo >= pf->n)
            return NULL;
        int64_t hi = lo + pf->chunk < pf->n ? lo + pf->chunk : pf->n;
        pf->fn(pf->ctx, lo, hi);
    }
}

static int hardwareThreads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void parallelFor(int64_t n, int64_t chunk, int threads, chunk_fn fn, void *ctx) {
    ParFor pf;
    pf.fn = fn;
    pf.ctx = ctx;
    pf.n = n;
    pf.chunk = chunk;
    atomic_init(&pf.cursor, 0);
    if (threads <= 0)
        threads = hardwareThreads();
    int64_t chunks = (n + chunk - 1) / chunk;
    if (chunks < threads)
        threads = (int)(chunks > 0 ? chunks : 1);
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int i = 1; i < threads; i++)
        pthread_create(&tid[i], NULL, parforWorker, &pf);
    parforWorker(&pf);
    for (int i = 1; i < threads; i++)
        pthread_join(tid[i], NULL);
    free(tid);
}

/* ------------------------------------------------------------------------------------------- API: stages */

static inline f3 ld3(const float *p, int64_t k) { return mk(p[3 * k], p[3 * k + 1], p[3 * k + 2]); }
static inline void st3(float *p, int64_t k, f3 v) {
    p[3 * k] = v.x;
    p[3 * k + 1] = v.y;
    p[3 * k + 2] = v.z;
}

void ora_camera_rays(const ora_scene *s, int64_t n, const float *x, const float *y, float *org, float *dir) {
    for (int64_t k = 0; k < n; k++) {
        f3 o, d;
        cameraRay(&s->camera, x[k], y[k], &o, &d);
        st3(org, k, o);
        st3(dir, k, d);
    }
}

void ora_pixel_rays(const ora_scene *s, int32_t W, int32_t H, int64_t n, const int32_t *pi, const int32_t *pj,
                    const float *phi1, const float *phi2, float *org, float *dir) {
    for (int64_t k = 0; k < n; k++) {
        /* NormalizedFrameBufferCoord, Render.cpp:29-37 */
        float dx = 1.0f / W, dy = 1.0f / H;
        float x = pi[k] * dx, y = pj[k] * dy;
        f3 o, d;
        cameraRay(&s->camera, x + phi1[k] * dx, y + phi2[k] * dy, &o, &d); /* Render.cpp:96 */
        st3(org, k, o);
        st3(dir, k, d);
    }
}

typedef struct {
    const ora_scene *s;
    const float *org, *dir, *tInit;
    float *t, *P, *N;
    int32_t *prim, *mat;
} IntersectCtx;

static void intersectChunk(void *vctx, int64_t lo, int64_t hi) {
    IntersectCtx *c = (IntersectCtx *)vctx;
    for (int64_t k = lo; k < hi; k++) {
        Hit hit;
        hit.t = c->tInit ? c->tInit[k] : INFINITY; /* IntersectionData::reset, Geometry.cpp:7-12 */
        hit.prim = -1;
        hit.P = hit.N = mk(0, 0, 0);
        hit.mat = -1;
        closestHit(c->s, ld3(c->org, k), ld3(c->dir, k), &hit);
        c->t[k] = hit.t;
        c->prim[k] = hit.prim;
        if (c->P)
            st3(c->P, k, hit.P);
        if (c->N)
            st3(c->N, k, hit.N);
        if (c->mat)
            c->mat[k] = hit.mat;
    }
}

void ora_intersect(const ora_scene *s, int64_t n, const float *org, const float *dir, const float *tInit, float *t,
                   int32_t *prim, float *P, float *N, int32_t *mat) {
    IntersectCtx c = {s, org, dir, tInit, t, P, N, prim, mat};
    parallelFor(n, 4096, 0, intersectChunk, &c);
}

void ora_bsdf_sample(const ora_scene *s, int64_t n, const int32_t *mat, const float *wo, const float *N,
                     const float *x, float *wi, float *pdf, float *f) {
    for (int64_t k = 0; k < n; k++) {
        Basis b = constructBasis(ld3(N, k));
        f3 w_in = mk(0, 0, 0);            /* Render.cpp:198 */
        float p = randomHemispherePDF();  /* Render.cpp:197 */
        f3 v = layeredSample(&s->materials[mat[k]], ld3(wo, k), ld3(x, k), &b, &w_in, &p);
        st3(wi, k, w_in);
        st3(f, k, v);
        pdf[k] = p;
    }
}

void ora_bsdf_eval(const ora_scene *s, int64_t n, const int32_t *mat, const float *wi, const float *wo,
                   const float *N, float *f, float *pdf) {
    for (int64_t k = 0; k < n; k++) {
        f3 normal = ld3(N, k);
        Basis b = constructBasis(normal);
        const Material *m = &s->materials[mat[k]];
        st3(f, k, layeredEval(m, ld3(wi, k), ld3(wo, k), normal));
        pdf[k] = layeredPdf(m, ld3(wi, k), ld3(wo, k), &b);
    }
}

static float russianRouletteFactor(f3 throughput, int32_t depth) { /* Render.cpp:153-165 */
    const float Base = 0.55f;
    if (depth < 3)
        return 0.99f;
    float power = std_clamp(mag2(throughput), 0.05f / Base, 0.99f);
    return Base * power;
}

void ora_rr_factor(int64_t n, const float *throughput, const int32_t *depth, float *prob) {
    for (int64_t k = 0; k < n; k++)
        prob[k] = russianRouletteFactor(ld3(throughput, k), depth[k]);
}

/* Render.cpp:199 `float3 samplePos(randomGen(), randomGen(), randomGen())`: the order of the three calls is
 * unspecified in C++; g++ 13.3 evaluates the arguments right to left, so the FIRST draw after the RR draw
 * feeds x(2), the second x(1), the third x(0).  tests/test_oracle.py asserts this against oracle/_ref. */
static const int32_t kSampleDrawOrder[3] = {2, 1, 0};

void ora_sample_draw_order(int32_t order[3]) { memcpy(order, kSampleDrawOrder, sizeof kSampleDrawOrder); }

/* Body of accumulateAndBounce for one ray, Render.cpp:173-216.  Returns 1 if the ray survives. */
static int bounce(const ora_scene *s, Xoshiro128Plus *rng, int32_t depth, f3 P, f3 N, int32_t matId, f3 *org,
                  f3 *dir, f3 *thr, f3 *rad) {
    f3 w_out = neg(*dir);
    const Material *mat = &s->materials[matId];
    float prob = russianRouletteFactor(*thr, depth);
    *rad = add(*rad, mul(*thr, mat->emission)); /* Render.cpp:67-69 */
    if (prob < prng_next(rng))
        return 0;
    Basis basis = constructBasis(N);
    float pdf = randomHemispherePDF();
    f3 w_in = mk(0, 0, 0);
    float draws[3];
    draws[0] = prng_next(rng);
    draws[1] = prng_next(rng);
    draws[2] = prng_next(rng);
    f3 x = mk(draws[kSampleDrawOrder[0]], draws[kSampleDrawOrder[1]], draws[kSampleDrawOrder[2]]);
    f3 f = layeredSample(mat, w_out, x, &basis, &w_in, &pdf);
    *org = add(P, scale(w_in, 0.0001f));
    *dir = w_in;
    float c = fabsf(dot(w_in, N));
    f3 fc = scale(f, c);
    float denom = pdf * prob;
    f3 k = mk(fc.x / denom, fc.y / denom, fc.z / denom); /* Color.cpp:11-17 */
    *thr = mul(*thr, k);
    return 1;
}

void ora_shade(const ora_scene *s, int64_t n, int32_t depth, uint64_t seedBase, float *uOut, const float *P,
               const float *N, const int32_t *mat, float *org, float *dir, float *thr, float *rad, uint8_t *alive) {
    for (int64_t k = 0; k < n; k++) {
        Xoshiro128Plus rng, peek;
        prng_seed(&rng, seedBase + (uint64_t)k);
        peek = rng;
        for (int c = 0; c < 4; c++)
            uOut[4 * k + c] = prng_next(&peek);
        f3 o = ld3(org, k), d = ld3(dir, k), T = ld3(thr, k), L = ld3(rad, k);
        alive[k] = (uint8_t)bounce(s, &rng, depth, ld3(P, k), ld3(N, k), mat[k], &o, &d, &T, &L);
        st3(org, k, o);
        st3(dir, k, d);
        st3(thr, k, T);
        st3(rad, k, L);
    }
}

float ora_gtr2(float c, float alpha) { return distributionGTR2(c, alpha); }
float ora_lambda_tr(float t, float alpha) { return lambdaTR(t, alpha); }
float ora_shadow_masking_tr(float ti, float to, float alpha) { return shadowMaskingTR(ti, to, alpha); }
float ora_schlick(float c, float n1, float n2) { return schlick(c, n1, n2); }

void ora_construct_basis(const float *N, float *out9) {
    Basis b = constructBasis(mk(N[0], N[1], N[2]));
    st3(out9, 0, b.T);
    st3(out9, 1, b.B);
    st3(out9, 2, b.N);
}

void ora_prng_floats(uint64_t seed, int64_t jumps, int64_t n, float *out) {
    Xoshiro128Plus root;
    prng_seed(&root, seed);
    Xoshiro128Plus g = cloneForThread(&root, (size_t)jumps);
    for (int64_t k = 0; k < n; k++)
        out[k] = prng_next(&g);
}

/* ------------------------------------------------------------------------------------------------ Tiles */

typedef struct { int32_t i0, j0, i1, j1; } Rect;

static Rect mkRect(int32_t ai, int32_t aj, int32_t bi, int32_t bj) { /* PixelRect(a, b), Math.hpp:240-241 */
    Rect r;
    r.i0 = ai < bi ? ai : bi;
    r.j0 = aj < bj ? aj : bj;
    r.i1 = ai > bi ? ai : bi;
    r.j1 = aj > bj ? aj : bj;
    return r;
}

/* FrameTiling ctor, Tiles.cpp:5-29 — including its behaviour when the frame is not a multiple of the tile
 * (the last column/row gets max = spill-1 measured from 0, Tiles.cpp:21-24). */
static int32_t frameTiling(int32_t W, int32_t H, int32_t tw, int32_t th, Rect *out) {
    if (W <= 0 || H <= 0 || tw <= 0 || th <= 0)
        return -1; /* the reference throws ExpectationException (Math.hpp:236) */
    int32_t numX = W / tw, numY = H / th;
    int32_t spillX = W % tw, spillY = H % th;
    if (spillX != 0)
        numX += 1;
    if (spillY != 0)
        numY += 1;
    int32_t number = 0;
    for (int32_t j = 0; j < numY; j++)
        for (int32_t i = 0; i < numX; i++) {
            int32_t minI = i * tw, minJ = j * th;
            int32_t maxI = (i + 1) * tw - 1, maxJ = (j + 1) * th - 1;
            if (i == numX - 1 && spillX != 0)
                maxI = spillX - 1;
            if (j == numY - 1 && spillY != 0)
                maxJ = spillY - 1;
            if (out)
                out[number] = mkRect(minI, minJ, maxI, maxJ);
            number++;
        }
    return number;
}

int32_t ora_frame_tiling(int32_t W, int32_t H, int32_t tileW, int32_t tileH, int32_t *rects) {
    return frameTiling(W, H, tileW, tileH, (Rect *)rects);
}

/* ----------------------------------------------------------------------------------------------- render */

typedef struct {
    const ora_scene *s;
    int32_t W, H, spp;
    Rect *tiles;
    Xoshiro128Plus *tileRng;
    float *mean, *variance;
    int32_t stride, outW; /* stride > 1: pixel (i, j) is stored at (j / stride) * outW + i / stride (ora_render_strided) */
    atomic_llong rays;
    atomic_int maxDepth;
} RenderCtx;

/* integrateTile, Render.cpp:220-255, with the per-pixel sample batch kept as arrays exactly like RayBatch so the
 * tile's PRNG is consumed in the reference's order: all camera jitters of the pixel first (Render.cpp:93-99), then
 * per bounce the survivors in active-list order (Render.cpp:173). */
static void renderTiles(void *vctx, int64_t lo, int64_t hi) {
    RenderCtx *c = (RenderCtx *)vctx;
    const ora_scene *s = c->s;
    int32_t const spp = c->spp;
    f3 *org = (f3 *)malloc(sizeof(f3) * (size_t)spp), *dir = (f3 *)malloc(sizeof(f3) * (size_t)spp);
    f3 *thr = (f3 *)malloc(sizeof(f3) * (size_t)spp), *rad = (f3 *)malloc(sizeof(f3) * (size_t)spp);
    Hit *hits = (Hit *)malloc(sizeof(Hit) * (size_t)spp);
    int32_t *active = (int32_t *)malloc(sizeof(int32_t) * (size_t)spp);
    long long rays = 0;
    int deepest = 0;
    for (int64_t tile = lo; tile < hi; tile++) {
        Rect r = c->tiles[tile];
        Xoshiro128Plus *rng = &c->tileRng[tile];
        for (int32_t j = r.j0; j <= r.j1; j++)
            for (int32_t i = r.i0; i <= r.i1; i++) {
                float dx = 1.0f / c->W, dy = 1.0f / c->H; /* Render.cpp:29-37 */
                float x = i * dx, y = j * dy;
                for (int32_t k = 0; k < spp; k++) { /* Render.cpp:55-61, 93-99 */
                    float phi1 = prng_next(rng);
                    float phi2 = prng_next(rng);
                    cameraRay(&s->camera, x + phi1 * dx, y + phi2 * dy, &org[k], &dir[k]);
                    thr[k] = mk(1.0f, 1.0f, 1.0f);
                    rad[k] = mk(0.0f, 0.0f, 0.0f);
                    active[k] = k;
                    hits[k].t = INFINITY;
                }
                int32_t nActive = spp, depth = 0;
                while (nActive > 0) { /* Render.cpp:237-243 */
                    rays += nActive;
                    int32_t kept = 0;
                    for (int32_t a = 0; a < nActive; a++) { /* Render.cpp:110-150 */
                        int32_t k = active[a];
                        closestHit(s, org[k], dir[k], &hits[k]);
                        if (hits[k].t < INFINITY)
                            active[kept++] = k;
                    }
                    nActive = kept;
                    kept = 0;
                    for (int32_t a = 0; a < nActive; a++) { /* Render.cpp:173-216 */
                        int32_t k = active[a];
                        if (bounce(s, rng, depth, hits[k].P, hits[k].N, hits[k].mat, &org[k], &dir[k], &thr[k], &rad[k]))
                            active[kept++] = k;
                    }
                    nActive = kept;
                    depth++;
                    for (int32_t k = 0; k < spp; k++) /* IntersectionData::reset */
                        hits[k].t = INFINITY;
                }
                if (depth > deepest)
                    deepest = depth;
                f3 color = mk(0, 0, 0); /* Render.cpp:245-250 */
                for (int32_t k = 0; k < spp; k++)
                    color = add(color, rad[k]);
                color = scale(color, 1.0f / spp);
                int64_t pix = (int64_t)j * c->W + i; /* FrameBuffer.hpp:66 */
                if (c->stride > 1)
                    pix = (int64_t)(j / c->stride) * c->outW + i / c->stride;
                st3(c->mean, pix, color);
                if (c->variance) {
                    for (int ch = 0; ch < 3; ch++) {
                        double sum = 0.0;
                        for (int32_t k = 0; k < spp; k++)
                            sum += ((float *)&rad[k])[ch];
                        double mu = sum / spp, q = 0.0;
                        for (int32_t k = 0; k < spp; k++) {
                            double e = ((float *)&rad[k])[ch] - mu;
                            q += e * e;
                        }
                        c->variance[3 * pix + ch] = spp > 1 ? (float)(q / (spp - 1)) : 0.0f;
                    }
                }
            }
    }
    atomic_fetch_add(&c->rays, rays);
    int seen = atomic_load(&c->maxDepth);
    while (deepest > seen && !atomic_compare_exchange_weak(&c->maxDepth, &seen, deepest)) {
    }
    free(org);
    free(dir);
    free(thr);
    free(rad);
    free(hits);
    free(active);
}

int ora_render(const ora_scene *s, int32_t W, int32_t H, int32_t spp, int32_t tileW, int32_t tileH, uint64_t seed,
               int32_t nthreads, float *mean, float *variance, double *stats) {
    if (spp <= 0 || W <= 0 || H <= 0)
        return 1;
    int32_t nTiles = frameTiling(W, H, tileW, tileH, NULL);
    if (nTiles <= 0)
        return 2;
    RenderCtx c;
    c.s = s;
    c.W = W;
    c.H = H;
    c.spp = spp;
    c.mean = mean;
    c.variance = variance;
    c.stride = 1;
    c.outW = W;
    atomic_init(&c.rays, 0);
    atomic_init(&c.maxDepth, 0);
    c.tiles = (Rect *)malloc(sizeof(Rect) * (size_t)nTiles);
    c.tileRng = (Xoshiro128Plus *)malloc(sizeof(Xoshiro128Plus) * (size_t)nTiles);
    frameTiling(W, H, tileW, tileH, c.tiles);
    /* Render.cpp:329-331: tile k's generator is the root jumped k times (incremental here, same states). */
    Xoshiro128Plus g;
    prng_seed(&g, seed);
    for (int32_t k = 0; k < nTiles; k++) {
        c.tileRng[k] = g;
        prng_jump(&g);
    }
    memset(mean, 0, sizeof(float) * 3u * (size_t)W * (size_t)H);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    parallelFor(nTiles, 1, nthreads, renderTiles, &c);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (stats) {
        stats[0] = (double)atomic_load(&c.rays);
        stats[1] = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        stats[2] = (double)W * H * spp;
        stats[3] = atomic_load(&c.maxDepth);
    }
    free(c.tiles);
    free(c.tileRng);
    return 0;
}

/* Every stride-th pixel of the W x H frame in each dimension, each as its own 1 x 1 tile (its generator is the root
 * jumped k times, k its index in the strided image): the integrand of a full-size frame at a fraction of the cost. */
int ora_render_strided(const ora_scene *s, int32_t W, int32_t H, int32_t spp, int32_t stride, uint64_t seed,
                       int32_t nthreads, float *mean, float *variance, double *stats) {
    if (spp <= 0 || W <= 0 || H <= 0 || stride <= 0)
        return 1;
    int32_t const outW = (W + stride - 1) / stride, outH = (H + stride - 1) / stride;
    int32_t const nTiles = outW * outH;
    RenderCtx c;
    c.s = s;
    c.W = W;
    c.H = H;
    c.spp = spp;
    c.mean = mean;
    c.variance = variance;
    c.stride = stride;
    c.outW = outW;
    atomic_init(&c.rays, 0);
    atomic_init(&c.maxDepth, 0);
    c.tiles = (Rect *)malloc(sizeof(Rect) * (size_t)nTiles);
    c.tileRng = (Xoshiro128Plus *)malloc(sizeof(Xoshiro128Plus) * (size_t)nTiles);
    Xoshiro128Plus g;
    prng_seed(&g, seed);
    for (int32_t k = 0; k < nTiles; k++) {
        int32_t const i = (k % outW) * stride, j = (k / outW) * stride;
        c.tiles[k] = mkRect(i, j, i, j);
        c.tileRng[k] = g;
        prng_jump(&g);
    }
    /* a stride of 1 stores through the full-frame index, like ora_render */
    memset(mean, 0, sizeof(float) * 3u * (size_t)outW * (size_t)outH);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    parallelFor(nTiles, 16, nthreads, renderTiles, &c);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (stats) {
        stats[0] = (double)atomic_load(&c.rays);
        stats[1] = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        stats[2] = (double)outW * outH * spp;
        stats[3] = atomic_load(&c.maxDepth);
    }
    free(c.tiles);
    free(c.tileRng);
    return 0;
}

/* ------------------------------------------------------------------------------------------------ Color */

static float srgbChannel(float x) { /* Color.cpp:64-78: note 12.95 (not 12.92) and the double pow */
    const float a = 0.055f;
    if ((double)x <= 0.0031308)
        return x * 12.95f;
    return (float)((1 + a) * pow(x, 1.0f / 2.4f) - a);
}

static uint8_t quantizeTo8bit(double v) { /* FrameBuffer.hpp:91-94 */
    v = round(255.0 * v);
    v = (v < 0.0) ? 0.0 : (255.0 < v) ? 255.0 : v;
    return (uint8_t)v;
}

void ora_to_srgb8(int64_t npixels, const float *rgb, uint8_t *out) {
    for (int64_t k = 0; k < 3 * npixels; k++)
        out[k] = quantizeTo8bit(srgbChannel(rgb[k]));
}
