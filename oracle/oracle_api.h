/* TEST INFRASTRUCTURE — not part of the product.
 *
 * One C interface, two implementations:
 *   oracle/_ref/libcornelis_ref.so     the UNMODIFIED reference sources under /root/reference compiled by
 *                                      oracle/build_ref.sh (+ oracle/ref_harness.cpp, which only calls them)
 *   oracle/libcornelis_oracle.so       oracle/cornelis_oracle.c, a plain-C restatement of the same algorithm
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may load either.
 * The product (cornelis_b200/, include/, the C-ABI library) never links, imports or calls anything here.
 *
 * All arrays are caller-owned.  3-vectors are packed xyz (stride 3 floats) unless noted.
 */
#ifndef CORNELIS_ORACLE_API_H
#define CORNELIS_ORACLE_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_scene ora_scene;

/* "reference" (compiled /root/reference) or "port" (C restatement). */
const char *ora_kind(void);

/* Scene description, flattened from the reference's SceneDescription (include/cornelis/SceneDescription.hpp:14-92):
 *   camera[8]      origin xyz, lookAt xyz, aspect, horizontalFov
 *   spheres[4*n]   center xyz, radius                 sphereMat[n]  material index or -1 (= none -> 0)
 *   planes[9*n]    normal xyz, point xyz, extents xyz planeMat[n]
 *   materials[11*n] albedo rgb, emissive rgb, roughness, reflectionTint rgb, ior — the USER materials;
 *                  index 0 is always the implicit default material (SceneDescription.hpp:89), so user
 *                  material k is scene material k+1.
 */
ora_scene *ora_scene_create(const float *camera,
                            const float *spheres, const int32_t *sphereMat, int32_t nSpheres,
                            const float *planes, const int32_t *planeMat, int32_t nPlanes,
                            const float *materials, int32_t nMaterials);
void ora_scene_destroy(ora_scene *scene);

/* cam(x, y) for explicit film coordinates (Camera.cpp:11-13).  org/dir: 3*n. */
void ora_camera_rays(const ora_scene *scene, int64_t n, const float *x, const float *y, float *org, float *dir);

/* generateCameraRays arithmetic for explicit jitter (Render.cpp:29-37, 85-100):
 * pixel (i,j) of a W x H frame, jitter (phi1, phi2) -> ray. */
void ora_pixel_rays(const ora_scene *scene, int32_t W, int32_t H, int64_t n, const int32_t *pi, const int32_t *pj,
                    const float *phi1, const float *phi2, float *org, float *dir);

/* Closest hit of n rays against every sphere then every plane (Render.cpp:110-140 over Geometry.cpp:34-178),
 * called one primitive at a time so that the primitive that last lowered t is known.
 * prim: sphere index, or nSpheres + plane index, or -1 for a miss (t = +inf).  P, N: 3*n; mat: n.
 * tInit may be NULL (= +inf everywhere). */
void ora_intersect(const ora_scene *scene, int64_t n, const float *org, const float *dir, const float *tInit,
                   float *t, int32_t *prim, float *P, float *N, int32_t *mat);

/* LayeredBRDF::generateDirection on explicit random numbers (Materials.hpp:279-293).
 * mat[n] scene material index; wo, N, x: 3*n in; wi, f: 3*n out; pdf: n out.
 * pdf is initialised to randomHemispherePDF() and wi to 0 exactly as Render.cpp:197-198 does. */
void ora_bsdf_sample(const ora_scene *scene, int64_t n, const int32_t *mat, const float *wo, const float *N,
                     const float *x, float *wi, float *pdf, float *f);

/* LayeredBRDF::operator() and ::pdf (Materials.hpp:255-277). */
void ora_bsdf_eval(const ora_scene *scene, int64_t n, const int32_t *mat, const float *wi, const float *wo,
                   const float *N, float *f, float *pdf);

/* russianRouletteFactor (Render.cpp:153-165). */
void ora_rr_factor(int64_t n, const float *throughput, const int32_t *depth, float *prob);

/* One accumulateAndBounce pass (Render.cpp:167-218) over n rays that all have a hit.  The reference takes its
 * random numbers from a PRNG object, so ray k is given its own generator PRNG(seedBase + k) and uOut[4*k..4*k+3]
 * reports that generator's next four draws in draw order (the pass consumes one if the ray is RR-killed, else
 * all four: RR first, then the three BSDF sample numbers — see ora_sample_draw_order).
 * In/out: org, dir, thr, rad (3*n each).  In: P, N (3*n), mat (n), depth (scalar).  Out: alive[n], uOut[4*n]. */
void ora_shade(const ora_scene *scene, int64_t n, int32_t depth, uint64_t seedBase, float *uOut, const float *P,
               const float *N, const int32_t *mat, float *org, float *dir, float *thr, float *rad, uint8_t *alive);

/* Microfacet helpers (Materials.cpp:16-42). */
float ora_gtr2(float cosThetaH, float alpha);
float ora_lambda_tr(float tanTheta, float alpha);
float ora_shadow_masking_tr(float tanI, float tanO, float alpha);
float ora_schlick(float cosTheta, float ior1, float ior2);

/* constructBasis (Math.hpp:424-434): out[9] = T xyz, B xyz, N xyz. */
void ora_construct_basis(const float *N, float *out9);

/* PRNG stream: PRNG(seed) jumped `jumps` times (PRNG.hpp:11-37), next n floats. */
void ora_prng_floats(uint64_t seed, int64_t jumps, int64_t n, float *out);

/* Full render by the reference's own loop: FrameTiling(W x H, tileW x tileH), per-tile PRNG =
 * cloneForThread(PRNG(seed), tileNumber), integrateTile per tile (Render.cpp:220-255, 327-343).
 *   mean[3*W*H]  the framebuffer, row-major j*W+i
 *   m2[3*W*H]    optional (NULL to skip): per-pixel sum over samples of (L_k - mean)^2 in double, returned as
 *                the unbiased sample variance (divided by spp-1); requires the instrumented loop
 *   stats[4]     optional: [0] rays intersected, [1] wall seconds, [2] pixel-samples, [3] max depth reached
 * nthreads <= 0 means all hardware threads.  Returns 0 on success. */
int ora_render(const ora_scene *scene, int32_t W, int32_t H, int32_t spp, int32_t tileW, int32_t tileH,
               uint64_t seed, int32_t nthreads, float *mean, float *variance, double *stats);

/* The same loop on every stride-th pixel of the W x H frame in each dimension: pixel (i' * stride, j' * stride) is
 * integrated exactly as integrateTile integrates it in the full frame (same NormalizedFrameBufferCoord, same camera),
 * as a 1 x 1 tile whose PRNG is cloneForThread(PRNG(seed), j' * outW + i').  mean / variance hold
 * 3 * outW * outH floats, outW = ceil(W / stride), outH = ceil(H / stride); stats[2] counts the strided pixel-samples.
 * This is how the headline 1920x1080 frame is pinned at 4096 spp without 8.5 G CPU samples. */
int ora_render_strided(const ora_scene *scene, int32_t W, int32_t H, int32_t spp, int32_t stride, uint64_t seed,
                       int32_t nthreads, float *mean, float *variance, double *stats);

/* Display transform + quantisation (Color.cpp:64-80, FrameBuffer.hpp:91-95). */
void ora_to_srgb8(int64_t npixels, const float *rgb, uint8_t *out);

/* FrameTiling bounds (Tiles.cpp:5-29): rects[4*k] = min.i, min.j, max.i, max.j.  Returns tile count
 * (call with rects = NULL to size). */
int32_t ora_frame_tiling(int32_t W, int32_t H, int32_t tileW, int32_t tileH, int32_t *rects);

/* How the three BSDF sample numbers map to consecutive PRNG draws in Render.cpp:199 as compiled
 * (C++ leaves the order unspecified): order[c] = which of the three draws after the RR draw feeds x(c). */
void ora_sample_draw_order(int32_t order[3]);

#ifdef __cplusplus
}
#endif
#endif
