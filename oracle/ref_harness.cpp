// TEST INFRASTRUCTURE — not part of the product.
//
// Exposes the UNMODIFIED reference implementation through oracle/oracle_api.h.  This translation unit pulls in
// the reference's own Render.cpp (found through -I/root/reference/src) so that its stage functions —
// generateCameraRays, intersect, russianRouletteFactor, accumulateAndBounce, integrateTile — and RayBatch are
// callable; every number returned below is computed by reference code, this file only marshals arrays.
// Built by oracle/build_ref.sh into oracle/_ref/libcornelis_ref.so (git-ignored).  Nothing from
// /root/reference is copied into this repository.
#include "Render.cpp" // the reference's src/Render.cpp — must come first

#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>

#include "oracle_api.h"

using namespace cornelis;

struct ora_scene {
    explicit ora_scene(SceneDescription const &d) : descr(d), data(descr) {}
    SceneDescription descr;
    mutable SceneData data; // the reference's stage functions take non-const references
};

namespace {
struct Rays : SoAObject<tags::PositionX,
                        tags::PositionY,
                        tags::PositionZ,
                        tags::DirectionX,
                        tags::DirectionY,
                        tags::DirectionZ> {
    explicit Rays(std::size_t n) : SoAObject(n) {}
};

template <typename F>
void forChunks(int64_t n, int64_t chunk, F body) {
    int64_t const numChunks = (n + chunk - 1) / chunk;
    std::atomic<int64_t> cursor{0};
    auto worker = [&] {
        for (;;) {
            int64_t c = cursor.fetch_add(1);
            if (c >= numChunks)
                return;
            int64_t lo = c * chunk;
            body(lo, std::min(n, lo + chunk));
        }
    };
    int threads = static_cast<int>(std::thread::hardware_concurrency());
    if (threads < 1)
        threads = 1;
    if (numChunks < threads)
        threads = static_cast<int>(numChunks);
    std::vector<std::thread> pool;
    for (int i = 1; i < threads; i++)
        pool.emplace_back(worker);
    worker();
    for (auto &t : pool)
        t.join();
}

inline float3 load3(float const *p, int64_t k) { return float3{p[3 * k], p[3 * k + 1], p[3 * k + 2]}; }
inline void store3(float *p, int64_t k, float3 const &v) {
    p[3 * k] = v(0);
    p[3 * k + 1] = v(1);
    p[3 * k + 2] = v(2);
}
inline void store3(float *p, int64_t k, RGB const &v) {
    p[3 * k] = v(0);
    p[3 * k + 1] = v(1);
    p[3 * k + 2] = v(2);
}
} // namespace

extern "C" {

const char *ora_kind(void) { return "reference"; }

ora_scene *ora_scene_create(const float *camera,
                            const float *spheres, const int32_t *sphereMat, int32_t nSpheres,
                            const float *planes, const int32_t *planeMat, int32_t nPlanes,
                            const float *materials, int32_t nMaterials) {
    SceneDescription d;
    PerspectiveCameraDescription cam;
    cam.origin = V3(camera[0], camera[1], camera[2]);
    cam.lookAt = V3(camera[3], camera[4], camera[5]);
    cam.aspect = camera[6];
    cam.horizontalFov = camera[7];
    d.setCamera(cam);
    for (int32_t m = 0; m < nMaterials; m++) {
        float const *p = materials + 11 * m;
        MaterialDescription md;
        md.albedo = RGB(p[0], p[1], p[2]);
        md.emissive = RGB(p[3], p[4], p[5]);
        md.roughness = p[6];
        md.reflectionTint = RGB(p[7], p[8], p[9]);
        md.ior = p[10];
        d.addMaterial(md);
    }
    for (int32_t s = 0; s < nSpheres; s++) {
        SphereDescription sd;
        sd.center = V3(spheres[4 * s], spheres[4 * s + 1], spheres[4 * s + 2]);
        sd.radius = spheres[4 * s + 3];
        if (sphereMat && sphereMat[s] >= 0)
            sd.material = static_cast<std::size_t>(sphereMat[s]);
        d.addSphere(sd);
    }
    for (int32_t q = 0; q < nPlanes; q++) {
        float const *p = planes + 9 * q;
        PlaneDescription pd;
        pd.normal = V3(p[0], p[1], p[2]);
        pd.point = V3(p[3], p[4], p[5]);
        pd.extents = V3(p[6], p[7], p[8]);
        if (planeMat && planeMat[q] >= 0)
            pd.material = static_cast<std::size_t>(planeMat[q]);
        d.addPlane(pd);
    }
    return new ora_scene(d);
}

void ora_scene_destroy(ora_scene *scene) { delete scene; }

void ora_camera_rays(const ora_scene *scene, int64_t n, const float *x, const float *y, float *org, float *dir) {
    for (int64_t k = 0; k < n; k++) {
        auto ray = scene->data.camera(x[k], y[k]);
        for (int c = 0; c < 3; c++) {
            org[3 * k + c] = ray.eye()[c];
            dir[3 * k + c] = ray.dir()[c];
        }
    }
}

void ora_pixel_rays(const ora_scene *scene, int32_t W, int32_t H, int64_t n, const int32_t *pi, const int32_t *pj,
                    const float *phi1, const float *phi2, float *org, float *dir) {
    for (int64_t k = 0; k < n; k++) {
        NormalizedFrameBufferCoord coord({pi[k], pj[k]}, {W, H});
        // Same expression as Render.cpp:96.
        auto ray = scene->data.camera(coord.x + phi1[k] * coord.dx, coord.y + phi2[k] * coord.dy);
        for (int c = 0; c < 3; c++) {
            org[3 * k + c] = ray.eye()[c];
            dir[3 * k + c] = ray.dir()[c];
        }
    }
}

void ora_intersect(const ora_scene *scene, int64_t n, const float *org, const float *dir, const float *tInit,
                   float *t, int32_t *prim, float *P, float *N, int32_t *mat) {
    SceneData &sd = scene->data;
    forChunks(n, 4096, [&](int64_t lo, int64_t hi) {
        std::size_t const m = static_cast<std::size_t>(hi - lo);
        Rays rays(m);
        IntersectionData isect(m);
        auto params = isect.get<tags::RayParam0>();
        std::vector<std::size_t> active(m);
        std::vector<float> prev(m);
        std::vector<int32_t> last(m, -1);
        for (std::size_t k = 0; k < m; k++) {
            setPosition(rays, k, load3(org, lo + k));
            setDirection(rays, k, load3(dir, lo + k));
            active[k] = k;
            if (tInit)
                params[k] = tInit[lo + k];
            prev[k] = params[k];
        }
        auto note = [&](int32_t id) {
            for (std::size_t k = 0; k < m; k++)
                if (params[k] != prev[k]) {
                    prev[k] = params[k];
                    last[k] = id;
                }
        };
        // Same primitive order as Render.cpp:115-140, one reference call per primitive.
        auto [Sx, Sy, Sz] = getPositions(sd.spheres);
        auto radius = sd.spheres.get<tags::Radius>();
        auto sMat = sd.spheres.get<tags::MaterialId>();
        int32_t const nS = static_cast<int32_t>(Sx.size());
        for (int32_t i = 0; i < nS; i++) {
            intersectSphere(getPositions(rays), getDirectionSpans(rays), float3(Sx[i], Sy[i], Sz[i]), radius[i],
                            sMat[i], isect, active);
            note(i);
        }
        auto [Px, Py, Pz] = getPositions(sd.planes);
        auto [PNx, PNy, PNz] = getNormalSpans(sd.planes);
        auto width = sd.planes.get<tags::WidthF>();
        auto height = sd.planes.get<tags::HeightF>();
        auto pMat = sd.planes.get<tags::MaterialId>();
        for (int32_t i = 0; i < static_cast<int32_t>(Px.size()); i++) {
            intersectPlane(getPositions(rays), getDirectionSpans(rays), float3(PNx[i], PNy[i], PNz[i]),
                           float3(Px[i], Py[i], Pz[i]), width[i], height[i], pMat[i], isect, active);
            note(nS + i);
        }
        auto [IPx, IPy, IPz] = getPositions(isect);
        auto [INx, INy, INz] = getNormalSpans(isect);
        auto iMat = isect.get<tags::MaterialId>();
        for (std::size_t k = 0; k < m; k++) {
            int64_t g = lo + static_cast<int64_t>(k);
            t[g] = params[k];
            prim[g] = last[k];
            bool const hit = last[k] >= 0;
            if (P)
                store3(P, g, hit ? float3{IPx[k], IPy[k], IPz[k]} : float3{0});
            if (N)
                store3(N, g, hit ? float3{INx[k], INy[k], INz[k]} : float3{0});
            if (mat)
                mat[g] = hit ? static_cast<int32_t>(iMat[k]) : -1;
        }
    });
}

void ora_bsdf_sample(const ora_scene *scene, int64_t n, const int32_t *mat, const float *wo, const float *N,
                     const float *x, float *wi, float *pdf, float *f) {
    for (int64_t k = 0; k < n; k++) {
        float3 const normal = load3(N, k);
        auto const &material = scene->data.materials[static_cast<std::size_t>(mat[k])];
        Basis basis = constructBasis(normal);
        BRDF const &brdf = material.brdf(float3{0}, normal);
        float p = randomHemispherePDF(); // Render.cpp:197
        float3 w_in{};                   // Render.cpp:198
        RGB value = brdf.generateDirection(load3(wo, k), load3(x, k), basis, w_in, p);
        store3(wi, k, w_in);
        store3(f, k, value);
        pdf[k] = p;
    }
}

void ora_bsdf_eval(const ora_scene *scene, int64_t n, const int32_t *mat, const float *wi, const float *wo,
                   const float *N, float *f, float *pdf) {
    for (int64_t k = 0; k < n; k++) {
        float3 const normal = load3(N, k);
        auto const &material = scene->data.materials[static_cast<std::size_t>(mat[k])];
        BRDF const &brdf = material.brdf(float3{0}, normal);
        store3(f, k, brdf(load3(wi, k), load3(wo, k), normal));
        pdf[k] = brdf.pdf(load3(wi, k), load3(wo, k), constructBasis(normal));
    }
}

void ora_rr_factor(int64_t n, const float *throughput, const int32_t *depth, float *prob) {
    for (int64_t k = 0; k < n; k++)
        prob[k] = russianRouletteFactor(RGB(throughput[3 * k], throughput[3 * k + 1], throughput[3 * k + 2]),
                                        depth[k]);
}

// Ray k uses PRNG(seedBase + k); uOut[4k..4k+3] receives the next four draws of that generator in draw order
// (the reference consumes one if the ray is RR-killed, else all four).
void ora_shade(const ora_scene *scene, int64_t n, int32_t depth, uint64_t seedBase, float *uOut,
                      const float *P, const float *N, const int32_t *mat, float *org, float *dir, float *thr,
                      float *rad, uint8_t *alive) {
    SceneData &sd = scene->data;
    for (int64_t k = 0; k < n; k++) {
        RayBatch batch(1);
        IntersectionData isect(1);
        setPosition(batch, 0, load3(org, k));
        setDirection(batch, 0, load3(dir, k));
        batch.get<PathThroughputTag>()[0] = RGB(thr[3 * k], thr[3 * k + 1], thr[3 * k + 2]);
        batch.get<LightInTag>()[0] = RGB(rad[3 * k], rad[3 * k + 1], rad[3 * k + 2]);
        setPosition(isect, 0, load3(P, k));
        setNormal(isect, 0, load3(N, k));
        isect.get<tags::MaterialId>()[0] = static_cast<std::size_t>(mat[k]);
        isect.get<tags::RayParam0>()[0] = 1.0f;

        PRNG prng(seedBase + static_cast<uint64_t>(k));
        PRNG peek(prng);
        for (int c = 0; c < 4; c++)
            uOut[4 * k + c] = peek();

        accumulateAndBounce(sd, batch, isect, prng, depth);

        alive[k] = batch.activeList.empty() ? 0 : 1;
        store3(org, k, batch.rayOrigin(0));
        store3(dir, k, batch.rayDir(0));
        store3(thr, k, batch.throughput(0));
        store3(rad, k, batch.get<LightInTag>()[0]);
    }
}

float ora_gtr2(float c, float alpha) { return models::distributionGTR2(c, alpha); }
float ora_lambda_tr(float tanTheta, float alpha) { return models::lambdaTR(tanTheta, alpha); }
float ora_shadow_masking_tr(float ti, float to, float alpha) { return models::shadowMaskingTR(ti, to, alpha); }
float ora_schlick(float c, float n1, float n2) { return models::schlick(c, n1, n2); }

void ora_construct_basis(const float *N, float *out9) {
    Basis b = constructBasis(float3{N[0], N[1], N[2]});
    store3(out9, 0, b.T);
    store3(out9, 1, b.B);
    store3(out9, 2, b.N);
}

void ora_prng_floats(uint64_t seed, int64_t jumps, int64_t n, float *out) {
    PRNG root(seed);
    PRNG g = cloneForThread(root, static_cast<std::size_t>(jumps));
    for (int64_t k = 0; k < n; k++)
        out[k] = g();
}

// The tile loop of RenderSession::render (Render.cpp:327-343) over an arbitrary list of tiles.  `stride` > 1 stores
// pixel (i, j) at (j / stride) * outW + i / stride of the output arrays (ora_render_strided).
static int renderTiles(const ora_scene *scene, int32_t W, int32_t H, int32_t spp, std::vector<TileInfo> &tiles,
                       int32_t stride, int32_t outW, int32_t outH, int32_t nthreads, float *mean, float *variance,
                       double *stats) {
    SceneData &sd = scene->data;
    RenderOptions options{spp}; // samplesAA is a const member (RenderOptions.hpp:15): aggregate-initialise
    RGBFrameBuffer fb(PixelRect(W, H));
    auto outIndex = [&](int64_t i, int64_t j) { return stride > 1 ? (j / stride) * outW + i / stride : j * W + i; };

    bool const instrumented = variance != nullptr || stats != nullptr;
    std::atomic<int64_t> raysTraced{0};
    std::atomic<int32_t> maxDepth{0};

    tbb::shim::requestedThreads().store(nthreads);
    auto const t0 = std::chrono::steady_clock::now();
    tbb::task_group group;
    group.run_and_wait([&] {
        tbb::parallel_for_each(std::begin(tiles), std::end(tiles), [&](TileInfo &tileInfo) {
            if (!instrumented) {
                integrateTile(tileInfo, options, sd, fb); // Render.cpp:343
                return;
            }
            // Instrumented twin of integrateTile (Render.cpp:220-255): identical calls in identical order, plus
            // ray counting and the per-sample second moment.  tests/ checks it reproduces integrateTile's bits.
            int64_t rays = 0;
            int32_t deepest = 0;
            for (auto j = tileInfo.bounds.min().j; j <= tileInfo.bounds.max().j; j++) {
                for (auto i = tileInfo.bounds.min().i; i <= tileInfo.bounds.max().i; i++) {
                    NormalizedFrameBufferCoord screenCoord({i, j}, {fb.width(), fb.height()});
                    RayBatch raybatch(options.samplesAA);
                    generateCameraRays(tileInfo, sd.camera, screenCoord, raybatch);
                    IntersectionData intersections(options.samplesAA);
                    int32_t depth = 0;
                    while (raybatch.activeList.size() > 0) {
                        rays += static_cast<int64_t>(raybatch.activeList.size());
                        intersect(sd, raybatch, intersections);
                        accumulateAndBounce(sd, raybatch, intersections, tileInfo.randomGen, depth++);
                        intersections.reset();
                    }
                    deepest = std::max(deepest, depth);
                    RGB color = RGB::black();
                    for (auto const &term : raybatch.get<LightInTag>())
                        color += term;
                    color = color * (1.0f / options.samplesAA);
                    fb(i, j) = color;
                    if (variance) {
                        auto L = raybatch.get<LightInTag>();
                        for (int c = 0; c < 3; c++) {
                            double s = 0.0;
                            for (auto const &term : L)
                                s += term(c);
                            double const mu = s / spp;
                            double q = 0.0;
                            for (auto const &term : L)
                                q += (term(c) - mu) * (term(c) - mu);
                            variance[3 * outIndex(i, j) + c] = spp > 1 ? static_cast<float>(q / (spp - 1)) : 0.0f;
                        }
                    }
                }
            }
            raysTraced += rays;
            int32_t seen = maxDepth.load();
            while (deepest > seen && !maxDepth.compare_exchange_weak(seen, deepest)) {
            }
        });
    });
    auto const t1 = std::chrono::steady_clock::now();
    tbb::shim::requestedThreads().store(0);

    if (stride > 1) {
        for (int64_t j = 0; j < H; j += stride)
            for (int64_t i = 0; i < W; i += stride)
                store3(mean, outIndex(i, j), fb.data()[j * W + i]);
    } else {
        for (int64_t k = 0; k < static_cast<int64_t>(W) * H; k++)
            store3(mean, k, fb.data()[k]);
    }
    if (stats) {
        stats[0] = static_cast<double>(raysTraced.load());
        stats[1] = std::chrono::duration<double>(t1 - t0).count();
        stats[2] = static_cast<double>(outW) * outH * spp;
        stats[3] = maxDepth.load();
    }
    return 0;
}

int ora_render(const ora_scene *scene, int32_t W, int32_t H, int32_t spp, int32_t tileW, int32_t tileH,
               uint64_t seed, int32_t nthreads, float *mean, float *variance, double *stats) {
    if (spp <= 0 || W <= 0 || H <= 0)
        return 1;
    PRNG rootRng(seed);
    // Render.cpp:327-331
    FrameTiling tiling(PixelRect(W, H), PixelRect{tileW, tileH});
    std::vector<TileInfo> tiles(std::begin(tiling), std::end(tiling));
    for (auto &tileInfo : tiles)
        tileInfo.randomGen = cloneForThread(rootRng, tileInfo.tileNumber);
    return renderTiles(scene, W, H, spp, tiles, 1, W, H, nthreads, mean, variance, stats);
}

int ora_render_strided(const ora_scene *scene, int32_t W, int32_t H, int32_t spp, int32_t stride, uint64_t seed,
                       int32_t nthreads, float *mean, float *variance, double *stats) {
    if (spp <= 0 || W <= 0 || H <= 0 || stride <= 0)
        return 1;
    int32_t const outW = (W + stride - 1) / stride, outH = (H + stride - 1) / stride;
    // one 1 x 1 tile per strided pixel; generators as Render.cpp:329-331 hands them to tiles: the root jumped k times
    // (incrementally here: cloneForThread(root, k) costs k jumps)
    PRNG g(seed);
    std::vector<TileInfo> tiles;
    tiles.reserve(static_cast<std::size_t>(outW) * outH);
    for (int32_t k = 0; k < outW * outH; k++) {
        int32_t const i = (k % outW) * stride, j = (k / outW) * stride;
        TileInfo t(static_cast<std::size_t>(k), PixelRect(PixelCoord{i, j}, PixelCoord{i, j}));
        t.randomGen = g;
        g = cloneForThread(g, 1);
        tiles.push_back(t);
    }
    return renderTiles(scene, W, H, spp, tiles, stride > 1 ? stride : 1, stride > 1 ? outW : W, stride > 1 ? outH : H,
                       nthreads, mean, variance, stats);
}

void ora_to_srgb8(int64_t npixels, const float *rgb, uint8_t *out) {
    for (int64_t k = 0; k < npixels; k++) {
        SRGB s = toSRGB(RGB(rgb[3 * k], rgb[3 * k + 1], rgb[3 * k + 2]));
        auto q = quantizeTo8bit(s);
        out[3 * k] = q[0];
        out[3 * k + 1] = q[1];
        out[3 * k + 2] = q[2];
    }
}

int32_t ora_frame_tiling(int32_t W, int32_t H, int32_t tileW, int32_t tileH, int32_t *rects) {
    FrameTiling tiling(PixelRect(W, H), PixelRect{tileW, tileH});
    if (rects) {
        for (std::size_t k = 0; k < tiling.size(); k++) {
            rects[4 * k] = tiling[k].bounds.min().i;
            rects[4 * k + 1] = tiling[k].bounds.min().j;
            rects[4 * k + 2] = tiling[k].bounds.max().i;
            rects[4 * k + 3] = tiling[k].bounds.max().j;
        }
    }
    return static_cast<int32_t>(tiling.size());
}

void ora_sample_draw_order(int32_t order[3]) {
    // Probe Render.cpp:199 as compiled: run one bounce with a known generator and find which permutation of the
    // three post-RR draws reproduces the sampled direction through generateDirection.
    static int32_t cached[3] = {-1, -1, -1};
    static std::once_flag once;
    std::call_once(once, [] {
        SceneDescription d;
        MaterialDescription md;
        md.albedo = RGB(.73f, .73f, .73f);
        d.addMaterial(md);
        SceneData sd(d);
        float3 const N{0.0f, 1.0f, 0.0f};
        float3 const wo = normalize(float3{0.3f, 0.8f, -0.2f});
        for (uint64_t seed = 1; seed < 64 && cached[0] < 0; seed++) {
            PRNG prng(seed);
            PRNG peek(prng);
            float u[4];
            for (float &v : u)
                v = peek();
            RayBatch batch(1);
            IntersectionData isect(1);
            setPosition(batch, 0, float3{0.0f, 1.0f, 0.0f});
            setDirection(batch, 0, -wo);
            setPosition(isect, 0, float3{0});
            setNormal(isect, 0, N);
            isect.get<tags::MaterialId>()[0] = 1;
            isect.get<tags::RayParam0>()[0] = 1.0f;
            accumulateAndBounce(sd, batch, isect, prng, 0);
            if (batch.activeList.empty())
                continue;
            float3 const got = batch.rayDir(0);
            int const perms[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
            int matches = 0, which = -1;
            for (int p = 0; p < 6; p++) {
                float3 x{u[1 + perms[p][0]], u[1 + perms[p][1]], u[1 + perms[p][2]]};
                float3 w_in{};
                float pdf = randomHemispherePDF();
                sd.materials[1].brdf(float3{0}, N).generateDirection(wo, x, constructBasis(N), w_in, pdf);
                if (w_in == got) {
                    matches++;
                    which = p;
                }
            }
            if (matches == 1)
                for (int c = 0; c < 3; c++)
                    cached[c] = perms[which][c];
        }
    });
    for (int c = 0; c < 3; c++)
        order[c] = cached[c];
}

} // extern "C"
