// The cornelis command line (reference src/cornelis.cpp): with no arguments it renders the Cornell box at 512x512,
// 4096 spp, and writes cornelisrender2.png, as the reference binary does.  The reference ignores argv; the flags
// below expose what it hard-codes.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include <cornelis/Render.hpp>
#include <cornelis/SceneDescription.hpp>

using namespace cornelis;

// The benchmark fixture: five finite planes, four spheres, five materials (values of reference src/cornelis.cpp:6-74).
static SceneDescription cornellBox(float aspect) {
    float const side = 555.0f;
    float const half = 550.0f / 2.0f;
    SceneDescription scene;
    PerspectiveCameraDescription cam;
    cam.origin = V3(0, half, -1100);
    cam.lookAt = V3(0, half, 0);
    cam.aspect = aspect;
    cam.horizontalFov = 0.7f;
    scene.setCamera(cam);

    auto matte = [](float r, float g, float b) {
        MaterialDescription m;
        m.albedo = RGB(r, g, b);
        return m;
    };
    auto const red = scene.addMaterial(matte(.65f, .05f, .05f));
    auto const white = scene.addMaterial(matte(.73f, .73f, .73f));
    auto const green = scene.addMaterial(matte(.12f, .45f, .15f));
    MaterialDescription goldDescr;
    goldDescr.albedo = RGB::black();
    goldDescr.roughness = 0.01f;
    goldDescr.reflectionTint = RGB(0.916f, 0.61f, 0.0f);
    goldDescr.ior = 0.470f;
    auto const gold = scene.addMaterial(goldDescr);
    MaterialDescription lightDescr;
    lightDescr.albedo = RGB::black();
    lightDescr.emissive = RGB(15, 15, 15);
    auto const light = scene.addMaterial(lightDescr);

    auto wall = [&](V3 normal, V3 point, std::size_t material) {
        PlaneDescription p;
        p.normal = normal;
        p.point = point;
        p.extents = V3(side, side, 0);
        p.material = material;
        scene.addPlane(p);
    };
    wall(V3(1, 0, 0), V3(-half, half, 0), green);   // left
    wall(V3(-1, 0, 0), V3(half, half, 0), red);     // right
    wall(V3(0, -1, 0), V3(0, side, 0), white);      // roof
    wall(V3(0, 1, 0), V3(0, 0, 0), white);          // floor
    wall(V3(0, 0, -1), V3(0, half, half), white);   // back

    auto ball = [&](V3 center, float radius, std::size_t material) {
        SphereDescription s;
        s.center = center;
        s.radius = radius;
        s.material = material;
        scene.addSphere(s);
    };
    ball(V3(0, side - 60.0f, 0), 60.0f, light);
    ball(V3(0, 50.0f, 0), 50.0f, red);
    ball(V3(-160, 100.0f, 0), 100.0f, white);
    ball(V3(160, 125.0f, 200), 125.0f, gold);
    return scene;
}

// BASELINE.json configs[3]-style scene: n random spheres over one 4000x4000 floor, 64 materials of the one material
// model with varied albedo / tint / roughness / ior, 2 % of them emissive.  SplitMix64 keyed by
// the seed; floats use the reference's 24-bit mapping.
static SceneDescription manySpheres(std::size_t n, std::uint64_t seed, float aspect) {
    std::uint64_t state = seed;
    auto uniform = [&state]() {
        std::uint64_t z = (state += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        z ^= z >> 31;
        return static_cast<float>(static_cast<std::uint32_t>(z >> 32) >> 8) * 0x1.0p-24f;
    };
    SceneDescription scene;
    PerspectiveCameraDescription cam;
    cam.origin = V3(0, 600, -2500);
    cam.lookAt = V3(0, 300, 0);
    cam.aspect = aspect;
    cam.horizontalFov = 0.7f;
    scene.setCamera(cam);
    std::size_t ids[64];
    for (std::size_t k = 0; k < 64; k++) {
        MaterialDescription m;
        m.albedo = RGB(0.1f + 0.8f * uniform(), 0.1f + 0.8f * uniform(), 0.1f + 0.8f * uniform());
        if (uniform() < 0.5f)
            m.reflectionTint = RGB(0.5f + 0.5f * uniform(), 0.5f + 0.5f * uniform(), 0.5f + 0.5f * uniform());
        m.roughness = 0.01f + 0.59f * uniform();
        m.ior = 1.2f + 0.8f * uniform();
        if (uniform() < 0.02f)
            m.emissive = RGB(15, 15, 15);
        ids[k] = scene.addMaterial(m);
    }
    PlaneDescription floor;
    floor.normal = V3(0, 1, 0);
    floor.point = V3(0, 0, 0);
    floor.extents = V3(4000, 4000, 0);
    floor.material = ids[1];
    scene.addPlane(floor);
    for (std::size_t i = 0; i < n; i++) {
        SphereDescription s;
        float const x = 2000.0f * uniform() - 1000.0f, y = 2000.0f * uniform(), z = 2000.0f * uniform() - 1000.0f;
        s.center = V3(x, y, z);
        s.radius = 5.0f + 20.0f * uniform();
        s.material = ids[i % 64];
        scene.addSphere(s);
    }
    return scene;
}

static void usage() {
    std::puts("usage: cornelis [--width W] [--height H] [--spp N] [--aspect A] [--seed S] [--max-depth D]\n"
              "                [--devices G] [--pool P] [--output file.png] [--no-save] [--drop-nonfinite] [--quiet]\n"
              "                [--scene cornell|spheres] [--spheres N] [--accel auto|none|grid]\n"
              "                [--progressive] [--time-budget SECONDS]   (spp is then the upper limit)\n"
              "                [--dump-raw file.f32]   (the float frame buffer: height*width*3 floats, row-major)\n"
              "defaults reproduce the reference CLI: the Cornell box, 512x512, 4096 spp, cornelisrender2.png");
}

int main(int argc, char *argv[]) {
    RenderOptions options;
    options.samplesAA = 4096;
    float aspect = -1.0f;
    bool quiet = false;
    std::string sceneName = "cornell";
    std::size_t sphereCount = 10000;
    std::string rawPath;
    for (int i = 1; i < argc; i++) {
        std::string const a = argv[i];
        auto next = [&]() -> char const * {
            if (i + 1 >= argc) {
                usage();
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "--width") options.width = std::atoi(next());
        else if (a == "--height") options.height = std::atoi(next());
        else if (a == "--spp") options.samplesAA = std::atoi(next());
        else if (a == "--aspect") aspect = static_cast<float>(std::atof(next()));
        else if (a == "--seed") options.seed = std::strtoull(next(), nullptr, 10);
        else if (a == "--max-depth") options.maxDepth = std::atoi(next());
        else if (a == "--devices") options.devices = std::atoi(next());
        else if (a == "--pool") options.poolPaths = std::atoi(next());
        else if (a == "--output") options.outputPath = next();
        else if (a == "--no-save") options.saveImage = false;
        else if (a == "--drop-nonfinite") options.dropNonFinite = true;
        else if (a == "--quiet") quiet = true;
        else if (a == "--progressive") options.progressive = true;
        else if (a == "--time-budget") options.timeBudgetSeconds = std::atof(next());
        else if (a == "--dump-raw") rawPath = next();
        else if (a == "--scene") sceneName = next();
        else if (a == "--spheres") sphereCount = static_cast<std::size_t>(std::atoll(next()));
        else if (a == "--accel") {
            std::string const m = next();
            options.acceleration = m == "none" ? Acceleration::None : m == "grid" ? Acceleration::Grid : Acceleration::Auto;
        }
        else {
            usage();
            return a == "--help" || a == "-h" ? 0 : 2;
        }
    }
    if (aspect <= 0.0f) // square pixels: the camera's aspect scales the vertical film vector
        aspect = options.width > 0 ? static_cast<float>(options.height) / static_cast<float>(options.width) : 1.0f;
    try {
        if (sceneName != "cornell" && sceneName != "spheres") {
            usage();
            return 2;
        }
        RenderSession session(sceneName == "spheres" ? manySpheres(sphereCount, options.seed, aspect) : cornellBox(aspect),
                              options);
        auto const t0 = std::chrono::steady_clock::now();
        int lastDecile = -1;
        session.render([&](RenderProgress const &p, RenderStatus const &status) {
            if (!quiet && p.samplesTotal) {
                int const decile = static_cast<int>(10 * p.samplesDone / p.samplesTotal);
                if (decile != lastDecile && status == RenderStatus::Running) {
                    lastDecile = decile;
                    std::fprintf(stderr, "%d%% done..\n", decile * 10);
                }
            }
            return RenderCommand::Continue;
        });
        double const wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        auto const &st = session.statistics();
        // Non-finite pixels: the reference's Oren-Nayar term is NaN when |w.z| rounds above 1 (about once per 1e8
        // samples) and the pixel then stays NaN; it is kept unless --drop-nonfinite.  In the PNG such a pixel is black:
        // std::clamp passes NaN through and the conversion to 8 bits gives 0 (formally undefined in C++; x86's cvttsd2si
        // and the device's cvt both yield 0, and so does the reference's binary).
        std::size_t nonFinite = 0;
        for (auto const &px : session.frameBuffer())
            if (!std::isfinite(px(0)) || !std::isfinite(px(1)) || !std::isfinite(px(2)))
                nonFinite++;
        if (!rawPath.empty()) {
            std::FILE *f = std::fopen(rawPath.c_str(), "wb");
            auto const &fb = session.frameBuffer();
            std::size_t const n = static_cast<std::size_t>(fb.width()) * static_cast<std::size_t>(fb.height());
            if (!f || std::fwrite(fb.data(), 3 * sizeof(float), n, f) != n) {
                std::fprintf(stderr, "cornelis: cannot write %s\n", rawPath.c_str());
                return 1;
            }
            std::fclose(f);
        }
        if (!quiet && nonFinite)
            std::printf("%zu non-finite pixel(s) (kept as the reference keeps them; black in the PNG; --drop-nonfinite "
                        "skips the offending paths)\n", nonFinite);
        if (!quiet)
            std::printf("%dx%d, %d spp, %d device(s): %.3f s wall, %.3f s on the GPU, %.1f Msamples/s, %.1f Mrays/s, "
                        "%.3f rays/sample, deepest path %u%s%s\n",
                        options.width, options.height, st.samplesPerPixel, options.devices, wall, st.gpuSeconds,
                        st.pixelSamples / st.gpuSeconds / 1e6, st.rays / st.gpuSeconds / 1e6,
                        static_cast<double>(st.rays) / static_cast<double>(st.pixelSamples), st.maxDepth,
                        options.saveImage ? ", wrote " : "", options.saveImage ? options.outputPath.c_str() : "");
    } catch (std::exception const &e) {
        std::fprintf(stderr, "cornelis: %s\n", e.what());
        return 1;
    }
    return 0;
}
