// RenderSession on top of the C-ABI (include/cornelis_cuda.h).
//
// Where the reference's RenderSession::render builds a FrameTiling, seeds one PRNG per tile and runs integrateTile on
// TBB workers (reference src/Render.cpp:302-363), this one hands the whole frame to the GPU wavefront: one scene handle
// per device, each device renders a contiguous range of the global sample indices of every pixel, the accumulators
// are summed on device 0 and resolved into the RGBFrameBuffer.  No CPU rendering path exists here.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cornelis/Render.hpp>
#include <cornelis_cuda.h>

namespace cornelis {

bool writePngRgb8(std::string const &path, int width, int height, unsigned char const *rgb);

namespace {

[[noreturn]] void raise(int code, char const *where) {
    throw RenderError(code, std::string(where) + ": " + cornelis_cuda_last_error());
}

struct DeviceScene {
    cornelis_cuda_scene *handle = nullptr;
    ~DeviceScene() {
        if (handle)
            cornelis_cuda_scene_destroy(handle);
    }
};

} // namespace

struct RenderSession::State {
    State(SceneDescription const &sc, RenderOptions opts)
        : sceneDescr(sc), options(std::move(opts)), fb(PixelRect(std::max(options.width, 1), std::max(options.height, 1))) {
        CORNELIS_EXPECTS(options.width > 0 && options.height > 0, "Frame cannot be a line or the empty rectangle.");
        CORNELIS_EXPECTS(options.devices > 0, "At least one device is needed.");
        int available = 0;
        if (int rc = cornelis_cuda_device_count(&available))
            raise(rc, "RenderSession");
        if (options.devices > available)
            throw RenderError(CORNELIS_ERR_INVALID_ARGUMENT, "RenderSession: more devices requested than this box has");
        upload();
    }

    // SceneDescription -> PODs of the C-ABI (the flattening SceneData does in reference src/Scene.cpp:5-53).
    void upload() {
        auto const cam = sceneDescr.camera();
        cornelis_camera_desc c{};
        for (int k = 0; k < 3; k++) {
            c.origin[k] = cam.origin[k];
            c.look_at[k] = cam.lookAt[k];
        }
        c.aspect = cam.aspect;
        c.horizontal_fov = cam.horizontalFov;
        std::vector<cornelis_material_desc> mats;
        for (auto const &m : sceneDescr.materials()) {
            cornelis_material_desc d{};
            for (int k = 0; k < 3; k++) {
                d.albedo[k] = m.albedo(k);
                d.emissive[k] = m.emissive(k);
                d.reflection_tint[k] = m.reflectionTint(k);
            }
            d.roughness = m.roughness;
            d.ior = m.ior;
            mats.push_back(d);
        }
        std::vector<cornelis_sphere_desc> spheres;
        for (auto const &s : sceneDescr.spheres()) {
            cornelis_sphere_desc d{};
            for (int k = 0; k < 3; k++)
                d.center[k] = s.center[k];
            d.radius = s.radius;
            d.material = s.material ? static_cast<int32_t>(*s.material) : -1;
            spheres.push_back(d);
        }
        std::vector<cornelis_plane_desc> planes;
        for (auto const &p : sceneDescr.planes()) {
            cornelis_plane_desc d{};
            for (int k = 0; k < 3; k++) {
                d.normal[k] = p.normal[k];
                d.point[k] = p.point[k];
                d.extents[k] = p.extents[k];
            }
            d.material = p.material ? static_cast<int32_t>(*p.material) : -1;
            planes.push_back(d);
        }
        devices.resize(static_cast<std::size_t>(options.devices));
        for (int dev = 0; dev < options.devices; dev++) {
            if (int rc = cornelis_cuda_scene_create(dev, &c, spheres.data(), spheres.size(), planes.data(), planes.size(),
                                                    mats.data(), mats.size(), &devices[static_cast<std::size_t>(dev)].handle))
                raise(rc, "cornelis_cuda_scene_create");
            if (options.acceleration != Acceleration::Auto)
                if (int rc = cornelis_cuda_scene_set_acceleration(devices[static_cast<std::size_t>(dev)].handle,
                                                                  static_cast<int>(options.acceleration)))
                    raise(rc, "cornelis_cuda_scene_set_acceleration");
        }
    }

    SceneDescription sceneDescr;
    RenderOptions options;
    RGBFrameBuffer fb;
    RenderStatistics stats;
    std::vector<DeviceScene> devices;
};

RenderSession::RenderSession(SceneDescription const &sc, RenderOptions options)
    : me_{std::make_unique<State>(sc, std::move(options))} {}

RenderSession::~RenderSession() = default;
RenderSession::RenderSession(RenderSession &&) noexcept = default;
RenderSession &RenderSession::operator=(RenderSession &&) noexcept = default;

void RenderSession::render() {
    render([](RenderProgress const &, RenderStatus const &) { return RenderCommand::Continue; });
}

RGBFrameBuffer const &RenderSession::frameBuffer() const { return me_->fb; }
RenderStatistics const &RenderSession::statistics() const { return me_->stats; }

namespace {
struct ProgressRelay {
    RenderSession::ProgressCallback *callback;
    std::vector<std::atomic<std::uint64_t>> *done;
    std::size_t device;
    std::uint64_t total;
    std::atomic<bool> *abort;
};

int relay(void *user, std::uint64_t samplesDone, std::uint64_t) {
    auto *r = static_cast<ProgressRelay *>(user);
    (*r->done)[r->device].store(samplesDone);
    RenderProgress p;
    for (auto const &d : *r->done)
        p.samplesDone += d.load();
    p.samplesTotal = r->total;
    if ((*r->callback)(p, RenderStatus::Running) != RenderCommand::Continue)
        r->abort->store(true);
    return r->abort->load() ? 1 : 0;
}
} // namespace

void RenderSession::render(ProgressCallback onProgress) {
    State &s = *me_;
    RenderOptions const &o = s.options;
    if (o.samplesAA <= 0) { // reference src/Render.cpp:310-313: message and silent return
        std::printf("AA Samples must be > 0 (not %d).\n", o.samplesAA);
        return;
    }
    std::size_t const n = s.devices.size();
    std::uint64_t const total = static_cast<std::uint64_t>(o.width) * static_cast<std::uint64_t>(o.height) *
                                static_cast<std::uint64_t>(o.samplesAA);
    std::vector<std::atomic<std::uint64_t>> done(n);
    for (auto &d : done)
        d.store(0);
    std::atomic<bool> abort{false};
    std::vector<int> rc(n, 0);
    std::vector<std::string> errors(n);
    std::vector<cornelis_render_stats> stats(n);

    auto work = [&](std::size_t dev) {
        // contiguous sample range of this device; the counter-based RNG makes the union identical for any count
        std::int32_t const first = static_cast<std::int32_t>(static_cast<std::int64_t>(o.samplesAA) * static_cast<std::int64_t>(dev) / static_cast<std::int64_t>(n));
        std::int32_t const last = static_cast<std::int32_t>(static_cast<std::int64_t>(o.samplesAA) * static_cast<std::int64_t>(dev + 1) / static_cast<std::int64_t>(n));
        if (last == first) {
            rc[dev] = -1; // nothing to do on this device
            return;
        }
        cornelis_render_params p{};
        p.width = o.width;
        p.height = o.height;
        p.samples = o.samplesAA;
        p.first_sample = first;
        p.sample_count = last - first;
        p.max_depth = o.maxDepth;
        p.seed = o.seed;
        p.flags = o.dropNonFinite ? CORNELIS_RENDER_DROP_NONFINITE : 0u;
        p.pool_paths = o.poolPaths;
        ProgressRelay r{&onProgress, &done, dev, total, &abort};
        rc[dev] = cornelis_cuda_render_accumulate(s.devices[dev].handle, &p, relay, &r, &stats[dev]);
        if (rc[dev])
            errors[dev] = cornelis_cuda_last_error();
    };
    if (n == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (std::size_t dev = 0; dev < n; dev++)
            pool.emplace_back(work, dev);
        for (auto &t : pool)
            t.join();
    }

    RenderProgress final;
    final.samplesTotal = total;
    bool aborted = false;
    for (std::size_t dev = 0; dev < n; dev++) {
        if (rc[dev] == CORNELIS_ERR_ABORTED)
            aborted = true;
        else if (rc[dev] > 0) {
            onProgress(final, RenderStatus::Failed);
            throw RenderError(rc[dev], "cornelis_cuda_render_accumulate: " + errors[dev]);
        }
    }

    s.stats = RenderStatistics{};
    std::vector<cornelis_cuda_scene *> rendered;
    for (std::size_t dev = 0; dev < n; dev++) {
        if (rc[dev] == -1)
            continue;
        rendered.push_back(s.devices[dev].handle);
        s.stats.pixelSamples += stats[dev].pixel_samples;
        s.stats.rays += stats[dev].rays;
        s.stats.passes += stats[dev].iterations;
        s.stats.kernelLaunches += stats[dev].kernel_launches;
        s.stats.maxDepth = std::max(s.stats.maxDepth, stats[dev].max_depth);
        s.stats.gpuSeconds = std::max(s.stats.gpuSeconds, 1e-3 * static_cast<double>(stats[dev].gpu_ms));
    }
    final.samplesDone = s.stats.pixelSamples;
    if (rendered.size() > 1)
        if (int e = cornelis_cuda_reduce_framebuffers(rendered.data(), static_cast<int>(rendered.size())))
            raise(e, "cornelis_cuda_reduce_framebuffers");
    // RGB is three packed floats, exactly the host_rgb layout of the C-ABI
    if (int e = cornelis_cuda_resolve(rendered.front(), o.samplesAA, reinterpret_cast<float *>(s.fb.data()), nullptr))
        raise(e, "cornelis_cuda_resolve");

    onProgress(final, aborted ? RenderStatus::Aborted : RenderStatus::Done);
    if (o.saveImage)
        saveImage(s.fb, o.outputPath);
}

void saveImage(RGBFrameBuffer const &fb, std::string const &path) {
    SRGBFrameBuffer display(PixelRect(fb.width(), fb.height()));
    std::transform(fb.begin(), fb.end(), display.begin(), toSRGB);
    auto const bytes = quantizeTo8bit(display);
    if (!writePngRgb8(path, bytes.width(), bytes.height(), reinterpret_cast<unsigned char const *>(bytes.data())))
        throw std::runtime_error("saveImage: cannot write " + path);
}

} // namespace cornelis
