// RenderSession: the public render API (reference include/cornelis/Render.hpp:10-49).  Same names, same calling
// convention; the work behind render() is the CUDA wavefront behind include/cornelis_cuda.h instead of the TBB tile
// loop of the reference's src/Render.cpp:302-363.  There is no CPU fallback: construction or render() throws
// RenderError when no CUDA device is usable.
#pragma once

#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>

#include <cornelis/FrameBuffer.hpp>
#include <cornelis/RenderOptions.hpp>
#include <cornelis/SceneDescription.hpp>

namespace cornelis {

enum class RenderCommand {
    Continue,
    Abort,
};
enum class RenderStatus { Running, Done, Aborted, Failed };

// Empty in the reference (Render.hpp:15); filled in here.
struct RenderProgress {
    std::uint64_t samplesDone = 0;  // camera paths started so far, over all devices
    std::uint64_t samplesTotal = 0; // width * height * samplesAA
};

struct RenderStatistics {
    std::uint64_t pixelSamples = 0;
    std::uint64_t rays = 0;
    std::uint64_t passes = 0;
    std::uint64_t kernelLaunches = 0;
    std::uint32_t maxDepth = 0;
    double gpuSeconds = 0.0; // slowest device (summed over the slices of a progressive render)
    std::int32_t samplesPerPixel = 0; // samples the frame buffer holds: samplesAA unless aborted or out of time budget
    std::uint32_t slices = 0;         // sample slices rendered (1 unless progressive)
};

class RenderError : public std::runtime_error {
  public:
    RenderError(int code, std::string const &what) : std::runtime_error(what), code_(code) {}
    int code() const noexcept { return code_; }

  private:
    int code_;
};

class RenderSession {
  public:
    using ProgressCallback = std::function<RenderCommand(RenderProgress const &, RenderStatus const &)>;

    RenderSession(SceneDescription const &, RenderOptions options);
    ~RenderSession();

    RenderSession(RenderSession &&) noexcept;
    RenderSession &operator=(RenderSession &&) noexcept;
    RenderSession(RenderSession const &) = delete;
    RenderSession &operator=(RenderSession const &) = delete;

    // Shorthand: render until completion.
    void render();

    // Starts the render and calls onProgress as it advances; blocks until the render stops.  onProgress is always
    // called at least once; it may be called from different threads (one per device) and has to be thread-safe.
    // Returning RenderCommand::Abort stops the render.  The last call reports Done (or Aborted / Failed).
    void render(ProgressCallback onProgress);

    // The image of the last render() (row-major, j = 0 on top).
    RGBFrameBuffer const &frameBuffer() const;
    RenderStatistics const &statistics() const;

  private:
    struct State;
    std::unique_ptr<State> me_;
};

// Display transform + 8-bit quantisation + PNG, what the reference's saveImage does (Render.cpp:257-265).
void saveImage(RGBFrameBuffer const &fb, std::string const &path);

} // namespace cornelis
