// What the communication layer (comm.cu) needs to know about a scene handle, whose definition stays private to api.cu.
#pragma once

#include <cstddef>
#include <string>

#include <cuda_runtime.h>

#include "../../include/cornelis_cuda.h"
#include "wavefront.h"

namespace cornelis_b200 {

// The accumulation images of a handle that has rendered: float4 per pixel (sum r, g, b, contributing paths), and the
// second-moment image when the last render accumulated variance.
struct FrameView {
    int device = 0;
    cudaStream_t stream = nullptr;
    float4 *accum = nullptr, *accum2 = nullptr; // accum2 is null unless haveVariance
    size_t npixels = 0;
    bool haveVariance = false;
    const LaunchShape *shape = nullptr;
};

// False when the handle is null or has not rendered a frame yet.
bool frameView(cornelis_cuda_scene *scene, FrameView &out);

// Records the calling thread's error message (cornelis_cuda_last_error) and returns `code`.
int failWith(cornelis_status code, const std::string &message);

} // namespace cornelis_b200
