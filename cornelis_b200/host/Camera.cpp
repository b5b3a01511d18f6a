// PerspectiveCamera on the host (reference src/Camera.cpp).  The GPU ray generator uses the same four vectors
// (cornelis_b200/csrc/api.cu makeCamera computes them with identical arithmetic), so rays produced here for a film
// position equal the ones the raygen kernel produces.
#include <cmath>

#include <cornelis/Camera.hpp>

namespace cornelis {

// Default camera: at the origin looking down +Z, hFov = 1 rad, aspect 1 (2 sin(0.5) = 0.9588510772).
PerspectiveCamera::PerspectiveCamera()
    : eye_(0.0f), corner_(-0.4794255386f, -0.4794255386f, 1.0f), u_(0.4794255386f * 2, 0.0f, 0.0f),
      v_(0.0f, 0.4794255386f * 2, 0.0f) {}

Ray PerspectiveCamera::operator()(float x, float y) const noexcept {
    V3 d = corner_ + x * u_ + y * v_;
    d.normalize();
    return Ray(eye_, d);
}

PerspectiveCamera PerspectiveCamera::lookAt(V3 const &from, V3 const &at, float aspectRatio, float hFov) {
    V3 const up(0.0f, 1.0f, 0.0f);
    V3 dir = at - from;
    dir.normalize();
    V3 u = up.cross(dir);
    V3 v = u.cross(dir);
    float const fovScale = static_cast<float>(2.0 * std::sin(hFov * 0.5));
    u *= fovScale;
    v *= aspectRatio * fovScale;
    PerspectiveCamera cam;
    cam.eye_ = from;
    cam.corner_ = dir - 0.5f * u - 0.5f * v; // halving is exact in float and in double alike
    cam.u_ = u;
    cam.v_ = v;
    return cam;
}

float horizontalFov35mm(float focalLength) {
    CORNELIS_EXPECTS(focalLength > 0.0f, "Does not support zero or negative focal lengths.");
    return static_cast<float>(2.0 * std::atan(36.0f / (2.0 * focalLength)));
}

} // namespace cornelis
