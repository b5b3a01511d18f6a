// Idealised perspective camera (reference include/cornelis/Camera.hpp, src/Camera.cpp): camera space +Z looks at
// the subject, film coordinates x, y in [0, 1] with y = 0 at the TOP of the image.
#pragma once

#include <memory>

#include <cornelis/Expects.hpp>
#include <cornelis/Math.hpp>

namespace cornelis {

class PerspectiveCamera {
  public:
    PerspectiveCamera();

    // World-space ray through film position (x, y).  Thread-safe.
    Ray operator()(float x, float y) const noexcept;

    // Camera at `from` looking at `at`.  hFov: horizontal field of view in radians.  NOTE: aspectRatio scales the
    // VERTICAL film vector (reference Camera.cpp:25), i.e. pass height / width for square pixels.
    static PerspectiveCamera lookAt(V3 const &from, V3 const &at, float aspectRatio, float hFov);

    V3 const &eye() const noexcept { return eye_; }
    V3 const &corner() const noexcept { return corner_; }
    V3 const &u() const noexcept { return u_; }
    V3 const &v() const noexcept { return v_; }

  private:
    V3 eye_, corner_, u_, v_;
};

using PerspectiveCameraPtr = std::shared_ptr<PerspectiveCamera>;

inline constexpr float HorizontalFovNormal = 1.011f; // ~43 mm lens on 35 mm film

// Horizontal field of view of a 35 mm camera lens.  Throws ExpectationException unless focalLength > 0.
float horizontalFov35mm(float focalLength);

} // namespace cornelis
