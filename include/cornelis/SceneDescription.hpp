// User-facing scene description (reference include/cornelis/SceneDescription.hpp:14-92): plain builders, material 0 is
// always a default material, objects without a material use it.
#pragma once

#include <cstddef>
#include <optional>
#include <vector>

#include <cornelis/Color.hpp>
#include <cornelis/Math.hpp>
#include <cornelis/Span.hpp>

namespace cornelis {

struct MaterialDescription {
    RGB albedo = RGB(0.5f, 0.5f, 0.5f);
    RGB emissive = RGB::black();
    float roughness = 0.2f;
    RGB reflectionTint = RGB::black();
    float ior = 1.5f;

    bool operator==(MaterialDescription const &o) const {
        return albedo == o.albedo && emissive == o.emissive && roughness == o.roughness &&
               reflectionTint == o.reflectionTint && ior == o.ior;
    }
};

struct ObjectDescription {
    std::optional<std::size_t> material;
    bool operator==(ObjectDescription const &o) const { return material == o.material; }
};

struct SphereDescription : public ObjectDescription {
    V3 center;
    float radius = 1.0f;
    bool operator==(SphereDescription const &o) const {
        return ObjectDescription::operator==(o) && center == o.center && radius == o.radius;
    }
};

struct PlaneDescription : public ObjectDescription {
    V3 normal{0.0f, 1.0f, 0.0f};
    V3 point{0.0f, 0.0f, 0.0f};
    V3 extents{1000.f, 1000.f, 0.0f}; // [0] width along the tangent, [1] height along the bi-tangent
    bool operator==(PlaneDescription const &o) const {
        return ObjectDescription::operator==(o) && normal == o.normal && point == o.point && extents == o.extents;
    }
};

struct PerspectiveCameraDescription : public ObjectDescription {
    V3 origin;
    V3 lookAt{0.0f, 0.0f, 1.0f};
    float aspect = 0.5f;
    float horizontalFov = 1.011f;
    bool operator==(PerspectiveCameraDescription const &o) const {
        return ObjectDescription::operator==(o) && origin == o.origin && lookAt == o.lookAt && aspect == o.aspect &&
               horizontalFov == o.horizontalFov;
    }
};

class SceneDescription {
  public:
    SceneDescription() = default;

    void setCamera(PerspectiveCameraDescription const &cam) { camera_ = cam; }

    std::size_t addMaterial(MaterialDescription const &mat) {
        materials_.push_back(mat);
        return materials_.size() - 1;
    }
    std::size_t addSphere(SphereDescription const &sphere) {
        spheres_.push_back(sphere);
        return spheres_.size() - 1;
    }
    std::size_t addPlane(PlaneDescription const &plane) {
        planes_.push_back(plane);
        return planes_.size() - 1;
    }

    span<const MaterialDescription> materials() const noexcept { return materials_; }
    span<const SphereDescription> spheres() const noexcept { return spheres_; }
    span<const PlaneDescription> planes() const noexcept { return planes_; }
    PerspectiveCameraDescription camera() const noexcept { return camera_; }

  private:
    PerspectiveCameraDescription camera_;
    std::vector<MaterialDescription> materials_ = {MaterialDescription{}};
    std::vector<SphereDescription> spheres_;
    std::vector<PlaneDescription> planes_;
};

} // namespace cornelis
