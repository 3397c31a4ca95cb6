// vkgs::Camera - same public surface, defaults and semantics as the reference's orbit camera
// (include/vkgs/scene/camera.h:8-58, src/vkgs/scene/camera.cc:25-70), without the glm dependency: matrices are
// column-major std::array<float,16> (m[c*4+r]), bit-compatible with glm::mat4.
#ifndef VKGS_SCENE_CAMERA_H
#define VKGS_SCENE_CAMERA_H

#include <array>
#include <cstdint>

#ifndef VKGS_API
#define VKGS_API __attribute__((visibility("default")))
#endif

namespace vkgs {

using Mat4 = std::array<float, 16>;
using Vec3 = std::array<float, 3>;

class VKGS_API Camera {
 public:
  static constexpr float min_fov() { return 40.f * 0.01745329251994329576923690768489f; }
  static constexpr float max_fov() { return 100.f * 0.01745329251994329576923690768489f; }

  Camera();
  ~Camera();

  float Near() const noexcept { return lens_.z_near; }
  float Far() const noexcept { return lens_.z_far; }

  void SetWindowSize(uint32_t width, uint32_t height);
  /** Set fov and dolly zoom.  fov: fov Y, in radians */
  void SetFov(float fov);

  Mat4 ProjectionMatrix() const;
  Mat4 ViewMatrix() const;
  Vec3 Eye() const;
  uint32_t width() const noexcept { return window_.w; }
  uint32_t height() const noexcept { return window_.h; }
  float fov() const noexcept { return lens_.fovy; }

  void Rotate(float x, float y);
  void Translate(float x, float y, float z = 0.f);
  void Zoom(float x);
  void DollyZoom(float scroll);

  // Extension (the reference reaches these only through the mouse handlers in Engine::Impl::Draw).
  void SetOrbit(const Vec3& center, float r, float phi, float theta);

 private:
  static constexpr float kDeg = 0.01745329251994329576923690768489f;

  struct Viewport {
    uint32_t w = 256, h = 256;
  } window_;
  struct Lens {
    float fovy = 60.f * kDeg, z_near = 0.01f, z_far = 100.f;
  } lens_;
  // eye = target + radius * (sin(polar) sin(azimuth), cos(polar), sin(polar) cos(azimuth)); +Y is up
  struct Orbit {
    Vec3 target = {0.f, 0.f, 0.f};
    float radius = 2.f, polar = 45.f * kDeg, azimuth = 45.f * kDeg;
  } orbit_;
  struct Sensitivity {
    float rotate = 0.01f, pan = 0.002f, zoom = 0.01f, dolly = kDeg;
  } sens_;
};

}  // namespace vkgs

#endif  // VKGS_SCENE_CAMERA_H
