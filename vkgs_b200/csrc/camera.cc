// vkgs::Camera (see include/vkgs/scene/camera.h).  Follows src/vkgs/scene/camera.cc:25-70; glm::perspective is the
// RH / depth -1..1 variant (third_party/glm/glm/ext/matrix_clip_space.inl:249-262) and glm::lookAt the RH variant.
#include <vkgs/scene/camera.h>

#include <algorithm>
#include <cmath>

#include "../../include/vkgsb.h"

namespace vkgs {

namespace {
constexpr float kPi = 3.14159265358979323846264338327950288f;
inline float Radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

Vec3 Normalize(const Vec3& v) {
  float inv = 1.f / std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);  // glm: v * inversesqrt(dot(v, v))
  return {v[0] * inv, v[1] * inv, v[2] * inv};
}
Vec3 Cross(const Vec3& a, const Vec3& b) {
  return {a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]};
}
float Dot(const Vec3& a, const Vec3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
}  // namespace

Camera::Camera() {}
Camera::~Camera() {}

void Camera::SetWindowSize(uint32_t width, uint32_t height) {
  window_.w = width;
  window_.h = height;
}

void Camera::SetFov(float fov) {
  orbit_.radius *= std::tan(lens_.fovy / 2.f) / std::tan(fov / 2.f);  // dolly zoom
  lens_.fovy = fov;
}

Mat4 Camera::ProjectionMatrix() const {
  float aspect = static_cast<float>(window_.w) / window_.h;
  const float t = std::tan(lens_.fovy / 2.f);
  Mat4 p{};  // glm::perspective(fovy, aspect, near, far)
  p[0 * 4 + 0] = 1.f / (aspect * t);
  p[1 * 4 + 1] = 1.f / t;
  p[2 * 4 + 2] = -(lens_.z_far + lens_.z_near) / (lens_.z_far - lens_.z_near);
  p[2 * 4 + 3] = -1.f;
  p[3 * 4 + 2] = -(2.f * lens_.z_far * lens_.z_near) / (lens_.z_far - lens_.z_near);
  // gl to vulkan: conversion * projection with conversion[1][1] = -1, [2][2] = 0.5, [3][2] = 0.5
  Mat4 out{};
  for (int c = 0; c < 4; ++c) {
    out[c * 4 + 0] = p[c * 4 + 0];
    out[c * 4 + 1] = -p[c * 4 + 1];
    out[c * 4 + 2] = 0.5f * p[c * 4 + 2] + 0.5f * p[c * 4 + 3];
    out[c * 4 + 3] = p[c * 4 + 3];
  }
  return out;
}

Vec3 Camera::Eye() const {
  const float sin_phi = std::sin(orbit_.polar), cos_phi = std::cos(orbit_.polar);
  const float sin_theta = std::sin(orbit_.azimuth), cos_theta = std::cos(orbit_.azimuth);
  return {orbit_.target[0] + orbit_.radius * (sin_phi * sin_theta), orbit_.target[1] + orbit_.radius * cos_phi, orbit_.target[2] + orbit_.radius * (sin_phi * cos_theta)};
}

Mat4 Camera::ViewMatrix() const {
  const Vec3 eye = Eye();
  const Vec3 f = Normalize({orbit_.target[0] - eye[0], orbit_.target[1] - eye[1], orbit_.target[2] - eye[2]});
  const Vec3 s = Normalize(Cross(f, {0.f, 1.f, 0.f}));
  const Vec3 u = Cross(s, f);
  Mat4 m{};
  m[0 * 4 + 0] = s[0]; m[1 * 4 + 0] = s[1]; m[2 * 4 + 0] = s[2];
  m[0 * 4 + 1] = u[0]; m[1 * 4 + 1] = u[1]; m[2 * 4 + 1] = u[2];
  m[0 * 4 + 2] = -f[0]; m[1 * 4 + 2] = -f[1]; m[2 * 4 + 2] = -f[2];
  m[3 * 4 + 0] = -Dot(s, eye);
  m[3 * 4 + 1] = -Dot(u, eye);
  m[3 * 4 + 2] = Dot(f, eye);
  m[3 * 4 + 3] = 1.f;
  return m;
}

void Camera::Rotate(float x, float y) {
  orbit_.azimuth -= sens_.rotate * x;
  float eps = Radians(0.1f);
  orbit_.polar = std::clamp(orbit_.polar - sens_.rotate * y, eps, kPi - eps);
}

void Camera::Translate(float x, float y, float z) {
  const float sin_phi = std::sin(orbit_.polar), cos_phi = std::cos(orbit_.polar);
  const float sin_theta = std::sin(orbit_.azimuth), cos_theta = std::cos(orbit_.azimuth);
  const float k = sens_.pan * orbit_.radius;
  const Vec3 ax = {cos_theta, 0.f, -sin_theta}, ay = {-cos_phi * sin_theta, sin_phi, -cos_phi * cos_theta};
  const Vec3 az = {sin_phi * sin_theta, cos_phi, sin_phi * cos_theta};
  for (int i = 0; i < 3; ++i) orbit_.target[i] += k * (-x * ax[i] + y * ay[i] + -z * az[i]);
}

void Camera::Zoom(float x) { orbit_.radius /= std::exp(sens_.zoom * x); }

void Camera::DollyZoom(float scroll) {
  float new_fov = std::clamp(lens_.fovy - scroll * sens_.dolly, min_fov(), max_fov());
  SetFov(new_fov);
}

void Camera::SetOrbit(const Vec3& center, float r, float phi, float theta) {
  orbit_.target = center;
  orbit_.radius = r;
  orbit_.polar = phi;
  orbit_.azimuth = theta;
}

}  // namespace vkgs

extern "C" int vkgsb_camera_orbit(uint32_t width, uint32_t height, float fovy, float r, float phi, float theta,
                                  const float center[3], vkgsb_camera* out) {
  if (!out || width == 0 || height == 0) return VKGSB_ERR_INVALID;
  vkgs::Camera cam;
  cam.SetWindowSize(width, height);
  cam.SetOrbit(center ? vkgs::Vec3{center[0], center[1], center[2]} : vkgs::Vec3{0.f, 0.f, 0.f}, r, phi, theta);
  if (fovy > 0.f) {
    // set the field of view without the dolly-zoom radius compensation: r is given explicitly
    float keep = r;
    cam.SetFov(fovy);
    cam.SetOrbit(center ? vkgs::Vec3{center[0], center[1], center[2]} : vkgs::Vec3{0.f, 0.f, 0.f}, keep, phi, theta);
  }
  const vkgs::Mat4 p = cam.ProjectionMatrix(), v = cam.ViewMatrix();
  const vkgs::Vec3 e = cam.Eye();
  for (int i = 0; i < 16; ++i) {
    out->projection[i] = p[i];
    out->view[i] = v[i];
    out->model[i] = (i % 5 == 0) ? 1.f : 0.f;
  }
  out->camera_position[0] = e[0];
  out->camera_position[1] = e[1];
  out->camera_position[2] = e[2];
  out->pad0 = 0.f;
  return VKGSB_OK;
}

// The viewer's mouse operations on a default camera, in the order Rotate, Zoom, SetFov (fov <= 0: keep), Translate,
// DollyZoom: what Engine::Impl::Draw does with ImGui's mouse deltas (engine.cc:767-818 -> camera.cc:47-70).
extern "C" int vkgsb_camera_apply(uint32_t width, uint32_t height, float rot_x, float rot_y, float zoom, float fov,
                                  float tx, float ty, float tz, float dolly, vkgsb_camera* out) {
  if (!out || width == 0 || height == 0) return VKGSB_ERR_INVALID;
  vkgs::Camera cam;
  cam.SetWindowSize(width, height);
  cam.Rotate(rot_x, rot_y);
  cam.Zoom(zoom);
  if (fov > 0.f) cam.SetFov(fov);
  cam.Translate(tx, ty, tz);
  if (dolly != 0.f) cam.DollyZoom(dolly);
  const vkgs::Mat4 p = cam.ProjectionMatrix(), v = cam.ViewMatrix();
  const vkgs::Vec3 e = cam.Eye();
  for (int i = 0; i < 16; ++i) {
    out->projection[i] = p[i];
    out->view[i] = v[i];
    out->model[i] = (i % 5 == 0) ? 1.f : 0.f;
  }
  out->camera_position[0] = e[0];
  out->camera_position[1] = e[1];
  out->camera_position[2] = e[2];
  out->pad0 = 0.f;
  return VKGSB_OK;
}
