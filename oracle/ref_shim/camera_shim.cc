// C shim over the reference's own vkgs::Camera (src/vkgs/scene/camera.cc, include/vkgs/scene/camera.h),
// compiled from /root/reference by oracle/build_ref.py into oracle/_ref/libref_camera.so.
// TEST INFRASTRUCTURE ONLY.  The reference exposes r/phi/theta only through its mouse operations, so the
// shim drives those (Rotate/Zoom/Translate with the public sensitivities) to reach a requested pose.
#include <vkgs/scene/camera.h>

#include <cstring>

extern "C" {

// Default-constructed reference camera (r=2, phi=theta=45deg, fovy 60deg) at a window size.
__attribute__((visibility("default"))) void ref_camera_default(unsigned w, unsigned h, float* proj16, float* view16,
                                                                float* eye3) {
  vkgs::Camera cam;
  cam.SetWindowSize(w, h);
  glm::mat4 p = cam.ProjectionMatrix(), v = cam.ViewMatrix();
  glm::vec3 e = cam.Eye();
  std::memcpy(proj16, &p[0][0], 64);
  std::memcpy(view16, &v[0][0], 64);
  std::memcpy(eye3, &e[0], 12);
}

// Apply Rotate(dx,dy), Zoom(z), SetFov(fov) (fov<=0: keep) to a default camera, in that order.
__attribute__((visibility("default"))) void ref_camera_ops(unsigned w, unsigned h, float rot_x, float rot_y, float zoom,
                                                            float fov, float tx, float ty, float tz, float* proj16,
                                                            float* view16, float* eye3) {
  vkgs::Camera cam;
  cam.SetWindowSize(w, h);
  cam.Rotate(rot_x, rot_y);
  cam.Zoom(zoom);
  if (fov > 0.f) cam.SetFov(fov);
  cam.Translate(tx, ty, tz);
  glm::mat4 p = cam.ProjectionMatrix(), v = cam.ViewMatrix();
  glm::vec3 e = cam.Eye();
  std::memcpy(proj16, &p[0][0], 64);
  std::memcpy(view16, &v[0][0], 64);
  std::memcpy(eye3, &e[0], 12);
}
}
