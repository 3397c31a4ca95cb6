// vkgs::Engine - drop-in for the reference's public class (include/vkgs/engine/engine.h:11-26): the same five
// methods with the same threading contract, implemented on libvkgsb (CUDA, headless) instead of Vulkan + GLFW.
// Extensions the north-star asks for (the reference has no programmatic camera, viewport or read-back):
// camera(), SetViewport(), SetModel(), SetOverlay(), DrawToImage(), stats().
#ifndef VKGS_ENGINE_ENGINE_H
#define VKGS_ENGINE_ENGINE_H

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include <vkgs/scene/camera.h>

#ifndef VKGS_API
#define VKGS_API __attribute__((visibility("default")))
#endif

namespace vkgs {

struct FrameStats {
  uint32_t total_point_count = 0, loaded_point_count = 0, visible_point_count = 0;
  float project_ms = 0, sort_ms = 0, bin_ms = 0, blend_ms = 0, total_ms = 0;
  uint64_t frame_counter = 0;
};

class VKGS_API Engine {
 public:
  // Throws std::runtime_error when no CUDA device is usable (the reference throws "No GPU found", context.cc:89).
  Engine();
  explicit Engine(int device, uint32_t max_splats = 1u << 23);
  ~Engine();

  void LoadSplats(const std::string& ply_filepath);       // starts the load, cancelling one in flight
  void LoadSplatsAsync(const std::string& ply_filepath);  // thread-safe; picked up by Run()

  // Headless render loop: draws offscreen frames at the current camera until Close(); re-entrant after Close.
  void Run();
  void Close();  // callable from any thread

  // ---- extensions ----
  Camera& camera();
  void SetViewport(uint32_t width, uint32_t height);  // default 1600 x 900 (viewer.cc:67)
  void SetModel(const Mat4& model);
  void SetBlendMode(int vkgsb_blend_mode_value);
  // The viewer's axis and grid (show_axis_ / show_grid_, engine.cc:1664-1665; geometry engine.cc:618-680) drawn under
  // the splats with depth test + write, the splats depth-tested against them.  Off until asked for: a headless
  // frame is the splats alone.
  void SetOverlay(bool show_axis, bool show_grid);
  void WaitForLoad();
  // One frame at the current camera; RGBA8, width*height*4 bytes.
  void DrawToImage(std::vector<uint8_t>* rgba);
  FrameStats stats() const;

 private:
  class Impl;
  std::shared_ptr<Impl> impl_;
};

}  // namespace vkgs

#endif  // VKGS_ENGINE_ENGINE_H
