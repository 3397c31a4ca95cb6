// vkgs::Engine facade over the C ABI.  Mirrors the control flow of the reference's Engine::Impl::Run()
// (engine.cc:551-601): poll the pending asynchronous path, start its load, size the camera to the viewport, draw -
// minus the window, swapchain and ImGui.  Errors surface as std::runtime_error from the constructor / DrawToImage
// and are otherwise swallowed by the loop, like the reference's ignored VkResults.
#include <vkgs/engine/engine.h>

#include <atomic>
#include <chrono>
#include <mutex>
#include <stdexcept>
#include <thread>

#include "../../include/vkgsb.h"

namespace vkgs {

class Engine::Impl {
 public:
  Impl(int device, uint32_t max_splats) {
    if (vkgsb_create(device, max_splats, &renderer_) != VKGSB_OK)
      throw std::runtime_error(std::string("vkgs::Engine: ") + vkgsb_last_error());
    for (int i = 0; i < 16; ++i) model_[i] = (i % 5 == 0) ? 1.f : 0.f;
  }
  ~Impl() { vkgsb_destroy(renderer_); }

  void LoadSplats(const std::string& path) { vkgsb_load_ply_async(renderer_, path.c_str()); }

  void LoadSplatsAsync(const std::string& path) {
    std::unique_lock<std::mutex> guard{mutex_};
    pending_ply_filepath_ = path;
  }

  void Run() {
    terminate_ = false;
    while (!terminate_) {
      std::string path;
      {
        std::unique_lock<std::mutex> guard{mutex_};
        path = std::move(pending_ply_filepath_);
        pending_ply_filepath_.clear();
      }
      if (!path.empty()) LoadSplats(path);
      // two frames in flight, like the reference's fences (engine.cc:1028-1035): frame i is issued once frame i - 2 has
      // finished, so camera / model / viewport changes, Close() and a concurrent DrawToImage never queue behind more
      // than two frames
      vkgsb_wait_frame(renderer_, 2);
      if (!DrawFrame(nullptr)) std::this_thread::sleep_for(std::chrono::milliseconds(1));  // nothing resident yet
    }
    vkgsb_sync(renderer_);
  }

  void Close() { terminate_ = true; }

  bool DrawFrame(void* host_dst) {
    std::unique_lock<std::mutex> guard{frame_mutex_};
    camera_.SetWindowSize(width_, height_);
    vkgsb_camera cam{};
    const Mat4 p = camera_.ProjectionMatrix(), v = camera_.ViewMatrix();
    const Vec3 e = camera_.Eye();
    for (int i = 0; i < 16; ++i) {
      cam.projection[i] = p[i];
      cam.view[i] = v[i];
      cam.model[i] = model_[i];
    }
    for (int i = 0; i < 3; ++i) cam.camera_position[i] = e[i];
    if (vkgsb_set_viewport(renderer_, width_, height_) != VKGSB_OK) return false;
    vkgsb_set_camera(renderer_, &cam);
    return vkgsb_draw(renderer_, host_dst, 0, nullptr) == VKGSB_OK;
  }

  vkgsb_renderer* renderer_ = nullptr;
  Camera camera_;
  Mat4 model_{};
  uint32_t width_ = 1600, height_ = 900;  // viewer.cc:67
  std::atomic_bool terminate_{false};
  std::mutex mutex_, frame_mutex_;
  std::string pending_ply_filepath_;
};

Engine::Engine() : impl_(std::make_shared<Impl>(0, 1u << 23)) {}
Engine::Engine(int device, uint32_t max_splats) : impl_(std::make_shared<Impl>(device, max_splats)) {}
Engine::~Engine() = default;

void Engine::LoadSplats(const std::string& ply_filepath) { impl_->LoadSplats(ply_filepath); }
void Engine::LoadSplatsAsync(const std::string& ply_filepath) { impl_->LoadSplatsAsync(ply_filepath); }
void Engine::Run() { impl_->Run(); }
void Engine::Close() { impl_->Close(); }

Camera& Engine::camera() { return impl_->camera_; }

void Engine::SetViewport(uint32_t width, uint32_t height) {
  std::unique_lock<std::mutex> guard{impl_->frame_mutex_};
  impl_->width_ = width;
  impl_->height_ = height;
}

void Engine::SetModel(const Mat4& model) {
  std::unique_lock<std::mutex> guard{impl_->frame_mutex_};
  impl_->model_ = model;
}

void Engine::SetBlendMode(int mode) { vkgsb_set_option(impl_->renderer_, VKGSB_OPT_BLEND_MODE, mode); }

void Engine::SetOverlay(bool show_axis, bool show_grid) {
  std::vector<float> pos, col;
  auto line = [&](float x0, float y0, float z0, float x1, float y1, float z1, float r, float g, float b) {
    const float p[6] = {x0, y0, z0, x1, y1, z1}, c[8] = {r, g, b, 1.f, r, g, b, 1.f};
    pos.insert(pos.end(), p, p + 6);
    col.insert(col.end(), c, c + 8);
  };
  if (show_axis) {  // engine.cc:618-627
    line(0, 0, 0, 1, 0, 0, 1, 0, 0);
    line(0, 0, 0, 0, 1, 0, 0, 1, 0);
    line(0, 0, 0, 0, 0, 1, 0, 0, 1);
  }
  if (show_grid) {  // engine.cc:635-680: 21 + 21 lines on y = 0
    constexpr int grid_size = 10;
    for (int i = -grid_size; i <= grid_size; ++i) {
      const float t = static_cast<float>(i) / grid_size;
      line(-1.f, 0, t, 1.f, 0, t, 0.5f, 0.5f, 0.5f);
      line(t, 0, -1.f, t, 0, 1.f, 0.5f, 0.5f, 0.5f);
    }
  }
  const float model[16] = {10, 0, 0, 0, 0, 10, 0, 0, 0, 0, 10, 0, 0, 0, 0, 1};  // engine.cc:1444-1448
  if (vkgsb_set_lines(impl_->renderer_, static_cast<uint32_t>(pos.size() / 6), pos.data(), col.data(), model) != VKGSB_OK)
    throw std::runtime_error(std::string("vkgs::Engine: ") + vkgsb_last_error());
}

void Engine::WaitForLoad() {
  if (vkgsb_wait_load(impl_->renderer_) != VKGSB_OK)
    throw std::runtime_error(std::string("vkgs::Engine: ") + vkgsb_last_error());
}

void Engine::DrawToImage(std::vector<uint8_t>* rgba) {
  rgba->resize(static_cast<size_t>(impl_->width_) * impl_->height_ * 4);
  if (!impl_->DrawFrame(rgba->data())) throw std::runtime_error(std::string("vkgs::Engine: ") + vkgsb_last_error());
}

FrameStats Engine::stats() const {
  vkgsb_stats s{};
  vkgsb_get_stats(impl_->renderer_, &s);
  FrameStats f;
  f.total_point_count = s.total_point_count;
  f.loaded_point_count = s.loaded_point_count;
  f.visible_point_count = s.visible_point_count;
  f.project_ms = s.ms_project;
  f.sort_ms = s.ms_sort;
  f.bin_ms = s.ms_bin;
  f.blend_ms = s.ms_blend;
  f.total_ms = s.ms_total;
  f.frame_counter = s.frame_counter;
  return f;
}

}  // namespace vkgs
