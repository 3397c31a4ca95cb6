// _pygs_cpp: native half of the `pygs` Python package (show / load / close), the drop-in for the reference's
// pybind11 module (binding/python/pygs_cpp/main.cc:16-73) on top of this repo's vkgs::Engine (CUDA, headless).
// Same contract: one process-wide engine created on the first show(), Engine::Run() on a background std::thread,
// load() forwards to Engine::LoadSplatsAsync, close() to Engine::Close, interpreter shutdown stops and joins.
// Extensions for headless use (the reference can only show a window): render(), set_orbit(), stats(), wait_loaded().
#include <pybind11/pybind11.h>

#include <condition_variable>
#include <cstdio>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

#include <vkgs/engine/engine.h>

namespace py = pybind11;

namespace {

// The render-loop session.  All state changes go through `lock`; `running` is true from show() until the loop
// thread has left Engine::Run().
class Session {
 public:
  ~Session() { Shutdown(); }

  void Show() {
    std::unique_lock<std::mutex> g(lock_);
    if (running_) {
      std::puts("[pygs] viewer is already running");
      return;
    }
    if (loop_.joinable()) loop_.join();  // a previous loop ended through close()
    EnsureEngine();
    running_ = true;
    loop_ = std::thread([this] {
      engine_->Run();
      std::puts("[pygs] bye");
      std::unique_lock<std::mutex> g2(lock_);
      running_ = false;
      idle_.notify_all();
    });
  }

  void Load(const std::string& path) {
    std::unique_lock<std::mutex> g(lock_);
    if (!engine_) return;  // the reference ignores load() before show() (main.cc:38-42)
    if (running_)
      engine_->LoadSplatsAsync(path);  // picked up by the loop, like the reference
    else
      engine_->LoadSplats(path);       // headless use without a loop: start the load right away
  }

  void Close() {
    std::unique_lock<std::mutex> g(lock_);
    if (engine_) engine_->Close();
  }

  void Shutdown() {
    {
      std::unique_lock<std::mutex> g(lock_);
      if (engine_) engine_->Close();
    }
    if (loop_.joinable()) loop_.join();
    std::unique_lock<std::mutex> g(lock_);
    engine_.reset();
  }

  vkgs::Engine& EngineForHeadless() {
    std::unique_lock<std::mutex> g(lock_);
    EnsureEngine();
    return *engine_;
  }

 private:
  void EnsureEngine() {
    if (!engine_) engine_ = std::make_unique<vkgs::Engine>();  // throws std::runtime_error without a CUDA device
  }

  std::mutex lock_;
  std::condition_variable idle_;
  std::unique_ptr<vkgs::Engine> engine_;
  std::thread loop_;
  bool running_ = false;
};

Session& session() {
  static Session* s = new Session();  // intentionally leaked; torn down by the module capsule below
  return *s;
}

py::bytes Render(uint32_t width, uint32_t height) {
  vkgs::Engine& e = session().EngineForHeadless();
  std::vector<uint8_t> rgba;
  {
    py::gil_scoped_release nogil;
    e.SetViewport(width, height);
    e.WaitForLoad();
    e.DrawToImage(&rgba);
  }
  return py::bytes(reinterpret_cast<const char*>(rgba.data()), rgba.size());
}

void SetOrbit(float cx, float cy, float cz, float r, float phi, float theta, float fovy) {
  vkgs::Engine& e = session().EngineForHeadless();
  e.camera().SetOrbit({cx, cy, cz}, r, phi, theta);
  if (fovy > 0.f) e.camera().SetFov(fovy);
}

py::dict Stats() {
  vkgs::FrameStats s = session().EngineForHeadless().stats();
  py::dict d;
  d["total_point_count"] = s.total_point_count;
  d["loaded_point_count"] = s.loaded_point_count;
  d["visible_point_count"] = s.visible_point_count;
  d["frame_counter"] = s.frame_counter;
  return d;
}

}  // namespace

PYBIND11_MODULE(_pygs_cpp, m) {
  m.def("show", [] { session().Show(); });
  m.def("load", [](const std::string& p) { session().Load(p); });
  m.def("close", [] { session().Close(); });
  m.def("render", &Render, py::arg("width"), py::arg("height"));
  m.def("set_orbit", &SetOrbit);
  m.def("stats", &Stats);
  m.def("wait_loaded", [] {
    vkgs::Engine& e = session().EngineForHeadless();
    py::gil_scoped_release nogil;
    e.WaitForLoad();
  });
  m.def("ensure_engine", [] { session().EngineForHeadless(); });
  // interpreter shutdown: stop the loop and join it before the CUDA context goes away
  m.add_object("_cleanup", py::capsule([] { session().Shutdown(); }));
}
