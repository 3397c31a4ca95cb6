// The reference's examples/vkgs_viewer.cc flow (construct vkgs::Engine, LoadSplats(path), Run()) against THIS repo's
// include/vkgs/engine/engine.h: the drop-in check for the C++ surface.  The reference's Run() blocks
// until the window closes; headless, another thread calls Close() after --run-ms, the way the window's close button
// would (engine.cc:603,1554).  With --out the frame at the reference's default camera is also written as raw RGBA8,
// so a test can compare it with the C ABI's image.
//   vkgs_viewer -i scene.ply [--run-ms 200] [--width 1600 --height 900] [--out frame.rgba]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include <vkgs/engine/engine.h>

int main(int argc, char** argv) {
  std::string input, out;
  int run_ms = 200;
  uint32_t width = 1600, height = 900;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "-i" || a == "--input") input = next();
    else if (a == "--out") out = next();
    else if (a == "--run-ms") run_ms = std::atoi(next());
    else if (a == "--width") width = static_cast<uint32_t>(std::atoi(next()));
    else if (a == "--height") height = static_cast<uint32_t>(std::atoi(next()));
    else {
      std::cerr << "usage: vkgs_viewer -i input.ply [--run-ms N] [--width W --height H] [--out frame.rgba]" << std::endl;
      return 1;
    }
  }
  try {
    vkgs::Engine engine;
    engine.SetViewport(width, height);
    if (!input.empty()) engine.LoadSplats(input);

    std::thread closer([&] {
      std::this_thread::sleep_for(std::chrono::milliseconds(run_ms));
      engine.Close();
    });
    engine.Run();  // returns after Close(), like the reference when the window is closed
    closer.join();

    const vkgs::FrameStats s = engine.stats();
    std::printf("frames %llu loaded %u / %u visible %u\n", static_cast<unsigned long long>(s.frame_counter),
                s.loaded_point_count, s.total_point_count, s.visible_point_count);
    if (!out.empty()) {
      engine.WaitForLoad();
      std::vector<uint8_t> rgba;
      engine.DrawToImage(&rgba);
      FILE* f = std::fopen(out.c_str(), "wb");
      if (!f || std::fwrite(rgba.data(), 1, rgba.size(), f) != rgba.size()) throw std::runtime_error("cannot write " + out);
      std::fclose(f);
    }
    // Run() is re-entrant after Close() (engine.cc:568,600)
    std::thread closer2([&] {
      std::this_thread::sleep_for(std::chrono::milliseconds(20));
      engine.Close();
    });
    engine.Run();
    closer2.join();
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return 2;
  }
  return 0;
}
