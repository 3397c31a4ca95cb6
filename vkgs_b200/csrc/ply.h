// 3DGS .ply header parsing: host logic of SplatLoadThread (src/vkgs/engine/splat_load_thread.cc:55-135).
#pragma once

#include <stdint.h>

#include <string>

namespace vkgsb {

struct PlyHeader {
  uint64_t vertex_count = 0;
  uint32_t stride_bytes = 0;  // bytes per vertex
  uint64_t body_offset = 0;   // file offset of the first vertex
  uint32_t offsets[60] = {0}; // float-unit offsets: 0-2 xyz, 3-5 scale, 6-9 rot_1,rot_2,rot_3,rot_0, 10..57 SH, 58 opacity,
                              // 59 = stride in floats (splat_load_thread.cc:114-135)
};

// Returns an empty string on success, otherwise the reason.  Stricter than the reference, which assumes every
// property is a 4-byte float and never checks the format line: here other scalar
// types are sized correctly, and the 59 properties the renderer reads must be float32 at 4-byte-aligned offsets.
std::string parse_ply_header(const std::string& path, PlyHeader* out);

}  // namespace vkgsb
