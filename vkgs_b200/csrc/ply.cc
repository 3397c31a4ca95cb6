#include "ply.h"

#include <fstream>
#include <sstream>
#include <unordered_map>

namespace vkgsb {

namespace {
int scalar_size(const std::string& t) {
  if (t == "float" || t == "float32" || t == "int" || t == "int32" || t == "uint" || t == "uint32") return 4;
  if (t == "double" || t == "float64") return 8;
  if (t == "short" || t == "int16" || t == "ushort" || t == "uint16") return 2;
  if (t == "char" || t == "int8" || t == "uchar" || t == "uint8") return 1;
  return -1;
}
}  // namespace

std::string parse_ply_header(const std::string& path, PlyHeader* out) {
  std::ifstream in(path, std::ios::binary);
  if (!in) return "cannot open " + path;

  std::unordered_map<std::string, int> offsets;       // byte offset of every vertex property
  std::unordered_map<std::string, bool> is_float;
  int offset = 0;
  bool in_vertex = false, saw_ply = false, saw_end = false, little_endian = false;
  uint64_t count = 0;
  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line == "end_header") {
      saw_end = true;
      break;
    }
    std::istringstream iss(line);
    std::string word;
    iss >> word;
    if (word == "ply") {
      saw_ply = true;
    } else if (word == "format") {
      std::string fmt;
      iss >> fmt;
      little_endian = (fmt == "binary_little_endian");
    } else if (word == "element") {
      std::string type;
      uint64_t c = 0;
      iss >> type >> c;
      in_vertex = (type == "vertex");
      if (in_vertex) count = c;
    } else if (word == "property" && in_vertex) {
      std::string type, name;
      iss >> type >> name;
      if (type == "list") return "list property in vertex element is not supported";
      int size = scalar_size(type);
      if (size < 0) return "unknown property type '" + type + "'";
      offsets[name] = offset;
      is_float[name] = (type == "float" || type == "float32");
      offset += size;
    }
  }
  if (!saw_ply || !saw_end) return "not a PLY file (missing 'ply' magic or 'end_header')";
  if (!little_endian) return "only 'format binary_little_endian' is supported";
  if (offset == 0 || offset % 4 != 0) return "vertex stride is not a multiple of 4 bytes";

  std::string missing;
  auto at = [&](const std::string& name) -> uint32_t {
    auto it = offsets.find(name);
    if (it == offsets.end() || !is_float[name] || it->second % 4 != 0) {
      if (missing.empty()) missing = name;
      return 0;
    }
    return static_cast<uint32_t>(it->second / 4);
  };
  uint32_t* o = out->offsets;
  o[0] = at("x"); o[1] = at("y"); o[2] = at("z");
  o[3] = at("scale_0"); o[4] = at("scale_1"); o[5] = at("scale_2");
  o[6] = at("rot_1"); o[7] = at("rot_2"); o[8] = at("rot_3"); o[9] = at("rot_0");  // (x,y,z,w) <- (w,x,y,z)
  o[10 + 0] = at("f_dc_0"); o[10 + 16] = at("f_dc_1"); o[10 + 32] = at("f_dc_2");
  for (int i = 0; i < 15; ++i) {
    o[10 + 1 + i] = at("f_rest_" + std::to_string(i));
    o[10 + 17 + i] = at("f_rest_" + std::to_string(15 + i));
    o[10 + 33 + i] = at("f_rest_" + std::to_string(30 + i));
  }
  o[58] = at("opacity");
  o[59] = static_cast<uint32_t>(offset / 4);
  if (!missing.empty()) return "vertex property '" + missing + "' is missing or not an aligned float32";

  out->vertex_count = count;
  out->stride_bytes = static_cast<uint32_t>(offset);
  out->body_offset = static_cast<uint64_t>(in.tellg());
  return std::string();
}

}  // namespace vkgsb
