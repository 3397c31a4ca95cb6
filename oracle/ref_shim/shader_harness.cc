// CPU harness that executes the reference's own GLSL shader sources (src/shader/*.comp|vert|frag) through the
// vendored glm, whose types and operators mirror GLSL.  oracle/build_ref.py derives gen/*.inc from the sources where
// they lie under /root/reference (directive / interface-block / swizzle syntax only; every arithmetic statement is
// the reference's), and compiles this file into oracle/_ref/libref_shaders.so.  TEST INFRASTRUCTURE ONLY: it
// validates oracle/vkgs_oracle.c and generates tests/golden/ fixtures (tests/golden/make_golden.py).
//
// Invocations run sequentially in ascending gl_GlobalInvocationID, so rank.comp's atomicAdd hands out slots in id
// order: one of the orders the reference can legally produce, and the deterministic one the oracle pins.
#define GLM_FORCE_SWIZZLE
#include <glm/glm.hpp>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

typedef _Float16 float16_t;
struct f16vec4 {
  float16_t v[4];
  operator glm::vec4() const { return glm::vec4((float)v[0], (float)v[1], (float)v[2], (float)v[3]); }
};
struct InvocationId {
  unsigned x, y, z;
};

// GLSL lets mat3 be built from any mix of scalars and vectors totalling 9 components (projection.comp:92);
// glm only has the all-scalar and all-column forms, so the shader namespaces see this thin subclass instead.
struct glsl_mat3 : glm::mat3 {
  glsl_mat3() {}
  glsl_mat3(const glm::mat3& m) : glm::mat3(m) {}
  explicit glsl_mat3(const glm::mat4& m) : glm::mat3(m) {}
  explicit glsl_mat3(float s) : glm::mat3(s) {}
  template <class A, class B, class... R>
  glsl_mat3(const A& a, const B& b, const R&... r) {
    float f[16];
    int n = 0;
    push(f, n, a);
    push(f, n, b);
    (push(f, n, r), ...);
    for (int c = 0; c < 3; ++c)
      for (int k = 0; k < 3; ++k) (*this)[c][k] = f[c * 3 + k];
  }
  static void push(float* f, int& n, float v) { f[n++] = v; }
  static void push(float* f, int& n, const glm::vec2& v) { f[n++] = v.x; f[n++] = v.y; }
  static void push(float* f, int& n, const glm::vec3& v) { f[n++] = v.x; f[n++] = v.y; f[n++] = v.z; }
};

#define SHADER_PRELUDE                                      \
  using namespace glm;                                      \
  typedef glsl_mat3 mat3;                                   \
  static InvocationId gl_GlobalInvocationID, gl_LocalInvocationID; \
  static int gl_VertexIndex;                                \
  static vec4 gl_Position;                                  \
  static inline uint atomicAdd(uint& mem, uint d) {         \
    uint old = mem;                                         \
    mem += d;                                               \
    return old;                                             \
  }                                                         \
  static inline void barrier() {}

#pragma GCC diagnostic ignored "-Wunused-variable"
#pragma GCC diagnostic ignored "-Wunused-function"

namespace parse_ply_comp {
SHADER_PRELUDE
#include "gen/parse_ply.comp.inc"
}  // namespace parse_ply_comp
namespace rank_comp {
SHADER_PRELUDE
#include "gen/rank.comp.inc"
}  // namespace rank_comp
namespace inverse_index_comp {
SHADER_PRELUDE
#include "gen/inverse_index.comp.inc"
}  // namespace inverse_index_comp
namespace projection_comp {
SHADER_PRELUDE
#include "gen/projection.comp.inc"
}  // namespace projection_comp
namespace splat_vert {
SHADER_PRELUDE
#include "gen/splat.vert.inc"
}  // namespace splat_vert
namespace splat_frag {
SHADER_PRELUDE
#include "gen/splat.frag.inc"
}  // namespace splat_frag

#define REF_API extern "C" __attribute__((visibility("default")))

static glm::mat4 m4(const float* p) {
  glm::mat4 m;
  std::memcpy(&m[0][0], p, 64);
  return m;
}

// parse_ply.comp dispatched over n vertices (engine.cc:1137-1152).
REF_API void ref_parse_ply(unsigned n, const unsigned* offsets60, const float* ply, float* pos, float* cov,
                           float* opacity, uint16_t* sh) {
  using namespace parse_ply_comp;
  std::memcpy(offsets, offsets60, 60 * 4);
  parse_ply_comp::ply = const_cast<float*>(ply);
  gaussian_position = pos;
  gaussian_cov3d = cov;
  gaussian_opacity = opacity;
  gaussian_sh = reinterpret_cast<float16_t*>(sh);
  // workgroup-shared offset table: the first 60 local invocations fill it before the barrier
  point_count = 0;
  for (unsigned l = 0; l < 60; ++l) {
    gl_LocalInvocationID.x = l;
    gl_GlobalInvocationID.x = l;
    shader_main();
  }
  point_count = n;
  for (unsigned id = 0; id < n; ++id) {
    gl_GlobalInvocationID.x = id;
    gl_LocalInvocationID.x = 64 + (id & 127);  // >= 60: table already resident
    shader_main();
  }
}

// fill(0) + rank.comp (engine.cc:1166-1194).  Returns visible_point_count.
REF_API unsigned ref_rank(unsigned n, const float* proj, const float* view, const float* model, const float* pos,
                          unsigned* key_out, unsigned* index_out) {
  using namespace rank_comp;
  projection = m4(proj);
  rank_comp::view = m4(view);
  rank_comp::model = m4(model);
  point_count = n;
  gaussian_position = const_cast<float*>(pos);
  visible_point_count = 0;
  key = key_out;
  rank_comp::index = index_out;
  for (unsigned id = 0; id < n; ++id) {
    gl_GlobalInvocationID.x = id;
    shader_main();
  }
  return visible_point_count;
}

// fill(-1) over N (engine.cc:1226) + inverse_index.comp.
REF_API void ref_inverse_index(unsigned n, unsigned visible, const unsigned* index_in, int* inverse_out) {
  using namespace inverse_index_comp;
  for (unsigned i = 0; i < n; ++i) inverse_out[i] = -1;
  visible_point_count = visible;
  inverse_index_comp::index = const_cast<unsigned*>(index_in);
  inverse_index = inverse_out;
  for (unsigned id = 0; id < n; ++id) {
    gl_GlobalInvocationID.x = id;
    shader_main();
  }
}

// projection.comp over N (engine.cc:1256-1274).  instances: N*12 floats; indirect12: the DrawIndirect block.
REF_API void ref_projection(unsigned n, unsigned visible, const float* proj, const float* view, const float* eye,
                            unsigned width, unsigned height, const float* model, const float* pos, const float* cov,
                            const float* opacity, const uint16_t* sh, const int* inverse, float* instances_out,
                            unsigned* indirect12) {
  using namespace projection_comp;
  projection = m4(proj);
  projection_comp::view = m4(view);
  projection_comp::model = m4(model);
  camera_position = glm::vec3(eye[0], eye[1], eye[2]);
  screen_size = glm::uvec2(width, height);
  point_count = n;
  visible_point_count = visible;
  gaussian_position = const_cast<float*>(pos);
  gaussian_cov3d = const_cast<float*>(cov);
  gaussian_opacity = const_cast<float*>(opacity);
  gaussian_sh = reinterpret_cast<f16vec4*>(const_cast<uint16_t*>(sh));
  inverse_map = const_cast<int*>(inverse);
  instances = reinterpret_cast<glm::vec4*>(instances_out);
  for (unsigned id = 0; id < n; ++id) {
    gl_GlobalInvocationID.x = id;
    shader_main();
  }
  if (indirect12) {
    unsigned v[12] = {indexCount, instanceCount, firstIndex, (unsigned)vertexOffset, firstInstance, 0, 0, 0,
                      vertexCount1, instanceCount1, firstVertex1, firstInstance1};
    std::memcpy(indirect12, v, sizeof v);
  }
}

struct VertOut {
  glm::vec4 pos, color;
  glm::vec2 position;
};
static VertOut run_vert(const float* instances_in, int vertex_index) {
  splat_vert::instances = reinterpret_cast<glm::vec4*>(const_cast<float*>(instances_in));
  splat_vert::gl_VertexIndex = vertex_index;
  splat_vert::shader_main();
  return {splat_vert::gl_Position, splat_vert::out_color, splat_vert::out_position};
}

// splat.vert for one vertex index (4 per splat, index buffer [0,1,2,2,1,3], engine.cc:607-616).
REF_API void ref_splat_vert(const float* instances_in, int vertex_index, float* gl_position4, float* color4,
                            float* position2) {
  VertOut o = run_vert(instances_in, vertex_index);
  std::memcpy(gl_position4, &o.pos[0], 16);
  std::memcpy(color4, &o.color[0], 16);
  std::memcpy(position2, &o.position[0], 8);
}

REF_API void ref_splat_frag(const float* color4, const float* position2, float* out4) {
  splat_frag::color = glm::vec4(color4[0], color4[1], color4[2], color4[3]);
  splat_frag::position = glm::vec2(position2[0], position2[1]);
  splat_frag::shader_main();
  std::memcpy(out4, &splat_frag::out_color[0], 16);
}

// Draw: the two triangles of every splat quad, rasterised at pixel centres with the varyings interpolated in
// double precision (the quad is a parallelogram with w = 1, so interpolation is affine), splat.frag per covered
// pixel, then SRC_ALPHA / ONE_MINUS_SRC_ALPHA blending for colour AND alpha (engine.cc:281-289) over a
// (0,0,0,1) clear (engine.cc:1382-1387) with depth LESS against 1.0 (graphics_pipeline.cc:79-81).
// The accumulator is fp32 (no UNORM8 re-quantisation: ROP rounding is driver-defined).  out: H*W*4 floats RGBA.
REF_API void ref_draw(const float* instances_in, unsigned visible, unsigned width, unsigned height, float* out) {
  for (size_t k = 0; k < (size_t)width * height; ++k) {
    out[4 * k + 0] = out[4 * k + 1] = out[4 * k + 2] = 0.f;
    out[4 * k + 3] = 1.f;
  }
  for (unsigned s = 0; s < visible; ++s) {
    VertOut v[4];
    for (int k = 0; k < 4; ++k) v[k] = run_vert(instances_in, (int)s * 4 + k);
    if (!(v[0].pos.z < 1.f)) continue;  // depth test LESS vs cleared 1.0
    // framebuffer coordinates of the vertices (viewport = full image, engine.cc:1419-1431)
    double X[4], Y[4];
    bool bad = false;
    for (int k = 0; k < 4; ++k) {
      X[k] = ((double)v[k].pos.x / v[k].pos.w + 1.0) * 0.5 * width;
      Y[k] = ((double)v[k].pos.y / v[k].pos.w + 1.0) * 0.5 * height;
      if (!(X[k] == X[k]) || !(Y[k] == Y[k])) bad = true;
    }
    if (bad) continue;  // NaN vertices: primitive discarded
    double xmin = X[0], xmax = X[0], ymin = Y[0], ymax = Y[0];
    for (int k = 1; k < 4; ++k) {
      xmin = X[k] < xmin ? X[k] : xmin; xmax = X[k] > xmax ? X[k] : xmax;
      ymin = Y[k] < ymin ? Y[k] : ymin; ymax = Y[k] > ymax ? Y[k] : ymax;
    }
    long i0 = (long)std::floor(xmin - 0.5), i1 = (long)std::ceil(xmax - 0.5);
    long j0 = (long)std::floor(ymin - 0.5), j1 = (long)std::ceil(ymax - 0.5);
    if (i0 < 0) i0 = 0;
    if (j0 < 0) j0 = 0;
    if (i1 > (long)width - 1) i1 = (long)width - 1;
    if (j1 > (long)height - 1) j1 = (long)height - 1;
    // parallelogram: P = P0 + u (P2 - P0) + v (P1 - P0), u,v in [0,1]   (vertex k = (k/2, k%2))
    double ax = X[2] - X[0], ay = Y[2] - Y[0], bx = X[1] - X[0], by = Y[1] - Y[0];
    double det = ax * by - bx * ay;
    if (det == 0.0) continue;
    for (long j = j0; j <= j1; ++j)
      for (long i = i0; i <= i1; ++i) {
        double dx = (i + 0.5) - X[0], dy = (j + 0.5) - Y[0];
        double u = (dx * by - bx * dy) / det, w = (ax * dy - dx * ay) / det;
        if (u < 0.0 || u > 1.0 || w < 0.0 || w > 1.0) continue;
        float p2[2] = {(float)(v[0].position.x + u * ((double)v[2].position.x - v[0].position.x) +
                               w * ((double)v[1].position.x - v[0].position.x)),
                       (float)(v[0].position.y + u * ((double)v[2].position.y - v[0].position.y) +
                               w * ((double)v[1].position.y - v[0].position.y))};
        float c4[4] = {v[0].color.x, v[0].color.y, v[0].color.z, v[0].color.w}, o[4];
        ref_splat_frag(c4, p2, o);
        float a = o[3] < 0.f ? 0.f : (o[3] > 1.f ? 1.f : o[3]);
        float* d = out + 4 * ((size_t)j * width + i);
        for (int k = 0; k < 3; ++k) {
          float sc = o[k] < 0.f ? 0.f : (o[k] > 1.f ? 1.f : o[k]);
          d[k] = sc * a + d[k] * (1.f - a);
        }
        d[3] = a * a + d[3] * (1.f - a);
      }
  }
}
