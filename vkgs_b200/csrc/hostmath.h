// Host-side matrix helpers for the per-frame parameter block.  Column-major float[16] (m[c*4+r]); a sum of products
// is one explicit fused chain (a0*b0, then fma per further term, left to right) in the order rank.comp:32 /
// projection.comp:85 write, exactly as oracle/vkgs_oracle.c pins it.  Build with -ffp-contract=off: std::fmaf is
// the only source of FMAs.
#pragma once

#include <cmath>
#include <cstring>

namespace vkgsb {

inline void mat4_mul(const float* A, const float* B, float* C) {
  float t[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r)
      t[c * 4 + r] = std::fmaf(A[3 * 4 + r], B[c * 4 + 3],
                               std::fmaf(A[2 * 4 + r], B[c * 4 + 2],
                                         std::fmaf(A[1 * 4 + r], B[c * 4 + 1], A[0 * 4 + r] * B[c * 4 + 0])));
  std::memcpy(C, t, sizeof t);
}

inline void mat4_vec(const float* M, const float* v, float* out) {
  float t[4];
  for (int i = 0; i < 4; ++i)
    t[i] = std::fmaf(M[3 * 4 + i], v[3], std::fmaf(M[2 * 4 + i], v[2], std::fmaf(M[1 * 4 + i], v[1], M[0 * 4 + i] * v[0])));
  std::memcpy(out, t, sizeof t);
}

// Adjugate / determinant, the cofactor scheme of glm::inverse(mat4).
inline void mat4_inverse(const float* m, float* out) {
  auto M = [&](int c, int r) { return m[c * 4 + r]; };
  float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
  float c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3), c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
  float c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
  float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
  float c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2), c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
  float c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
  float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
  float c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2), c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
  float c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
  const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
  const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
  const float v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
  const float v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, v3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
  const float sa[4] = {+1.f, -1.f, +1.f, -1.f}, sb[4] = {-1.f, +1.f, -1.f, +1.f};
  float inv[16];
  for (int i = 0; i < 4; ++i) {
    inv[0 * 4 + i] = ((v1[i] * f0[i] - v2[i] * f1[i]) + v3[i] * f2[i]) * sa[i];
    inv[1 * 4 + i] = ((v0[i] * f0[i] - v2[i] * f3[i]) + v3[i] * f4[i]) * sb[i];
    inv[2 * 4 + i] = ((v0[i] * f1[i] - v1[i] * f3[i]) + v3[i] * f5[i]) * sa[i];
    inv[3 * 4 + i] = ((v0[i] * f2[i] - v1[i] * f4[i]) + v2[i] * f5[i]) * sb[i];
  }
  float det = (M(0, 0) * inv[0] + M(0, 1) * inv[4]) + (M(0, 2) * inv[8] + M(0, 3) * inv[12]);
  float rdet = 1.0f / det;
  for (int i = 0; i < 16; ++i) out[i] = inv[i] * rdet;
}

}  // namespace vkgsb
