// Opaque line layer under the splats.  Replaces the reference's axis / grid draw (engine.cc:1440-1469: LINE_LIST
// pipeline with depth test AND depth write, engine.cc:398-415; color.vert: gl_Position = projection * view * model *
// position; color.frag: premultiplied colour) and the depth interaction it exists for (DETAILS.md:7): the splats are
// depth-tested LESS against what the lines wrote and do not write depth themselves (engine.cc:298-299).
//
// One warp per line.  The layer is one 64-bit word per pixel, depth bits in the upper half and the UNORM8 colour in
// the lower, so "depth test LESS + write" is a single atomicMin; cleared to all ones (= no line) every frame.
// Vulkan leaves non-strict line rasterisation to the implementation and the reference has no test for it (parity
// unpinned); the rule here is the one oracle/vkgs_oracle.c (vko_raster_lines) states, operation for operation.
// Compiled with -fmad=false like project.cu: the fma chains are explicit.
#include "common.cuh"
#include "kernels.h"

namespace vkgsb {

__device__ __forceinline__ void mat4_vec1(const float* M, float x, float y, float z, float* r) {
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = fmaf(M[3 * 4 + i], 1.f, fmaf(M[2 * 4 + i], z, fmaf(M[1 * 4 + i], y, M[0 * 4 + i] * x)));
}
__device__ __forceinline__ uint32_t q8_rne(float x) {  // the oracle's q8(): rint(x * 255), NaN and negatives to 0
  const float v = rintf(x * 255.f);
  if (!(v > 0.f)) return 0u;
  return v > 255.f ? 255u : static_cast<uint32_t>(v);
}

__global__ void __launch_bounds__(128)
k_lines(const FrameParams* __restrict__ fpp, uint32_t n_lines, const float* __restrict__ pos,
        const float* __restrict__ col, unsigned long long* __restrict__ layer) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t l = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (l >= n_lines) return;
  const uint32_t W = fpp->width, H = fpp->height;
  float pvm[16];  // projection * view * model of the lines, composed on the host left to right (FrameParams::pvm_lines)
#pragma unroll
  for (int i = 0; i < 16; ++i) pvm[i] = fpp->pvm_lines[i];
  float c0[4], c1[4];
  mat4_vec1(pvm, pos[6 * l + 0], pos[6 * l + 1], pos[6 * l + 2], c0);
  mat4_vec1(pvm, pos[6 * l + 3], pos[6 * l + 4], pos[6 * l + 5], c1);
  float t0 = 0.f, t1 = 1.f;
#pragma unroll
  for (int pl = 0; pl < 6; ++pl) {
    float d0, d1;
    switch (pl) {
      case 0: d0 = c0[3] + c0[0]; d1 = c1[3] + c1[0]; break;
      case 1: d0 = c0[3] - c0[0]; d1 = c1[3] - c1[0]; break;
      case 2: d0 = c0[3] + c0[1]; d1 = c1[3] + c1[1]; break;
      case 3: d0 = c0[3] - c0[1]; d1 = c1[3] - c1[1]; break;
      case 4: d0 = c0[2]; d1 = c1[2]; break;
      default: d0 = c0[3] - c0[2]; d1 = c1[3] - c1[2]; break;
    }
    if ((d0 < 0.f && d1 < 0.f) || !(d0 == d0) || !(d1 == d1)) return;
    if (d0 < 0.f) {
      const float t = d0 / (d0 - d1);
      if (t > t0) t0 = t;
    } else if (d1 < 0.f) {
      const float t = d0 / (d0 - d1);
      if (t < t1) t1 = t;
    }
  }
  if (!(t0 < t1)) return;
  float e0[4], e1[4], q0[4], q1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float d = c1[k] - c0[k];
    e0[k] = fmaf(t0, d, c0[k]);
    e1[k] = fmaf(t1, d, c0[k]);
    const float a = col[8 * l + k], b = col[8 * l + 4 + k], dc = b - a;
    q0[k] = fmaf(t0, dc, a);
    q1[k] = fmaf(t1, dc, a);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {  // color.frag: premultiplied alpha
    q0[k] = q0[k] * q0[3];
    q1[k] = q1[k] * q1[3];
  }
  const float hw = 0.5f * static_cast<float>(W), hh = 0.5f * static_cast<float>(H);
  const float iw0 = 1.f / e0[3], iw1 = 1.f / e1[3];
  const float sx0 = fmaf(e0[0] * iw0, hw, hw), sy0 = fmaf(e0[1] * iw0, hh, hh);
  const float sx1 = fmaf(e1[0] * iw1, hw, hw), sy1 = fmaf(e1[1] * iw1, hh, hh);
  float z0 = e0[2] * iw0, z1 = e1[2] * iw1;
  const bool xmajor = fabsf(sx1 - sx0) >= fabsf(sy1 - sy0);
  float a0 = xmajor ? sx0 : sy0, a1 = xmajor ? sx1 : sy1, m0 = xmajor ? sy0 : sx0, m1 = xmajor ? sy1 : sx1;
  if (a0 > a1) {  // walk from the smaller major coordinate
    float t;
    t = a0; a0 = a1; a1 = t;
    t = m0; m0 = m1; m1 = t;
    t = z0; z0 = z1; z1 = t;
#pragma unroll
    for (int k = 0; k < 4; ++k) { t = q0[k]; q0[k] = q1[k]; q1[k] = t; }
  }
  if (!(a1 > a0)) return;
  const float inv = 1.f / (a1 - a0), dm = m1 - m0, dz = z1 - z0;
  const float lim_a = xmajor ? static_cast<float>(W) : static_cast<float>(H);
  float fa = ceilf(a0 - 0.5f), fb = ceilf(a1 - 0.5f) - 1.f;  // pixel centres c + 0.5 in [a0, a1)
  if (fa < 0.f) fa = 0.f;
  if (fb > lim_a - 1.f) fb = lim_a - 1.f;
  if (!(fa <= fb)) return;
  const int ca = static_cast<int>(fa), cb = static_cast<int>(fb), lim_m = xmajor ? static_cast<int>(H) : static_cast<int>(W);
  for (int c = ca + static_cast<int>(lane); c <= cb; c += 32) {
    const float u = ((static_cast<float>(c) + 0.5f) - a0) * inv;
    const float mf = floorf(fmaf(u, dm, m0));
    if (!(mf >= 0.f && mf <= static_cast<float>(lim_m - 1))) continue;
    const int m = static_cast<int>(mf);
    const float z = fmaf(u, dz, z0);
    if (!(z >= 0.f && z <= 1.f)) continue;
    uint32_t rgba = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) rgba |= q8_rne(fmaf(u, q1[k] - q0[k], q0[k])) << (8 * k);
    const unsigned long long packed = (static_cast<unsigned long long>(__float_as_uint(z)) << 32) | rgba;
    const size_t pix = xmajor ? static_cast<size_t>(m) * W + static_cast<size_t>(c) : static_cast<size_t>(c) * W + static_cast<size_t>(m);
    atomicMin(layer + pix, packed);  // depth test LESS + write; equal depths: the smaller packed colour (the oracle's rule)
  }
}

void launch_lines(const FrameParams* d_fp, uint32_t n_lines, const float* d_pos, const float* d_col, uint32_t width,
                  uint32_t height, unsigned long long* d_layer, cudaStream_t stream) {
  cudaMemsetAsync(d_layer, 0xff, static_cast<size_t>(width) * height * sizeof(unsigned long long), stream);
  if (n_lines == 0) return;
  k_lines<<<(n_lines + 3) / 4, 128, 0, stream>>>(d_fp, n_lines, d_pos, d_col, d_layer);
}

}  // namespace vkgsb
