// Stage 3b: alpha-blend rasterisation.  Replaces the instanced-quad graphics pipeline: splat.vert:10-26 (quad =
// centre +- RS*(+-3,+-3)), splat.frag:8-12 (alpha = opacity * exp(-|p|^2/2)), and the fixed-function state around it:
// SRC_ALPHA / ONE_MINUS_SRC_ALPHA for colour AND alpha (engine.cc:281-289), clear (0,0,0,1) (engine.cc:1382-1387),
// depth LESS / no write (graphics_pipeline.cc:79-81; applied in bin.cu), B8G8R8A8_UNORM target (render_pass.cc:15).
//
// One CTA per 16x16 tile walks the tile's nearest-first splat list in batches staged through shared memory.
// Per-fragment arithmetic is the pinned form shared with oracle/vkgs_oracle.c (tile-origin-relative):
//   px = fma(A00, lx, fma(A01, ly, bx)),  py = fma(A10, lx, fma(A11, ly, by)),  covered <=> |px|<=3 && |py|<=3
// with A = (diag(W/2,H/2) * RS)^-1 and b = A * (tile_origin - centre_px) built from non-fused mul/add.
//
// VKGSB_BLEND_FP32   front-to-back: C += c*a*T, A += a*a*T, T *= 1-a; a pixel retires at T < 1e-4, a tile when all
//                    its pixels have.  Exact-arithmetic identical to the reference's back-to-front recurrence
//                    (C <- c*a + C*(1-a), A <- a*a + A*(1-a), A0 = 1); one UNORM8 rounding at the end.
// VKGSB_BLEND_UNORM8 back-to-front, destination re-quantised after every splat like an 8-bit ROP:
//                    q <- rint(fma(255*src, a, q*(1-a))); no early exit possible.
#include "common.cuh"
#include "kernels.h"

namespace vkgsb {

constexpr int kBlendThreads = kTile * kTile;  // one pixel per thread
constexpr int kBatch = 256;
constexpr float kTransmittanceCut = 1e-4f;

struct __align__(16) Staged {
  float a00, a01, a10, a11;
  float bx, by, r, g;
  float b, op, pad0, pad1;
};

// Per (tile, splat) setup from the 12-float instance record; mirrors raster_setup() + the tile terms in the oracle.
__device__ __forceinline__ Staged stage_splat(const float4 r0, const float4 r1, const float4 r2, float hw, float hh,
                                              float tile_x, float tile_y) {
  Staged s;
  const float cpx = fmaf(r0.x, hw, hw - 0.5f), cpy = fmaf(r0.y, hh, hh - 0.5f);
  const float m00 = __fmul_rn(r1.x, hw), m10 = __fmul_rn(r1.y, hh), m01 = __fmul_rn(r1.z, hw), m11 = __fmul_rn(r1.w, hh);
  const float det = __fsub_rn(__fmul_rn(m00, m11), __fmul_rn(m01, m10));
  s.a00 = __fdiv_rn(m11, det);
  s.a01 = __fdiv_rn(-m01, det);
  s.a10 = __fdiv_rn(-m10, det);
  s.a11 = __fdiv_rn(m00, det);
  const float ox = __fsub_rn(tile_x, cpx), oy = __fsub_rn(tile_y, cpy);
  s.bx = __fadd_rn(__fmul_rn(s.a00, ox), __fmul_rn(s.a01, oy));
  s.by = __fadd_rn(__fmul_rn(s.a10, ox), __fmul_rn(s.a11, oy));
  s.r = __saturatef(r2.x);  // source colour is clamped by the UNORM target; NaN -> 0
  s.g = __saturatef(r2.y);
  s.b = __saturatef(r2.z);
  s.op = r2.w;
  s.pad0 = s.pad1 = 0.f;
  return s;
}

__device__ __forceinline__ uint32_t quantize8(float x) {  // RNE, saturating
  return static_cast<uint32_t>(__float2int_rn(__saturatef(x) * 255.f));
}

__device__ __forceinline__ void store_pixel(uint8_t* image, uint32_t width, uint32_t x, uint32_t y, uint32_t r, uint32_t g,
                                            uint32_t b, uint32_t a, int bgra) {
  const uint32_t w = bgra ? (b | (g << 8) | (r << 16) | (a << 24)) : (r | (g << 8) | (b << 16) | (a << 24));
  reinterpret_cast<uint32_t*>(image)[static_cast<size_t>(y) * width + x] = w;
}

template <int MODE>
__global__ void __launch_bounds__(kBlendThreads)
k_blend(const FrameParams* __restrict__ fpp, const uint2* __restrict__ ranges, const uint32_t* __restrict__ pair_slot,
        const float4* __restrict__ inst, int bgra, uint8_t* __restrict__ image) {
  __shared__ Staged s_splat[kBatch];

  const uint32_t width = fpp->width, height = fpp->height, tiles_x = fpp->tiles_x;
  const uint32_t band_y0 = fpp->band_y0, band_y1 = fpp->band_y1, tile_y0 = fpp->tile_y0;
  const uint32_t tile = blockIdx.x, tx = tile % tiles_x, ty = tile / tiles_x + tile_y0;
  const uint32_t tid = threadIdx.x, lx = tid % kTile, ly = tid / kTile;
  const uint32_t x = tx * kTile + lx, y = ty * kTile + ly;
  const bool inside = x < width && y >= band_y0 && y < band_y1 && y < height;
  const float hw = 0.5f * static_cast<float>(width), hh = 0.5f * static_cast<float>(height);
  const float tile_x = static_cast<float>(tx * kTile), tile_y = static_cast<float>(ty * kTile);
  const float flx = static_cast<float>(lx), fly = static_cast<float>(ly);
  const uint2 range = ranges[tile];

  if (MODE == VKGSB_BLEND_FP32_MODE) {
    float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, ca = 0.f;
    bool done = !inside;
    for (uint32_t b0 = range.x; b0 < range.y; b0 += kBatch) {
      if (__syncthreads_and(done)) break;
      const uint32_t cnt = min(static_cast<uint32_t>(kBatch), range.y - b0);
      if (tid < cnt) {
        const uint32_t slot = __ldg(pair_slot + b0 + tid);
        s_splat[tid] = stage_splat(__ldg(inst + slot * 3 + 0), __ldg(inst + slot * 3 + 1), __ldg(inst + slot * 3 + 2), hw,
                                   hh, tile_x, tile_y);
      }
      __syncthreads();
      if (!done) {
        for (uint32_t j = 0; j < cnt; ++j) {
          const Staged s = s_splat[j];
          const float px = fmaf(s.a00, flx, fmaf(s.a01, fly, s.bx));
          const float py = fmaf(s.a10, flx, fmaf(s.a11, fly, s.by));
          if (!(fabsf(px) <= 3.f && fabsf(py) <= 3.f)) continue;
          float al = s.op * __expf(-0.5f * fmaf(py, py, px * px));
          al = __saturatef(al);
          const float w = al * T;
          cr = fmaf(s.r, w, cr);
          cg = fmaf(s.g, w, cg);
          cb = fmaf(s.b, w, cb);
          ca = fmaf(al, w, ca);
          T -= w;
          if (T < kTransmittanceCut) {
            done = true;
            break;
          }
        }
      }
    }
    if (inside) store_pixel(image, width, x, y, quantize8(cr), quantize8(cg), quantize8(cb), quantize8(ca + T), bgra);
  } else {
    // back-to-front over the nearest-first list: batches from the tail, entries in reverse
    float qr = 0.f, qg = 0.f, qb = 0.f, qa = 255.f;
    uint32_t remaining = range.y - range.x;
    while (remaining > 0) {
      const uint32_t cnt = min(static_cast<uint32_t>(kBatch), remaining);
      const uint32_t b0 = range.x + remaining - cnt;
      __syncthreads();
      if (tid < cnt) {
        const uint32_t slot = __ldg(pair_slot + b0 + tid);
        s_splat[tid] = stage_splat(__ldg(inst + slot * 3 + 0), __ldg(inst + slot * 3 + 1), __ldg(inst + slot * 3 + 2), hw,
                                   hh, tile_x, tile_y);
      }
      __syncthreads();
      if (inside) {
        for (int j = static_cast<int>(cnt) - 1; j >= 0; --j) {
          const Staged s = s_splat[j];
          const float px = fmaf(s.a00, flx, fmaf(s.a01, fly, s.bx));
          const float py = fmaf(s.a10, flx, fmaf(s.a11, fly, s.by));
          if (!(fabsf(px) <= 3.f && fabsf(py) <= 3.f)) continue;
          float al = s.op * __expf(-0.5f * fmaf(py, py, px * px));
          al = __saturatef(al);
          const float om = __fsub_rn(1.f, al);
          qr = rintf(fmaf(__fmul_rn(255.f, s.r), al, __fmul_rn(qr, om)));
          qg = rintf(fmaf(__fmul_rn(255.f, s.g), al, __fmul_rn(qg, om)));
          qb = rintf(fmaf(__fmul_rn(255.f, s.b), al, __fmul_rn(qb, om)));
          qa = rintf(fmaf(__fmul_rn(255.f, al), al, __fmul_rn(qa, om)));
        }
      }
      remaining -= cnt;
    }
    if (inside)
      store_pixel(image, width, x, y, static_cast<uint32_t>(fminf(fmaxf(qr, 0.f), 255.f)),
                  static_cast<uint32_t>(fminf(fmaxf(qg, 0.f), 255.f)), static_cast<uint32_t>(fminf(fmaxf(qb, 0.f), 255.f)),
                  static_cast<uint32_t>(fminf(fmaxf(qa, 0.f), 255.f)), bgra);
  }
}

void launch_blend(const FrameParams* d_fp, const FrameParams& h_fp, const uint2* d_ranges, const uint32_t* d_pair_slot,
                  const float* d_inst, int blend_mode, int bgra, uint8_t* d_image, cudaStream_t stream) {
  const uint32_t ntiles = h_fp.tiles_x * (h_fp.tile_y1 - h_fp.tile_y0);
  if (ntiles == 0) return;
  if (blend_mode == 0)
    k_blend<VKGSB_BLEND_FP32_MODE><<<ntiles, kBlendThreads, 0, stream>>>(d_fp, d_ranges, d_pair_slot,
                                                                       reinterpret_cast<const float4*>(d_inst), bgra, d_image);
  else
    k_blend<VKGSB_BLEND_UNORM8_MODE><<<ntiles, kBlendThreads, 0, stream>>>(d_fp, d_ranges, d_pair_slot,
                                                                         reinterpret_cast<const float4*>(d_inst), bgra, d_image);
}

}  // namespace vkgsb
