// Stage 3b: alpha-blend rasterisation.  Replaces the instanced-quad graphics pipeline: splat.vert:10-26 (quad =
// centre +- RS*(+-3,+-3)), splat.frag:8-12 (alpha = opacity * exp(-|p|^2/2)), and the fixed-function state around it:
// SRC_ALPHA / ONE_MINUS_SRC_ALPHA for colour AND alpha (engine.cc:281-289), clear (0,0,0,1) (engine.cc:1382-1387),
// depth LESS / no write (graphics_pipeline.cc:79-81; applied in bin.cu), B8G8R8A8_UNORM target (render_pass.cc:15).
//
// One 1024-thread CTA per 64x64-pixel bin streams the nearest-first splat list of the COARSE bin it lies in (bin.cu)
// in batches of 1024 through shared memory.  While staging, each thread turns its splat's pixel bounding box into a 32-bit mask over the bin's
// 4x8 sub-tiles (16x8 pixels, one per warp); each warp then picks its splats out of the batch with one ballot per
// 32 entries and shades them, 4 pixels per lane.  So a splat only costs the warps it can touch, and a warp whose 128
// pixels are all saturated drops out of the masks; the CTA stops when every warp has.
// Per-fragment arithmetic is the pinned form shared with oracle/vkgs_oracle.c (16-pixel-tile-origin-relative):
//   px = fma(A00, lx, fma(A01, ly, bx)),  py = fma(A10, lx, fma(A11, ly, by)),  covered <=> |px|<=3 && |py|<=3
// with A = (diag(W/2,H/2) * RS)^-1 (bin.cu) and b = A * (tile_origin - centre_px) from non-fused mul/add.
//
// VKGSB_BLEND_FP32   front-to-back: C += c*a*T, A += a*a*T, T *= 1-a; a pixel retires at T < 1e-4.  In exact
//                    arithmetic identical to the reference's back-to-front recurrence (C <- c*a + C*(1-a),
//                    A <- a*a + A*(1-a), A0 = 1); one UNORM8 rounding at the end.
// VKGSB_BLEND_UNORM8 back-to-front, destination re-quantised after every splat like an 8-bit ROP:
//                    q <- rint(fma(255*src, a, q*(1-a))); no early exit possible.
#include "common.cuh"
#include "kernels.h"

namespace vkgsb {

constexpr int kBlendThreads = 1024;
constexpr int kBatch = 1024;
constexpr int kPix = 4;  // pixels per lane: rows ly0 + 2k of a 16x8 sub-tile
constexpr float kTransmittanceCut = 1e-4f;
constexpr size_t kBlendSmem = kBatch * (3 * sizeof(float4) + sizeof(uint32_t));
constexpr size_t kBlendSmemLayer = kBlendSmem + kBatch * sizeof(float);  // + ndc.z of the staged entries

__device__ __forceinline__ uint32_t quantize8(float x) {  // RNE, saturating
  return static_cast<uint32_t>(__float2int_rn(__saturatef(x) * 255.f));
}
__device__ __forceinline__ uint32_t clamp255(float q) { return static_cast<uint32_t>(fminf(fmaxf(q, 0.f), 255.f)); }

__device__ __forceinline__ uint32_t pack_pixel(uint32_t r, uint32_t g, uint32_t b, uint32_t a, int bgra) {
  return bgra ? (b | (g << 8) | (r << 16) | (a << 24)) : (r | (g << 8) | (b << 16) | (a << 24));
}

// 32-bit mask (bit = row * 4 + col) of the bin's sub-tiles a pixel box [x0,x1] x [y0,y1] touches.
__device__ __forceinline__ uint32_t subtile_mask(uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1, uint32_t bin_x,
                                                 uint32_t bin_y) {
  // entries come from the coarse bin's list: most boxes miss this 64x64 bin altogether (also the empty box x0 > x1)
  if (x1 < bin_x || x0 >= bin_x + kBinW || y1 < bin_y || y0 >= bin_y + kBinH || x0 > x1 || y0 > y1) return 0u;
  const int c0 = max(static_cast<int>(x0) - static_cast<int>(bin_x), 0) / kSubW;
  const int c1 = min(static_cast<int>(x1) - static_cast<int>(bin_x), kBinW - 1) / kSubW;
  const int r0 = max(static_cast<int>(y0) - static_cast<int>(bin_y), 0) / kSubH;
  const int r1 = min(static_cast<int>(y1) - static_cast<int>(bin_y), kBinH - 1) / kSubH;
  if (c1 < c0 || r1 < r0) return 0u;
  const uint32_t cols = ((1u << (c1 - c0 + 1)) - 1u) << c0;                                       // 4 bits
  const uint32_t rows = static_cast<uint32_t>(((1ull << (4 * (r1 + 1))) - (1ull << (4 * r0)))) & 0x11111111u;
  return rows * cols;  // no carries: cols <= 0xF
}

// LAYER: an opaque line layer lies under the splats (lines.cu; the reference's axis / grid, engine.cc:1440-1469).  A
// fragment is kept only if the splat's ndc.z is LESS than the layer's depth at the pixel (engine.cc:298-299), and the
// result is composited over the layer's colour instead of the clear colour.
template <int MODE, bool LAYER>
__global__ void __launch_bounds__(kBlendThreads, 1)
k_blend(const FrameParams* __restrict__ fpp, const uint2* __restrict__ ranges, const uint32_t* __restrict__ pair_rank,
        const float4* __restrict__ rrec, int bgra, const unsigned long long* __restrict__ layer,
        const float* __restrict__ zndc, uint8_t* __restrict__ image) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float4* s_q0 = reinterpret_cast<float4*>(smem_raw);
  float4* s_q1 = s_q0 + kBatch;
  float4* s_q2 = s_q1 + kBatch;
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_q2 + kBatch);
  float* s_z = reinterpret_cast<float*>(s_mask + kBatch);  // LAYER only
  __shared__ uint32_t s_alive;

  const uint32_t width = fpp->width, bins_x = fpp->bins_x;
  const uint32_t band_y0 = fpp->band_y0, band_y1 = fpp->band_y1;
  const uint32_t bin = blockIdx.x, bin_x = (bin % bins_x) * kBinW, bin_y = (bin / bins_x + fpp->bin_y0) * kBinH;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t sub_x = bin_x + (warp % kSubCols) * kSubW, sub_y = bin_y + (warp / kSubCols) * kSubH;
  const uint32_t x = sub_x + (lane & 15u), y_first = sub_y + (lane >> 4);
  const uint32_t org_y = sub_y & ~static_cast<uint32_t>(kTile - 1);  // 16-aligned origin (sub_x already is)
  const float tile_x = static_cast<float>(sub_x), tile_y = static_cast<float>(org_y);
  const float flx = static_cast<float>(lane & 15u);
  float fly[kPix];
  bool inside[kPix];
#pragma unroll
  for (int k = 0; k < kPix; ++k) {
    const uint32_t y = y_first + 2 * k;
    fly[k] = static_cast<float>(y - org_y);
    inside[k] = x < width && y >= band_y0 && y < band_y1;
  }
  float ldepth[kPix];   // the layer's depth at the pixel (1.0 = cleared: every splat with ndc.z < 1 passes)
  uint32_t lrgba[kPix]; // and its colour; (0,0,0,255) = the clear colour (engine.cc:1382-1387)
#pragma unroll
  for (int k = 0; k < kPix; ++k) {
    ldepth[k] = 1.f;
    lrgba[k] = 0xff000000u;
    if (LAYER && inside[k]) {
      const unsigned long long w = __ldg(layer + static_cast<size_t>(y_first + 2 * k) * width + x);
      if (w != ~0ull) {
        ldepth[k] = __uint_as_float(static_cast<uint32_t>(w >> 32));
        lrgba[k] = static_cast<uint32_t>(w) | 0xff000000u;  // premultiplied over the opaque clear colour: alpha -> 1
      }
    }
  }
  uint2 range = ranges[((bin_y >> fpp->cshift_y) - fpp->cbin_y0) * fpp->cbins_x + (bin_x >> fpp->cshift_x)];
  if (range.y < range.x) range.y = range.x;  // a bin no pair reached keeps end = 0 (bin.cu)
  const uint32_t wbit = 1u << warp;

  if (MODE == VKGSB_BLEND_FP32_MODE) {
    float T[kPix], cr[kPix], cg[kPix], cb[kPix], ca[kPix];
    bool done[kPix];
#pragma unroll
    for (int k = 0; k < kPix; ++k) {
      T[k] = 1.f; cr[k] = cg[k] = cb[k] = ca[k] = 0.f;
      done[k] = !inside[k];
    }
    bool warp_done = __all_sync(0xffffffffu, done[0] && done[1] && done[2] && done[3]);
    if (tid == 0) s_alive = 0xffffffffu;
    __syncthreads();
    if (warp_done && lane == 0) atomicAnd(&s_alive, ~wbit);
    __syncthreads();

    for (uint32_t b0 = range.x; b0 < range.y; b0 += kBatch) {
      const uint32_t alive = s_alive;
      if (alive == 0u) break;
      const uint32_t cnt = min(static_cast<uint32_t>(kBatch), range.y - b0);
      if (tid < cnt) {
        const uint32_t rank = __ldg(pair_rank + b0 + tid);
        const float4 q0 = __ldg(rrec + rank * 3 + 0), q1 = __ldg(rrec + rank * 3 + 1), q2 = __ldg(rrec + rank * 3 + 2);
        const uint32_t bxw = __float_as_uint(q2.z), byw = __float_as_uint(q2.w);
        s_q0[tid] = q0; s_q1[tid] = q1; s_q2[tid] = q2;
        if (LAYER) s_z[tid] = __ldg(zndc + rank);
        s_mask[tid] = subtile_mask(bxw & 0xffffu, bxw >> 16, byw & 0xffffu, byw >> 16, bin_x, bin_y) & alive;
      }
      __syncthreads();
      if (!warp_done) {
        for (uint32_t g0 = 0; g0 < cnt && !warp_done; g0 += 32) {
          const uint32_t m = (g0 + lane < cnt) ? s_mask[g0 + lane] : 0u;
          uint32_t bits = __ballot_sync(0xffffffffu, (m & wbit) != 0u);
          while (bits) {
            const uint32_t j = g0 + __ffs(bits) - 1;
            bits &= bits - 1;
            const float4 q0 = s_q0[j], q1 = s_q1[j];
            const float2 q2 = *reinterpret_cast<const float2*>(&s_q2[j]);
            const float ox = __fsub_rn(tile_x, q1.x), oy = __fsub_rn(tile_y, q1.y);
            const float bx = __fadd_rn(__fmul_rn(q0.x, ox), __fmul_rn(q0.y, oy));
            const float by = __fadd_rn(__fmul_rn(q0.z, ox), __fmul_rn(q0.w, oy));
            const float z = LAYER ? s_z[j] : 0.f;
#pragma unroll
            for (int k = 0; k < kPix; ++k) {
              const float px = fmaf(q0.x, flx, fmaf(q0.y, fly[k], bx));
              const float py = fmaf(q0.z, flx, fmaf(q0.w, fly[k], by));
              if (!done[k] && fabsf(px) <= 3.f && fabsf(py) <= 3.f && (!LAYER || z < ldepth[k])) {
                const float al = __saturatef(q2.y * __expf(-0.5f * fmaf(py, py, px * px)));
                const float w = al * T[k];
                cr[k] = fmaf(q1.z, w, cr[k]);
                cg[k] = fmaf(q1.w, w, cg[k]);
                cb[k] = fmaf(q2.x, w, cb[k]);
                ca[k] = fmaf(al, w, ca[k]);
                T[k] -= w;
                done[k] = T[k] < kTransmittanceCut;
              }
            }
            if (__all_sync(0xffffffffu, done[0] && done[1] && done[2] && done[3])) {
              warp_done = true;
              if (lane == 0) atomicAnd(&s_alive, ~wbit);
              break;
            }
          }
        }
      }
      __syncthreads();
    }
    uint32_t* img = reinterpret_cast<uint32_t*>(image);
#pragma unroll
    for (int k = 0; k < kPix; ++k)
      if (inside[k]) {
        if (LAYER) {  // what is left of the transmittance shows the layer (UNORM8 in the target when the splats start)
          cr[k] = fmaf(static_cast<float>(lrgba[k] & 255u) * (1.f / 255.f), T[k], cr[k]);
          cg[k] = fmaf(static_cast<float>((lrgba[k] >> 8) & 255u) * (1.f / 255.f), T[k], cg[k]);
          cb[k] = fmaf(static_cast<float>((lrgba[k] >> 16) & 255u) * (1.f / 255.f), T[k], cb[k]);
        }
        img[static_cast<size_t>(y_first + 2 * k) * width + x] =
            pack_pixel(quantize8(cr[k]), quantize8(cg[k]), quantize8(cb[k]), quantize8(ca[k] + T[k]), bgra);
      }
  } else {
    // back-to-front over the nearest-first list: batches from the tail, entries in reverse
    float qr[kPix], qg[kPix], qb[kPix], qa[kPix];
#pragma unroll
    for (int k = 0; k < kPix; ++k) {
      qr[k] = static_cast<float>(lrgba[k] & 255u);
      qg[k] = static_cast<float>((lrgba[k] >> 8) & 255u);
      qb[k] = static_cast<float>((lrgba[k] >> 16) & 255u);
      qa[k] = 255.f;
    }
    uint32_t remaining = range.y - range.x;
    while (remaining > 0) {
      const uint32_t cnt = min(static_cast<uint32_t>(kBatch), remaining);
      const uint32_t b0 = range.x + remaining - cnt;
      __syncthreads();
      if (tid < cnt) {
        const uint32_t rank = __ldg(pair_rank + b0 + tid);
        const float4 q0 = __ldg(rrec + rank * 3 + 0), q1 = __ldg(rrec + rank * 3 + 1), q2 = __ldg(rrec + rank * 3 + 2);
        const uint32_t bxw = __float_as_uint(q2.z), byw = __float_as_uint(q2.w);
        s_q0[tid] = q0; s_q1[tid] = q1; s_q2[tid] = q2;
        if (LAYER) s_z[tid] = __ldg(zndc + rank);
        s_mask[tid] = subtile_mask(bxw & 0xffffu, bxw >> 16, byw & 0xffffu, byw >> 16, bin_x, bin_y);
      }
      __syncthreads();
      for (int g0 = static_cast<int>((cnt - 1) & ~31u); g0 >= 0; g0 -= 32) {
        const uint32_t m = (g0 + lane < cnt) ? s_mask[g0 + lane] : 0u;
        uint32_t bits = __ballot_sync(0xffffffffu, (m & wbit) != 0u);
        while (bits) {
          const uint32_t top = 31u - __clz(bits);
          const uint32_t j = g0 + top;
          bits &= ~(1u << top);
          const float4 q0 = s_q0[j], q1 = s_q1[j];
          const float2 q2 = *reinterpret_cast<const float2*>(&s_q2[j]);
          const float ox = __fsub_rn(tile_x, q1.x), oy = __fsub_rn(tile_y, q1.y);
          const float bx = __fadd_rn(__fmul_rn(q0.x, ox), __fmul_rn(q0.y, oy));
          const float by = __fadd_rn(__fmul_rn(q0.z, ox), __fmul_rn(q0.w, oy));
          const float r255 = __fmul_rn(255.f, q1.z), g255 = __fmul_rn(255.f, q1.w), b255 = __fmul_rn(255.f, q2.x);
          const float z = LAYER ? s_z[j] : 0.f;
#pragma unroll
          for (int k = 0; k < kPix; ++k) {
            const float px = fmaf(q0.x, flx, fmaf(q0.y, fly[k], bx));
            const float py = fmaf(q0.z, flx, fmaf(q0.w, fly[k], by));
            if (fabsf(px) <= 3.f && fabsf(py) <= 3.f && (!LAYER || z < ldepth[k])) {
              const float al = __saturatef(q2.y * __expf(-0.5f * fmaf(py, py, px * px)));
              const float om = __fsub_rn(1.f, al);
              qr[k] = rintf(fmaf(r255, al, __fmul_rn(qr[k], om)));
              qg[k] = rintf(fmaf(g255, al, __fmul_rn(qg[k], om)));
              qb[k] = rintf(fmaf(b255, al, __fmul_rn(qb[k], om)));
              qa[k] = rintf(fmaf(__fmul_rn(255.f, al), al, __fmul_rn(qa[k], om)));
            }
          }
        }
      }
      remaining -= cnt;
    }
    uint32_t* img = reinterpret_cast<uint32_t*>(image);
#pragma unroll
    for (int k = 0; k < kPix; ++k)
      if (inside[k])
        img[static_cast<size_t>(y_first + 2 * k) * width + x] =
            pack_pixel(clamp255(qr[k]), clamp255(qg[k]), clamp255(qb[k]), clamp255(qa[k]), bgra);
  }
}

void blend_configure() {
  cudaFuncSetAttribute(k_blend<VKGSB_BLEND_FP32_MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBlendSmem);
  cudaFuncSetAttribute(k_blend<VKGSB_BLEND_UNORM8_MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBlendSmem);
  cudaFuncSetAttribute(k_blend<VKGSB_BLEND_FP32_MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBlendSmemLayer);
  cudaFuncSetAttribute(k_blend<VKGSB_BLEND_UNORM8_MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBlendSmemLayer);
}

void launch_blend(const FrameParams* d_fp, const FrameParams& h_fp, const uint2* d_ranges, const uint32_t* d_pair_rank,
                  const float* d_rrec, int blend_mode, int bgra, const unsigned long long* d_layer, const float* d_zndc,
                  uint8_t* d_image, cudaStream_t stream) {
  const uint32_t nbins = h_fp.bins_x * (h_fp.bin_y1 - h_fp.bin_y0);
  if (nbins == 0) return;
  const float4* rrec = reinterpret_cast<const float4*>(d_rrec);
  const bool layer = d_layer != nullptr && d_zndc != nullptr;
  if (blend_mode == VKGSB_BLEND_FP32_MODE) {
    if (layer)
      k_blend<VKGSB_BLEND_FP32_MODE, true><<<nbins, kBlendThreads, kBlendSmemLayer, stream>>>(
          d_fp, d_ranges, d_pair_rank, rrec, bgra, d_layer, d_zndc, d_image);
    else
      k_blend<VKGSB_BLEND_FP32_MODE, false><<<nbins, kBlendThreads, kBlendSmem, stream>>>(
          d_fp, d_ranges, d_pair_rank, rrec, bgra, nullptr, nullptr, d_image);
  } else {
    if (layer)
      k_blend<VKGSB_BLEND_UNORM8_MODE, true><<<nbins, kBlendThreads, kBlendSmemLayer, stream>>>(
          d_fp, d_ranges, d_pair_rank, rrec, bgra, d_layer, d_zndc, d_image);
    else
      k_blend<VKGSB_BLEND_UNORM8_MODE, false><<<nbins, kBlendThreads, kBlendSmem, stream>>>(
          d_fp, d_ranges, d_pair_rank, rrec, bgra, nullptr, nullptr, d_image);
  }
}

}  // namespace vkgsb
