// Stage 3b: alpha-blend rasterisation.  Replaces the instanced-quad graphics pipeline: splat.vert:10-26 (quad =
// centre +- RS*(+-3,+-3)), splat.frag:8-12 (alpha = opacity * exp(-|p|^2/2)), and the fixed-function state around it:
// SRC_ALPHA / ONE_MINUS_SRC_ALPHA for colour AND alpha (engine.cc:281-289), clear (0,0,0,1) (engine.cc:1382-1387),
// depth LESS / no write (graphics_pipeline.cc:79-81; applied in bin.cu), B8G8R8A8_UNORM target (render_pass.cc:15).
//
// One CTA of 4*ROWS warps owns a region of 64 x 8*ROWS pixels and streams the nearest-first splat list of the COARSE
// bin it lies in (bin.cu) in batches of 1024 through shared memory.  While staging, each thread turns its splat's pixel
// bounding box into a mask over the region's 4 x ROWS sub-tiles (16x8 pixels, one per warp); each warp then picks its
// splats out of the batch with one ballot per 32 entries and shades them, 4 pixels per lane.  So a splat only costs the
// warps it can touch, and a warp whose 128 pixels are all saturated drops out of the masks; the CTA stops when every
// warp has.
// Per-fragment arithmetic is the pinned form shared with oracle/vkgs_oracle.c (16-pixel-tile-origin-relative):
//   px = fma(A00, lx, fma(A01, ly, bx)),  py = fma(A10, lx, fma(A11, ly, by)),  covered <=> |px|<=3 && |py|<=3
// with A = (diag(W/2,H/2) * RS)^-1 (bin.cu) and b = A * (tile_origin - centre_px) from non-fused mul/add.
//
// VKGSB_BLEND_FP32   front-to-back: C += c*a*T, A += a*a*T, T *= 1-a; a pixel retires at T < 1e-4.  In exact
//                    arithmetic identical to the reference's back-to-front recurrence (C <- c*a + C*(1-a),
//                    A <- a*a + A*(1-a), A0 = 1); one UNORM8 rounding at the end.
// VKGSB_BLEND_UNORM8 the reference's target semantics: back-to-front, destination re-quantised after every splat like an
//                    8-bit ROP, q <- rint(fma(255*src, a, q*(1-a))).  The recurrence cannot stop early, but it can
//                    START late, with a certificate: every step is monotone non-decreasing in q (a in [0,1]), so a run
//                    of splats F satisfies F(0) <= F(q) <= F(255) for every possible destination value q, and a step
//                    never widens a gap (F(q+g) - F(q) <= g).  Phase A (front to back, transmittance only) finds per
//                    pixel the list position where T < tau; phase B walks back to front from THERE with the bracket
//                    lo = 0, hi = 255 in every channel, both ends through the exact recurrence, until the ends are at
//                    most 1 apart in every channel of the warp's 128 pixels (a few opaque splats), then carries lo
//                    alone.  At the front lo is then within 1/255 of what the full walk over everything behind would
//                    have produced - equal to it wherever the ends met - whatever lies behind.  A warp whose ends are
//                    still further apart (faint layers keep an 8-bit destination 'stuck', tests/
//                    test_unorm8_bracket_cpu.py) tries again with tau^2 (Control::blend_retries), finally from the far
//                    end of the list, where the start value is the background: always certified, fast where splats
//                    are opaque.
#include "common.cuh"
#include "kernels.h"

namespace vkgsb {

#ifndef VKGSB_BLEND_ROWS_FP32
#define VKGSB_BLEND_ROWS_FP32 8
#endif
#ifndef VKGSB_BLEND_ROWS_UNORM8
#define VKGSB_BLEND_ROWS_UNORM8 4
#endif
constexpr int kRowsFp32 = VKGSB_BLEND_ROWS_FP32;      // region = 64 x 8*ROWS pixels, 4*ROWS warps
constexpr int kRowsUnorm8 = VKGSB_BLEND_ROWS_UNORM8;  // the bracket state needs > 64 registers: half the threads per CTA
constexpr int kBatch = 1024;
constexpr int kPix = 4;  // pixels per lane: rows ly0 + 2k of a 16x8 sub-tile
constexpr float kTransmittanceCut = 1e-4f;
constexpr uint32_t kNoCut = 0xffffffffu;
constexpr float kRound = 12582912.f;  // 1.5 * 2^23: (x + kRound) - kRound == rintf(x) for 0 <= x < 2^22 (RNE, ties to even)

constexpr size_t blend_smem(bool layer) {
  return kBatch * (3 * sizeof(float4) + sizeof(uint32_t)) + (layer ? kBatch * sizeof(float) : 0);
}

__device__ __forceinline__ uint32_t quantize8(float x) {  // RNE, saturating
  return static_cast<uint32_t>(__float2int_rn(__saturatef(x) * 255.f));
}
__device__ __forceinline__ uint32_t clamp255(float q) { return static_cast<uint32_t>(fminf(fmaxf(q, 0.f), 255.f)); }
__device__ __forceinline__ float rint255(float x) { return __fsub_rn(__fadd_rn(x, kRound), kRound); }

__device__ __forceinline__ uint32_t pack_pixel(uint32_t r, uint32_t g, uint32_t b, uint32_t a, int bgra) {
  return bgra ? (b | (g << 8) | (r << 16) | (a << 24)) : (r | (g << 8) | (b << 16) | (a << 24));
}

// Mask (bit = row * 4 + col) of the region's sub-tiles a pixel box [x0,x1] x [y0,y1] touches; the region is
// 64 x 8*ROWS pixels at (reg_x, reg_y).
template <int ROWS>
__device__ __forceinline__ uint32_t subtile_mask(uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1, uint32_t reg_x,
                                                 uint32_t reg_y) {
  constexpr int RW = kBinW, RH = kSubH * ROWS;
  // entries come from the coarse bin's list: most boxes miss this region altogether (also the empty box x0 > x1)
  if (x1 < reg_x || x0 >= reg_x + RW || y1 < reg_y || y0 >= reg_y + RH || x0 > x1 || y0 > y1) return 0u;
  const int c0 = max(static_cast<int>(x0) - static_cast<int>(reg_x), 0) / kSubW;
  const int c1 = min(static_cast<int>(x1) - static_cast<int>(reg_x), RW - 1) / kSubW;
  const int r0 = max(static_cast<int>(y0) - static_cast<int>(reg_y), 0) / kSubH;
  const int r1 = min(static_cast<int>(y1) - static_cast<int>(reg_y), RH - 1) / kSubH;
  if (c1 < c0 || r1 < r0) return 0u;
  const uint32_t cols = ((1u << (c1 - c0 + 1)) - 1u) << c0;                                       // 4 bits
  const uint32_t rows = static_cast<uint32_t>(((1ull << (4 * (r1 + 1))) - (1ull << (4 * r0)))) & 0x11111111u;
  return rows * cols;  // no carries: cols <= 0xF
}

// LAYER: an opaque line layer lies under the splats (lines.cu; the reference's axis / grid, engine.cc:1440-1469).  A
// fragment is kept only if the splat's ndc.z is LESS than the layer's depth at the pixel (engine.cc:298-299), and the
// result is composited over the layer's colour instead of the clear colour.
// COUNT: also count the fragments shaded (pixel x splat pairs inside the +-3 sigma square that reach the blend
// arithmetic) into Control::fragment_count - the work unit of this stage; off on the timed path.
template <int MODE, bool LAYER, int ROWS, bool COUNT>
__global__ void __launch_bounds__(128 * ROWS, MODE == VKGSB_BLEND_FP32_MODE ? 1024 / (128 * ROWS) : 512 / (128 * ROWS))
k_blend(const FrameParams* __restrict__ fpp, Control* __restrict__ ctrl, const uint2* __restrict__ ranges,
        const uint32_t* __restrict__ pair_rank, const float4* __restrict__ rrec, int bgra, uint32_t reg_y0,
        const unsigned long long* __restrict__ layer, const float* __restrict__ zndc, uint8_t* __restrict__ image) {
  constexpr int THREADS = 128 * ROWS, RH = kSubH * ROWS;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float4* s_q0 = reinterpret_cast<float4*>(smem_raw);
  float4* s_q1 = s_q0 + kBatch;
  float4* s_q2 = s_q1 + kBatch;
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_q2 + kBatch);
  float* s_z = reinterpret_cast<float*>(s_mask + kBatch);  // LAYER only
  __shared__ uint32_t s_alive, s_redo, s_from;

  const uint32_t width = fpp->width, bins_x = fpp->bins_x;
  const uint32_t band_y0 = fpp->band_y0, band_y1 = fpp->band_y1;
  const uint32_t reg = blockIdx.x, reg_x = (reg % bins_x) * kBinW, reg_y = (reg / bins_x + reg_y0) * RH;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t sub_x = reg_x + (warp % kSubCols) * kSubW, sub_y = reg_y + (warp / kSubCols) * kSubH;
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
  // the region lies inside one coarse bin (coarse bins are >= 128 x 128 pixels, aligned)
  uint2 range = ranges[((reg_y >> fpp->cshift_y) - fpp->cbin_y0) * fpp->cbins_x + (reg_x >> fpp->cshift_x)];
  if (range.y < range.x) range.y = range.x;  // a bin no pair reached keeps end = 0 (bin.cu)
  const uint32_t wbit = 1u << warp;

  // entries [b0, b0 + cnt) of the list -> shared memory, with their sub-tile masks (& keep)
  auto stage = [&](uint32_t b0, uint32_t cnt, uint32_t keep) {
    for (uint32_t i = tid; i < cnt; i += THREADS) {
      const uint32_t rank = __ldg(pair_rank + b0 + i);
      const float4 q0 = __ldg(rrec + rank * 3 + 0), q1 = __ldg(rrec + rank * 3 + 1), q2 = __ldg(rrec + rank * 3 + 2);
      const uint32_t bxw = __float_as_uint(q2.z), byw = __float_as_uint(q2.w);
      s_q0[i] = q0; s_q1[i] = q1; s_q2[i] = q2;
      if (LAYER) s_z[i] = __ldg(zndc + rank);
      s_mask[i] = subtile_mask<ROWS>(bxw & 0xffffu, bxw >> 16, byw & 0xffffu, byw >> 16, reg_x, reg_y) & keep;
    }
  };
  // per-entry constants of the pinned fragment form
  struct Entry {
    float4 q0, q1;
    float2 q2;
    float bx, by, z;
  };
  auto entry = [&](uint32_t j) {
    Entry e;
    e.q0 = s_q0[j];
    e.q1 = s_q1[j];
    e.q2 = *reinterpret_cast<const float2*>(&s_q2[j]);
    const float ox = __fsub_rn(tile_x, e.q1.x), oy = __fsub_rn(tile_y, e.q1.y);
    e.bx = __fadd_rn(__fmul_rn(e.q0.x, ox), __fmul_rn(e.q0.y, oy));
    e.by = __fadd_rn(__fmul_rn(e.q0.z, ox), __fmul_rn(e.q0.w, oy));
    e.z = LAYER ? s_z[j] : 0.f;
    return e;
  };
  uint32_t nfrag = 0;
  // alpha of entry e at pixel k of this lane - 0 when the pixel is not `live`, outside the +-3 sigma square or behind the
  // layer.  Branch-free: a zero alpha makes every blend step below an exact no-op, and one divergent region per pixel
  // cost more than the arithmetic it skipped.  exp(-d/2) = 2^(-d * log2(e) / 2) with the ftz form of ex2 (__expf's
  // own instruction without its denormal-range rescaling: alphas below 2^-126 become 0).
  auto alpha = [&](const Entry& e, int k, bool live) {
    const float px = fmaf(e.q0.x, flx, fmaf(e.q0.y, fly[k], e.bx));
    const float py = fmaf(e.q0.z, flx, fmaf(e.q0.w, fly[k], e.by));
    const bool cov = live && fabsf(px) <= 3.f && fabsf(py) <= 3.f && (!LAYER || e.z < ldepth[k]);
    float ex;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-0.72134752044448170368f * fmaf(py, py, px * px)));
    if (COUNT && cov) ++nfrag;
    return cov ? __saturatef(e.q2.y * ex) : 0.f;
  };
  // the frame's destination travels in the parameter block, so one recorded graph serves every destination
  uint32_t* img = reinterpret_cast<uint32_t*>(fpp->dst_image ? reinterpret_cast<uint8_t*>(fpp->dst_image) : image);

  if (MODE == VKGSB_BLEND_FP32_MODE) {
    float T[kPix], cr[kPix], cg[kPix], cb[kPix], ca[kPix];
    bool done[kPix];
#pragma unroll
    for (int k = 0; k < kPix; ++k) {
      T[k] = 1.f; cr[k] = cg[k] = cb[k] = ca[k] = 0.f;
      done[k] = !inside[k];
    }
    bool warp_done = __all_sync(0xffffffffu, done[0] && done[1] && done[2] && done[3]);
    if (tid == 0) s_alive = 0xffffffffu >> (32 - 4 * ROWS);
    __syncthreads();
    if (warp_done && lane == 0) atomicAnd(&s_alive, ~wbit);
    __syncthreads();

    for (uint32_t b0 = range.x; b0 < range.y; b0 += kBatch) {
      const uint32_t alive = s_alive;
      if (alive == 0u) break;
      const uint32_t cnt = min(static_cast<uint32_t>(kBatch), range.y - b0);
      stage(b0, cnt, alive);
      __syncthreads();
      if (!warp_done) {
        for (uint32_t g0 = 0; g0 < cnt && !warp_done; g0 += 32) {
          const uint32_t m = (g0 + lane < cnt) ? s_mask[g0 + lane] : 0u;
          uint32_t bits = __ballot_sync(0xffffffffu, (m & wbit) != 0u);
          while (bits) {
            const uint32_t j = g0 + __ffs(bits) - 1;
            bits &= bits - 1;
            const Entry e = entry(j);
#pragma unroll
            for (int k = 0; k < kPix; ++k) {
              const float al = alpha(e, k, !done[k]);
              const float w = al * T[k];
              cr[k] = fmaf(e.q1.z, w, cr[k]);
              cg[k] = fmaf(e.q1.w, w, cg[k]);
              cb[k] = fmaf(e.q2.x, w, cb[k]);
              ca[k] = fmaf(al, w, ca[k]);
              T[k] -= w;
              done[k] = done[k] || T[k] < kTransmittanceCut;
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
    // State of a pixel across attempts: its transmittance so far (phase A resumes where it stopped) and the list
    // position of the splat that took it below the current cut (kNoCut: not reached - the walk of that pixel starts at
    // the far end of the list, from the real background, and is exact).
    float T[kPix];
    uint32_t cut[kPix];
    // A pixel is FINAL once its own bracket has been certified; its value then never depends on what the other pixels of
    // the warp need (a later, deeper attempt only serves the pixels still open), so a pixel's result is a function of
    // its own splat list alone - bands, regions and groups all give the same bits.
    uint32_t fin[kPix];
    bool isfin[kPix];
#pragma unroll
    for (int k = 0; k < kPix; ++k) {
      T[k] = 1.f;
      cut[k] = kNoCut;
      fin[k] = lrgba[k] | 0xff000000u;
      isfin[k] = !inside[k];
    }
    const bool any_inside = __any_sync(0xffffffffu, inside[0] || inside[1] || inside[2] || inside[3]);
    bool wfinal = !any_inside || range.y == range.x;  // warp-uniform: this warp's pixels are written
    uint32_t wnext = range.x;                         // warp-uniform: first list position phase A has not examined
    uint32_t last_b0 = range.x;                       // CTA-uniform: the deepest batch phase A has staged
    float tau = fpp->unorm8_cut;
    if (range.y == range.x) {  // nothing in the list: the background
#pragma unroll
      for (int k = 0; k < kPix; ++k)
        if (inside[k]) img[static_cast<size_t>(y_first + 2 * k) * width + x] =
            pack_pixel(lrgba[k] & 255u, (lrgba[k] >> 8) & 255u, (lrgba[k] >> 16) & 255u, 255u, bgra);
    }
    if (tid == 0) {
      s_alive = 0u;
      s_redo = 0u;
      s_from = range.y;
    }
    __syncthreads();
    if (!wfinal && lane == 0) atomicOr(&s_redo, wbit);  // s_redo = the warps still working
    __syncthreads();
    for (uint32_t attempt = 0; s_redo != 0u; ++attempt) {
      // ---- phase A, front to back, transmittance only, resumed: until every pixel of every working warp is below
      //      tau (or the list ends)
      bool done[kPix];
#pragma unroll
      for (int k = 0; k < kPix; ++k) {
        done[k] = isfin[k] || T[k] < tau;
        if (!done[k]) cut[k] = kNoCut;  // set again when it crosses this attempt's (lower) tau
      }
      bool warp_done = wfinal || wnext >= range.y || __all_sync(0xffffffffu, done[0] && done[1] && done[2] && done[3]);
      if (!warp_done && lane == 0) {
        atomicOr(&s_alive, wbit);
        atomicMin(&s_from, wnext);
      }
      __syncthreads();
      if (s_alive != 0u) {
        for (uint32_t b0 = range.x + ((s_from - range.x) / kBatch) * kBatch; b0 < range.y; b0 += kBatch) {
          const uint32_t alive = s_alive;
          if (alive == 0u) break;
          last_b0 = max(last_b0, b0);
          const uint32_t cnt = min(static_cast<uint32_t>(kBatch), range.y - b0);
          stage(b0, cnt, alive);
          __syncthreads();
          if (!warp_done) {
            for (uint32_t g0 = 0; g0 < cnt && !warp_done; g0 += 32) {
              const uint32_t gi = b0 + g0 + lane;
              const uint32_t m = (g0 + lane < cnt && gi >= wnext) ? s_mask[g0 + lane] : 0u;
              uint32_t bits = __ballot_sync(0xffffffffu, (m & wbit) != 0u);
              while (bits) {
                const uint32_t j = g0 + __ffs(bits) - 1;
                bits &= bits - 1;
                const Entry e = entry(j);
#pragma unroll
                for (int k = 0; k < kPix; ++k) {
                  const float al = alpha(e, k, !done[k]);
                  T[k] = __fmul_rn(T[k], __fsub_rn(1.f, al));
                  if (!done[k] && T[k] < tau) {
                    done[k] = true;
                    cut[k] = b0 + j;
                  }
                }
                if (__all_sync(0xffffffffu, done[0] && done[1] && done[2] && done[3])) {
                  warp_done = true;
                  wnext = b0 + j + 1;
                  if (lane == 0) atomicAnd(&s_alive, ~wbit);
                  break;
                }
              }
            }
            if (!warp_done) wnext = b0 + cnt;
          }
          __syncthreads();
        }
      }
      // ---- phase B, back to front from the cuts: both ends of the bracket through the exact recurrence
      float lo[kPix][4], hi[kPix][4];
#pragma unroll
      for (int k = 0; k < kPix; ++k) {
        const bool open = cut[k] != kNoCut;
        lo[k][0] = open ? 0.f : static_cast<float>(lrgba[k] & 255u);
        lo[k][1] = open ? 0.f : static_cast<float>((lrgba[k] >> 8) & 255u);
        lo[k][2] = open ? 0.f : static_cast<float>((lrgba[k] >> 16) & 255u);
        lo[k][3] = open ? 0.f : 255.f;
        hi[k][0] = open ? 255.f : lo[k][0];
        hi[k][1] = open ? 255.f : lo[k][1];
        hi[k][2] = open ? 255.f : lo[k][2];
        hi[k][3] = 255.f;
      }
      // the warp's deepest start; entries behind it are skipped without a look
      uint32_t wcut = 0;
#pragma unroll
      for (int k = 0; k < kPix; ++k) wcut = max(wcut, isfin[k] ? 0u : cut[k]);
      wcut = __reduce_max_sync(0xffffffffu, wcut);
      // warp-uniform: in every channel of every pixel the ends are at most 1 apart.  A ROP step never widens the gap
      // (|F(q + g) - F(q)| <= g for 0 <= a <= 1), so from here on one end is enough: the other stays within 1 of it.
      bool met = false;
      for (uint32_t b0 = last_b0;; b0 -= kBatch) {
        const uint32_t cnt = min(static_cast<uint32_t>(kBatch), range.y - b0);
        __syncthreads();  // the previous batch has been consumed
        stage(b0, cnt, s_redo);
        __syncthreads();
        if (!wfinal && wcut >= b0) {
          for (int g0 = static_cast<int>((cnt - 1) & ~31u); g0 >= 0; g0 -= 32) {
            const uint32_t gi = b0 + g0 + lane;
            const uint32_t m = (g0 + lane < cnt && gi <= wcut) ? s_mask[g0 + lane] : 0u;
            uint32_t bits = __ballot_sync(0xffffffffu, (m & wbit) != 0u);
            while (bits) {
              const uint32_t top = 31u - __clz(bits);
              const uint32_t j = g0 + top;
              bits &= ~(1u << top);
              const Entry e = entry(j);
              const float s255[3] = {__fmul_rn(255.f, e.q1.z), __fmul_rn(255.f, e.q1.w), __fmul_rn(255.f, e.q2.x)};
              if (!met) {
#pragma unroll
                for (int k = 0; k < kPix; ++k) {
                  const float al = alpha(e, k, !isfin[k] && b0 + j <= cut[k]);  // 0 before the pixel's own start: a no-op
                  const float om = __fsub_rn(1.f, al), a255 = __fmul_rn(255.f, al);
#pragma unroll
                  for (int c = 0; c < 4; ++c) {
                    const float sc = c < 3 ? s255[c < 3 ? c : 0] : a255;
                    lo[k][c] = rint255(fmaf(sc, al, __fmul_rn(lo[k][c], om)));
                    hi[k][c] = rint255(fmaf(sc, al, __fmul_rn(hi[k][c], om)));
                  }
                }
                bool close = true;  // a pixel still waiting for its (nearer) cut holds 0 / 255: not close
#pragma unroll
                for (int k = 0; k < kPix; ++k)
#pragma unroll
                  for (int c = 0; c < 4; ++c) close = close && (isfin[k] || hi[k][c] - lo[k][c] <= 1.f);
                met = __all_sync(0xffffffffu, close);
              } else {
#pragma unroll
                for (int k = 0; k < kPix; ++k) {
                  const float al = alpha(e, k, true);
                  const float om = __fsub_rn(1.f, al), a255 = __fmul_rn(255.f, al);
#pragma unroll
                  for (int c = 0; c < 4; ++c) {
                    const float sc = c < 3 ? s255[c < 3 ? c : 0] : a255;
                    lo[k][c] = rint255(fmaf(sc, al, __fmul_rn(lo[k][c], om)));
                  }
                }
              }
            }
          }
        }
        if (b0 == range.x) break;
      }
      // ---- the certificate, per pixel: ends at most 1 apart in every channel -> lo is within 1/255 of the exact
      //      recurrence's result (equal to it where the ends met) and the pixel is final.  A warp with pixels still
      //      open tries again from a deeper cut - for those pixels only.
      if (!wfinal) {
        bool open = false;
#pragma unroll
        for (int k = 0; k < kPix; ++k) {
          if (isfin[k]) continue;
          bool close = true;
          if (!met) {
#pragma unroll
            for (int c = 0; c < 4; ++c) close = close && hi[k][c] - lo[k][c] <= 1.f;
          }
          if (close) {
            fin[k] = clamp255(lo[k][0]) | (clamp255(lo[k][1]) << 8) | (clamp255(lo[k][2]) << 16) | (clamp255(lo[k][3]) << 24);
            isfin[k] = true;
          } else {
            open = true;
          }
        }
        if (__any_sync(0xffffffffu, open)) {
          if (lane == 0) atomicAdd(&ctrl->blend_retries, 1u);
        } else {
#pragma unroll
          for (int k = 0; k < kPix; ++k)
            if (inside[k])
              img[static_cast<size_t>(y_first + 2 * k) * width + x] =
                  pack_pixel(fin[k] & 255u, (fin[k] >> 8) & 255u, (fin[k] >> 16) & 255u, fin[k] >> 24, bgra);
          wfinal = true;
        }
      }
      __syncthreads();  // everyone has read s_redo (staging) before it changes
      if (tid == 0) {
        s_alive = 0u;
        s_from = range.y;
      }
      if (wfinal && lane == 0) atomicAnd(&s_redo, ~wbit);
      // tau -> tau^2 -> tau^4 -> 0 (never reached: every cut opens at the far end, where the start value is known)
      tau = attempt >= 2 ? 0.f : tau * tau;
      __syncthreads();
    }
  }
  if (COUNT) {
    nfrag = __reduce_add_sync(0xffffffffu, nfrag);
    if (lane == 0 && nfrag) atomicAdd(&ctrl->fragment_count, static_cast<unsigned long long>(nfrag));
  }
}

template <int MODE, bool LAYER, int ROWS, bool COUNT>
static void blend_launch(const FrameParams* d_fp, const FrameParams& h_fp, Control* d_ctrl, const uint2* d_ranges,
                         const uint32_t* d_pair_rank, const float4* rrec, int bgra, const unsigned long long* d_layer,
                         const float* d_zndc, uint8_t* d_image, cudaStream_t stream) {
  constexpr uint32_t RH = kSubH * ROWS;
  if (h_fp.band_y1 <= h_fp.band_y0) return;
  const uint32_t ry0 = h_fp.band_y0 / RH, ry1 = (h_fp.band_y1 + RH - 1) / RH;
  const uint32_t nreg = h_fp.bins_x * (ry1 - ry0);
  if (nreg == 0) return;
  k_blend<MODE, LAYER, ROWS, COUNT><<<nreg, 128 * ROWS, blend_smem(LAYER), stream>>>(d_fp, d_ctrl, d_ranges, d_pair_rank, rrec,
                                                                                   bgra, ry0, d_layer, d_zndc, d_image);
}

template <int MODE, int ROWS>
static void blend_opt_in() {
  cudaFuncSetAttribute(k_blend<MODE, false, ROWS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, blend_smem(false));
  cudaFuncSetAttribute(k_blend<MODE, false, ROWS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, blend_smem(false));
  cudaFuncSetAttribute(k_blend<MODE, true, ROWS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, blend_smem(true));
  cudaFuncSetAttribute(k_blend<MODE, true, ROWS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, blend_smem(true));
}
void blend_configure() {  // once per device: opt in to > 48 KB dynamic shared memory
  blend_opt_in<VKGSB_BLEND_FP32_MODE, kRowsFp32>();
  blend_opt_in<VKGSB_BLEND_UNORM8_MODE, kRowsUnorm8>();
}

template <int MODE, int ROWS>
static void blend_dispatch(const FrameParams* d_fp, const FrameParams& h_fp, Control* d_ctrl, const uint2* d_ranges,
                           const uint32_t* d_pair_rank, const float4* rrec, int bgra, bool count,
                           const unsigned long long* d_layer, const float* d_zndc, uint8_t* d_image, cudaStream_t stream) {
  const bool layer = d_layer != nullptr && d_zndc != nullptr;
  if (layer) {
    if (count) blend_launch<MODE, true, ROWS, true>(d_fp, h_fp, d_ctrl, d_ranges, d_pair_rank, rrec, bgra, d_layer, d_zndc, d_image, stream);
    else blend_launch<MODE, true, ROWS, false>(d_fp, h_fp, d_ctrl, d_ranges, d_pair_rank, rrec, bgra, d_layer, d_zndc, d_image, stream);
  } else {
    if (count) blend_launch<MODE, false, ROWS, true>(d_fp, h_fp, d_ctrl, d_ranges, d_pair_rank, rrec, bgra, nullptr, nullptr, d_image, stream);
    else blend_launch<MODE, false, ROWS, false>(d_fp, h_fp, d_ctrl, d_ranges, d_pair_rank, rrec, bgra, nullptr, nullptr, d_image, stream);
  }
}

void launch_blend(const FrameParams* d_fp, const FrameParams& h_fp, Control* d_ctrl, const uint2* d_ranges,
                  const uint32_t* d_pair_rank, const float* d_rrec, int blend_mode, int bgra, bool count_fragments,
                  const unsigned long long* d_layer, const float* d_zndc, uint8_t* d_image, cudaStream_t stream) {
  const float4* rrec = reinterpret_cast<const float4*>(d_rrec);
  if (blend_mode == VKGSB_BLEND_FP32_MODE)
    blend_dispatch<VKGSB_BLEND_FP32_MODE, kRowsFp32>(d_fp, h_fp, d_ctrl, d_ranges, d_pair_rank, rrec, bgra, count_fragments,
                                                     d_layer, d_zndc, d_image, stream);
  else
    blend_dispatch<VKGSB_BLEND_UNORM8_MODE, kRowsUnorm8>(d_fp, h_fp, d_ctrl, d_ranges, d_pair_rank, rrec, bgra, count_fragments,
                                                         d_layer, d_zndc, d_image, stream);
}

}  // namespace vkgsb
