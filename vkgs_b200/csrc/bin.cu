// Stage 3a: coarse binning.  No counterpart in the reference, whose instanced quads go through the hardware
// rasteriser in sorted order (vkCmdDrawIndexedIndirect, engine.cc:1472-1480); here the globally sorted splat list is
// turned into per-bin lists that keep that order:
//   k_make_pairs   one thread per sorted rank, walked NEAREST FIRST.  Builds the splat's raster record once
//                  (pixel-space inverse footprint + clamped colour + pixel bounding box of the +-3 sigma quad of
//                  splat.vert:19-25), stores it at its rank, and emits one (bin, rank) pair per 64x64-pixel bin the
//                  box touches, at offsets from an ordered (decoupled look-back) scan - the pair list is rank-major.
//   stable onesweep sort of the pairs by bin id (sort.cu) -> each bin's pairs are contiguous, still nearest-first.
//   k_bin_ranges   [begin,end) of every bin in the sorted pair list.
// Bins are deliberately coarse: a first version binned straight to 16x16 tiles and spent 3.6 ms of a 4.1 ms frame
// sorting 9e7 pairs of which early termination consumed a few percent (profiles/r01_notes.md).  The blend stage
// refines a bin's list to 16x8 sub-tiles on chip.
#include "common.cuh"
#include "kernels.h"

namespace vkgsb {

constexpr int kPairThreads = 256;

uint32_t pairs_num_blocks(uint32_t max_visible) { return (max_visible + kPairThreads - 1) / kPairThreads; }

__global__ void __launch_bounds__(kPairThreads)
k_make_pairs(const FrameParams* __restrict__ fpp, Control* __restrict__ ctrl, unsigned long long* __restrict__ scan_desc,
             const uint32_t* __restrict__ sorted_slots, const float4* __restrict__ inst, uint64_t max_pairs,
             float4* __restrict__ rrec, uint32_t* __restrict__ pair_bin, uint32_t* __restrict__ pair_rank) {
  __shared__ uint32_t s_off[kPairThreads + 1];  // exclusive offsets of this block's splats
  __shared__ uint32_t s_rect[kPairThreads];     // bx0 | by0 << 8 | bw << 16 | bh << 24   (<= 60 x 34 bins at 3840 x 2160)
  __shared__ uint32_t s_wsum[kPairThreads / 32];
  __shared__ uint32_t s_ticket, s_base;

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t V = ctrl->visible_count;
  const uint32_t nblocks = (V + kPairThreads - 1) / kPairThreads;
  if (tid == 0) s_ticket = atomicAdd(&ctrl->pairs_ticket, 1u);
  __syncthreads();
  const uint32_t ticket = s_ticket;
  if (ticket >= nblocks) return;
  const uint32_t width = fpp->width, height = fpp->height, bins_x = fpp->bins_x, bin_y0 = fpp->bin_y0;
  const float band_lo = static_cast<float>(fpp->band_y0), band_hi = static_cast<float>(fpp->band_y1) - 1.f;

  const uint32_t i = ticket * kPairThreads + tid;  // i-th nearest splat
  uint32_t count = 0, rect = 0;
  if (i < V) {
    const uint32_t slot = sorted_slots[V - 1 - i];  // ascending key = far -> near (rank.comp:39): walk it backwards
    const float4 r0 = __ldg(inst + slot * 3 + 0), r1 = __ldg(inst + slot * 3 + 1), r2 = __ldg(inst + slot * 3 + 2);
    // pixel frame: pixel i has its centre at coordinate i  =>  cpx = (ndc.x + 1) * W/2 - 1/2
    const float hw = 0.5f * static_cast<float>(width), hh = 0.5f * static_cast<float>(height);
    const float cpx = fmaf(r0.x, hw, hw - 0.5f), cpy = fmaf(r0.y, hh, hh - 0.5f);
    const float m00 = __fmul_rn(r1.x, hw), m10 = __fmul_rn(r1.y, hh), m01 = __fmul_rn(r1.z, hw), m11 = __fmul_rn(r1.w, hh);
    const float det = __fsub_rn(__fmul_rn(m00, m11), __fmul_rn(m01, m10));
    const float a00 = __fdiv_rn(m11, det), a01 = __fdiv_rn(-m01, det), a10 = __fdiv_rn(-m10, det), a11 = __fdiv_rn(m00, det);
    // conservative pixel bounding box of centre +- RS*(+-3,+-3); the exact |p| <= 3 test is per pixel in the blend.
    // NaN lanes (D == 0 / negative eigenvalue, SURVEY.md §7 hard part 6) and depth >= 1 (LESS against the cleared
    // 1.0, graphics_pipeline.cc:79-81) fail the comparisons and emit nothing.
    const float ex = 3.f * (fabsf(m00) + fabsf(m01)), ey = 3.f * (fabsf(m10) + fabsf(m11));
    const float fx0 = fmaxf(ceilf(cpx - ex - 0.01f), 0.f), fx1 = fminf(floorf(cpx + ex + 0.01f), static_cast<float>(width) - 1.f);
    const float fy0 = fmaxf(ceilf(cpy - ey - 0.01f), band_lo), fy1 = fminf(floorf(cpy + ey + 0.01f), band_hi);
    uint32_t x0 = 1, x1 = 0, y0 = 1, y1 = 0;
    if (r0.z < 1.f && fx0 <= fx1 && fy0 <= fy1 && det == det && fabsf(det) <= 3.0e38f && ex <= 3.0e38f && ey <= 3.0e38f) {
      x0 = static_cast<uint32_t>(fx0); x1 = static_cast<uint32_t>(fx1);
      y0 = static_cast<uint32_t>(fy0); y1 = static_cast<uint32_t>(fy1);
      const uint32_t bx0 = x0 / kBinW, by0 = y0 / kBinH - bin_y0;
      const uint32_t bw = x1 / kBinW - bx0 + 1, bh = y1 / kBinH - bin_y0 - by0 + 1;
      count = bw * bh;
      rect = bx0 | (by0 << 8) | (bw << 16) | (bh << 24);
    }
    rrec[i * 3 + 0] = make_float4(a00, a01, a10, a11);
    rrec[i * 3 + 1] = make_float4(cpx, cpy, __saturatef(r2.x), __saturatef(r2.y));  // UNORM target clamps the source
    rrec[i * 3 + 2] = make_float4(__saturatef(r2.z), r2.w, __uint_as_float(x0 | (x1 << 16)), __uint_as_float(y0 | (y1 << 16)));
  }
  // block exclusive scan of count
  uint32_t v = count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= static_cast<uint32_t>(o)) v += t;
  }
  if (lane == 31) s_wsum[warp] = v;
  __syncthreads();
  uint32_t wb = 0;
  for (uint32_t w = 0; w < warp; ++w) wb += s_wsum[w];
  const uint32_t excl = wb + v - count;
  s_off[tid] = excl;
  s_rect[tid] = rect;
  if (tid == kPairThreads - 1) s_off[kPairThreads] = excl + count;
  __syncthreads();
  const uint32_t total = s_off[kPairThreads];
  if (warp == 0) {
    uint32_t base = scan_lookback_warp(scan_desc, ticket, total);
    if (lane == 0) {
      s_base = base;
      if (ticket == nblocks - 1) {
        uint64_t d = static_cast<uint64_t>(base) + total;
        ctrl->pair_count = static_cast<uint32_t>(d < max_pairs ? d : max_pairs);
        if (d > max_pairs) ctrl->pair_overflow = 1u;
      }
    }
  }
  __syncthreads();
  const uint64_t base = s_base;

  // load-balanced expansion: output element e belongs to the splat whose offset interval contains it
  for (uint32_t e = tid; e < total; e += kPairThreads) {
    uint32_t lo = 0, hi = kPairThreads;  // last s with s_off[s] <= e
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const uint32_t mid = (lo + hi) >> 1;
      if (s_off[mid] <= e) lo = mid; else hi = mid;
    }
    const uint32_t k = e - s_off[lo], rc = s_rect[lo];
    const uint32_t bx0 = rc & 255u, by0 = (rc >> 8) & 255u, bw = (rc >> 16) & 255u;
    const uint64_t g = base + e;
    if (g < max_pairs) {  // overflow drops the farthest pairs (the list is nearest-first)
      pair_bin[g] = (by0 + k / bw) * bins_x + bx0 + k % bw;
      pair_rank[g] = ticket * kPairThreads + lo;
    }
  }
}

// Bin boundaries in the bin-sorted pair list.  ranges must be zero on entry (empty bins stay [0,0)).
__global__ void __launch_bounds__(256)
k_bin_ranges(const Control* __restrict__ ctrl, const uint32_t* __restrict__ bin_sorted, uint2* __restrict__ ranges) {
  const uint32_t D = ctrl->pair_count;
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < D; i += gridDim.x * 256) {
    const uint32_t t = bin_sorted[i];
    if (i == 0 || bin_sorted[i - 1] != t) ranges[t].x = i;
    if (i == D - 1 || bin_sorted[i + 1] != t) ranges[t].y = i + 1;
  }
}

// Parity tap: ids and instance records in sorted (far -> near) order.
__global__ void __launch_bounds__(256)
k_gather_sorted(const Control* __restrict__ ctrl, const uint32_t* __restrict__ sorted_slots,
                const uint32_t* __restrict__ vis_id, const float4* __restrict__ inst, uint32_t* __restrict__ ids_out,
                float4* __restrict__ inst_out) {
  const uint32_t V = ctrl->visible_count;
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < V; i += gridDim.x * 256) {
    const uint32_t s = sorted_slots[i];
    if (ids_out) ids_out[i] = vis_id[s];
    if (inst_out) {
      inst_out[i * 3 + 0] = inst[s * 3 + 0];
      inst_out[i * 3 + 1] = inst[s * 3 + 1];
      inst_out[i * 3 + 2] = inst[s * 3 + 2];
    }
  }
}

void launch_make_pairs(const FrameParams* d_fp, Control* d_ctrl, unsigned long long* d_scan_desc,
                       const uint32_t* d_sorted_slots, const float* d_inst, uint32_t max_visible, uint64_t max_pairs,
                       float* d_rrec, uint32_t* d_pair_bin, uint32_t* d_pair_rank, cudaStream_t stream) {
  uint32_t nb = pairs_num_blocks(max_visible);
  if (nb == 0) return;
  k_make_pairs<<<nb, kPairThreads, 0, stream>>>(d_fp, d_ctrl, d_scan_desc, d_sorted_slots,
                                                reinterpret_cast<const float4*>(d_inst), max_pairs,
                                                reinterpret_cast<float4*>(d_rrec), d_pair_bin, d_pair_rank);
}

void launch_bin_ranges(const Control* d_ctrl, const uint32_t* d_pair_bin_sorted, uint64_t max_pairs, uint2* d_ranges,
                       cudaStream_t stream) {
  uint64_t want = (max_pairs + 255) / 256;
  int blocks = static_cast<int>(want < 148 * 16 ? (want ? want : 1) : 148 * 16);
  k_bin_ranges<<<blocks, 256, 0, stream>>>(d_ctrl, d_pair_bin_sorted, d_ranges);
}

void launch_gather_sorted(const Control* d_ctrl, const uint32_t* d_sorted_slots, const uint32_t* d_vis_id,
                          const float* d_inst, uint32_t max_visible, uint32_t* d_ids_out, float* d_inst_out,
                          cudaStream_t stream) {
  uint32_t want = (max_visible + 255) / 256;
  int blocks = static_cast<int>(want < 148 * 8 ? (want ? want : 1) : 148 * 8);
  k_gather_sorted<<<blocks, 256, 0, stream>>>(d_ctrl, d_sorted_slots, d_vis_id, reinterpret_cast<const float4*>(d_inst),
                                              d_ids_out, reinterpret_cast<float4*>(d_inst_out));
}

}  // namespace vkgsb
