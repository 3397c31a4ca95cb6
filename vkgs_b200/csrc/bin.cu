// Stage 3a: coarse binning.  No counterpart in the reference, whose instanced quads go through the hardware
// rasteriser in sorted order (vkCmdDrawIndexedIndirect, engine.cc:1472-1480); here the globally sorted splat list is
// turned into per-bin lists that keep that order:
//   k_make_pairs   one thread per sorted rank, walked NEAREST FIRST: reads the pixel bounding box k_project left in
//                  the splat's raster record and emits one (bin, slot) pair per 64x64-pixel bin the box touches, at
//                  offsets from an ordered (decoupled look-back) scan - the pair list is rank-major.  Also counts
//                  pairs per bin.
//   k_bin_scan     per-bin counts -> every bin's [begin,end) in the sorted list + the onesweep digit histograms,
//                  so neither a histogram pass nor a boundary search ever re-reads the pairs.
//   stable onesweep sort of the pairs by bin id (sort.cu) -> each bin's pairs are contiguous, still nearest-first.
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
             const uint32_t* __restrict__ sorted_slots, const float4* __restrict__ rrec, uint64_t max_pairs,
             uint32_t* __restrict__ pair_bin, uint32_t* __restrict__ pair_slot) {
  __shared__ uint32_t s_off[kPairThreads + 1];  // exclusive offsets of this block's splats
  __shared__ uint32_t s_rect[kPairThreads];     // bx0 | by0 << 8 | bw << 16 | bh << 24   (<= 64 x 64 bins)
  __shared__ uint32_t s_slot[kPairThreads];
  __shared__ uint32_t s_wsum[kPairThreads / 32];
  __shared__ uint32_t s_bins[kMaxBins];         // this block's pairs per bin
  __shared__ uint32_t s_ticket, s_base;

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t V = ctrl->visible_count;
  const uint32_t nblocks = (V + kPairThreads - 1) / kPairThreads;
  if (tid == 0) s_ticket = atomicAdd(&ctrl->pairs_ticket, 1u);
  __syncthreads();
  const uint32_t ticket = s_ticket;
  if (ticket >= nblocks) return;
  const uint32_t bins_x = fpp->bins_x, bin_y0 = fpp->bin_y0, nbins = bins_x * (fpp->bin_y1 - bin_y0);
  for (uint32_t b = tid; b < nbins; b += kPairThreads) s_bins[b] = 0u;

  const uint32_t i = ticket * kPairThreads + tid;  // i-th nearest splat
  uint32_t count = 0, rect = 0, slot = 0;
  if (i < V) {
    slot = sorted_slots[V - 1 - i];  // ascending key = far -> near (rank.comp:39): walk it backwards
    const float4 q2 = __ldg(rrec + slot * 3 + 2);
    const uint32_t bxw = __float_as_uint(q2.z), byw = __float_as_uint(q2.w);
    const uint32_t x0 = bxw & 0xffffu, x1 = bxw >> 16, y0 = byw & 0xffffu, y1 = byw >> 16;
    if (x0 <= x1 && y0 <= y1) {  // empty box: culled by depth, NaN lane, or outside the band
      const uint32_t bx0 = x0 / kBinW, by0 = y0 / kBinH - bin_y0;
      const uint32_t bw = x1 / kBinW - bx0 + 1, bh = y1 / kBinH - bin_y0 - by0 + 1;
      count = bw * bh;
      rect = bx0 | (by0 << 8) | (bw << 16) | (bh << 24);
    }
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
  s_slot[tid] = slot;
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
      const uint32_t bin = (by0 + k / bw) * bins_x + bx0 + k % bw;
      pair_bin[g] = bin;
      pair_slot[g] = s_slot[lo];
      atomicAdd(&s_bins[bin], 1u);
    }
  }
  __syncthreads();
  for (uint32_t b = tid; b < nbins; b += kPairThreads) {
    const uint32_t c = s_bins[b];
    if (c) atomicAdd(&ctrl->bin_count[b], c);
  }
}

// From the per-bin pair counts: every bin's [begin,end) in the bin-sorted pair list (an exclusive scan) and the
// digit histograms the onesweep passes need.  All blocks also clear the look-back words those passes will use.
__global__ void __launch_bounds__(1024)
k_bin_scan(const FrameParams* __restrict__ fpp, Control* __restrict__ ctrl, uint2* __restrict__ ranges,
           uint32_t* __restrict__ lookback, uint32_t max_parts, int npass) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_hist[2 * 256];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t D = ctrl->pair_count;
  const uint32_t nparts = (D + 4095u) / 4096u;
  for (int p = 0; p < npass; ++p) {
    uint32_t* lb = lookback + static_cast<size_t>(p) * max_parts * 256;
    for (size_t i = static_cast<size_t>(blockIdx.x) * 1024 + tid; i < static_cast<size_t>(nparts) * 256;
         i += static_cast<size_t>(gridDim.x) * 1024)
      lb[i] = 0u;
  }
  if (blockIdx.x != 0) return;
  const uint32_t nbins = fpp->bins_x * (fpp->bin_y1 - fpp->bin_y0);
  if (tid < 512) s_hist[tid] = 0u;
  uint32_t c[4], sum = 0;  // 4 consecutive bins per thread (kMaxBins = 4096)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t b = tid * 4 + k;
    c[k] = b < nbins ? ctrl->bin_count[b] : 0u;
    sum += c[k];
  }
  uint32_t v = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= static_cast<uint32_t>(o)) v += t;
  }
  if (lane == 31) s_warp[warp] = v;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane], x = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= static_cast<uint32_t>(o)) x += t;
    }
    s_warp[lane] = x - w;
  }
  __syncthreads();
  uint32_t start = s_warp[warp] + v - sum;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t b = tid * 4 + k;
    if (b < nbins) {
      ranges[b] = make_uint2(start, start + c[k]);
      if (c[k]) {
        atomicAdd(&s_hist[b & 255u], c[k]);
        if (npass > 1) atomicAdd(&s_hist[256 + ((b >> 8) & 255u)], c[k]);
      }
    }
    start += c[k];
  }
  __syncthreads();
  if (tid < 256u * static_cast<uint32_t>(npass)) ctrl->hist_bin[tid] = s_hist[tid];
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
                       const uint32_t* d_sorted_slots, const float* d_rrec, uint32_t max_visible, uint64_t max_pairs,
                       uint32_t* d_pair_bin, uint32_t* d_pair_slot, cudaStream_t stream) {
  uint32_t nb = pairs_num_blocks(max_visible);
  if (nb == 0) return;
  k_make_pairs<<<nb, kPairThreads, 0, stream>>>(d_fp, d_ctrl, d_scan_desc, d_sorted_slots,
                                                reinterpret_cast<const float4*>(d_rrec), max_pairs, d_pair_bin,
                                                d_pair_slot);
}

void launch_bin_scan(const FrameParams* d_fp, Control* d_ctrl, uint2* d_ranges, uint32_t* d_lookback, uint64_t max_pairs,
                     int npass, cudaStream_t stream) {
  const uint32_t max_parts = sort_max_parts(static_cast<uint32_t>(max_pairs));
  int blocks = static_cast<int>(max_parts / 64 + 1);
  if (blocks > 64) blocks = 64;
  k_bin_scan<<<blocks, 1024, 0, stream>>>(d_fp, d_ctrl, d_ranges, d_lookback, max_parts, npass);
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
