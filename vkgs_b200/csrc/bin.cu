// Stage 3a: coarse binning.  No counterpart in the reference, whose instanced quads go through the hardware
// rasteriser in sorted order (vkCmdDrawIndexedIndirect, engine.cc:1472-1480); here the globally sorted splat list is
// split into one list per coarse bin (<= 256 bins of >= 128x128 pixels, common.cuh) that keeps that order - a stable
// multi-split with a small, known bin count, done by counting and direct placement instead of radix passes, and
// balanced by (bin, splat) PAIRS, not by splats: the few splats nearest to the camera cover every bin, so the first
// 1024 ranks can own more pairs than the next 100 000.
//   k_bin_tiles     one CTA per tile of 1024 sorted ranks, walked NEAREST FIRST: reads the coarse-bin box k_project
//                   left for each splat (4 B, an L2-resident array) -> pairs per tile.
//   k_bin_tile_scan one CTA: prefix of the tile "costs" (pairs + a fixed overhead per tile), the pair-capacity cut
//                   (the farthest pairs are dropped once the running pair count would exceed max_pairs), the number
//                   of work items = cost / kBinQuota.
//   k_bin_count     one CTA per work item = kBinQuota consecutive cost units = a slice of one or a few tiles' pairs:
//                   pairs per (item, bin) -> item_bin[bin][item].
//   k_bin_colscan   one CTA per bin: exclusive scan of its row over the items (where each item's pairs start in the
//                   bin's list) and the bin total.
//   k_bin_place     one CTA per item again: re-enumerates its pairs, ranks them stably (warp match on the bin id +
//                   per-warp bin cursors) and stores every splat slot straight at its final position.  No (bin, slot)
//                   pair list ever exists in memory and nothing spins on another CTA.
// History (profiles/): binning straight to 16x16 tiles spent 3.6 ms of a 4.1 ms frame sorting 9e7 pairs of which early
// termination consumed a few percent; 64x64 bins still moved 1e7 pairs through two onesweep passes (0.35 ms of a
// 0.8 ms frame); a decoupled look-back over the pair offsets serialised into ~150 L2 round trips because the whole
// list is one wave of tiles; one CTA per 1024 ranks left the nearest tile's 1e4..1e5 pairs to a single CTA (0.18 ms).
// The blend stage filters a coarse bin's list down to its own 64x64 pixels and refines that to 16x8 sub-tiles on chip.
#include "common.cuh"
#include "kernels.h"

namespace vkgsb {

constexpr int kBinThreads = 256;
constexpr int kBinWarps = kBinThreads / 32;
constexpr int kBinItems = 4;                        // ranks per thread
constexpr int kBinTile = kBinThreads * kBinItems;   // ranks per tile: 1024
constexpr uint32_t kBinQuota = 4096;                // cost units per work item
constexpr uint32_t kBinTileCost = 256;              // fixed cost of touching a tile (bounds tiles per item)

uint32_t bin_num_tiles(uint32_t max_visible) { return (max_visible + kBinTile - 1) / kBinTile; }
uint32_t bin_max_items(uint32_t max_visible, uint64_t max_pairs) {
  return static_cast<uint32_t>((max_pairs + static_cast<uint64_t>(kBinTileCost) * bin_num_tiles(max_visible)) / kBinQuota + 2);
}

__device__ __forceinline__ uint32_t rect_pairs(uint32_t rect) { return (rect >> 16 & 255u) * (rect >> 24); }

// Tile t's ranks, thread-major (thread i holds ranks 4i..4i+3 of the tile, nearest first).
__device__ __forceinline__ void load_tile(uint32_t t, uint32_t V, const uint32_t* __restrict__ sorted_slots,
                                          const uint32_t* __restrict__ bin_rect, uint32_t slot[kBinItems],
                                          uint32_t rect[kBinItems]) {
  const uint32_t r0 = t * kBinTile + threadIdx.x * kBinItems;
#pragma unroll
  for (int it = 0; it < kBinItems; ++it) {
    const uint32_t i = r0 + it;
    slot[it] = (i < V) ? __ldg(sorted_slots + (V - 1 - i)) : 0u;  // ascending key = far -> near (rank.comp:39): backwards
  }
#pragma unroll
  for (int it = 0; it < kBinItems; ++it)
    rect[it] = (r0 + it < V) ? __ldg(bin_rect + slot[it]) : 0u;  // 0: empty box (depth cull, NaN lane, outside the band)
}

__global__ void __launch_bounds__(kBinThreads)
k_bin_tiles(const Control* __restrict__ ctrl, const uint32_t* __restrict__ sorted_slots,
            const uint32_t* __restrict__ bin_rect, uint32_t* __restrict__ tile_pairs) {
  __shared__ uint32_t s_sum[kBinWarps];
  const uint32_t V = ctrl->visible_count, t = blockIdx.x;
  if (t * kBinTile >= V) return;
  uint32_t slot[kBinItems], rect[kBinItems];
  load_tile(t, V, sorted_slots, bin_rect, slot, rect);
  uint32_t c = 0;
#pragma unroll
  for (int it = 0; it < kBinItems; ++it) c += rect_pairs(rect[it]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31u) == 0) s_sum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t a = 0;
    for (int w = 0; w < kBinWarps; ++w) a += s_sum[w];
    tile_pairs[t] = a;
  }
}

// Exclusive block scan of one value per thread (1024 threads).
__device__ __forceinline__ unsigned long long block_scan_1024(unsigned long long v, unsigned long long* s_warp /*[32]*/,
                                                              unsigned long long* total) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  unsigned long long x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long u = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= static_cast<uint32_t>(o)) x += u;
  }
  __syncthreads();
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  unsigned long long base = 0, all = 0;
  for (uint32_t w = 0; w < 32; ++w) {
    const unsigned long long sw = s_warp[w];
    if (w < warp) base += sw;
    all += sw;
  }
  *total = all;
  return base + x - v;
}

// tile_pairs[t] -> tile_cost[t] = exclusive prefix of (pairs + kBinTileCost) over the kept tiles, tile_cost[cut] = end.
__global__ void __launch_bounds__(1024)
k_bin_tile_scan(Control* __restrict__ ctrl, uint64_t max_pairs, const uint32_t* __restrict__ tile_pairs,
                uint32_t* __restrict__ tile_cost) {
  __shared__ unsigned long long s_warp[32];
  __shared__ uint32_t s_cut;
  __shared__ unsigned long long s_kept;
  const uint32_t tid = threadIdx.x;
  const uint32_t V = ctrl->visible_count;
  const uint32_t ntiles = (V + kBinTile - 1) / kBinTile;
  if (tid == 0) {
    s_cut = ntiles;
    s_kept = 0ull;
  }
  __syncthreads();
  // ---- pass 1: capacity cut = first tile whose inclusive pair count exceeds max_pairs (the prefix is monotone)
  unsigned long long carry = 0;
  for (uint32_t t0 = 0; t0 < ntiles; t0 += 1024) {
    const uint32_t t = t0 + tid;
    const unsigned long long v = t < ntiles ? tile_pairs[t] : 0u;
    unsigned long long all;
    const unsigned long long ex = carry + block_scan_1024(v, s_warp, &all);
    if (t < ntiles && ex <= max_pairs && ex + v > max_pairs) {
      s_cut = t;
      s_kept = ex;
    }
    carry += all;
  }
  __syncthreads();
  const uint32_t cut = s_cut;
  const bool overflow = cut != ntiles;
  const unsigned long long pairs = overflow ? max_pairs : carry;
  const uint32_t partial = overflow ? static_cast<uint32_t>(max_pairs - s_kept) : 0u;  // pairs kept of tile `cut`
  // ---- pass 2: cost prefix over the kept tiles
  carry = 0;
  for (uint32_t t0 = 0; t0 < cut; t0 += 1024) {
    const uint32_t t = t0 + tid;
    const unsigned long long v = t < cut ? static_cast<unsigned long long>(tile_pairs[t]) + kBinTileCost : 0ull;
    unsigned long long all;
    const unsigned long long ex = carry + block_scan_1024(v, s_warp, &all);
    if (t < cut) tile_cost[t] = static_cast<uint32_t>(ex);
    carry += all;
  }
  if (tid == 0) {
    tile_cost[cut] = static_cast<uint32_t>(carry);
    uint32_t kept_tiles = cut;
    if (partial) {  // the list is nearest-first: the cut tile keeps its first `partial` pairs, farther ones are dropped
      carry += kBinTileCost + partial;
      tile_cost[cut + 1] = static_cast<uint32_t>(carry);
      kept_tiles = cut + 1;
    }
    ctrl->tile_cut = kept_tiles;
    ctrl->bin_cost = static_cast<uint32_t>(carry);
    ctrl->bin_items = static_cast<uint32_t>((carry + kBinQuota - 1) / kBinQuota);
    ctrl->pair_count = static_cast<uint32_t>(pairs);
    ctrl->pair_overflow = overflow ? 1u : 0u;
  }
}

// Shared memory of one work item.
struct BinShared {
  uint32_t off[kBinTile + 1];   // exclusive pair offsets of the current tile's ranks
  uint32_t rect[kBinTile];      // bx0 | by0 << 8 | bw << 16 | bh << 24
  uint32_t slot[kBinTile];
  uint32_t wsum[kBinWarps];
};

// Loads tile t into shared memory with the exclusive scan of its ranks' pair counts.  Ends with a barrier.
__device__ __forceinline__ void stage_tile(BinShared& sh, uint32_t t, uint32_t V, const uint32_t* __restrict__ sorted_slots,
                                           const uint32_t* __restrict__ bin_rect) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  uint32_t slot[kBinItems], rect[kBinItems], cnt[kBinItems];
  load_tile(t, V, sorted_slots, bin_rect, slot, rect);
  uint32_t sum = 0;
#pragma unroll
  for (int it = 0; it < kBinItems; ++it) {
    cnt[it] = rect_pairs(rect[it]);
    sum += cnt[it];
  }
  uint32_t x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= static_cast<uint32_t>(o)) x += u;
  }
  __syncthreads();  // the previous slice is done with the arrays
  if (lane == 31) sh.wsum[warp] = x;
  __syncthreads();
  uint32_t run = x - sum;
  for (uint32_t w = 0; w < warp; ++w) run += sh.wsum[w];
#pragma unroll
  for (int it = 0; it < kBinItems; ++it) {
    sh.off[tid * kBinItems + it] = run;
    sh.rect[tid * kBinItems + it] = rect[it];
    sh.slot[tid * kBinItems + it] = slot[it];
    run += cnt[it];
  }
  if (tid == kBinThreads - 1) sh.off[kBinTile] = run;
  __syncthreads();
}

// Pair e of the staged tile -> (index of its rank in the tile, bin id): e belongs to the rank whose offset interval
// contains it.
__device__ __forceinline__ uint32_t pair_bin_of(const BinShared& sh, uint32_t e, uint32_t cbins_x, uint32_t* which) {
  uint32_t lo = 0, hi = kBinTile;  // last s with off[s] <= e
#pragma unroll
  for (int it = 0; it < 10; ++it) {
    const uint32_t mid = (lo + hi) >> 1;
    if (sh.off[mid] <= e) lo = mid; else hi = mid;
  }
  const uint32_t k = e - sh.off[lo], rc = sh.rect[lo];
  const uint32_t bx0 = rc & 255u, by0 = (rc >> 8) & 255u, bw = (rc >> 16) & 255u;
  *which = lo;
  return (by0 + k / bw) * cbins_x + bx0 + k % bw;
}

// Last tile t in [0, cut) with tile_cost[t] <= a.  256-ary search with block-wide vote counts; all threads return it.
__device__ __forceinline__ uint32_t find_tile(const uint32_t* __restrict__ tile_cost, uint32_t cut, uint32_t a) {
  uint32_t lo = 0, hi = cut;
  while (hi - lo > 1) {
    const uint32_t step = (hi - lo + kBinThreads - 1) / kBinThreads;
    const uint32_t pos = lo + threadIdx.x * step;
    const int below = __syncthreads_count(pos < hi && __ldg(tile_cost + pos) <= a);  // monotone: a prefix of the threads
    lo = lo + (below - 1) * step;
    hi = min(lo + step, hi);
  }
  return lo;
}

// The work item's slices: for every tile whose cost interval [cb, cb + kBinTileCost + pairs) meets the item's
// [A, B), the tile's pair range [lo, hi) that falls inside.  f(t, lo, hi) is called by all threads, block-uniformly.
template <class F>
__device__ __forceinline__ void for_each_slice(const Control* __restrict__ ctrl, const uint32_t* __restrict__ tile_cost,
                                               uint32_t item, F f) {
  const uint32_t cut = ctrl->tile_cut;
  const uint32_t A = item * kBinQuota, B = min(A + kBinQuota, ctrl->bin_cost);
  for (uint32_t t = find_tile(tile_cost, cut, A); t < cut; ++t) {
    const uint32_t cb = __ldg(tile_cost + t), ce = __ldg(tile_cost + t + 1);
    if (cb >= B) break;
    const uint32_t p0 = cb + kBinTileCost;  // cost position of the tile's pair 0
    const uint32_t lo = A > p0 ? A - p0 : 0u, hi = min(B, ce) > p0 ? min(B, ce) - p0 : 0u;
    if (hi > lo) f(t, lo, hi);
  }
}

__global__ void __launch_bounds__(kBinThreads)
k_bin_count(const FrameParams* __restrict__ fpp, const Control* __restrict__ ctrl,
            const uint32_t* __restrict__ sorted_slots, const uint32_t* __restrict__ bin_rect,
            const uint32_t* __restrict__ tile_cost, uint32_t item_stride, uint32_t* __restrict__ item_bin) {
  __shared__ BinShared sh;
  __shared__ uint32_t s_cnt[kMaxCoarseBins];
  const uint32_t tid = threadIdx.x;
  const uint32_t V = ctrl->visible_count, cbins_x = fpp->cbins_x, ncbins = fpp->ncbins, nitems = ctrl->bin_items;
  for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    s_cnt[tid] = 0u;
    for_each_slice(ctrl, tile_cost, item, [&](uint32_t t, uint32_t lo, uint32_t hi) {
      stage_tile(sh, t, V, sorted_slots, bin_rect);
#pragma unroll 2
      for (uint32_t e = lo + tid; e < hi; e += kBinThreads) {
        uint32_t which;
        atomicAdd(&s_cnt[pair_bin_of(sh, e, cbins_x, &which)], 1u);
      }
    });
    __syncthreads();
    if (tid < ncbins) item_bin[static_cast<size_t>(tid) * item_stride + item] = s_cnt[tid];
    __syncthreads();
  }
}

// Row b of item_bin: exclusive scan over the items, in place; bin_total[b] = the bin's pair count.
__global__ void __launch_bounds__(1024)
k_bin_colscan(const Control* __restrict__ ctrl, uint32_t item_stride, uint32_t* __restrict__ item_bin,
              uint32_t* __restrict__ bin_total) {
  __shared__ unsigned long long s_warp[32];
  const uint32_t tid = threadIdx.x, nitems = ctrl->bin_items;
  uint32_t* row = item_bin + static_cast<size_t>(blockIdx.x) * item_stride;
  unsigned long long carry = 0;
  for (uint32_t i0 = 0; i0 < nitems; i0 += 1024) {
    const uint32_t i = i0 + tid;
    const unsigned long long v = i < nitems ? row[i] : 0u;
    unsigned long long all;
    const unsigned long long ex = carry + block_scan_1024(v, s_warp, &all);
    if (i < nitems) row[i] = static_cast<uint32_t>(ex);
    carry += all;
  }
  if (tid == 0) bin_total[blockIdx.x] = static_cast<uint32_t>(carry);
}

__global__ void __launch_bounds__(kBinThreads)
k_bin_place(const FrameParams* __restrict__ fpp, const Control* __restrict__ ctrl,
            const uint32_t* __restrict__ sorted_slots, const uint32_t* __restrict__ bin_rect,
            const uint32_t* __restrict__ tile_cost, uint32_t item_stride, const uint32_t* __restrict__ item_bin,
            const uint32_t* __restrict__ bin_total, uint2* __restrict__ ranges, uint32_t* __restrict__ bin_slots) {
  __shared__ BinShared sh;
  __shared__ uint32_t s_wcnt[kBinWarps][kMaxCoarseBins];  // pairs per (warp, bin) of the slice, then each warp's cursor
  __shared__ uint32_t s_cursor[kMaxCoarseBins];           // next free position of every bin's list for this item
  __shared__ uint32_t s_scan[kBinWarps];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t V = ctrl->visible_count, cbins_x = fpp->cbins_x, ncbins = fpp->ncbins, nitems = ctrl->bin_items;
  if (blockIdx.x >= nitems) return;
  // ---- where each bin's list starts: exclusive scan of the totals
  uint32_t begin;
  {
    const uint32_t mine = tid < ncbins ? bin_total[tid] : 0u;
    uint32_t x = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= static_cast<uint32_t>(o)) x += u;
    }
    if (lane == 31) s_scan[warp] = x;
    __syncthreads();
    begin = x - mine;
    for (uint32_t w = 0; w < warp; ++w) begin += s_scan[w];
    if (blockIdx.x == 0) ranges[tid] = make_uint2(begin, begin + mine);
  }
  uint32_t* wcnt = s_wcnt[warp];
  for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
  __syncthreads();  // the previous item's last slice is done with the cursors
  s_cursor[tid] = tid < ncbins ? begin + item_bin[static_cast<size_t>(tid) * item_stride + item] : 0u;
  for_each_slice(ctrl, tile_cost, item, [&](uint32_t t, uint32_t lo, uint32_t hi) {
    __syncthreads();  // every warp is done with the previous slice's cursors
#pragma unroll
    for (int w = 0; w < kBinWarps; ++w) s_wcnt[w][tid] = 0u;
    stage_tile(sh, t, V, sorted_slots, bin_rect);  // barriers inside order the zeroing too
    // the slice's pairs in 8 contiguous warp shares, each a multiple of 32
    const uint32_t share = ((hi - lo + kBinWarps * 32 - 1) / (kBinWarps * 32)) * 32;
    const uint32_t w_lo = min(lo + warp * share, hi), w_hi = min(w_lo + share, hi);
    // ---- pass A: pairs per (warp, bin)
    for (uint32_t e = w_lo + lane; e < w_hi; e += 32) {
      uint32_t which;
      atomicAdd(&wcnt[pair_bin_of(sh, e, cbins_x, &which)], 1u);
    }
    __syncthreads();
    // ---- thread b: bin b across the warps -> every warp's first position in the bin's list
    {
      uint32_t run = s_cursor[tid];
#pragma unroll
      for (int w = 0; w < kBinWarps; ++w) {
        const uint32_t c = s_wcnt[w][tid];
        s_wcnt[w][tid] = run;
        run += c;
      }
      s_cursor[tid] = run;
    }
    __syncthreads();
    // ---- pass B: the warp walks its share in order, 32 pairs at a time; lanes with the same bin are ranked by lane
    //      (= pair order) and the bin's cursor advances by the group size: stable
    for (uint32_t e0 = w_lo; e0 < w_hi; e0 += 32) {
      const uint32_t e = e0 + lane;
      const bool valid = e < w_hi;
      uint32_t which = 0, bin = 0xffffffffu;
      if (valid) bin = pair_bin_of(sh, e, cbins_x, &which);
      const uint32_t peers = __match_any_sync(0xffffffffu, bin);
      const uint32_t leader = __ffs(peers) - 1;
      uint32_t prev = 0;
      if (valid && lane == leader) {
        prev = wcnt[bin];
        wcnt[bin] = prev + __popc(peers);
      }
      prev = __shfl_sync(0xffffffffu, prev, leader);
      if (valid) bin_slots[prev + __popc(peers & ((1u << lane) - 1u))] = sh.slot[which];
      __syncwarp();
    }
  });
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

void launch_bin(const FrameParams* d_fp, uint32_t ncbins, Control* d_ctrl, const uint32_t* d_sorted_slots,
                const uint32_t* d_bin_rect, uint32_t max_visible, uint64_t max_pairs, const BinScratch& w,
                uint2* d_ranges, uint32_t* d_bin_slots, cudaStream_t stream) {
  const uint32_t tiles = bin_num_tiles(max_visible);
  if (tiles == 0 || ncbins == 0) return;
  const uint32_t max_items = bin_max_items(max_visible, max_pairs);
  const uint32_t items = max_items < 148u * 8u ? max_items : 148u * 8u;  // persistent: CTAs stride over the items
  k_bin_tiles<<<tiles, kBinThreads, 0, stream>>>(d_ctrl, d_sorted_slots, d_bin_rect, w.tile_pairs);
  k_bin_tile_scan<<<1, 1024, 0, stream>>>(d_ctrl, max_pairs, w.tile_pairs, w.tile_cost);
  k_bin_count<<<items, kBinThreads, 0, stream>>>(d_fp, d_ctrl, d_sorted_slots, d_bin_rect, w.tile_cost, w.item_stride,
                                                 w.item_bin);
  k_bin_colscan<<<ncbins, 1024, 0, stream>>>(d_ctrl, w.item_stride, w.item_bin, w.bin_total);
  k_bin_place<<<items, kBinThreads, 0, stream>>>(d_fp, d_ctrl, d_sorted_slots, d_bin_rect, w.tile_cost, w.item_stride,
                                                 w.item_bin, w.bin_total, d_ranges, d_bin_slots);
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
