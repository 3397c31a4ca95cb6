// Stage 3a: coarse binning.  No counterpart in the reference, whose instanced quads go through the hardware
// rasteriser in sorted order (vkCmdDrawIndexedIndirect, engine.cc:1472-1480); here the globally sorted splat list is
// split into one list per coarse bin (<= 256 bins of >= 128x128 pixels, common.cuh) that keeps that order - a stable
// multi-split with a small, known bin count, done by counting and direct placement instead of radix passes.
// A splat's box in coarse bins is a rectangle, so "pairs per bin" of any set of splats is four +-1 corner updates per
// splat into a 2-D difference array and one 2-D prefix sum; only the placement itself touches every (bin, splat) pair.
// Placement is balanced by PAIRS, not by splats: the few splats nearest to the camera cover every bin, so the first
// 1024 ranks can own more pairs than the next 100 000.
//   k_bin_tiles  one CTA per tile of 1024 sorted ranks, walked NEAREST FIRST: reads the coarse-bin box k_project left
//                for each splat (4 B, an L2-resident array) -> pairs per (bin, tile) and per tile.
//   k_bin_scan   one CTA per bin: exclusive scan of the bin's row over the tiles (where each tile's pairs start in the
//                bin's list); one more CTA: the work items - a tile is one item, a tile with more than kBinQuota pairs
//                is split into several - and the pair-capacity cut (the farthest pairs are dropped once the running
//                pair count would exceed max_pairs).
//   k_bin_place  one CTA per work item: enumerates its pairs 32 at a time per warp, ranks them stably (warp match on
//                the bin id + per-warp bin cursors) and stores every splat slot straight at its final position.  No
//                (bin, slot) pair list ever exists in memory and nothing spins on another CTA.
// History (profiles/): binning straight to 16x16 tiles spent 3.6 ms of a 4.1 ms frame sorting 9e7 pairs of which early
// termination consumed a few percent; 64x64 bins still moved 1e7 pairs through two onesweep passes (0.35 ms of a
// 0.8 ms frame); a decoupled look-back over the pair offsets serialised into ~150 L2 round trips because the whole
// list is one wave of tiles; one CTA per 1024 ranks left the nearest tile's 1e4..1e5 pairs to a single CTA (0.18 ms);
// a binary search per pair cost 150 instructions per pair (0.13 ms); items cut at fixed pair counts across tile
// borders re-staged every tile ~3 times and paid a prologue per item (35 M instructions for 5 M pairs).
// The blend stage filters a coarse bin's list down to its own 64x64 pixels and refines that to 16x8 sub-tiles on chip.
#include "common.cuh"
#include "kernels.h"

namespace vkgsb {

constexpr int kBinThreads = 256;
constexpr int kBinWarps = kBinThreads / 32;
constexpr int kBinItems = 4;                        // ranks per thread
constexpr int kBinTile = kBinThreads * kBinItems;   // ranks per tile: 1024
constexpr uint32_t kBinQuota = 8192;                // most pairs one work item places
constexpr int kDiffMax = kMaxCoarseBins + 2 * kMaxCoarseBins + 8;  // (cols + 1) * (rows + 1) <= bins + cols + rows + 1

uint32_t bin_num_tiles(uint32_t max_visible) { return (max_visible + kBinTile - 1) / kBinTile; }
uint32_t bin_max_items(uint32_t max_visible, uint64_t max_pairs) {
  return static_cast<uint32_t>(max_pairs / kBinQuota + bin_num_tiles(max_visible) + 1);
}
size_t bin_slots_capacity(uint64_t max_pairs) {  // the cut tile is placed behind whole-tile offsets: one tile of slack
  return static_cast<size_t>(max_pairs) + static_cast<size_t>(kBinTile) * kMaxCoarseBins;
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

// ---- 2-D difference array over the coarse-bin grid: +1 on every bin of a box = 4 corner updates -----------------------
__device__ __forceinline__ void diff_add(uint32_t* diff, uint32_t W /* cbins_x + 1 */, uint32_t rc) {
  const uint32_t bx0 = rc & 255u, by0 = (rc >> 8) & 255u, bw = (rc >> 16) & 255u, bh = rc >> 24;
  atomicAdd(&diff[by0 * W + bx0], 1u);
  atomicAdd(&diff[by0 * W + bx0 + bw], 0xffffffffu);
  atomicAdd(&diff[(by0 + bh) * W + bx0], 0xffffffffu);
  atomicAdd(&diff[(by0 + bh) * W + bx0 + bw], 1u);
}
// In place: diff -> counts, diff[y * W + x] = number of boxes covering bin (x, y).  Barriers inside (block-uniform).
__device__ __forceinline__ void diff_prefix(uint32_t* diff, uint32_t W, uint32_t rows /* cbins_y */) {
  __syncthreads();
  if (threadIdx.x <= rows) {  // along x
    uint32_t* r = diff + threadIdx.x * W;
    uint32_t run = 0;
    for (uint32_t x = 0; x < W; ++x) {
      run += r[x];
      r[x] = run;
    }
  }
  __syncthreads();
  for (uint32_t x = threadIdx.x; x < W; x += kBinThreads) {  // along y
    uint32_t run = 0;
    for (uint32_t y = 0; y <= rows; ++y) {
      run += diff[y * W + x];
      diff[y * W + x] = run;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kBinThreads)
k_bin_tiles(const FrameParams* __restrict__ fpp, const Control* __restrict__ ctrl,
            const uint32_t* __restrict__ sorted_slots, const uint32_t* __restrict__ bin_rect, uint32_t tile_stride,
            uint32_t* __restrict__ tile_bin, uint32_t* __restrict__ tile_pairs) {
  __shared__ uint32_t s_diff[kDiffMax];
  __shared__ uint32_t s_sum[kBinWarps];
  const uint32_t tid = threadIdx.x;
  const uint32_t V = ctrl->visible_count, t = blockIdx.x;
  if (t * kBinTile >= V) return;
  const uint32_t cbins_x = fpp->cbins_x, ncbins = fpp->ncbins, W = cbins_x + 1, rows = ncbins / cbins_x;
  for (uint32_t i = tid; i < W * (rows + 1); i += kBinThreads) s_diff[i] = 0u;
  uint32_t slot[kBinItems], rect[kBinItems];
  load_tile(t, V, sorted_slots, bin_rect, slot, rect);
  __syncthreads();
  uint32_t c = 0;
#pragma unroll
  for (int it = 0; it < kBinItems; ++it)
    if (rect[it]) {
      c += rect_pairs(rect[it]);
      diff_add(s_diff, W, rect[it]);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((tid & 31u) == 0) s_sum[tid >> 5] = c;
  diff_prefix(s_diff, W, rows);
  if (tid < ncbins) tile_bin[static_cast<size_t>(tid) * tile_stride + t] = s_diff[(tid / cbins_x) * W + tid % cbins_x];
  if (tid == 0) {
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

__device__ __forceinline__ uint32_t tile_parts(unsigned long long pairs) {
  return static_cast<uint32_t>((pairs + kBinQuota - 1) / kBinQuota);
}

// CTA b < nbins: row b of tile_bin -> exclusive scan over the kept tiles, in place; bin_total[b].
// CTA nbins:     tile_item[t] = exclusive prefix of the tiles' work-item counts (+ end), Control fields.
// Every CTA walks tile_pairs once for the capacity cut; the row / item scans ride in the same iterations.
__global__ void __launch_bounds__(1024)
k_bin_scan(Control* __restrict__ ctrl, uint32_t nbins, uint64_t max_pairs, uint32_t tile_stride,
           const uint32_t* __restrict__ tile_pairs, uint32_t* __restrict__ tile_bin, uint32_t* __restrict__ tile_item,
           uint32_t* __restrict__ bin_total) {
  __shared__ unsigned long long s_warp[32];
  __shared__ uint32_t s_cut;
  __shared__ unsigned long long s_before, s_mine_before, s_mine_cut;
  const uint32_t tid = threadIdx.x;
  const bool items_cta = blockIdx.x == nbins;
  const uint32_t V = ctrl->visible_count;
  const uint32_t ntiles = (V + kBinTile - 1) / kBinTile;
  uint32_t* row = items_cta ? tile_item : tile_bin + static_cast<size_t>(blockIdx.x) * tile_stride;
  if (tid == 0) {
    s_cut = ntiles;
    s_before = 0ull;
    s_mine_before = 0ull;
  }
  __syncthreads();
  unsigned long long carry = 0, mine_carry = 0;
  for (uint32_t t0 = 0; t0 < ntiles; t0 += 1024) {
    const uint32_t t = t0 + tid;
    const unsigned long long v = t < ntiles ? tile_pairs[t] : 0u;
    const unsigned long long mine = items_cta ? tile_parts(v) : (t < ntiles ? row[t] : 0u);
    unsigned long long all, mine_all;
    const unsigned long long ex = carry + block_scan_1024(v, s_warp, &all);
    const unsigned long long mine_ex = mine_carry + block_scan_1024(mine, s_warp, &mine_all);
    if (t < ntiles) row[t] = static_cast<uint32_t>(mine_ex);
    // capacity cut = first tile whose inclusive pair count exceeds max_pairs (the prefix is monotone: one thread hits)
    if (t < ntiles && ex <= max_pairs && ex + v > max_pairs) {
      s_cut = t;
      s_before = ex;
      s_mine_before = mine_ex;
      s_mine_cut = mine;
    }
    carry += all;
    mine_carry += mine_all;
  }
  __syncthreads();
  const uint32_t cut = s_cut;
  const bool overflow = cut != ntiles;
  // The list is nearest-first: the cut tile keeps its first `partial` pairs, everything farther is dropped.
  const uint32_t partial = overflow ? static_cast<uint32_t>(max_pairs - s_before) : 0u;
  if (tid != 0) return;
  if (!items_cta) {
    // the cut tile counts in full here: its bin offsets leave room for pairs that are not placed (bin_slots_capacity)
    bin_total[blockIdx.x] = static_cast<uint32_t>(!overflow ? mine_carry : s_mine_before + (partial ? s_mine_cut : 0ull));
    return;
  }
  const uint32_t kept = cut + (partial ? 1u : 0u);
  const unsigned long long items = overflow ? s_mine_before + tile_parts(partial) : mine_carry;
  tile_item[kept] = static_cast<uint32_t>(items);
  ctrl->tile_cut = kept;
  ctrl->partial_pairs = partial;
  ctrl->bin_items = static_cast<uint32_t>(items);
  ctrl->pair_count = static_cast<uint32_t>(overflow ? max_pairs : carry);
  ctrl->pair_overflow = overflow ? 1u : 0u;
}

// Shared memory of one work item: the current tile, compacted to the ranks that own at least one pair.
struct BinShared {
  uint32_t off[kBinTile + 1];   // exclusive pair offsets; off[m] = the tile's pair count
  uint32_t rect[kBinTile];      // bx0 | by0 << 8 | bw << 16 | bh << 24
  uint32_t slot[kBinTile];
  uint32_t wsum[kBinWarps];
  uint32_t m;                   // ranks kept
};

// Loads tile t into shared memory: ordered compaction of its non-empty ranks + exclusive scan of their pair counts (one
// packed scan: kept ranks << 20 | pairs; a tile has < 2^18 pairs).  Ends with a barrier.
__device__ __forceinline__ void stage_tile(BinShared& sh, uint32_t t, uint32_t V, const uint32_t* __restrict__ sorted_slots,
                                           const uint32_t* __restrict__ bin_rect) {
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  uint32_t slot[kBinItems], rect[kBinItems], cnt[kBinItems];
  load_tile(t, V, sorted_slots, bin_rect, slot, rect);
  uint32_t sum = 0;
#pragma unroll
  for (int it = 0; it < kBinItems; ++it) {
    cnt[it] = rect_pairs(rect[it]);
    sum += cnt[it] + (cnt[it] ? (1u << 20) : 0u);
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
  for (int it = 0; it < kBinItems; ++it)
    if (cnt[it]) {
      const uint32_t i = run >> 20;
      sh.off[i] = run & 0xfffffu;
      sh.rect[i] = rect[it];
      sh.slot[i] = slot[it];
      run += cnt[it] + (1u << 20);
    }
  if (tid == kBinThreads - 1) {
    sh.off[run >> 20] = run & 0xfffffu;
    sh.m = run >> 20;
  }
  __syncthreads();
}

// Index of the kept rank that owns pair e of the staged tile: last s in [0, m) with off[s] <= e.
__device__ __forceinline__ uint32_t rank_of_pair(const BinShared& sh, uint32_t e) {
  uint32_t lo = 0, hi = sh.m;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (sh.off[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

// k-th bin of a box, row-major.  k / bw by a float reciprocal: exact, since frac((k + 0.5) / bw) is at least 0.5 / 255
// away from an integer and the quotient is below 256 (error < 1e-4).
__device__ __forceinline__ uint32_t bin_of(uint32_t rc, uint32_t k, uint32_t cbins_x) {
  const uint32_t bx0 = rc & 255u, by0 = (rc >> 8) & 255u, bw = (rc >> 16) & 255u;
  const uint32_t q = __float2uint_rz(__fdividef(__uint2float_rn(k) + 0.5f, __uint2float_rn(bw)));
  return (by0 + q) * cbins_x + bx0 + (k - q * bw);
}

// One warp step over the 32 pairs [e0, e0 + 32) of the staged tile.  `ra` (warp-uniform) is the kept rank that owns
// pair e0 and is advanced to the owner of e0 + 32.  Every kept rank owns at least one pair, so the window touches at
// most 32 rank boundaries: each lane fetches one end offset, the boundaries become a bit mask, and a lane's rank is
// ra + the number of boundaries at or below its position.
__device__ __forceinline__ uint32_t step_ranks(const BinShared& sh, uint32_t e0, uint32_t* ra) {
  const uint32_t lane = threadIdx.x & 31u, m = sh.m;
  const bool real = *ra + 1u + lane <= m;
  const uint32_t end = sh.off[real ? *ra + 1u + lane : m];  // ascending over the lanes, > e0
  const uint32_t d = end - e0;
  const uint32_t mask = __reduce_or_sync(0xffffffffu, (real && d < 32u) ? (1u << d) : 0u);  // bit 0 is never set
  const uint32_t rank = *ra + __popc(mask & (0xffffffffu >> (31u - lane)));
  *ra += __popc(__ballot_sync(0xffffffffu, real && d <= 32u));
  return rank;
}

// Last tile t in [0, cut) with tile_item[t] <= a.  256-ary search with block-wide vote counts; all threads return it.
__device__ __forceinline__ uint32_t find_tile(const uint32_t* __restrict__ tile_item, uint32_t cut, uint32_t a) {
  uint32_t lo = 0, hi = cut;
  while (hi - lo > 1) {
    const uint32_t step = (hi - lo + kBinThreads - 1) / kBinThreads;
    const uint32_t pos = lo + threadIdx.x * step;
    const int below = __syncthreads_count(pos < hi && __ldg(tile_item + pos) <= a);  // monotone: a prefix of the threads
    lo = lo + (below - 1) * step;
    hi = min(lo + step, hi);
  }
  return lo;
}

__global__ void __launch_bounds__(kBinThreads)
k_bin_place(const FrameParams* __restrict__ fpp, const Control* __restrict__ ctrl,
            const uint32_t* __restrict__ sorted_slots, const uint32_t* __restrict__ bin_rect,
            const uint32_t* __restrict__ tile_item, uint32_t tile_stride, const uint32_t* __restrict__ tile_bin,
            const uint32_t* __restrict__ bin_total, uint32_t* __restrict__ ranges /* uint2[bins] as words */,
            uint32_t* __restrict__ bin_slots) {
  __shared__ BinShared sh;
  __shared__ uint32_t s_wcnt[kBinWarps][kMaxCoarseBins];  // pairs per (warp, bin) of the item, then each warp's cursor
  __shared__ uint32_t s_diff[kDiffMax];
  __shared__ uint32_t s_scan[kBinWarps];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t V = ctrl->visible_count, nitems = ctrl->bin_items, cut = ctrl->tile_cut;
  const uint32_t partial = ctrl->pair_overflow ? ctrl->partial_pairs : 0u;
  const uint32_t cbins_x = fpp->cbins_x, ncbins = fpp->ncbins, W = cbins_x + 1, rows = ncbins / cbins_x;
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
    if (blockIdx.x == 0) ranges[2 * tid] = begin;  // .y grows with the cursors (atomicMax below); zeroed per frame
  }
  uint32_t* wcnt = s_wcnt[warp];
  for (uint32_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const uint32_t t = find_tile(tile_item, cut, item);
    const uint32_t part = item - __ldg(tile_item + t), nparts = __ldg(tile_item + t + 1) - __ldg(tile_item + t);
    __syncthreads();  // every warp is done with the previous item's cursors
#pragma unroll
    for (int w = 0; w < kBinWarps; ++w) s_wcnt[w][tid] = 0u;
    stage_tile(sh, t, V, sorted_slots, bin_rect);  // barriers inside order the zeroing too
    uint32_t pairs = sh.off[sh.m];
    if (partial && t == cut - 1) pairs = min(pairs, partial);  // the capacity cut: nearest `partial` pairs of this tile
    // part `part` of `nparts` equal slices [lo, hi) of the tile's pairs
    const uint32_t per = ((pairs + nparts - 1) / nparts + 255u) & ~255u;
    const uint32_t lo = min(part * per, pairs), hi = min(lo + per, pairs);
    // ---- every bin's cursor = list begin + pairs of earlier tiles + this tile's pairs before lo
    uint32_t cursor = tid < ncbins ? begin + tile_bin[static_cast<size_t>(tid) * tile_stride + t] : 0u;
    if (lo > 0) {  // block-uniform: only parts 1.. of a split tile
      for (uint32_t i = tid; i < W * (rows + 1); i += kBinThreads) s_diff[i] = 0u;
      __syncthreads();
      const uint32_t r_lo = rank_of_pair(sh, lo);  // ranks [0, r_lo) lie wholly before the slice
      for (uint32_t i = tid; i < r_lo; i += kBinThreads) diff_add(s_diff, W, sh.rect[i]);
      diff_prefix(s_diff, W, rows);
      if (tid < ncbins) cursor += s_diff[(tid / cbins_x) * W + tid % cbins_x];
      __syncthreads();
      // ... and the first lo - off[r_lo] pairs of rank r_lo itself (distinct bins)
      s_diff[tid] = 0u;
      __syncthreads();
      const uint32_t rc = sh.rect[r_lo];
      for (uint32_t k = tid; k < lo - sh.off[r_lo]; k += kBinThreads) s_diff[bin_of(rc, k, cbins_x)] = 1u;
      __syncthreads();
      cursor += s_diff[tid];
    }
    // the slice's pairs in 8 contiguous warp shares, each a multiple of 32
    const uint32_t share = ((hi - lo + kBinWarps * 32 - 1) / (kBinWarps * 32)) * 32;
    const uint32_t w_lo = min(lo + warp * share, hi), w_hi = min(w_lo + share, hi);
    const uint32_t ra0 = w_lo < w_hi ? rank_of_pair(sh, w_lo) : 0u;
    // ---- pass A: pairs per (warp, bin)
    uint32_t ra = ra0;
    for (uint32_t e0 = w_lo; e0 < w_hi; e0 += 32) {
      const uint32_t rank = step_ranks(sh, e0, &ra);
      const uint32_t e = e0 + lane;
      if (e < w_hi) atomicAdd(&wcnt[bin_of(sh.rect[rank], e - sh.off[rank], cbins_x)], 1u);
    }
    __syncthreads();
    // ---- thread b: bin b across the warps -> every warp's first position in the bin's list
#pragma unroll
    for (int w = 0; w < kBinWarps; ++w) {
      const uint32_t c = s_wcnt[w][tid];
      s_wcnt[w][tid] = cursor;
      cursor += c;
    }
    // the farthest item that touched a bin leaves its end; a bin nobody reached keeps .y = 0 <= .x: empty
    if (tid < ncbins && hi > lo) atomicMax(&ranges[2 * tid + 1], cursor);
    __syncthreads();
    // ---- pass B: the warp walks its share in order, 32 pairs at a time; lanes with the same bin are ranked by lane
    //      (= pair order) and the bin's cursor advances by the group size: stable
    ra = ra0;
    for (uint32_t e0 = w_lo; e0 < w_hi; e0 += 32) {
      const uint32_t rank = step_ranks(sh, e0, &ra);
      const uint32_t e = e0 + lane;
      const bool valid = e < w_hi;
      const uint32_t bin = valid ? bin_of(sh.rect[rank], e - sh.off[rank], cbins_x) : 0xffffffffu;
      // lanes with the same bin, from 8 ballots (bins < 256): match.any costs a round per distinct value in the warp,
      // and 32 consecutive pairs of depth-ordered splats land in ~20 different bins
      uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const bool bit = (bin >> b) & 1u;
        const uint32_t v = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? v : ~v;
      }
      const uint32_t leader = __ffs(peers) - 1;
      uint32_t prev = 0;
      if (valid && lane == leader) {
        prev = wcnt[bin];
        wcnt[bin] = prev + __popc(peers);
      }
      prev = __shfl_sync(0xffffffffu, prev, leader);
      if (valid) bin_slots[prev + __popc(peers & ((1u << lane) - 1u))] = sh.slot[rank];
      __syncwarp();
    }
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
  const uint32_t items = max_items < static_cast<uint32_t>(sm_count()) * 8u ? max_items : static_cast<uint32_t>(sm_count()) * 8u;  // persistent: CTAs stride over the items
  k_bin_tiles<<<tiles, kBinThreads, 0, stream>>>(d_fp, d_ctrl, d_sorted_slots, d_bin_rect, w.tile_stride, w.tile_bin,
                                                 w.tile_pairs);
  k_bin_scan<<<ncbins + 1, 1024, 0, stream>>>(d_ctrl, ncbins, max_pairs, w.tile_stride, w.tile_pairs, w.tile_bin,
                                              w.tile_item, w.bin_total);
  k_bin_place<<<items, kBinThreads, 0, stream>>>(d_fp, d_ctrl, d_sorted_slots, d_bin_rect, w.tile_item, w.tile_stride,
                                                 w.tile_bin, w.bin_total, reinterpret_cast<uint32_t*>(d_ranges),
                                                 d_bin_slots);
}

// Splat centres per image row of the last frame (from the raster records): what a screen-band partition balances on.
__global__ void __launch_bounds__(256)
k_row_histogram(const Control* __restrict__ ctrl, const float4* __restrict__ rrec, uint32_t height,
                uint32_t* __restrict__ hist) {
  const uint32_t V = ctrl->visible_count;
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < V; i += gridDim.x * 256) {
    const float cpy = rrec[i * 3 + 1].y;  // pixel-frame centre row (project.cu: raster_record)
    if (cpy == cpy) {
      const float r = fminf(fmaxf(rintf(cpy), 0.f), static_cast<float>(height - 1));
      atomicAdd(&hist[static_cast<uint32_t>(r)], 1u);
    }
  }
}

void launch_row_histogram(const Control* d_ctrl, const float* d_rrec, uint32_t max_visible, uint32_t height,
                          uint32_t* d_hist, cudaStream_t stream) {
  cudaMemsetAsync(d_hist, 0, height * sizeof(uint32_t), stream);
  uint32_t want = (max_visible + 255) / 256;
  int blocks = static_cast<int>(want < static_cast<uint32_t>(sm_count()) * 8 ? (want ? want : 1) : static_cast<uint32_t>(sm_count()) * 8);
  k_row_histogram<<<blocks, 256, 0, stream>>>(d_ctrl, reinterpret_cast<const float4*>(d_rrec), height, d_hist);
}

void launch_gather_sorted(const Control* d_ctrl, const uint32_t* d_sorted_slots, const uint32_t* d_vis_id,
                          const float* d_inst, uint32_t max_visible, uint32_t* d_ids_out, float* d_inst_out,
                          cudaStream_t stream) {
  uint32_t want = (max_visible + 255) / 256;
  int blocks = static_cast<int>(want < static_cast<uint32_t>(sm_count()) * 8 ? (want ? want : 1) : static_cast<uint32_t>(sm_count()) * 8);
  k_gather_sorted<<<blocks, 256, 0, stream>>>(d_ctrl, d_sorted_slots, d_vis_id, reinterpret_cast<const float4*>(d_inst),
                                              d_ids_out, reinterpret_cast<float4*>(d_inst_out));
}

}  // namespace vkgsb
