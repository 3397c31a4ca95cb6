// Stage 2: onesweep LSD radix sort of 32-bit keys + 32-bit values, element count read on the device.
//
// Replaces third_party/vulkan_radix_sort (vrdxCmdSortKeyValueIndirect, src/vk_radix_sort.cc:249-416 with
// upsweep/spine/downsweep.slang): same contract - ascending, STABLE, 8-bit digits, count taken from a device buffer,
// surplus workgroups exit - but one-sweep instead of reduce-then-scan: a single histogram kernel for all digits
// (4 B/key) and one scatter kernel per digit (16 B/pair), 68 B/pair over 4 passes against the reference's 80
// (SURVEY.md §8a-5).  Partition prefixes come from an 8-ary tree of partition aggregates (a depth sort is ONE wave of
// partitions that all post at once: no chained look-back), not from a spine kernel.
#include "common.cuh"
#include "kernels.h"

namespace vkgsb {

constexpr int kSortThreads = 256;  // thread t owns the digits [t * DPT, (t + 1) * DPT) in the scan / aggregate-tree steps
constexpr int kSortItems = 16;
constexpr int kSortPart = kSortThreads * kSortItems;  // 4096 pairs per partition (PARTITION_SIZE, constants.slang:5)
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kMaxBins = 1024;  // digit bins of all passes together: 4 x 256 (the reference's layout) or 256 + 256 + 512
constexpr uint32_t kFlagAggregate = 1u << 30, kFlagMask = 3u << 30, kValueMask = ~kFlagMask;

__host__ __device__ inline uint32_t parts_of(uint32_t n) { return (n + kSortPart - 1) / kSortPart; }
uint32_t sort_max_parts(uint32_t max_n) { return parts_of(max_n); }

// The partitions' digit counts form an 8-ary tree (k_sort_onesweep): level 0 = one row per partition, level k = one row
// per aligned group of 8 rows of level k-1, up to a single row.  off[k] = first row of level k.
struct SortLevels {
  uint32_t off[8];
  int K;
  uint32_t total;
};
__host__ __device__ inline SortLevels sort_levels(uint32_t nparts) {
  SortLevels lv;
  lv.K = 0;
  uint32_t rows = nparts ? nparts : 1u, o = 0;
  while (true) {
    lv.off[lv.K++] = o;
    o += rows;
    if (rows <= 1u || lv.K == 8) break;
    rows = (rows + 7u) / 8u;
  }
  lv.total = o;
  return lv;
}
// rows x (all passes' bins) words + per pass one arrival counter per row
size_t sort_lookback_bytes(uint32_t max_n) {
  return static_cast<size_t>(sort_levels(sort_max_parts(max_n)).total) * (kMaxBins + 4) * sizeof(uint32_t);
}

__host__ __device__ inline int pass_bits(const SortArgs& a, int p) { return a.bits[p] ? a.bits[p] : 8; }
// first bit and first histogram bin of pass p
__host__ __device__ inline void pass_layout(const SortArgs& a, int p, int* shift, uint32_t* bin0) {
  int sh = a.begin_bit;
  uint32_t b = 0;
  for (int q = 0; q < p; ++q) {
    sh += pass_bits(a, q);
    b += 1u << pass_bits(a, q);
  }
  *shift = sh;
  *bin0 = b;
}

// All digit histograms in one pass over the keys; also clears the look-back words the scatter passes will use.
__global__ void __launch_bounds__(256) k_sort_hist(SortArgs a) {
  __shared__ uint32_t s_hist[kMaxBins];
  const uint32_t n = min(*a.d_count, a.max_n);
  const uint32_t tid = threadIdx.x;
  for (int i = tid; i < kMaxBins; i += 256) s_hist[i] = 0;
  __syncthreads();

  // the tree of partition aggregates and its arrival counters: laid out for the actual partition count
  const uint32_t nparts = (n + kSortPart - 1) / kSortPart;
  const size_t used = static_cast<size_t>(sort_levels(nparts).total) * (kMaxBins + 4);
  for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + tid; i < used; i += static_cast<size_t>(gridDim.x) * 256)
    a.lookback[i] = 0u;

  // One shared-memory atomic per key and digit - except where a whole warp holds the same digit (the upper digits of
  // depth keys: a handful of exponent values), which would be a 32-way same-address pile-up: one vote finds that case
  // and a single lane adds the warp's count.  (Aggregating every digit with match.any cost a round per distinct value -
  // ~30 for a dense digit - and made this kernel a third of the whole sort on uniform keys.)
  const uint32_t n4 = n / 4;
  const uint4* k4 = reinterpret_cast<const uint4*>(a.keys);
  const uint32_t lane = tid & 31u;
  for (uint32_t i0 = blockIdx.x * 256; i0 < n4; i0 += gridDim.x * 256) {  // i0 is warp-uniform: full warps vote
    const uint32_t i = i0 + tid;
    const bool ok = i < n4;
    uint4 k = ok ? __ldg(k4 + i) : make_uint4(0, 0, 0, 0);
    const uint32_t ks[4] = {k.x, k.y, k.z, k.w};
    const uint32_t active = __ballot_sync(0xffffffffu, ok);
    if (!ok) continue;
    const uint32_t leader = __ffs(active) - 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int sh = a.begin_bit;
      uint32_t b0 = 0;
      for (int p = 0; p < a.npass; ++p) {
        const int bits = pass_bits(a, p);
        const uint32_t d = (ks[j] >> sh) & ((1u << bits) - 1u);
        const uint32_t d0 = __shfl_sync(active, d, leader);
        if (__all_sync(active, d == d0)) {
          if (lane == leader) atomicAdd(&s_hist[b0 + d], __popc(active));
        } else {
          atomicAdd(&s_hist[b0 + d], 1u);
        }
        sh += bits;
        b0 += 1u << bits;
      }
    }
  }
  if (blockIdx.x == 0 && tid < (n & 3u)) {
    uint32_t k = a.keys[n4 * 4 + tid];
    int sh = a.begin_bit;
    uint32_t b0 = 0;
    for (int p = 0; p < a.npass; ++p) {
      const int bits = pass_bits(a, p);
      atomicAdd(&s_hist[b0 + ((k >> sh) & ((1u << bits) - 1u))], 1u);
      sh += bits;
      b0 += 1u << bits;
    }
  }
  __syncthreads();
  for (int i = tid; i < kMaxBins; i += 256) {
    uint32_t c = s_hist[i];
    if (c) atomicAdd(&a.hist[i], c);
  }
}

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One digit pass of BITS bits: rank within the partition (warp match + per-warp histograms), chained scan across
// partitions, shared-memory reorder, coalesced scatter.  Stable: order inside a partition is (warp, item, lane) == input
// order.  BITS = 8 is the reference's digit; BITS = 9 lets the 25-bit depth keys of the frame path finish in 3 passes.
template <int BITS, bool MATCH>
__global__ void __launch_bounds__(kSortThreads, 4) k_sort_onesweep(SortArgs a, int pass, int shift, uint32_t bin0) {
  constexpr int BINS = 1 << BITS, DPT = BINS / kSortThreads;
  constexpr uint32_t MASK = BINS - 1;
  static_assert(DPT >= 1 && kSortWarps * BINS <= kSortPart, "the value stage reuses the per-warp histograms' storage");
  __shared__ uint32_t s_keys[kSortPart];
  __shared__ uint32_t s_vals[kSortPart];  // also the per-warp digit counters [kSortWarps][BINS] until the values land
  __shared__ uint32_t s_gbase[BINS];      // global index of local sorted position 0 of digit d, minus its local base
  __shared__ uint32_t s_lbase[BINS];
  __shared__ uint32_t s_scan[kSortWarps];
  __shared__ uint32_t s_cnt[BINS];  // the partition's digit counts, posted before the (much longer) ranking
  __shared__ uint32_t s_part, s_arrive;
  uint32_t* s_whist = s_vals;

  const uint32_t n = min(*a.d_count, a.max_n);
  const uint32_t nparts = (n + kSortPart - 1) / kSortPart;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t* __restrict__ src_k = (pass & 1) ? a.keys_alt : a.keys;
  const uint32_t* __restrict__ src_v = (pass & 1) ? a.vals_alt : a.vals;
  uint32_t* __restrict__ dst_k = (pass & 1) ? a.keys : a.keys_alt;
  uint32_t* __restrict__ dst_v = (pass & 1) ? a.vals : a.vals_alt;
  // tree of partition aggregates (below): rows of all levels back to back, BINS words each; arrival counters behind
  // the passes' rows
  const SortLevels lv = sort_levels(nparts);
  uint32_t* __restrict__ lookback = a.lookback + static_cast<size_t>(bin0) * lv.total;
  uint32_t* __restrict__ counters = a.lookback + static_cast<size_t>(kMaxBins) * lv.total + static_cast<size_t>(pass) * lv.total;
  const bool store_keys = !(a.values_only && pass == a.npass - 1);

  // block-wide exclusive scan of one value per thread
  auto block_excl = [&](uint32_t v) {
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= static_cast<uint32_t>(o)) incl += t;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t w = 0; w < warp; ++w) wbase += s_scan[w];
    __syncthreads();
    return wbase + incl - v;
  };

  // exclusive scan of this pass's global histogram (every block recomputes it: BINS words)
  uint32_t gexcl[DPT];
  {
    uint32_t c[DPT], sum = 0;
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
      c[j] = a.hist[bin0 + tid * DPT + j];
      sum += c[j];
    }
    uint32_t run = block_excl(sum);
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
      gexcl[j] = run;
      run += c[j];
    }
  }

  while (true) {
    if (tid == 0) s_part = atomicAdd(&a.tickets[pass], 1u);
    for (int i = tid; i < kSortWarps * BINS; i += kSortThreads) s_whist[i] = 0u;
    for (int i = tid; i < BINS; i += kSortThreads) s_cnt[i] = 0u;
    __syncthreads();
    const uint32_t part = s_part;
    if (part >= nparts) return;
    const uint32_t pbase = part * kSortPart;
    const uint32_t valid = min(static_cast<uint32_t>(kSortPart), n - pbase);

    // ---- load (warp-striped: lane-consecutive addresses) and rank
    uint32_t key[kSortItems], pos[kSortItems];
    const uint32_t wbase = warp * (32 * kSortItems);
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t li = wbase + i * 32 + lane;
      key[i] = (li < valid) ? __ldg(src_k + pbase + li) : 0xffffffffu;  // padding ranks after every real key
    }
    // ---- the partition's digit counts, posted BEFORE the ranking: every partition's prefix needs every earlier
    // partition's counts, so in a single wave the pass is as slow as the slowest partition's time-to-post.  Counting is
    // 16 shared-memory atomics per thread; the ranking below (ballots, serial per-warp counters) is ten times that and
    // now overlaps the tree reduction.
#pragma unroll
    for (int i = 0; i < kSortItems; ++i)
      if (wbase + i * 32 + lane < valid) atomicAdd(&s_cnt[(key[i] >> shift) & MASK], 1u);
    __syncthreads();
    // An 8-ary tree of aggregates gives the exclusive prefix over the partitions, per digit.  Level 0 holds one word per
    // (partition, digit); a level-k word is the sum of an aligned group of 8 words of level k-1, published by whichever
    // CTA arrives LAST in that group (an arrival counter per group).  A prefix is then <= 7 earlier siblings per level -
    // words that depend on nobody's prefix.  The usual chained look-back serialises here: a depth sort is ONE wave of
    // ~500 partitions that all post at the same moment, and the first inclusive value crept forward 8 partitions per
    // L2 round trip (30 us of a 37 us pass; with every partition resident 110 us).
    uint32_t total[DPT];
    {
      const uint32_t* off = lv.off;
#pragma unroll
      for (int j = 0; j < DPT; ++j) {
        total[j] = s_cnt[tid * DPT + j];
        st_relaxed(lookback + (static_cast<size_t>(off[0]) + part) * BINS + tid * DPT + j, kFlagAggregate | total[j]);
      }
      // climb: the last arriver of a complete group sums it and arrives, in turn, at the group's parent
      uint32_t idx = part;
      for (int k = 0; k + 1 < lv.K; ++k) {
        __threadfence();
        __syncthreads();
        if (tid == 0) s_arrive = atomicAdd(&counters[off[k + 1] + (idx >> 3)], 1u);
        __syncthreads();
        if (s_arrive != 7u) break;  // CTA-uniform; a partial group at the end of a level never completes and is never read
        __threadfence();
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
          const uint32_t* row = lookback + (static_cast<size_t>(off[k]) + (idx & ~7u)) * BINS + tid * DPT + j;
          uint32_t w[8], sum = 0;
#pragma unroll
          for (int m = 0; m < 8; ++m) w[m] = ld_relaxed(row + static_cast<size_t>(m) * BINS);
#pragma unroll
          for (int m = 0; m < 8; ++m) sum += w[m] & kValueMask;
          st_relaxed(lookback + (static_cast<size_t>(off[k + 1]) + (idx >> 3)) * BINS + tid * DPT + j, kFlagAggregate | sum);
        }
        idx >>= 3;
      }
    }

    // ---- rank
    uint32_t* wh = s_whist + warp * BINS;
    // lanes holding the same digit, from BITS ballots (match.any costs one round per distinct value in the warp -
    // ~30 of them for a dense digit: it took a quarter of the kernel's stall samples); all 16 items first, so the
    // votes pipeline
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t d = (key[i] >> shift) & MASK;
      uint32_t m = 0xffffffffu;
      if (MATCH) {  // the caller's hint: this digit takes a handful of values (a depth key's top bits) - few rounds
        m = __match_any_sync(0xffffffffu, d);
      } else {
#pragma unroll
        for (int b = 0; b < BITS; ++b) {
          const bool bit = (d >> b) & 1u;
          const uint32_t v = __ballot_sync(0xffffffffu, bit);
          m &= bit ? v : ~v;
        }
      }
      pos[i] = m;
    }
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t d = (key[i] >> shift) & MASK;
      const uint32_t peers = pos[i];
      const uint32_t leader = __ffs(peers) - 1;
      uint32_t prev = 0;
      if (lane == leader) {
        prev = wh[d];
        wh[d] = prev + __popc(peers);
      }
      prev = __shfl_sync(0xffffffffu, prev, leader);
      pos[i] = prev + __popc(peers & ((1u << lane) - 1u));
      __syncwarp();
    }
    __syncthreads();

    // ---- this thread's digits across warps -> per-warp exclusive bases
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
      const uint32_t d = tid * DPT + j;
      uint32_t t = 0;
#pragma unroll
      for (int w = 0; w < kSortWarps; ++w) {
        uint32_t c = s_whist[w * BINS + d];
        s_whist[w * BINS + d] = t;
        t += c;
      }
    }
    // ---- descend the tree: earlier siblings of this partition's ancestor at every level (posted long ago by now)
    uint32_t excl[DPT];
#pragma unroll
    for (int j = 0; j < DPT; ++j) excl[j] = 0;
    {
      const uint32_t* off = lv.off;
      // descend: earlier siblings of this partition's ancestor at every level
      for (int k = 0; k < lv.K; ++k) {
        const uint32_t me = part >> (3 * k), nsib = me & 7u;
        if (nsib == 0u) continue;
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
          const uint32_t* row = lookback + (static_cast<size_t>(off[k]) + (me & ~7u)) * BINS + tid * DPT + j;
          uint32_t w[7];
#pragma unroll
          for (int m = 0; m < 7; ++m) w[m] = static_cast<uint32_t>(m) < nsib ? ld_relaxed(row + static_cast<size_t>(m) * BINS) : kFlagAggregate;
#pragma unroll
          for (int m = 0; m < 7; ++m) {
            while ((w[m] & kFlagMask) == 0u) w[m] = ld_relaxed(row + static_cast<size_t>(m) * BINS);  // not posted yet
            excl[j] += w[m] & kValueMask;
          }
        }
      }
    }
    // ---- block-local exclusive scan of totals over digits
    {
      uint32_t sum = 0;
#pragma unroll
      for (int j = 0; j < DPT; ++j) sum += total[j];
      uint32_t run = block_excl(sum);
#pragma unroll
      for (int j = 0; j < DPT; ++j) {
        s_lbase[tid * DPT + j] = run;
        s_gbase[tid * DPT + j] = gexcl[j] + excl[j] - run;
        run += total[j];
      }
    }
    __syncthreads();

    // ---- reorder through shared memory
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t d = (key[i] >> shift) & MASK;
      pos[i] += s_lbase[d] + wh[d];
      s_keys[pos[i]] = key[i];
    }
    // values: issue the loads now so they overlap the key scatter
    uint32_t val[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t li = wbase + i * 32 + lane;
      val[i] = (li < valid) ? ((a.vals_identity && pass == 0) ? pbase + li : __ldg(src_v + pbase + li)) : 0u;
    }
    __syncthreads();  // every wh[] has been read: its storage now takes the values
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t j = i * kSortThreads + tid;
      if (j < valid) {
        const uint32_t k = s_keys[j];
        if (store_keys) dst_k[s_gbase[(k >> shift) & MASK] + j] = k;
      }
    }
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) s_vals[pos[i]] = val[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t j = i * kSortThreads + tid;
      if (j < valid) dst_v[s_gbase[(s_keys[j] >> shift) & MASK] + j] = s_vals[j];
    }
    __syncthreads();  // s_part / s_whist / s_keys are rewritten by the next iteration
  }
}

void launch_sort(const SortArgs& a, cudaStream_t stream) {
  if (a.max_n == 0) return;
  const uint32_t parts = sort_max_parts(a.max_n);
  int hist_blocks = static_cast<int>(min(static_cast<uint32_t>(sm_count() * 8), (a.max_n + 4095u) / 4096u));
  if (!a.have_hist) k_sort_hist<<<hist_blocks, 256, 0, stream>>>(a);
  int blocks = static_cast<int>(min(parts, static_cast<uint32_t>(sm_count() * 4)));
  for (int p = 0; p < a.npass; ++p) {
    int shift;
    uint32_t bin0;
    pass_layout(a, p, &shift, &bin0);
    const bool clustered = (a.clustered_passes >> p) & 1u;
    if (pass_bits(a, p) == 9) {
      if (clustered) k_sort_onesweep<9, true><<<blocks, kSortThreads, 0, stream>>>(a, p, shift, bin0);
      else k_sort_onesweep<9, false><<<blocks, kSortThreads, 0, stream>>>(a, p, shift, bin0);
    } else {
      if (clustered) k_sort_onesweep<8, true><<<blocks, kSortThreads, 0, stream>>>(a, p, shift, bin0);
      else k_sort_onesweep<8, false><<<blocks, kSortThreads, 0, stream>>>(a, p, shift, bin0);
    }
  }
}

}  // namespace vkgsb
