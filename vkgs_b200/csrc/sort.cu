// Stage 2: onesweep LSD radix sort of 32-bit keys + 32-bit values, element count read on the device.
//
// Replaces third_party/vulkan_radix_sort (vrdxCmdSortKeyValueIndirect, src/vk_radix_sort.cc:249-416 with
// upsweep/spine/downsweep.slang): same contract - ascending, STABLE, 8-bit digits, count taken from a device buffer,
// surplus workgroups exit - but one-sweep instead of reduce-then-scan: a single histogram kernel for all digits
// (4 B/key) and one chained-scan scatter kernel per digit (16 B/pair), 68 B/pair over 4 passes against the
// reference's 80.  Partition prefixes travel through a decoupled look-back array instead of the
// spine kernel.
#include "common.cuh"
#include "kernels.h"

namespace vkgsb {

constexpr int kSortThreads = 256;  // == radix: thread d owns digit d in the scan / look-back steps
constexpr int kSortItems = 16;
constexpr int kSortPart = kSortThreads * kSortItems;  // 4096 pairs per partition (PARTITION_SIZE, constants.slang:5)
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kLookBatch = 8;
constexpr uint32_t kFlagAggregate = 1u << 30, kFlagInclusive = 2u << 30, kFlagMask = 3u << 30, kValueMask = ~kFlagMask;

__host__ __device__ inline uint32_t parts_of(uint32_t n) { return (n + kSortPart - 1) / kSortPart; }
uint32_t sort_max_parts(uint32_t max_n) { return parts_of(max_n); }
size_t sort_lookback_bytes(uint32_t max_n, int npass) {
  return static_cast<size_t>(npass) * sort_max_parts(max_n) * 256 * sizeof(uint32_t);
}

// All digit histograms in one pass over the keys; also clears the look-back words the scatter passes will use.
__global__ void __launch_bounds__(256) k_sort_hist(SortArgs a) {
  __shared__ uint32_t s_hist[4 * 256];
  const uint32_t n = min(*a.d_count, a.max_n);
  const uint32_t tid = threadIdx.x;
  for (int i = tid; i < a.npass * 256; i += 256) s_hist[i] = 0;
  __syncthreads();

  const uint32_t nparts = (n + kSortPart - 1) / kSortPart, max_parts = parts_of(a.max_n);
  for (int p = 0; p < a.npass; ++p) {
    uint32_t* lb = a.lookback + static_cast<size_t>(p) * max_parts * 256;
    for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + tid; i < static_cast<size_t>(nparts) * 256;
         i += static_cast<size_t>(gridDim.x) * 256)
      lb[i] = 0u;
  }

  // Depth keys and tile ids are heavily clustered in their upper digits (a handful of exponent values; runs of
  // neighbouring tiles), so lanes are aggregated with match.any first: one shared-memory atomic per distinct digit
  // per warp instead of a 32-way same-address conflict.
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
#pragma unroll
    for (int j = 0; j < 4; ++j)
      for (int p = 0; p < a.npass; ++p) {
        const uint32_t d = (ks[j] >> (a.begin_bit + 8 * p)) & 255u;
        const uint32_t peers = __match_any_sync(active, d);
        if (lane == __ffs(peers) - 1) atomicAdd(&s_hist[p * 256 + d], __popc(peers));
      }
  }
  if (blockIdx.x == 0 && tid < (n & 3u)) {
    uint32_t k = a.keys[n4 * 4 + tid];
    for (int p = 0; p < a.npass; ++p) atomicAdd(&s_hist[p * 256 + ((k >> (a.begin_bit + 8 * p)) & 255u)], 1u);
  }
  __syncthreads();
  for (int i = tid; i < a.npass * 256; i += 256) {
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

// One digit pass: rank within the partition (warp match + per-warp histograms), chained scan across partitions,
// shared-memory reorder, coalesced scatter.  Stable: order inside a partition is (warp, item, lane) == input order.
__global__ void __launch_bounds__(kSortThreads) k_sort_onesweep(SortArgs a, int pass) {
  __shared__ uint32_t s_whist[kSortWarps * 256];
  __shared__ uint32_t s_keys[kSortPart];
  __shared__ uint32_t s_vals[kSortPart];
  __shared__ uint32_t s_gbase[256];  // global index of local sorted position 0 of digit d, minus its local base
  __shared__ uint32_t s_lbase[256];
  __shared__ uint32_t s_scan[kSortWarps];
  __shared__ uint32_t s_part;

  const uint32_t n = min(*a.d_count, a.max_n);
  const uint32_t nparts = (n + kSortPart - 1) / kSortPart, max_parts = parts_of(a.max_n);
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const int shift = a.begin_bit + 8 * pass;
  const uint32_t* __restrict__ src_k = (pass & 1) ? a.keys_alt : a.keys;
  const uint32_t* __restrict__ src_v = (pass & 1) ? a.vals_alt : a.vals;
  uint32_t* __restrict__ dst_k = (pass & 1) ? a.keys : a.keys_alt;
  uint32_t* __restrict__ dst_v = (pass & 1) ? a.vals : a.vals_alt;
  uint32_t* __restrict__ lookback = a.lookback + static_cast<size_t>(pass) * max_parts * 256;
  const bool store_keys = !(a.values_only && pass == a.npass - 1);

  // exclusive scan of this pass's global histogram (every block recomputes it: 256 words)
  uint32_t gexcl;
  {
    uint32_t c = a.hist[pass * 256 + tid], v = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= static_cast<uint32_t>(o)) v += t;
    }
    if (lane == 31) s_scan[warp] = v;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t w = 0; w < warp; ++w) wbase += s_scan[w];
    gexcl = wbase + v - c;
    __syncthreads();
  }

  while (true) {
    if (tid == 0) s_part = atomicAdd(&a.tickets[pass], 1u);
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) s_whist[w * 256 + tid] = 0u;
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
    uint32_t* wh = s_whist + warp * 256;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t d = (key[i] >> shift) & 255u;
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
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

    // ---- thread d: digit d across warps -> per-warp exclusive bases, block total
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      uint32_t c = s_whist[w * 256 + tid];
      s_whist[w * 256 + tid] = total;
      total += c;
    }
    // ---- chained scan over partitions (decoupled look-back), one word per (partition, digit)
    uint32_t excl = 0;
    {
      uint32_t* mine = lookback + static_cast<size_t>(part) * 256 + tid;
      if (part == 0) {
        st_relaxed(mine, kFlagInclusive | total);
      } else {
        st_relaxed(mine, kFlagAggregate | total);
        // Walk back kLookBatch predecessors per round trip: when the whole input is one wave of partitions (a 2 M-key
        // depth sort is 500 of them) every aggregate appears at about the same time and a one-at-a-time walk would
        // serialise hundreds of L2 latencies.
        int64_t q = static_cast<int64_t>(part) - 1;
        bool fin = false;
        while (!fin) {
          uint32_t w[kLookBatch];
#pragma unroll
          for (int j = 0; j < kLookBatch; ++j)
            w[j] = (q - j >= 0) ? ld_relaxed(lookback + static_cast<size_t>(q - j) * 256 + tid) : kFlagInclusive;
          int used = 0;
#pragma unroll
          for (int j = 0; j < kLookBatch; ++j) {
            const uint32_t f = w[j] & kFlagMask;
            if (!fin && used == j && f != 0u) {
              excl += w[j] & kValueMask;
              used = j + 1;
              fin = f == kFlagInclusive;
            }
          }
          q -= used;
        }
        st_relaxed(mine, kFlagInclusive | (excl + total));
      }
    }
    // ---- block-local exclusive scan of totals over digits
    uint32_t lexcl;
    {
      uint32_t v = total;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= static_cast<uint32_t>(o)) v += t;
      }
      if (lane == 31) s_scan[warp] = v;
      __syncthreads();
      uint32_t wb = 0;
      for (uint32_t w = 0; w < warp; ++w) wb += s_scan[w];
      lexcl = wb + v - total;
    }
    s_lbase[tid] = lexcl;
    s_gbase[tid] = gexcl + excl - lexcl;
    __syncthreads();

    // ---- reorder through shared memory
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t d = (key[i] >> shift) & 255u;
      pos[i] += s_lbase[d] + wh[d];
      s_keys[pos[i]] = key[i];
    }
    // values: issue the loads now so they overlap the key scatter
    uint32_t val[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t li = wbase + i * 32 + lane;
      val[i] = (li < valid) ? __ldg(src_v + pbase + li) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t j = i * kSortThreads + tid;
      if (j < valid) {
        const uint32_t k = s_keys[j];
        if (store_keys) dst_k[s_gbase[(k >> shift) & 255u] + j] = k;
      }
    }
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) s_vals[pos[i]] = val[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
      const uint32_t j = i * kSortThreads + tid;
      if (j < valid) dst_v[s_gbase[(s_keys[j] >> shift) & 255u] + j] = s_vals[j];
    }
    __syncthreads();  // s_part / s_whist / s_keys are rewritten by the next iteration
  }
}

void launch_sort(const SortArgs& a, cudaStream_t stream) {
  if (a.max_n == 0) return;
  const uint32_t parts = sort_max_parts(a.max_n);
  int hist_blocks = static_cast<int>(min(static_cast<uint32_t>(148 * 8), (a.max_n + 4095u) / 4096u));
  if (!a.have_hist) k_sort_hist<<<hist_blocks, 256, 0, stream>>>(a);
  int blocks = static_cast<int>(min(parts, static_cast<uint32_t>(148 * 4)));
  for (int p = 0; p < a.npass; ++p) k_sort_onesweep<<<blocks, kSortThreads, 0, stream>>>(a, p);
}

}  // namespace vkgsb
