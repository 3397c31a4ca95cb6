// Stage 3a: tile binning.  No counterpart in the reference, whose instanced quads go through the hardware
// rasteriser in sorted order (vkCmdDrawIndexedIndirect, engine.cc:1472-1480); here the globally sorted splat list
// is turned into per-tile lists that keep that order:
//   k_make_pairs   one thread per sorted rank, walked NEAREST FIRST: bounding box of the +-3 sigma quad
//                  (splat.vert:19-25) clipped to the viewport / band -> tile rectangle -> (tile, slot) pairs written
//                  at offsets from an ordered (decoupled look-back) scan, so the pair list is rank-major.
//   stable 2-pass onesweep sort of the pairs by tile id (sort.cu) -> each tile's pairs are contiguous and still
//                  nearest-first.
//   k_tile_ranges  [begin,end) of every tile in the sorted pair list.
#include "common.cuh"
#include "kernels.h"

namespace vkgsb {

constexpr int kPairThreads = 256;

uint32_t pairs_num_blocks(uint32_t max_visible) { return (max_visible + kPairThreads - 1) / kPairThreads; }

struct TileRect {
  uint32_t x0, y0, w, h;  // in tiles; w*h == 0 -> nothing
};

// Conservative pixel bounding box of the quad centre +- RS*(+-3,+-3) in the pixel frame where pixel i has its
// centre at coordinate i (cpx = (ndc.x+1)*W/2 - 1/2).  The exact coverage test is per pixel in the blend stage.
__device__ __forceinline__ TileRect splat_tile_rect(const FrameParams& fp, const float4 r0, const float4 r1) {
  TileRect t{0, 0, 0, 0};
  if (!(r0.z < 1.f)) return t;  // depth LESS against the cleared 1.0 (graphics_pipeline.cc:79-81)
  const float hw = 0.5f * static_cast<float>(fp.width), hh = 0.5f * static_cast<float>(fp.height);
  const float cpx = fmaf(r0.x, hw, hw - 0.5f), cpy = fmaf(r0.y, hh, hh - 0.5f);
  const float ex = 3.f * (fabsf(r1.x * hw) + fabsf(r1.z * hw));  // |m00| + |m01|
  const float ey = 3.f * (fabsf(r1.y * hh) + fabsf(r1.w * hh));  // |m10| + |m11|
  // NaN lanes (D == 0, negative eigenvalue: SURVEY.md §7 hard part 6) fail every comparison below and vanish.
  const float fx0 = ceilf(cpx - ex - 0.01f), fx1 = floorf(cpx + ex + 0.01f);
  const float fy0 = ceilf(cpy - ey - 0.01f), fy1 = floorf(cpy + ey + 0.01f);
  const float bx0 = 0.f, bx1 = static_cast<float>(fp.width) - 1.f;
  const float by0 = static_cast<float>(fp.band_y0), by1 = static_cast<float>(fp.band_y1) - 1.f;
  if (!(fx0 <= fx1 && fy0 <= fy1 && fx1 >= bx0 && fx0 <= bx1 && fy1 >= by0 && fy0 <= by1)) return t;
  const uint32_t x0 = static_cast<uint32_t>(fmaxf(fx0, bx0)), x1 = static_cast<uint32_t>(fminf(fx1, bx1));
  const uint32_t y0 = static_cast<uint32_t>(fmaxf(fy0, by0)), y1 = static_cast<uint32_t>(fminf(fy1, by1));
  t.x0 = x0 / kTile;
  t.y0 = y0 / kTile;
  t.w = x1 / kTile - t.x0 + 1;
  t.h = y1 / kTile - t.y0 + 1;
  return t;
}

__global__ void __launch_bounds__(kPairThreads)
k_make_pairs(const FrameParams* __restrict__ fpp, Control* __restrict__ ctrl, unsigned long long* __restrict__ scan_desc,
             const uint32_t* __restrict__ sorted_slots, const float4* __restrict__ inst, uint64_t max_pairs,
             uint32_t* __restrict__ pair_tile, uint32_t* __restrict__ pair_slot) {
  __shared__ FrameParams fp;
  __shared__ uint32_t s_off[kPairThreads + 1];  // exclusive offsets of this block's splats
  __shared__ uint32_t s_slot[kPairThreads];
  __shared__ uint32_t s_rect[kPairThreads];     // x0 | y0 << 10 | w << 20  (w <= 241 tiles at 3840 px)
  __shared__ uint32_t s_wsum[kPairThreads / 32];
  __shared__ uint32_t s_ticket, s_base;

  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t V = ctrl->visible_count;
  const uint32_t nblocks = (V + kPairThreads - 1) / kPairThreads;
  if (tid == 0) s_ticket = atomicAdd(&ctrl->pairs_ticket, 1u);
  for (uint32_t i = tid; i < sizeof(FrameParams) / 4; i += kPairThreads)
    reinterpret_cast<uint32_t*>(&fp)[i] = reinterpret_cast<const uint32_t*>(fpp)[i];
  __syncthreads();
  const uint32_t ticket = s_ticket;
  if (ticket >= nblocks) return;

  const uint32_t i = ticket * kPairThreads + tid;  // i-th nearest splat
  uint32_t count = 0, slot = 0;
  TileRect rc{0, 0, 0, 0};
  if (i < V) {
    slot = sorted_slots[V - 1 - i];  // ascending key = far -> near (rank.comp:39), so walk it backwards
    const float4 r0 = __ldg(inst + slot * 3 + 0), r1 = __ldg(inst + slot * 3 + 1);
    rc = splat_tile_rect(fp, r0, r1);
    count = rc.w * rc.h;
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
  s_slot[tid] = slot;
  s_rect[tid] = rc.x0 | (rc.y0 << 10) | (rc.w << 20);
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
    uint32_t lo = 0, hi = kPairThreads;  // find last s with s_off[s] <= e
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const uint32_t mid = (lo + hi) >> 1;
      if (s_off[mid] <= e) lo = mid; else hi = mid;
    }
    const uint32_t k = e - s_off[lo], rect = s_rect[lo];
    const uint32_t x0 = rect & 1023u, y0 = (rect >> 10) & 1023u, w = rect >> 20;
    const uint32_t ty = y0 + k / w, tx = x0 + k % w;
    const uint64_t g = base + e;
    if (g < max_pairs) {  // overflow drops the farthest pairs (the list is nearest-first)
      pair_tile[g] = (ty - fp.tile_y0) * fp.tiles_x + tx;
      pair_slot[g] = s_slot[lo];
    }
  }
}

// Tile boundaries in the tile-sorted pair list.  ranges must be zero on entry (empty tiles stay [0,0)).
__global__ void __launch_bounds__(256)
k_tile_ranges(const Control* __restrict__ ctrl, const uint32_t* __restrict__ tile_sorted, uint2* __restrict__ ranges) {
  const uint32_t D = ctrl->pair_count;
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < D; i += gridDim.x * 256) {
    const uint32_t t = tile_sorted[i];
    if (i == 0 || tile_sorted[i - 1] != t) ranges[t].x = i;
    if (i == D - 1 || tile_sorted[i + 1] != t) ranges[t].y = i + 1;
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
                       uint32_t* d_pair_tile, uint32_t* d_pair_slot, cudaStream_t stream) {
  uint32_t nb = pairs_num_blocks(max_visible);
  if (nb == 0) return;
  k_make_pairs<<<nb, kPairThreads, 0, stream>>>(d_fp, d_ctrl, d_scan_desc, d_sorted_slots,
                                                reinterpret_cast<const float4*>(d_inst), max_pairs, d_pair_tile,
                                                d_pair_slot);
}

void launch_tile_ranges(const Control* d_ctrl, const uint32_t* d_pair_tile_sorted, uint64_t max_pairs, uint2* d_ranges,
                        cudaStream_t stream) {
  uint64_t want = (max_pairs + 255) / 256;
  int blocks = static_cast<int>(want < 148 * 16 ? (want ? want : 1) : 148 * 16);
  k_tile_ranges<<<blocks, 256, 0, stream>>>(d_ctrl, d_pair_tile_sorted, d_ranges);
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
