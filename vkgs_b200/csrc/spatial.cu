// Load-time spatial order of the resident scene + the bounding box of every tile of 256 splats.
//
// No counterpart in the reference, which keeps the file's order (splat_load_thread.cc:138-159) and tests every centre
// every frame (rank.comp:27-42).  A trained 3DGS file carries no spatial order, so the ~1/3 of the splats a frame sees
// are scattered over the whole 128-byte-per-splat payload array: every visible splat is an isolated DRAM line and every
// tile of consecutive splats straddles the frustum.  Here the scene is stored ONCE, at load, in Morton order of the
// centres:
//   * a tile of 256 consecutive splats is a small box in space -> k_cull_classify (project.cu) decides most tiles from
//     the box alone (all outside / all inside) and only the tiles cut by the frustum are tested per splat;
//   * the visible splats are long runs of consecutive ids -> k_project streams their payload lines.
// Nothing observable depends on the order except how equal depth keys are ordered, which the reference leaves to a race
// (atomicAdd slots, rank.comp:38): the frame resolves ties by STORED index.  vkgsb_read_scene returns the stored order,
// vkgsb_read_order the file index of every stored splat; VKGSB_OPT_SPATIAL_ORDER = 0 keeps the file's order.
//
// Deterministic: bounds from integer min / max and integer moments (atomics on integers commute), keys from those, a
// stable radix sort -> two renderers loading the same file store the same order.
#include "common.cuh"
#include "kernels.h"

namespace vkgsb {

// order-preserving integer image of a float (finite values)
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ bool finite3(float x, float y, float z) {
  return fabsf(x) <= 3.0e38f && fabsf(y) <= 3.0e38f && fabsf(z) <= 3.0e38f;  // false for NaN
}

__global__ void k_stats_init(SpatialStats* st) {
  if (threadIdx.x < 3) {
    st->min_ord[threadIdx.x] = 0xffffffffu;
    st->max_ord[threadIdx.x] = 0u;
    st->sum[threadIdx.x] = 0ull;
    st->sumsq[threadIdx.x] = 0ull;
  }
  if (threadIdx.x == 0) st->finite = 0ull;
}

// pass 1: bounding box of the finite centres
__global__ void __launch_bounds__(256) k_stats_minmax(SceneStorage sc, uint32_t n, SpatialStats* st) {
  uint32_t mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const float p[3] = {sc.x[i], sc.y[i], sc.z[i]};
    if (!finite3(p[0], p[1], p[2])) continue;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const uint32_t o = f2ord(p[a]);
      mn[a] = min(mn[a], o);
      mx[a] = max(mx[a], o);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = __reduce_min_sync(0xffffffffu, mn[a]);
    mx[a] = __reduce_max_sync(0xffffffffu, mx[a]);
    if ((threadIdx.x & 31u) == 0) {
      atomicMin(&st->min_ord[a], mn[a]);
      atomicMax(&st->max_ord[a], mx[a]);
    }
  }
}

// a centre's coordinate on axis a as a fraction of the bounding box, [0, 1]
__device__ __forceinline__ float box_fraction(const SpatialStats* st, int a, float v) {
  const float lo = ord2f(st->min_ord[a]), hi = ord2f(st->max_ord[a]);
  const float ext = hi - lo;
  return ext > 0.f ? __saturatef((v - lo) / ext) : 0.f;
}

// pass 2: first and second moments of the box fractions in 16-bit fixed point (integer sums: order-independent)
__global__ void __launch_bounds__(256) k_stats_moments(SceneStorage sc, uint32_t n, SpatialStats* st) {
  unsigned long long s[3] = {0, 0, 0}, q[3] = {0, 0, 0}, cnt = 0;
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const float p[3] = {sc.x[i], sc.y[i], sc.z[i]};
    if (!finite3(p[0], p[1], p[2])) continue;
    ++cnt;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const unsigned long long u = static_cast<unsigned long long>(box_fraction(st, a, p[a]) * 65535.f + 0.5f);
      s[a] += u;
      q[a] += u * u;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      s[a] += __shfl_xor_sync(0xffffffffu, s[a], o);
      q[a] += __shfl_xor_sync(0xffffffffu, q[a], o);
    }
  }
  if ((threadIdx.x & 31u) == 0) {
    atomicAdd(&st->finite, cnt);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicAdd(&st->sum[a], s[a]);
      atomicAdd(&st->sumsq[a], q[a]);
    }
  }
}

__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 10 bits -> every third bit
  v &= 1023u;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

// pass 3: the sort key = size class (2 bits) | 30-bit Morton code of the centre quantised to 1024 cells per axis over
// mean +- 3 sigma of the box fractions (a few far outliers must not squeeze the bulk of the scene into a handful of
// cells); outliers clamp to the border cells, non-finite centres go last.  vals = 0, 1, 2, ...
// Size classes: splats whose 3-sigma radius exceeds 1/16 (class 2) or 1/128 (class 1) of that window's extent are stored
// in front of the rest, each class in Morton order.  A large splat reaches screen rows far from its centre: mixed in
// with its small neighbours it would make their tile "cannot be decided from the box" for every screen band
// (classify_tile takes the tile's largest lambda_max) and leave one or two visible splats per tile all over the band's
// frustum.
__global__ void __launch_bounds__(256) k_spatial_keys(SceneStorage sc, uint32_t n, const SpatialStats* st, uint32_t* keys,
                                                      uint32_t* vals) {
  __shared__ float s_lo[3], s_scale[3], s_ext[3];
  if (threadIdx.x < 3) {
    const int a = threadIdx.x;
    const double cnt = static_cast<double>(st->finite ? st->finite : 1ull);
    const double m = static_cast<double>(st->sum[a]) / cnt / 65535.0;
    const double v = static_cast<double>(st->sumsq[a]) / cnt / (65535.0 * 65535.0) - m * m;
    const double sd = sqrt(v > 0.0 ? v : 0.0);
    const double lo = fmax(m - 3.0 * sd, 0.0), hi = fmin(m + 3.0 * sd, 1.0);
    s_lo[a] = static_cast<float>(lo);
    s_scale[a] = hi > lo ? static_cast<float>(1024.0 / (hi - lo)) : 0.f;
    s_ext[a] = static_cast<float>(hi - lo) * (ord2f(st->max_ord[a]) - ord2f(st->min_ord[a]));  // the window in world units
  }
  __syncthreads();
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float p[3] = {sc.x[i], sc.y[i], sc.z[i]};
  const float ext = fmaxf(fmaxf(s_ext[0], s_ext[1]), s_ext[2]), r2 = 9.f * sc.tr[i];  // (3 sigma)^2; NaN -> class 0
  const uint32_t cls = static_cast<uint32_t>(r2 * 16384.f > ext * ext) + static_cast<uint32_t>(r2 * 256.f > ext * ext);
  uint32_t key = 0x3fffffffu;
  if (finite3(p[0], p[1], p[2])) {
    uint32_t c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float t = (box_fraction(st, a, p[a]) - s_lo[a]) * s_scale[a];
      c[a] = static_cast<uint32_t>(fminf(fmaxf(t, 0.f), 1023.f));
    }
    key = spread3(c[0]) | (spread3(c[1]) << 1) | (spread3(c[2]) << 2);
  }
  keys[i] = ((2u - cls) << 30) | key;  // the large ones first
  vals[i] = i;
}

void launch_spatial_keys(const SceneStorage& sc, uint32_t n, SpatialStats* d_stats, uint32_t* d_keys, uint32_t* d_vals,
                         cudaStream_t stream) {
  if (n == 0) return;
  const uint32_t blocks = (n + 255u) / 256u, wave = static_cast<uint32_t>(sm_count()) * 8u;
  k_stats_init<<<1, 32, 0, stream>>>(d_stats);
  k_stats_minmax<<<blocks < wave ? blocks : wave, 256, 0, stream>>>(sc, n, d_stats);
  k_stats_moments<<<blocks < wave ? blocks : wave, 256, 0, stream>>>(sc, n, d_stats);
  k_spatial_keys<<<blocks, 256, 0, stream>>>(sc, n, d_stats, d_keys, d_vals);
}

// dst[i] = src[order[i]] for the four planar arrays (x, y, z, lambda_max) -> tmp[a * n + i]
__global__ void __launch_bounds__(256) k_gather_planar(SceneStorage sc, const uint32_t* __restrict__ order, uint32_t n,
                                                       float* __restrict__ tmp) {
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = order[i];
  const size_t N = n;
  tmp[i] = sc.x[s];
  tmp[N + i] = sc.y[s];
  tmp[2 * N + i] = sc.z[s];
  tmp[3 * N + i] = sc.tr[s];
}
// payload lines: 8 threads move one 128-byte line
__global__ void __launch_bounds__(256) k_gather_payload(const uint4* __restrict__ src, const uint32_t* __restrict__ order,
                                                        uint32_t n, uint4* __restrict__ dst) {
  const size_t t = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  const size_t i = t >> 3;
  if (i >= n) return;
  dst[t] = src[static_cast<size_t>(order[i]) * 8 + (t & 7u)];
}

int spatial_permute(const SceneStorage& sc, const uint32_t* d_order, uint32_t n, void* d_tmp, cudaStream_t stream) {
  if (n == 0) return 0;
  const size_t N = n;
  float* tf = static_cast<float*>(d_tmp);
  k_gather_planar<<<(n + 255u) / 256u, 256, 0, stream>>>(sc, d_order, n, tf);
  cudaMemcpyAsync(sc.x, tf, N * 4, cudaMemcpyDeviceToDevice, stream);
  cudaMemcpyAsync(sc.y, tf + N, N * 4, cudaMemcpyDeviceToDevice, stream);
  cudaMemcpyAsync(sc.z, tf + 2 * N, N * 4, cudaMemcpyDeviceToDevice, stream);
  cudaMemcpyAsync(sc.tr, tf + 3 * N, N * 4, cudaMemcpyDeviceToDevice, stream);
  const size_t threads = N * 8;
  k_gather_payload<<<static_cast<unsigned int>((threads + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(sc.payload), d_order, n, static_cast<uint4*>(d_tmp));
  cudaMemcpyAsync(sc.payload, d_tmp, N * sizeof(SplatPayload), cudaMemcpyDeviceToDevice, stream);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

__global__ void __launch_bounds__(256) k_iota(uint32_t* v, uint32_t first, uint32_t count) {
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i < count) v[first + i] = first + i;
}
void launch_iota(uint32_t* d_v, uint32_t first, uint32_t count, cudaStream_t stream) {
  if (count) k_iota<<<(count + 255u) / 256u, 256, 0, stream>>>(d_v, first, count);
}

// Tile boxes: one warp per tile of 256 splats.  box[2t] = (min x, min y, min z, max lambda_max),
// box[2t + 1] = (max x, max y, max z, 0).  A tile holding a non-finite centre gets min x = NaN: never decided from the box.
__global__ void __launch_bounds__(256) k_tile_boxes(SceneStorage sc, uint32_t n, uint32_t first_tile, uint32_t n_tiles,
                                                    float4* __restrict__ box) {
  const uint32_t lane = threadIdx.x & 31u, w = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= n_tiles) return;
  const uint32_t t = first_tile + w;
  float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f}, trm = 0.f;
  bool fin = true;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const uint32_t id = t * 256u + it * 32u + lane;
    if (id < n) {
      const float p[3] = {sc.x[id], sc.y[id], sc.z[id]};
      const float tr = sc.tr[id];
      fin = fin && finite3(p[0], p[1], p[2]) && tr <= 3.0e38f;  // false for a NaN lambda_max
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        lo[a] = fminf(lo[a], p[a]);
        hi[a] = fmaxf(hi[a], p[a]);
      }
      trm = fmaxf(trm, tr);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    trm = fmaxf(trm, __shfl_xor_sync(0xffffffffu, trm, o));
  }
  fin = __all_sync(0xffffffffu, fin);
  if (lane == 0) {
    box[2 * static_cast<size_t>(t)] = make_float4(fin ? lo[0] : __uint_as_float(0x7fc00000u), lo[1], lo[2], trm);
    box[2 * static_cast<size_t>(t) + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
  }
}

void launch_tile_boxes(const SceneStorage& sc, uint32_t n, uint32_t first_tile, uint32_t n_tiles, float4* d_box,
                       cudaStream_t stream) {
  if (n_tiles == 0) return;
  k_tile_boxes<<<(n_tiles + 7u) / 8u, 256, 0, stream>>>(sc, n, first_tile, n_tiles, d_box);
}

}  // namespace vkgsb
