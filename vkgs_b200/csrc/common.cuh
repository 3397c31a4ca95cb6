// Shared device-side declarations of the vkgs_b200 frame pipeline (sm_100a).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vkgsb {

constexpr int kTile = 16;              // origin granularity of the pinned fragment arithmetic (oracle: TILE = 16)
// Two-level binning.  The sorted splat list is split ONCE into coarse bins (FrameParams::cshift_*: 128x128 pixels,
// widened on the host until the image has at most kMaxCoarseBins of them, so the split is a single 8-bit radix pass);
// one CTA of the blend stage owns a kBinW x kBinH pixel bin, streams the list of the coarse bin it lies in, keeps the
// entries whose pixel box touches it and refines them to 32 sub-tiles of kSubW x kSubH pixels, one per warp
// (4 pixels per lane).
constexpr int kBinW = 64, kBinH = 64, kSubW = 16, kSubH = 8;
constexpr uint32_t kFlagKeepInstances = 1u;  // also write the reference-format instance records (parity tap)
constexpr uint32_t kFlagBandCull = 4u;       // band rendering: k_project drops splats whose footprint cannot reach the band
constexpr uint32_t kFlagDepthLayer = 2u;     // an opaque line layer is drawn under the splats (lines.cu): k_project also
                                             // writes ndc.z per slot, the blend stage depth-tests against the layer
constexpr int kSubCols = kBinW / kSubW, kSubRows = kBinH / kSubH;  // 4 x 8 = 32 sub-tiles
constexpr int kMaxCoarseBins = 256;
constexpr int kMaxImageDim = 8192;
static_assert(kSubCols * kSubRows == 32, "one sub-tile per warp of a 1024-thread CTA");
constexpr int VKGSB_BLEND_FP32_MODE = 0, VKGSB_BLEND_UNORM8_MODE = 1;  // == enum vkgsb_blend_mode (include/vkgsb.h)

// ---- resident scene (HBM layout) ---------------------------------------------------------------
// Positions are planar so the cull pass streams 12 B/splat fully coalesced; everything only a *visible* splat needs
// is one 128-byte line.
struct __align__(16) SplatPayload {
  float cov[6];   // c00 c10 c20 c11 c21 c22  (parse_ply.comp:80-85)
  float opacity;  // post-sigmoid
  float pad;
  __half sh[48];  // channel-major [3][16]     (projection.comp:31-33,165-168)
};
static_assert(sizeof(SplatPayload) == 128, "payload must be one 128-byte line");

struct Scene {
  const float* x;
  const float* y;
  const float* z;
  const float* tr;  // largest eigenvalue of the 3-D covariance (max scale^2): only the band cull reads it
  const SplatPayload* payload;
  uint32_t n;
  // bounding boxes of the tiles of 256 consecutive splats (spatial.cu): box[2t] = (min x, min y, min z, max lambda_max
  // of the tile), box[2t + 1] = (max x, max y, max z, 0); min x = NaN when the tile holds a non-finite centre
  const float4* box;
};

// ---- per-frame parameter block (device copy of uniforms.h:10-15 + derived values) ----------------------------
struct FrameParams {
  float proj[16];
  float view[16];
  float model[16];
  float pvm[16];        // (proj*view)*model composed on the host, rank.comp:32
  float vm[16];         // view*model                        (projection.comp:98-102 composed on the host)
  float w3[9];          // mat3(view)*mat3(model), m[c*3+r]  (projection.comp:95-101)
  float ps[4];          // mat2(proj), m[c*2+r]              (projection.comp:112)
  float lpx, lpy;       // 1/W/W, 1/H/H                      (projection.comp:116-117)
  float pad2;
  float cam_model[3];   // inverse(model)*eye / w, projection.comp:85-86 hoisted
  uint32_t flags;       // kFlagKeepInstances
  uint32_t pad0;
  uint32_t width, height;
  uint32_t bins_x, bins_y;    // bin grid of the whole image
  uint32_t band_y0, band_y1;  // rows [y0,y1) this renderer bins and blends
  uint32_t bin_y0, bin_y1;    // bin rows covering the band
  uint32_t cshift_x, cshift_y;  // log2 of the coarse bin's width / height in pixels
  uint32_t cbins_x;             // coarse bins per row
  uint32_t cbin_y0;             // first coarse row of the band
  uint32_t ncbins;              // coarse bins covering the band: cbins_x * rows (<= kMaxCoarseBins)
  uint32_t l2_pin_splats;       // centres of splats [0, l2_pin_splats) are kept in L2 across frames (project.cu)
  uint32_t pad1[2];
  float pvm_lines[16];          // (proj*view)*model of the line layer (color.vert, engine.cc:1444-1448)
  // band cull (kFlagBandCull): the footprint's pixel half-height ey satisfies
  //   ey^2 <= bc_a * (bc_p + x_ndc^2 + y_ndc^2) * lambda_max(Sigma) / w^2 + bc_b
  float bc_a, bc_b, bc_p;
  float unorm8_cut;  // VKGSB_BLEND_UNORM8: transmittance below which the first attempt starts its back-to-front walk
  uint32_t epoch;    // frame number (band groups: the value of the hand-shake flags, project.cu)
  unsigned long long dst_image;  // where the blend stage writes this frame's pixels: the renderer's own image, or the
                                 // caller's device destination (possibly another GPU's memory, mapped over NVLink)
};
static_assert(sizeof(FrameParams) % 8 == 0, "FrameParams carries a 64-bit pointer");

// ---- control block: everything the host zeroes with one memset per frame --------------------------------------
struct Control {
  uint32_t visible_count;   // V  (VisiblePointCount, rank.comp:38)
  uint32_t pair_count;      // D
  uint32_t pair_overflow;
  uint32_t blend_retries;    // UNORM8 blend: warp attempts whose bracket stayed open at the front (blend.cu)
  unsigned long long fragment_count;  // fragments shaded by the blend stage (VKGSB_OPT_COUNT_FRAGMENTS)
  uint32_t tile_cut;        // binning tiles [0, tile_cut) fit in max_pairs (bin.cu)
  uint32_t partial_pairs;   // pairs kept of the last kept tile when the capacity cut falls inside it
  uint32_t bin_items;       // work items of k_bin_count / k_bin_place
  uint32_t sort_ticket[4];  // depth passes
  uint32_t pad[3];
  uint32_t hist_depth[4 * 256];
};

// ---- decoupled look-back descriptor for the two ordered block scans (visible slots, pair offsets) --------------
// 64-bit word: low 32 = value, high 32 = status (0 = invalid, 1 = aggregate, 2 = inclusive prefix).
constexpr uint32_t kScanAggregate = 1u, kScanInclusive = 2u;

__device__ __forceinline__ void scan_publish(unsigned long long* d, uint32_t status, uint32_t value) {
  unsigned long long w = (static_cast<unsigned long long>(status) << 32) | value;
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(d), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long scan_peek(const unsigned long long* d) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(d) : "memory");
  return w;
}

// Ordered scan over ticketed tiles in two steps, so that a tile's aggregate can be published long before its prefix is
// needed.  Both are called by ONE full warp with a warp-uniform ticket.
//   scan_post(desc, ticket, total)     publish this tile's aggregate (tile 0: its inclusive value) - never waits.
//   scan_resolve(desc, ticket, total)  exclusive prefix of `total` over all tiles with a smaller ticket: walks back 32
//                                      descriptors at a time until one holds an inclusive value, then publishes this
//                                      tile's inclusive value.  Every lane returns the prefix.
__device__ __forceinline__ void scan_post(unsigned long long* desc, uint32_t ticket, uint32_t total) {
  if ((threadIdx.x & 31u) == 0) scan_publish(desc + ticket, ticket == 0 ? kScanInclusive : kScanAggregate, total);
}
// scan_resolve can be split so that the first round trip to the descriptors overlaps other work:
//   w0 = scan_peek_first(desc, ticket); ...independent work...; prefix = scan_resolve_from(desc, ticket, total, w0);
__device__ __forceinline__ unsigned long long scan_peek_first(const unsigned long long* desc, uint32_t ticket) {
  const int64_t idx = static_cast<int64_t>(ticket) - 1 - (threadIdx.x & 31u);
  // lanes before descriptor 0 act as a zero inclusive terminator
  return idx >= 0 ? scan_peek(desc + idx) : (static_cast<unsigned long long>(kScanInclusive) << 32);
}
__device__ __forceinline__ uint32_t scan_resolve_from(unsigned long long* desc, uint32_t ticket, uint32_t total,
                                                      unsigned long long w) {
  const uint32_t lane = threadIdx.x & 31u;
  if (ticket == 0) return 0u;
  uint32_t prefix = 0;
  int64_t hi = static_cast<int64_t>(ticket) - 1;  // newest descriptor not yet consumed
  while (true) {
    const int64_t idx = hi - lane;
    uint32_t st = static_cast<uint32_t>(w >> 32);
    while (st == 0u) {  // not posted yet (idx >= 0 here: the terminator lanes carry kScanInclusive)
      w = scan_peek(desc + idx);
      st = static_cast<uint32_t>(w >> 32);
    }
    const uint32_t val = static_cast<uint32_t>(w);
    // first lane (nearest predecessor side) holding an inclusive value terminates the walk
    uint32_t incl_mask = __ballot_sync(0xffffffffu, st == kScanInclusive);
    uint32_t first = __ffs(incl_mask) - 1;
    if (incl_mask == 0u) first = 32u;
    uint32_t contrib = (lane <= first) ? val : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
    prefix += contrib;
    if (incl_mask != 0u) break;
    hi -= 32;
    w = hi - lane >= 0 ? scan_peek(desc + (hi - lane)) : (static_cast<unsigned long long>(kScanInclusive) << 32);
  }
  if (lane == 0) scan_publish(desc + ticket, kScanInclusive, prefix + total);
  return prefix;
}
__device__ __forceinline__ uint32_t scan_resolve(unsigned long long* desc, uint32_t ticket, uint32_t total) {
  if (ticket == 0) return 0u;
  return scan_resolve_from(desc, ticket, total, scan_peek_first(desc, ticket));
}
// Both steps back to back.
__device__ __forceinline__ uint32_t scan_lookback_warp(unsigned long long* desc, uint32_t ticket, uint32_t block_total) {
  scan_post(desc, ticket, block_total);
  return scan_resolve(desc, ticket, block_total);
}

}  // namespace vkgsb
