// Host-side launch interface of the stage kernels (one .cu per stage).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "common.cuh"

namespace vkgsb {

// ---- load.cu: parse_ply.comp equivalent + parity tap ---------------------------------------------------------
struct SceneStorage {
  float* x;
  float* y;
  float* z;
  float* tr;  // largest eigenvalue of the 3-D covariance
  SplatPayload* payload;
};
// rows: `count` PLY vertices already on the device; offsets: 60-entry table on the device.  Writes splats
// [first, first+count).
void launch_activate(const float* d_rows, const uint32_t* d_offsets, uint32_t first, uint32_t count,
                     const SceneStorage& dst, cudaStream_t stream);
// back to the reference layout (engine.cc:1639-1651); any destination may be null.
void launch_export_scene(const SceneStorage& src, uint32_t n, float* d_pos, float* d_cov, float* d_opacity,
                         uint16_t* d_sh, cudaStream_t stream);

// ---- spatial.cu: load-time spatial (Morton) order of the scene + tile bounding boxes ------------------------------
struct SpatialStats {                 // of the finite centres
  uint32_t min_ord[3], max_ord[3];    // bounding box (order-preserving integer images of the floats)
  unsigned long long sum[3], sumsq[3], finite;  // moments of the box fractions in 16-bit fixed point; their count
};
// d_keys[i] = 30-bit Morton key of splat i's centre, d_vals[i] = i (sorted by key -> the stored order)
void launch_spatial_keys(const SceneStorage& sc, uint32_t n, SpatialStats* d_stats, uint32_t* d_keys, uint32_t* d_vals,
                         cudaStream_t stream);
// every array of the scene: new[i] = old[d_order[i]], through d_tmp (n * 128 bytes); nonzero on a launch error
int spatial_permute(const SceneStorage& sc, const uint32_t* d_order, uint32_t n, void* d_tmp, cudaStream_t stream);
void launch_iota(uint32_t* d_v, uint32_t first, uint32_t count, cudaStream_t stream);
// boxes of tiles [first_tile, first_tile + n_tiles) of a scene of n splats
void launch_tile_boxes(const SceneStorage& sc, uint32_t n, uint32_t first_tile, uint32_t n_tiles, float4* d_box,
                       cudaStream_t stream);

// ---- project.cu ------------------------------------------------------------------------------------------------
int sm_count();                          // SMs of the current device (cached)
uint32_t project_num_tiles(uint32_t n);  // tiles of 256 splats
// What k_cull leaves for k_project: a visibility bit per splat and a 32-ary tree of visible counts over the tiles.
struct CullIndex {
  uint32_t* mask;      // [tiles * 8]  bit (id & 31) of word id >> 5
  uint32_t* tile_cnt;  // [tiles]      visible splats per tile
  uint32_t* lvl_a;     // [na] sums over 32 tiles      } zero on entry
  uint32_t* lvl_b;     // [nb] sums over 32 A entries  }
  uint32_t* lvl_c;     // [nc] sums over 32 B entries  }
  uint32_t* mixed_cnt;   // [1] zero on entry: tiles the frame tests per splat (cut by the frustum / near the band) ...
  uint32_t* mixed_tile;  // [tiles] ... and which
};
struct CullIndexLayout {
  uint32_t tiles, na, nb, nc;
};
CullIndexLayout cull_index_layout(uint32_t max_splats);
void project_configure();  // once per device: opt in to > 48 KB dynamic shared memory
// ---- band group: the cull of a frame shared out over the W renderers that each draw one screen band of it ------------
constexpr int kMaxGroup = 16;
// Flags of one member, written by the others over NVLink (system-scope release / acquire); epochs are frame numbers.
struct GroupFlags {
  unsigned long long arrive[2][kMaxGroup];  // [parity][source]: source's share of that frame's cull has landed here
  unsigned long long consumed[2];           // [parity]: this member's k_project of that frame has read its cull index
  unsigned int timeout;                     // a wait gave up (a member did not issue the frame): the frame is garbage
};
struct GroupParams {
  uint32_t rank, world;
  uint32_t tile0, tile1;             // this member's share of the scene: CTA tiles [tile0, tile1) of 2048 splats
  uint32_t edges[kMaxGroup + 1];     // band g = rows [edges[g], edges[g + 1])
  CullIndex peer[kMaxGroup];         // member g's cull index of this parity (own entry: local memory)
  GroupFlags* flags[kMaxGroup];      // member g's flags
};
// member `rank`'s share: every band's visibility bits for its splats, written into that band's member
void launch_cull_group(const Scene& scene, const FrameParams* d_fp, const GroupParams& gp, int parity, cudaStream_t stream);
// destination side: wait for every member's share, then build the upper levels of the count tree (zero on entry)
void launch_group_tree(const FrameParams* d_fp, const GroupParams& gp, int parity, uint32_t n, cudaStream_t stream);
// after k_project: this member's cull index of the frame may be overwritten
void launch_group_consumed(const FrameParams* d_fp, GroupFlags* own, int parity, cudaStream_t stream);

// The cull: k_cull_classify decides every tile it can from its bounding box, k_cull_mixed tests the splats of the
// others -> cull index `ix` (its upper levels and mixed_cnt zero on entry).  Reads only the parameter block and the
// scene, so a frame's cull may run while the previous frame is still in its later stages.
void launch_cull(const Scene& scene, const FrameParams* d_fp, const CullIndex& ix, cudaStream_t stream);
// k_project.  d_rrec: 3 x float4 raster record per visible slot; d_inst: 12-float instance record (written only when
// FrameParams.flags has kFlagKeepInstances).  Also accumulates the depth-key digit histograms in Control and writes
// Control::visible_count.
void launch_project(const Scene& scene, const FrameParams* d_fp, Control* d_ctrl, const CullIndex& ix, uint32_t* d_keys,
                    float* d_rrec, uint32_t* d_bin_rect, float* d_inst, float* d_zndc, cudaStream_t stream);
// parity taps: splat id of every visible slot of the last frame (from its cull index) -> d_vis_id
void launch_expand_ids(const CullIndex& ix, uint32_t n, uint32_t* d_vis_id, cudaStream_t stream);

// ---- sort.cu: onesweep LSD radix sort, count read on the device ----------------------------------------------------
struct SortArgs {
  const uint32_t* d_count;  // element count (device)
  uint32_t max_n;           // capacity the launch is sized for
  uint32_t* keys;           // in; the result lands here when the pass count is even, in keys_alt / vals_alt when odd
  uint32_t* vals;
  uint32_t* keys_alt;       // ping-pong scratch, max_n each
  uint32_t* vals_alt;
  uint32_t* hist;           // the passes' digit histograms back to back (256 or 512 bins each, <= 1024 in all);
                            // zero on entry, or already filled when have_hist
  uint32_t* tickets;        // [npass], zero on entry
  uint32_t* lookback;       // sort_lookback_bytes(max_n); cleared by the histogram kernel, or by the caller when
                            // have_hist
  bool have_hist;           // the producer of the keys already built the digit histograms: no histogram pass
  int begin_bit;            // first pass digit starts here
  int npass;                // 1 .. 4
  uint8_t bits[4];          // digit width per pass: 8 (0 reads as 8: the reference's digit) or 9; sum of 2^bits <= 1024
  bool values_only;         // the caller only reads the sorted values: the last pass does not store keys
  bool vals_identity;       // the values are 0, 1, 2, ...: the first pass generates them instead of reading `vals`
  uint32_t clustered_passes;  // bit p: pass p's digit takes only a handful of values (ranked with match.any, one round
                              // per distinct value, instead of one ballot per bit)
};
uint32_t sort_max_parts(uint32_t max_n);
size_t sort_lookback_bytes(uint32_t max_n);
void launch_sort(const SortArgs& a, cudaStream_t stream);

// ---- bin.cu: the sorted list split into one nearest-first list of splat slots per coarse bin ------------------------
uint32_t bin_num_tiles(uint32_t max_visible);                      // tiles of 1024 sorted ranks
uint32_t bin_max_items(uint32_t max_visible, uint64_t max_pairs);  // upper bound of the pair-balanced work items
size_t bin_slots_capacity(uint64_t max_pairs);                     // entries d_bin_slots must hold
struct BinScratch {
  uint32_t* tile_pairs;  // [tile_stride]
  uint32_t* tile_item;   // [tile_stride + 1]
  uint32_t* tile_bin;    // [kMaxCoarseBins][tile_stride]
  uint32_t* bin_total;   // [kMaxCoarseBins]
  uint32_t tile_stride;  // >= bin_num_tiles(max_visible) of every launch
};
// d_ranges[b] = [begin, end) of coarse bin b in d_bin_slots (end <= begin: empty); must be zero on entry.
void launch_bin(const FrameParams* d_fp, uint32_t ncbins, Control* d_ctrl, const uint32_t* d_sorted_slots,
                const uint32_t* d_bin_rect, uint32_t max_visible, uint64_t max_pairs, const BinScratch& scratch,
                uint2* d_ranges, uint32_t* d_bin_slots, cudaStream_t stream);

// ---- blend.cu ------------------------------------------------------------------------------------------------------
void blend_configure();  // once per device: opt in to > 48 KB dynamic shared memory
// ---- lines.cu: opaque line layer (depth bits << 32 | rgba8 per pixel; all ones = no line) --------------------------
void launch_lines(const FrameParams* d_fp, uint32_t n_lines, const float* d_pos, const float* d_col, uint32_t width,
                  uint32_t height, unsigned long long* d_layer, cudaStream_t stream);

// d_ranges: [begin,end) of every coarse bin in d_pair_slot
// d_layer / d_zndc: both null, or the line layer and the splats' ndc.z by slot (depth test LESS against the layer)
void launch_blend(const FrameParams* d_fp, const FrameParams& h_fp, Control* d_ctrl, const uint2* d_ranges,
                  const uint32_t* d_pair_slot, const float* d_rrec, int blend_mode, int bgra, bool count_fragments,
                  const unsigned long long* d_layer, const float* d_zndc, uint8_t* d_image, cudaStream_t stream);

// ---- misc ----------------------------------------------------------------------------------------------------------
void launch_row_histogram(const Control* d_ctrl, const float* d_rrec, uint32_t max_visible, uint32_t height,
                          uint32_t* d_hist, cudaStream_t stream);
void launch_gather_sorted(const Control* d_ctrl, const uint32_t* d_sorted_slots, const uint32_t* d_vis_id,
                          const float* d_inst, uint32_t max_visible, uint32_t* d_ids_out, float* d_inst_out,
                          cudaStream_t stream);

}  // namespace vkgsb
