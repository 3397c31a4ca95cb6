// C ABI (include/vkgsb.h) and frame orchestration: the CUDA counterpart of Engine::Impl (engine.cc:114-524 buffers,
// 742-1378 Draw) without the window, swapchain and UI.  One renderer = one device, one render stream, one loader
// thread.  Storage is allocated once in vkgsb_create (the reference pre-allocates for MAX_SPLAT_COUNT,
// engine.cc:483-502); per frame the host sends the parameter block (256 B in the reference's UBO + push constant,
// here one kernel-argument upload), clears a ~0.6 MB control region and replays a CUDA graph of the stage kernels.
// Counts (visible splats, pairs) never come back to the host on the critical path, like the reference's indirect
// dispatch (engine.cc:1218-1219, projection.comp:64-75).
#include <atomic>
#include <condition_variable>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vkgsb.h"
#include "hostmath.h"
#include "kernels.h"
#include "ply.h"

using namespace vkgsb;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define CU_TRY(expr)                                                                                  \
  do {                                                                                                \
    cudaError_t e_ = (expr);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return fail(VKGSB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));                \
  } while (0)

__global__ void k_set_params(FrameParams p, FrameParams* dst) {
  if (threadIdx.x == 0) *dst = p;
}

constexpr uint32_t kChunkVertices = 65536;  // splat_load_thread.cc:140

}  // namespace

struct vkgsb_renderer {
  int device = 0;
  uint32_t max_splats = 0, max_width = 0, max_height = 0;
  uint64_t max_pairs = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t load_stream = nullptr;

  // resident scene
  SceneStorage scene{};
  std::atomic<uint32_t> scene_n{0};
  // spatial order (spatial.cu): the scene is stored in Morton order of the centres; order[i] = index in the file of stored
  // splat i; boxes = the bounding box of every tile of 256 stored splats (what the cull classifies tiles from)
  float4* boxes = nullptr;
  uint32_t* order = nullptr;
  SpatialStats* d_stats = nullptr;
  int spatial_order = 1;  // VKGSB_OPT_SPATIAL_ORDER

  // Per-frame work buffers, TWO sets used by alternate frames (index p = the frame's parity): consecutive frames of a
  // batch then run side by side on the sets' own streams (vkgsb_draw_batch), and a single frame's cull runs beside the
  // previous frame's later stages (cull_stream).
  uint32_t *keys[2] = {}, *slots[2] = {}, *keys_alt[2] = {}, *slots_alt[2] = {}, *vis_id = nullptr;
  float* inst[2] = {};         // reference-format instance records: allocated when VKGSB_OPT_KEEP_INSTANCES is first set
  float* rrec[2] = {};         // raster records by compacted slot (project.cu)
  // opaque line layer (vkgsb_set_lines): geometry, the per-pixel depth | colour words, the splats' ndc.z by slot
  uint32_t n_lines = 0;
  float *line_pos = nullptr, *line_col = nullptr, *zndc[2] = {};
  unsigned long long* layer[2] = {};
  float line_model[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  uint32_t* bin_rect[2] = {};   // coarse-bin box by compacted slot
  uint32_t* bin_slots[2] = {};  // splat slots by coarse bin, nearest first (bin.cu); capacity max_pairs
  BinScratch bin[2]{};
  uint32_t* lookback_depth[2] = {};
  uint8_t* zero_region[2] = {};  // Control | coarse-bin ranges
  size_t zero_bytes = 0;
  Control* ctrl[2] = {};
  cudaStream_t slot_stream[2] = {nullptr, nullptr};  // the work-buffer sets' own streams (batches)
  cudaStream_t slot_last_stream[2] = {nullptr, nullptr};  // the stream that last drew with set p
  // Two copies, used by alternate frames, of what a frame's cull needs and leaves: the parameter block and the cull index
  // (k_cull -> k_project, project.cu).  Frame f + 1's cull then runs on cull_stream while frame f is still in its
  // projection / sort / binning / blend on the frame's stream.
  CullIndex cull[2]{};
  uint32_t* cull_tree[2] = {nullptr, nullptr};  // the upper levels of cull[p]'s count tree + its mixed-tile count: zeroed before every cull
  uint32_t* mixed_tile[2] = {nullptr, nullptr};  // cull[p]'s list of tiles tested per splat
  size_t cull_tree_bytes = 0;
  cudaStream_t cull_stream = nullptr;
  cudaEvent_t cull_done[2] = {nullptr, nullptr};
  // band group (vkgsb_group_*): mask + tile counts of both parities and the hand-shake flags live in ONE allocation the
  // other members map; `gp[p]` holds every member's pointers for parity p
  uint8_t* group_block = nullptr;
  size_t group_off_mask[2] = {0, 0}, group_off_cnt[2] = {0, 0}, group_off_flags = 0, group_bytes = 0;
  GroupFlags* group_flags = nullptr;
  bool grouped = false;
  GroupParams gp[2]{};
  void* group_peer_base[kMaxGroup] = {nullptr};
  bool group_peer_ipc[kMaxGroup] = {false};
  uint64_t group_epoch = 0;  // frames drawn as a member; parity of a group frame = its epoch & 1 (the same on every member)
  int last_parity = 0;
  uint2* ranges[2] = {};
  FrameParams* d_fp[2] = {nullptr, nullptr};
  uint8_t* image[2] = {};
  // draw_batch to host memory: a finished frame leaves image[p] over PCIe on copy_stream while the next frames render
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t frame_done[2] = {nullptr, nullptr}, copy_done[2] = {nullptr, nullptr}, batch_start = nullptr;
  uint32_t* h_counts[2] = {};  // pinned: visible, pairs, overflow, ... of the set's last frame

  // load staging
  float* d_rows[2] = {nullptr, nullptr};
  float* h_rows[2] = {nullptr, nullptr};
  size_t rows_capacity_bytes = 0;
  uint32_t* d_offsets = nullptr;
  cudaEvent_t chunk_done[2] = {nullptr, nullptr};

  // frame state
  vkgsb_camera cam{};
  bool have_cam = false;
  uint32_t width = 0, height = 0;
  int blend_mode = VKGSB_BLEND_FP32, pixel_format = VKGSB_FORMAT_RGBA8, stage_timing = 0, keep_instances = 0;
  int count_fragments = 0;
  float unorm8_cut = 1e-4f;
  int l2_pin_mb = 72;  // VKGSB_OPT_L2_PIN_MB
  bool last_frame_has_instances = false;
  uint32_t band_y0 = 0, band_y1 = 0;
  int band_cull = 1;  // VKGSB_OPT_BAND_CULL
  FrameParams h_fp{};
  // per parity: the cull graph (cull_stream) and the rest of the frame (the frame's stream); valid for one
  // (n, viewport, band, mode, format)
  cudaGraphExec_t graph_cull[2] = {nullptr, nullptr}, graph_main[2] = {nullptr, nullptr};
  bool graph_valid[2] = {false, false};
  uint32_t graph_n[2] = {0, 0};
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t flight[4] = {nullptr, nullptr, nullptr, nullptr};  // end of frame f on the stream that drew it, at [f & 3]
  bool ev_recorded = false;
  uint64_t frame_counter = 0;
  std::mutex draw_mutex;

  // loader thread (SplatLoadThread)
  std::thread loader;
  std::mutex load_mutex;
  std::condition_variable load_cv;
  std::string pending_path;
  bool loader_exit = false;
  std::atomic<bool> cancel{false};
  std::atomic<uint32_t> total_points{0}, loaded_points{0};
  std::atomic<int> load_state{0};  // 0 idle, 1 loading, 2 done, <0 error
  std::string load_error;
};

namespace {

int set_device(vkgsb_renderer* r) {
  CU_TRY(cudaSetDevice(r->device));
  return VKGSB_OK;
}

void invalidate_graph(vkgsb_renderer* r) { r->graph_valid[0] = r->graph_valid[1] = false; }

// Every frame issued so far has finished - on the renderer's stream or on a caller's.
cudaError_t drain_frames(vkgsb_renderer* r) {
  cudaError_t e = cudaStreamSynchronize(r->stream);
  if (e == cudaSuccess && r->cull_stream) e = cudaStreamSynchronize(r->cull_stream);
  if (e == cudaSuccess && r->frame_counter) e = cudaEventSynchronize(r->flight[r->frame_counter & 3]);
  if (e == cudaSuccess && r->frame_counter > 1) e = cudaEventSynchronize(r->flight[(r->frame_counter - 1) & 3]);
  for (cudaStream_t st : {r->slot_stream[0], r->slot_stream[1], r->copy_stream})
    if (e == cudaSuccess && st) e = cudaStreamSynchronize(st);
  return e;
}

int ensure_row_staging(vkgsb_renderer* r, uint32_t stride_bytes) {
  size_t need = static_cast<size_t>(kChunkVertices) * stride_bytes;
  if (need <= r->rows_capacity_bytes) return VKGSB_OK;
  for (int i = 0; i < 2; ++i) {
    if (r->d_rows[i]) cudaFree(r->d_rows[i]);
    if (r->h_rows[i]) cudaFreeHost(r->h_rows[i]);
    r->d_rows[i] = nullptr;
    r->h_rows[i] = nullptr;
  }
  r->rows_capacity_bytes = 0;
  for (int i = 0; i < 2; ++i) {
    CU_TRY(cudaMalloc(&r->d_rows[i], need));
    CU_TRY(cudaMallocHost(&r->h_rows[i], need));
  }
  r->rows_capacity_bytes = need;
  return VKGSB_OK;
}

// The resident scene (n splats, file order, no frame in flight) -> Morton order of the centres (spatial.cu) + tile boxes.
// Uses work-buffer set 0 for the sort and a temporary of n * 128 bytes; when that cannot be allocated the scene simply
// stays in file order (slower frames, same pixels).
int spatial_reorder(vkgsb_renderer* r, uint32_t n) {
  cudaStream_t ls = r->load_stream;
  void* tmp = nullptr;
  if (cudaMalloc(&tmp, static_cast<size_t>(n) * sizeof(SplatPayload)) != cudaSuccess) {
    cudaGetLastError();
    return VKGSB_OK;
  }
  auto run = [&]() -> cudaError_t {
    cudaError_t e = cudaMemsetAsync(r->zero_region[0], 0, r->zero_bytes, ls);
    if (e != cudaSuccess) return e;
    launch_spatial_keys(r->scene, n, r->d_stats, r->keys[0], r->slots[0], ls);
    e = cudaMemcpyAsync(&r->ctrl[0]->visible_count, &n, sizeof(uint32_t), cudaMemcpyHostToDevice, ls);
    if (e != cudaSuccess) return e;
    SortArgs a{};
    a.d_count = &r->ctrl[0]->visible_count;
    a.max_n = n;
    a.keys = r->keys[0]; a.vals = r->slots[0]; a.keys_alt = r->keys_alt[0]; a.vals_alt = r->slots_alt[0];
    a.hist = r->ctrl[0]->hist_depth; a.tickets = r->ctrl[0]->sort_ticket; a.lookback = r->lookback_depth[0];
    a.begin_bit = 0; a.npass = 4;  // 4 x 8 bits: an even pass count, the result lands in keys / vals
    launch_sort(a, ls);
    e = cudaMemcpyAsync(r->order, r->slots[0], static_cast<size_t>(n) * 4, cudaMemcpyDeviceToDevice, ls);
    if (e != cudaSuccess) return e;
    if (spatial_permute(r->scene, r->order, n, tmp, ls)) return cudaErrorLaunchFailure;
    launch_tile_boxes(r->scene, n, 0, (n + 255u) / 256u, r->boxes, ls);
    e = cudaStreamSynchronize(ls);
    return e != cudaSuccess ? e : cudaGetLastError();
  };
  const cudaError_t e = run();
  cudaFree(tmp);
  if (e != cudaSuccess) {
    r->scene_n.store(0);  // the arrays may be half permuted: nothing is resident
    return fail(VKGSB_ERR_CUDA, std::string("spatial reorder: ") + cudaGetErrorString(e));
  }
  return VKGSB_OK;
}

// Streamed ingest shared by upload_splats (memory source) and load_ply (file source): 65 536-vertex chunks through
// two pinned buffers, H2D and activation of chunk k overlapping the host fill of chunk k+1.
// fill(dst, first_vertex, count) returns false on a short read.
template <class Fill>
int ingest(vkgsb_renderer* r, uint64_t n64, uint32_t stride_bytes, const uint32_t offsets[60], Fill fill) {
  if (set_device(r)) return VKGSB_ERR_CUDA;
  if (n64 > r->max_splats)
    return fail(VKGSB_ERR_CAPACITY, "scene has " + std::to_string(n64) + " splats, renderer was created for " +
                                        std::to_string(r->max_splats));
  const uint32_t n = static_cast<uint32_t>(n64);
  if (int e = ensure_row_staging(r, stride_bytes)) return e;
  {
    // the resident scene is rewritten in place: retire it and drain frames that still read it
    std::lock_guard<std::mutex> g(r->draw_mutex);
    r->scene_n.store(0);
    invalidate_graph(r);
    CU_TRY(drain_frames(r));
  }
  r->total_points.store(n);
  r->loaded_points.store(0);
  CU_TRY(cudaMemcpyAsync(r->d_offsets, offsets, 60 * sizeof(uint32_t), cudaMemcpyHostToDevice, r->load_stream));
  CU_TRY(cudaStreamSynchronize(r->load_stream));  // offsets may live on the caller's stack
  int buf = 0;
  for (uint64_t start = 0; start < n; start += kChunkVertices, buf ^= 1) {
    if (r->cancel.load()) {
      cudaStreamSynchronize(r->load_stream);
      return fail(VKGSB_ERR_CANCELLED, "load cancelled");
    }
    const uint32_t count = static_cast<uint32_t>(std::min<uint64_t>(kChunkVertices, n - start));
    CU_TRY(cudaEventSynchronize(r->chunk_done[buf]));  // staging buffer free again: chunk start - 2 chunks is resident
    if (start >= 2ull * kChunkVertices) {
      // like the reference, which draws the loaded_point_count splats parsed so far (engine.cc:1137-1160): everything
      // before the chunk still in flight is published to the frame path
      std::lock_guard<std::mutex> g(r->draw_mutex);
      r->scene_n.store(static_cast<uint32_t>(start - kChunkVertices));
    }
    if (!fill(r->h_rows[buf], start, count)) {
      cudaStreamSynchronize(r->load_stream);
      return fail(VKGSB_ERR_IO, "short read at vertex " + std::to_string(start));
    }
    CU_TRY(cudaMemcpyAsync(r->d_rows[buf], r->h_rows[buf], static_cast<size_t>(count) * stride_bytes,
                           cudaMemcpyHostToDevice, r->load_stream));
    launch_activate(r->d_rows[buf], r->d_offsets, static_cast<uint32_t>(start), count, r->scene, r->load_stream);
    // in file order until the whole scene is resident: boxes of the chunk's tiles (a chunk is a whole number of tiles)
    launch_tile_boxes(r->scene, static_cast<uint32_t>(start + count), static_cast<uint32_t>(start / 256u), (count + 255u) / 256u,
                      r->boxes, r->load_stream);
    launch_iota(r->order, static_cast<uint32_t>(start), count, r->load_stream);
    CU_TRY(cudaEventRecord(r->chunk_done[buf], r->load_stream));
    r->loaded_points.store(static_cast<uint32_t>(start + count));
  }
  CU_TRY(cudaStreamSynchronize(r->load_stream));
  CU_TRY(cudaGetLastError());
  {
    std::lock_guard<std::mutex> g(r->draw_mutex);
    if (r->spatial_order && n > 1) {
      // frames drawn during the load read the file-ordered prefix: drain them, then store the scene in spatial order
      CU_TRY(drain_frames(r));
      if (int e = spatial_reorder(r, n)) return e;
    }
    r->scene_n.store(n);
    invalidate_graph(r);
  }
  return VKGSB_OK;
}

int load_file(vkgsb_renderer* r, const std::string& path) {
  PlyHeader h;
  std::string err = parse_ply_header(path, &h);
  if (!err.empty()) return fail(VKGSB_ERR_IO, path + ": " + err);
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return fail(VKGSB_ERR_IO, "cannot open " + path);
  // 64-bit offsets: the reference's `offset * start` is 32-bit and wraps past 4 GiB (splat_load_thread.cc:145)
  if (fseeko(f, static_cast<off_t>(h.body_offset), SEEK_SET) != 0) {
    std::fclose(f);
    return fail(VKGSB_ERR_IO, "seek failed in " + path);
  }
  int rc = ingest(r, h.vertex_count, h.stride_bytes, h.offsets, [&](float* dst, uint64_t, uint32_t count) {
    size_t want = static_cast<size_t>(count) * h.stride_bytes;
    return std::fread(dst, 1, want, f) == want;
  });
  std::fclose(f);
  return rc;
}

void loader_main(vkgsb_renderer* r) {
  while (true) {
    std::string path;
    {
      std::unique_lock<std::mutex> g(r->load_mutex);
      r->load_cv.wait(g, [r] { return r->loader_exit || !r->pending_path.empty(); });
      if (r->loader_exit) return;
      path = std::move(r->pending_path);
      r->pending_path.clear();
      r->cancel.store(false);
      r->load_state.store(1);
    }
    int rc = load_file(r, path);
    {
      std::unique_lock<std::mutex> g(r->load_mutex);
      if (rc != VKGSB_OK) r->load_error = g_last_error;
      // a newer request supersedes this result
      if (r->pending_path.empty()) r->load_state.store(rc == VKGSB_OK ? 2 : -rc);
    }
    r->load_cv.notify_all();
  }
}

void fill_params(vkgsb_renderer* r) {
  FrameParams& p = r->h_fp;
  std::memcpy(p.proj, r->cam.projection, 64);
  std::memcpy(p.view, r->cam.view, 64);
  std::memcpy(p.model, r->cam.model, 64);
  float pv[16];
  mat4_mul(p.proj, p.view, pv);  // projection * view * model, left to right (rank.comp:32)
  mat4_mul(pv, p.model, p.pvm);
  float inv[16], e[4] = {r->cam.camera_position[0], r->cam.camera_position[1], r->cam.camera_position[2], 1.f}, cm[4];
  mat4_inverse(p.model, inv);  // inverse(model) * vec4(camera_position, 1), hoisted (projection.comp:85-86)
  mat4_vec(inv, e, cm);
  mat4_mul(p.view, p.model, p.vm);
  for (int c = 0; c < 3; ++c)
    for (int rr = 0; rr < 3; ++rr)  // mat3(view) * mat3(model), same chain as the oracle's mat3_mul
      p.w3[c * 3 + rr] = std::fmaf(p.view[2 * 4 + rr], p.model[c * 4 + 2],
                                   std::fmaf(p.view[1 * 4 + rr], p.model[c * 4 + 1], p.view[0 * 4 + rr] * p.model[c * 4 + 0]));
  p.ps[0] = p.proj[0]; p.ps[1] = p.proj[1]; p.ps[2] = p.proj[4]; p.ps[3] = p.proj[5];
  const float fw = static_cast<float>(r->width), fh = static_cast<float>(r->height);
  p.lpx = 1.f / fw / fw;
  p.lpy = 1.f / fh / fh;
  p.pad2 = 0.f;
  p.cam_model[0] = cm[0] / cm[3];
  p.cam_model[1] = cm[1] / cm[3];
  p.cam_model[2] = cm[2] / cm[3];
  p.width = r->width;
  p.height = r->height;
  p.flags = (r->keep_instances ? kFlagKeepInstances : 0u) | (r->n_lines ? kFlagDepthLayer : 0u);
  mat4_mul(pv, r->line_model, p.pvm_lines);  // projection * view * model of the lines (color.vert)
  p.pad0 = 0u;
  p.bins_x = (r->width + kBinW - 1) / kBinW;
  p.bins_y = (r->height + kBinH - 1) / kBinH;
  p.band_y0 = std::min(r->band_y0, r->height);
  p.band_y1 = (r->band_y1 == 0 || r->band_y1 > r->height) ? r->height : r->band_y1;
  if (p.band_y1 < p.band_y0) p.band_y1 = p.band_y0;
  p.bin_y0 = p.band_y0 / kBinH;
  p.bin_y1 = (p.band_y1 + kBinH - 1) / kBinH;
  if (p.band_y1 == p.band_y0) p.bin_y1 = p.bin_y0;
  // coarse bins: 128 x 128 pixels, widened (x first) until the band has at most kMaxCoarseBins of them
  p.cshift_x = p.cshift_y = 7;
  auto coarse = [&](uint32_t* rows) {
    p.cbins_x = ((r->width - 1) >> p.cshift_x) + 1;
    p.cbin_y0 = p.band_y0 >> p.cshift_y;
    *rows = p.band_y1 > p.band_y0 ? ((p.band_y1 - 1) >> p.cshift_y) - p.cbin_y0 + 1 : 0;
    return p.cbins_x * *rows;
  };
  uint32_t crows = 0;
  while (coarse(&crows) > static_cast<uint32_t>(kMaxCoarseBins)) {
    if (p.cshift_x <= p.cshift_y) ++p.cshift_x; else ++p.cshift_y;
  }
  p.ncbins = p.cbins_x * crows;
  p.pad1[0] = p.pad1[1] = 0u;
  p.l2_pin_splats = static_cast<uint32_t>(std::min<uint64_t>(r->scene_n.load(), (static_cast<uint64_t>(r->l2_pin_mb) << 20) / 12u));
  // Band rendering (one GPU of a screen-band partition): splats whose footprint cannot reach the band
  // are dropped at the cull, so sort / projection / binning shrink with the band.  The footprint's pixel box has
  // half-height ey = 3 * (H/2) * (|RS10| + |RS11|) <= 3 * (H/2) * sqrt(trace(cov2d))   (RS RS^T = cov2d), and
  //   trace(cov2d) = trace(K Sigma K^T) + lpx + lpy <= |K|_F^2 * lambda_max(Sigma) + lpx + lpy,
  //   |K|_F <= |mat2(proj) J|_F * |W|_2,   |mat2(proj) J|_F^2 = (P00^2 + P11^2 + x_ndc^2 + y_ndc^2) / w^2   (w = -z_view).
  // Valid for a projection of the reference's shape (camera.cc:25-35: diagonal 2x2 block, w = -z); any other
  // projection matrix keeps the whole visible set.
  {
    const float* P = p.proj;
    const bool shaped = P[1] == 0.f && P[2] == 0.f && P[3] == 0.f && P[4] == 0.f && P[6] == 0.f && P[7] == 0.f &&
                        P[8] == 0.f && P[9] == 0.f && P[11] == -1.f && P[12] == 0.f && P[13] == 0.f && P[15] == 0.f;
    const bool banded = p.band_y0 > 0u || p.band_y1 < p.height;
    // |W|_2^2 = largest eigenvalue of G = W^T W, bounded from above two ways: Gershgorin (the largest absolute row sum
    // of G - exact for a uniformly scaled rotation, the usual model matrix) and the trace (the Frobenius norm of W).
    // Both are upper bounds for every W, so their minimum is one.
    double G[9], wf2 = 0.0, gersh = 0.0;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        G[a * 3 + b] = 0.0;
        for (int k = 0; k < 3; ++k) G[a * 3 + b] += static_cast<double>(p.w3[a * 3 + k]) * p.w3[b * 3 + k];
      }
    for (int i = 0; i < 9; ++i) wf2 += static_cast<double>(p.w3[i]) * p.w3[i];
    for (int a = 0; a < 3; ++a) gersh = std::max(gersh, std::fabs(G[a * 3]) + std::fabs(G[a * 3 + 1]) + std::fabs(G[a * 3 + 2]));
    double w2 = std::min(wf2, gersh) * (1.0 + 1e-6);
    if (!(w2 > 0.0)) w2 = wf2;  // degenerate or non-finite model: NaN keeps everything
    const float hh = 0.5f * fh;
    p.bc_a = static_cast<float>(9.0 * hh * hh * w2);
    p.bc_b = 9.f * hh * hh * (p.lpx + p.lpy);
    p.bc_p = P[0] * P[0] + P[5] * P[5];
    p.unorm8_cut = r->unorm8_cut;
    if (shaped && banded && r->band_cull) p.flags |= kFlagBandCull;
  }
}

// The cull of a frame (parity p) on `s`: count tree cleared, k_cull.
int record_cull(vkgsb_renderer* r, int p, cudaStream_t s, bool clear = true) {
  const uint32_t n = r->scene_n.load();
  Scene sc{r->scene.x, r->scene.y, r->scene.z, r->scene.tr, r->scene.payload, n, r->boxes};
  if (clear) CU_TRY(cudaMemsetAsync(r->cull_tree[p], 0, r->cull_tree_bytes, s));
  if (r->grouped) {
    // this member's share of the scene against every band, written into the bands' members; then wait for the others'
    // shares of this band and build the count tree
    GroupParams gp = r->gp[p];
    const uint32_t nct = (n + 2047u) / 2048u;
    gp.tile0 = static_cast<uint32_t>(static_cast<uint64_t>(nct) * gp.rank / gp.world);
    gp.tile1 = static_cast<uint32_t>(static_cast<uint64_t>(nct) * (gp.rank + 1) / gp.world);
    launch_cull_group(sc, r->d_fp[p], gp, p, s);
    launch_group_tree(r->d_fp[p], gp, p, n, s);
  } else {
    launch_cull(sc, r->d_fp[p], r->cull[p], s);
  }
  CU_TRY(cudaGetLastError());
  return VKGSB_OK;
}

// What a frame clears and draws before its splat stages, on `s`.
int record_clears(vkgsb_renderer* r, int p, cudaStream_t s) {
  const uint32_t n = r->scene_n.load();
  CU_TRY(cudaMemsetAsync(r->zero_region[p], 0, r->zero_bytes, s));
  // look-back words of the depth sort: only partitions of the <= n visible splats can be touched
  CU_TRY(cudaMemsetAsync(r->lookback_depth[p], 0, sort_lookback_bytes(n), s));
  if (r->n_lines) launch_lines(r->d_fp[p], r->n_lines, r->line_pos, r->line_col, r->width, r->height, r->layer[p], s);
  return VKGSB_OK;
}

// The stages behind the cull on `s`.  With `timed`, CUDA events bracket the stages (the caller records ev[0] before the
// cull and ev[5] behind it).
int record_stages(vkgsb_renderer* r, int p, cudaStream_t s, bool timed) {
  const uint32_t n = r->scene_n.load();
  Scene sc{r->scene.x, r->scene.y, r->scene.z, r->scene.tr, r->scene.payload, n, r->boxes};
  FrameParams* d_fp = r->d_fp[p];
  // the depth sort runs an odd number of passes: its input goes to the ping-pong side, its result lands in keys / slots
  launch_project(sc, d_fp, r->ctrl[p], r->cull[p], r->keys_alt[p], r->rrec[p], r->bin_rect[p], r->inst[p], r->n_lines ? r->zndc[p] : nullptr, s);
  if (r->grouped) launch_group_consumed(d_fp, r->group_flags, p, s);  // the others may write the next frame of this parity
  if (timed) CU_TRY(cudaEventRecord(r->ev[1], s));
  SortArgs depth{};
  depth.d_count = &r->ctrl[p]->visible_count;
  depth.max_n = n;
  depth.keys = r->keys_alt[p]; depth.vals = r->slots_alt[p]; depth.keys_alt = r->keys[p]; depth.vals_alt = r->slots[p];
  depth.hist = r->ctrl[p]->hist_depth; depth.tickets = r->ctrl[p]->sort_ticket; depth.lookback = r->lookback_depth[p];
  // k_project emits the depth key as the integer (1 - z) * 2^24 in [0, 2^24] (the float 1 - z is always a multiple of
  // 2^-24, so this is exact and ordered like the reference's float bits): 25 live bits = 3 passes of 8 + 8 + 9 bits
  // instead of the reference's 4 x 8 over the full word
  depth.begin_bit = 0; depth.npass = 3;
  depth.bits[0] = 8; depth.bits[1] = 8; depth.bits[2] = 9;
  depth.clustered_passes = 1u << 2;  // bits 16..24 of (1 - z) * 2^24: a few values per warp
  depth.have_hist = true;  // k_project accumulated the three digit histograms
  depth.vals_identity = true;  // the value is the compacted slot: generated by the first pass
  launch_sort(depth, s);
  if (timed) CU_TRY(cudaEventRecord(r->ev[2], s));
  launch_bin(d_fp, r->h_fp.ncbins, r->ctrl[p], r->slots[p], r->bin_rect[p], n, r->max_pairs, r->bin[p], r->ranges[p], r->bin_slots[p], s);
  if (timed) CU_TRY(cudaEventRecord(r->ev[3], s));
  launch_blend(d_fp, r->h_fp, r->ctrl[p], r->ranges[p], r->bin_slots[p], r->rrec[p], r->blend_mode,
               r->pixel_format == VKGSB_FORMAT_BGRA8, r->count_fragments != 0, r->n_lines ? r->layer[p] : nullptr, r->n_lines ? r->zndc[p] : nullptr,
               r->image[p], s);
  if (timed) CU_TRY(cudaEventRecord(r->ev[4], s));
  CU_TRY(cudaMemcpyAsync(r->h_counts[p], r->ctrl[p], 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaGetLastError());
  return VKGSB_OK;
}

// `direct_dst`: a device destination the blend stage writes straight into (null: the renderer's own image)
int run_frame(vkgsb_renderer* r, cudaStream_t s, void* direct_dst = nullptr) {
  if (!r->have_cam) return fail(VKGSB_ERR_INVALID, "vkgsb_set_camera has not been called");
  if (r->width == 0 || r->height == 0) return fail(VKGSB_ERR_INVALID, "vkgsb_set_viewport has not been called");
  const uint32_t n = r->scene_n.load();
  // nothing resident: an error, unless a load is under way or lines are set - the reference's window shows the clear
  // colour and the axis / grid while the first chunks are parsed
  if (n == 0 && r->n_lines == 0 && r->load_state.load() != 1) return fail(VKGSB_ERR_NO_SCENE, "no splats loaded");
  fill_params(r);
  r->h_fp.dst_image = reinterpret_cast<unsigned long long>(direct_dst);
  const uint64_t f = r->frame_counter + 1;  // this frame
  // a group frame's parity is the same on every member (they write into each other's parity buffers)
  if (r->grouped) r->group_epoch++;
  const int p = static_cast<int>((r->grouped ? r->group_epoch : f) & 1);
  r->h_fp.epoch = static_cast<uint32_t>(r->group_epoch);
  r->last_parity = p;
  // frame f - 2 used the same set of work buffers: a frame on another stream than that one waits for it
  if (f > 2 && s != r->slot_last_stream[p]) CU_TRY(cudaStreamWaitEvent(s, r->flight[(f - 2) & 3], 0));
  if (r->stage_timing) {
    // eager, everything on the frame's stream, events between the stages
    k_set_params<<<1, 32, 0, s>>>(r->h_fp, r->d_fp[p]);
    CU_TRY(cudaMemsetAsync(r->cull_tree[p], 0, r->cull_tree_bytes, s));
    if (int e = record_clears(r, p, s)) return e;
    CU_TRY(cudaEventRecord(r->ev[0], s));
    if (int e = record_cull(r, p, s, false)) return e;
    CU_TRY(cudaEventRecord(r->ev[5], s));
    if (int e = record_stages(r, p, s, true)) return e;
    r->ev_recorded = true;
  } else {
    if (!r->graph_valid[p] || r->graph_n[p] != n) {
      for (cudaGraphExec_t* g : {&r->graph_cull[p], &r->graph_main[p]})
        if (*g) {
          cudaGraphExecDestroy(*g);
          *g = nullptr;
        }
      cudaGraph_t graph = nullptr;
      CU_TRY(cudaStreamBeginCapture(r->cull_stream, cudaStreamCaptureModeThreadLocal));
      int e = record_cull(r, p, r->cull_stream);
      cudaError_t ce = cudaStreamEndCapture(r->cull_stream, &graph);
      if (e) return e;
      CU_TRY(ce);
      CU_TRY(cudaGraphInstantiate(&r->graph_cull[p], graph, 0));
      cudaGraphDestroy(graph);
      CU_TRY(cudaStreamBeginCapture(r->stream, cudaStreamCaptureModeThreadLocal));
      e = record_clears(r, p, r->stream);
      if (!e) e = record_stages(r, p, r->stream, false);
      ce = cudaStreamEndCapture(r->stream, &graph);
      if (e) return e;
      CU_TRY(ce);
      CU_TRY(cudaGraphInstantiate(&r->graph_main[p], graph, 0));
      cudaGraphDestroy(graph);
      r->graph_valid[p] = true;
      r->graph_n[p] = n;
    }
    // The cull reads the scene and its parity's parameter block only: it runs on cull_stream as soon as frame f - 2
    // (the last reader of that parameter block and cull index) has finished - that is, beside frame f - 1's later stages
    // when frames are issued back to back.  The rest of the frame follows on the frame's stream.
    if (f > 2) CU_TRY(cudaStreamWaitEvent(r->cull_stream, r->flight[(f - 2) & 3], 0));
    k_set_params<<<1, 32, 0, r->cull_stream>>>(r->h_fp, r->d_fp[p]);
    CU_TRY(cudaGraphLaunch(r->graph_cull[p], r->cull_stream));
    CU_TRY(cudaEventRecord(r->cull_done[p], r->cull_stream));
    CU_TRY(cudaStreamWaitEvent(s, r->cull_done[p], 0));
    CU_TRY(cudaGraphLaunch(r->graph_main[p], s));
    r->ev_recorded = false;
  }
  r->frame_counter++;
  r->slot_last_stream[p] = s;
  CU_TRY(cudaEventRecord(r->flight[r->frame_counter & 3], s));
  r->last_frame_has_instances = r->keep_instances != 0;
  return VKGSB_OK;
}

}  // namespace

extern "C" {

const char* vkgsb_last_error(void) { return g_last_error.c_str(); }

// for the other translation units of the library (interop.cu): the calling thread's error text
__attribute__((visibility("hidden"))) void vkgsb_set_last_error(const char* msg) { g_last_error = msg ? msg : ""; }

int vkgsb_device_count(int* count) {
  if (!count) return fail(VKGSB_ERR_INVALID, "count is null");
  *count = 0;
  CU_TRY(cudaGetDeviceCount(count));
  return VKGSB_OK;
}

int vkgsb_create(int device, uint32_t max_splats, vkgsb_renderer** out) {
  vkgsb_config c{};
  c.struct_size = sizeof(c);
  c.device = device;
  c.max_splats = max_splats;
  return vkgsb_create_ex(&c, out);
}

int vkgsb_create_ex(const vkgsb_config* cfg, vkgsb_renderer** out) {
  if (!cfg || !out) return fail(VKGSB_ERR_INVALID, "null argument");
  if (cfg->struct_size != sizeof(vkgsb_config)) return fail(VKGSB_ERR_INVALID, "vkgsb_config.struct_size mismatch");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(VKGSB_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                    " (this renderer has no CPU path)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(VKGSB_ERR_INVALID, "device ordinal out of range");
  auto* r = new vkgsb_renderer();
  r->device = cfg->device;
  r->max_splats = cfg->max_splats ? cfg->max_splats : (1u << 23);
  r->max_width = cfg->max_width ? cfg->max_width : 3840;
  r->max_height = cfg->max_height ? cfg->max_height : 2160;
  r->max_pairs = cfg->max_pairs ? cfg->max_pairs : 8ull * r->max_splats;
  if (r->max_pairs > (1ull << 31)) r->max_pairs = 1ull << 31;
  if (r->max_pairs < 4096) r->max_pairs = 4096;

  auto bail = [&](const std::string& what, cudaError_t ce) {
    std::string msg = what + ": " + cudaGetErrorString(ce);
    vkgsb_destroy(r);
    return fail(VKGSB_ERR_CUDA, msg);
  };
#define ALLOC(ptr, bytes)                                                  \
  do {                                                                     \
    cudaError_t ce_ = cudaMalloc(reinterpret_cast<void**>(&(ptr)), bytes); \
    if (ce_ != cudaSuccess) return bail("cudaMalloc " #ptr, ce_);          \
  } while (0)

  if ((e = cudaSetDevice(r->device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&r->load_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&r->cull_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  for (auto& ev : r->cull_done)
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
  for (int i = 0; i < 2; ++i) {
    if ((e = cudaEventCreateWithFlags(&r->frame_done[i], cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
    if ((e = cudaEventCreateWithFlags(&r->copy_done[i], cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
  }
  const size_t N = r->max_splats, P = r->max_pairs;
  ALLOC(r->scene.x, N * 4); ALLOC(r->scene.y, N * 4); ALLOC(r->scene.z, N * 4); ALLOC(r->scene.tr, N * 4);
  ALLOC(r->scene.payload, N * sizeof(SplatPayload));
  ALLOC(r->vis_id, N * 4);
  ALLOC(r->order, N * 4);
  ALLOC(r->boxes, static_cast<size_t>(project_num_tiles(r->max_splats)) * 2 * sizeof(float4));
  ALLOC(r->d_stats, sizeof(SpatialStats));
  const size_t ctrl_bytes = (sizeof(Control) + 255) & ~size_t(255);
  r->zero_bytes = ctrl_bytes + kMaxCoarseBins * sizeof(uint2);
  for (int p = 0; p < 2; ++p) {
    ALLOC(r->keys[p], N * 4); ALLOC(r->slots[p], N * 4); ALLOC(r->keys_alt[p], N * 4); ALLOC(r->slots_alt[p], N * 4);
    ALLOC(r->rrec[p], N * 48);
    ALLOC(r->bin_rect[p], N * 4);
    ALLOC(r->bin_slots[p], bin_slots_capacity(P) * 4);
    r->bin[p].tile_stride = bin_num_tiles(r->max_splats);
    ALLOC(r->bin[p].tile_pairs, static_cast<size_t>(r->bin[p].tile_stride) * 4);
    ALLOC(r->bin[p].tile_item, (static_cast<size_t>(r->bin[p].tile_stride) + 1) * 4);
    ALLOC(r->bin[p].tile_bin, static_cast<size_t>(kMaxCoarseBins) * r->bin[p].tile_stride * 4);
    ALLOC(r->bin[p].bin_total, kMaxCoarseBins * 4);
    ALLOC(r->lookback_depth[p], sort_lookback_bytes(r->max_splats));
    ALLOC(r->zero_region[p], r->zero_bytes);
    r->ctrl[p] = reinterpret_cast<Control*>(r->zero_region[p]);
    r->ranges[p] = reinterpret_cast<uint2*>(r->zero_region[p] + ctrl_bytes);
    ALLOC(r->image[p], static_cast<size_t>(r->max_width) * r->max_height * 4);
  }
  const CullIndexLayout cl = cull_index_layout(r->max_splats);
  r->cull_tree_bytes = ((static_cast<size_t>(cl.na) + cl.nb + cl.nc + 1 + 63) & ~size_t(63)) * 4;  // + the mixed-tile count
  {
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    size_t o = 0;
    for (int p = 0; p < 2; ++p) {
      r->group_off_mask[p] = o;
      o += up(static_cast<size_t>(cl.tiles) * 8 * 4);
      r->group_off_cnt[p] = o;
      o += up(static_cast<size_t>(cl.tiles) * 4);
    }
    r->group_off_flags = o;
    o += up(sizeof(GroupFlags));
    r->group_bytes = o;
    ALLOC(r->group_block, r->group_bytes);
    if ((e = cudaMemset(r->group_block + r->group_off_flags, 0, sizeof(GroupFlags))) != cudaSuccess) return bail("cudaMemset", e);
    r->group_flags = reinterpret_cast<GroupFlags*>(r->group_block + r->group_off_flags);
  }
  for (int p = 0; p < 2; ++p) {
    ALLOC(r->cull_tree[p], r->cull_tree_bytes);
    r->cull[p].lvl_a = r->cull_tree[p];
    r->cull[p].lvl_b = r->cull[p].lvl_a + cl.na;
    r->cull[p].lvl_c = r->cull[p].lvl_b + cl.nb;
    r->cull[p].mixed_cnt = r->cull[p].lvl_c + cl.nc;
    ALLOC(r->mixed_tile[p], static_cast<size_t>(cl.tiles) * 4);
    r->cull[p].mixed_tile = r->mixed_tile[p];
    r->cull[p].mask = reinterpret_cast<uint32_t*>(r->group_block + r->group_off_mask[p]);
    r->cull[p].tile_cnt = reinterpret_cast<uint32_t*>(r->group_block + r->group_off_cnt[p]);
    ALLOC(r->d_fp[p], sizeof(FrameParams));
  }
  ALLOC(r->d_offsets, 60 * 4);
#undef ALLOC
  for (int p = 0; p < 2; ++p) {
    if ((e = cudaMallocHost(reinterpret_cast<void**>(&r->h_counts[p]), 64)) != cudaSuccess) return bail("cudaMallocHost", e);
    std::memset(r->h_counts[p], 0, 64);
    if ((e = cudaStreamCreateWithFlags(&r->slot_stream[p], cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  }
  if ((e = cudaEventCreateWithFlags(&r->batch_start, cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
  for (auto& ev : r->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
  for (auto& ev : r->chunk_done)
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  for (auto& ev : r->flight)
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  blend_configure();
  project_configure();
  r->loader = std::thread(loader_main, r);
  *out = r;
  return VKGSB_OK;
}

void vkgsb_destroy(vkgsb_renderer* r) {
  if (!r) return;
  if (r->loader.joinable()) {
    {
      std::unique_lock<std::mutex> g(r->load_mutex);
      r->loader_exit = true;
      r->cancel.store(true);
    }
    r->load_cv.notify_all();
    r->loader.join();
  }
  cudaSetDevice(r->device);
  if (r->stream) drain_frames(r);
  if (r->grouped) vkgsb_group_leave(r);
  if (r->load_stream) cudaStreamSynchronize(r->load_stream);
  for (cudaGraphExec_t g : {r->graph_cull[0], r->graph_cull[1], r->graph_main[0], r->graph_main[1]})
    if (g) cudaGraphExecDestroy(g);
  std::vector<void*> dev = {r->scene.x, r->scene.y, r->scene.z, r->scene.tr, r->scene.payload, r->vis_id, r->group_block,
                            r->order, r->boxes, r->d_stats, r->mixed_tile[0], r->mixed_tile[1],
                            r->d_offsets, r->d_rows[0], r->d_rows[1], r->line_pos, r->line_col};
  for (int p = 0; p < 2; ++p)
    for (void* q : {static_cast<void*>(r->keys[p]), static_cast<void*>(r->slots[p]), static_cast<void*>(r->keys_alt[p]),
                    static_cast<void*>(r->slots_alt[p]), static_cast<void*>(r->inst[p]), static_cast<void*>(r->rrec[p]),
                    static_cast<void*>(r->bin_rect[p]), static_cast<void*>(r->bin_slots[p]), static_cast<void*>(r->bin[p].tile_pairs),
                    static_cast<void*>(r->bin[p].tile_item), static_cast<void*>(r->bin[p].tile_bin),
                    static_cast<void*>(r->bin[p].bin_total), static_cast<void*>(r->lookback_depth[p]),
                    static_cast<void*>(r->zero_region[p]), static_cast<void*>(r->cull_tree[p]), static_cast<void*>(r->d_fp[p]),
                    static_cast<void*>(r->image[p]), static_cast<void*>(r->zndc[p]), static_cast<void*>(r->layer[p])})
      dev.push_back(q);
  for (void* p : dev)
    if (p) cudaFree(p);
  for (int p = 0; p < 2; ++p) {
    if (r->h_counts[p]) cudaFreeHost(r->h_counts[p]);
    if (r->slot_stream[p]) cudaStreamDestroy(r->slot_stream[p]);
  }
  if (r->batch_start) cudaEventDestroy(r->batch_start);
  for (auto* p : r->h_rows)
    if (p) cudaFreeHost(p);
  for (auto& ev : r->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : r->chunk_done)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : r->flight)
    if (ev) cudaEventDestroy(ev);
  for (int i = 0; i < 2; ++i) {
    if (r->frame_done[i]) cudaEventDestroy(r->frame_done[i]);
    if (r->copy_done[i]) cudaEventDestroy(r->copy_done[i]);
  }
  if (r->stream) cudaStreamDestroy(r->stream);
  if (r->load_stream) cudaStreamDestroy(r->load_stream);
  if (r->copy_stream) cudaStreamDestroy(r->copy_stream);
  if (r->cull_stream) cudaStreamDestroy(r->cull_stream);
  for (auto& ev : r->cull_done)
    if (ev) cudaEventDestroy(ev);
  delete r;
}

int vkgsb_set_option(vkgsb_renderer* r, int option, int64_t value) {
  if (!r) return fail(VKGSB_ERR_INVALID, "renderer is null");
  std::lock_guard<std::mutex> g(r->draw_mutex);
  switch (option) {
    case VKGSB_OPT_STAGE_TIMING: r->stage_timing = value != 0; break;
    case VKGSB_OPT_BLEND_MODE:
      if (value != VKGSB_BLEND_FP32 && value != VKGSB_BLEND_UNORM8) return fail(VKGSB_ERR_INVALID, "bad blend mode");
      r->blend_mode = static_cast<int>(value);
      break;
    case VKGSB_OPT_PIXEL_FORMAT:
      if (value != VKGSB_FORMAT_RGBA8 && value != VKGSB_FORMAT_BGRA8) return fail(VKGSB_ERR_INVALID, "bad pixel format");
      r->pixel_format = static_cast<int>(value);
      break;
    case VKGSB_OPT_KEEP_INSTANCES:
      if (value != 0)
        for (int p = 0; p < 2; ++p)
          if (!r->inst[p]) {
            if (set_device(r)) return VKGSB_ERR_CUDA;
            CU_TRY(cudaMalloc(&r->inst[p], static_cast<size_t>(r->max_splats) * 48));
          }
      r->keep_instances = value != 0;
      break;
    case VKGSB_OPT_BAND_Y0: r->band_y0 = static_cast<uint32_t>(value); break;
    case VKGSB_OPT_BAND_Y1: r->band_y1 = static_cast<uint32_t>(value); break;
    case VKGSB_OPT_BAND_CULL: r->band_cull = value != 0; break;
    case VKGSB_OPT_COUNT_FRAGMENTS: r->count_fragments = value != 0; break;
    case VKGSB_OPT_L2_PIN_MB:
      if (value < 0 || value > 4096) return fail(VKGSB_ERR_INVALID, "L2 pin size out of range");
      r->l2_pin_mb = static_cast<int>(value);
      break;
    case VKGSB_OPT_SPATIAL_ORDER: r->spatial_order = value != 0; break;
    case VKGSB_OPT_UNORM8_CUT_EXP:
      if (value < 1 || value > 18) return fail(VKGSB_ERR_INVALID, "unorm8 cut exponent must be in [1, 18]");
      r->unorm8_cut = std::pow(10.f, -static_cast<float>(value));
      break;
    default: return fail(VKGSB_ERR_INVALID, "unknown option");
  }
  invalidate_graph(r);
  return VKGSB_OK;
}

int vkgsb_load_ply_async(vkgsb_renderer* r, const char* path) {
  if (!r || !path) return fail(VKGSB_ERR_INVALID, "null argument");
  {
    std::unique_lock<std::mutex> g(r->load_mutex);
    r->cancel.store(true);  // Cancel(); Start(path)   (engine.cc:541-544)
    r->pending_path = path;
    r->load_state.store(1);
  }
  r->load_cv.notify_all();
  return VKGSB_OK;
}

int vkgsb_wait_load(vkgsb_renderer* r) {
  if (!r) return fail(VKGSB_ERR_INVALID, "renderer is null");
  std::unique_lock<std::mutex> g(r->load_mutex);
  r->load_cv.wait(g, [r] { return r->load_state.load() != 1; });
  int st = r->load_state.load();
  if (st < 0) return fail(-st, r->load_error);
  return VKGSB_OK;
}

int vkgsb_load_ply(vkgsb_renderer* r, const char* path) {
  if (int e = vkgsb_load_ply_async(r, path)) return e;
  return vkgsb_wait_load(r);
}

int vkgsb_load_progress(vkgsb_renderer* r, uint32_t* total, uint32_t* loaded, int* state) {
  if (!r) return fail(VKGSB_ERR_INVALID, "renderer is null");
  if (total) *total = r->total_points.load();
  if (loaded) *loaded = r->loaded_points.load();
  if (state) *state = r->load_state.load();
  return VKGSB_OK;
}

int vkgsb_cancel_load(vkgsb_renderer* r) {
  if (!r) return fail(VKGSB_ERR_INVALID, "renderer is null");
  r->cancel.store(true);
  return VKGSB_OK;
}

int vkgsb_upload_splats(vkgsb_renderer* r, uint32_t n, const float* rows, const uint32_t offsets[60]) {
  if (!r || !rows || !offsets) return fail(VKGSB_ERR_INVALID, "null argument");
  if (n == 0) return fail(VKGSB_ERR_INVALID, "n is 0");
  const uint32_t stride = offsets[59];
  for (int i = 0; i < 59; ++i)
    if (offsets[i] >= stride) return fail(VKGSB_ERR_INVALID, "offset table entry beyond the row stride");
  if (int e = vkgsb_wait_load(r); e != VKGSB_OK && e != VKGSB_ERR_CANCELLED && e != VKGSB_ERR_IO) return e;
  r->cancel.store(false);
  int rc = ingest(r, n, stride * 4u, offsets, [&](float* dst, uint64_t first, uint32_t count) {
    std::memcpy(dst, rows + first * stride, static_cast<size_t>(count) * stride * 4u);
    return true;
  });
  r->load_state.store(rc == VKGSB_OK ? 2 : -rc);
  return rc;
}

int vkgsb_set_camera(vkgsb_renderer* r, const vkgsb_camera* cam) {
  if (!r || !cam) return fail(VKGSB_ERR_INVALID, "null argument");
  r->cam = *cam;
  r->have_cam = true;
  return VKGSB_OK;
}

int vkgsb_set_viewport(vkgsb_renderer* r, uint32_t width, uint32_t height) {
  if (!r) return fail(VKGSB_ERR_INVALID, "renderer is null");
  if (width == 0 || height == 0) return fail(VKGSB_ERR_INVALID, "empty viewport");
  if (static_cast<size_t>(width) * height > static_cast<size_t>(r->max_width) * r->max_height ||
      width > static_cast<uint32_t>(kMaxImageDim) || height > static_cast<uint32_t>(kMaxImageDim))
    return fail(VKGSB_ERR_CAPACITY, "viewport larger than the renderer was created for");
  if (width != r->width || height != r->height) {
    std::lock_guard<std::mutex> g(r->draw_mutex);
    r->width = width;
    r->height = height;
    invalidate_graph(r);
  }
  return VKGSB_OK;
}

int vkgsb_set_lines(vkgsb_renderer* r, uint32_t n_lines, const float* positions, const float* colors,
                    const float model[16]) {
  if (!r) return fail(VKGSB_ERR_INVALID, "renderer is null");
  if (n_lines && (!positions || !colors)) return fail(VKGSB_ERR_INVALID, "null line geometry");
  if (n_lines > (1u << 20)) return fail(VKGSB_ERR_CAPACITY, "at most 2^20 lines");
  if (set_device(r)) return VKGSB_ERR_CUDA;
  std::lock_guard<std::mutex> g(r->draw_mutex);
  CU_TRY(drain_frames(r));  // frames in flight still read the old geometry
  invalidate_graph(r);
  if (r->line_pos) cudaFree(r->line_pos);
  if (r->line_col) cudaFree(r->line_col);
  r->line_pos = r->line_col = nullptr;
  r->n_lines = 0;
  if (n_lines == 0) return VKGSB_OK;
  for (int p = 0; p < 2; ++p) {
    if (!r->layer[p]) CU_TRY(cudaMalloc(&r->layer[p], static_cast<size_t>(r->max_width) * r->max_height * sizeof(unsigned long long)));
    if (!r->zndc[p]) CU_TRY(cudaMalloc(&r->zndc[p], static_cast<size_t>(r->max_splats) * sizeof(float)));
  }
  CU_TRY(cudaMalloc(&r->line_pos, static_cast<size_t>(n_lines) * 6 * sizeof(float)));
  CU_TRY(cudaMalloc(&r->line_col, static_cast<size_t>(n_lines) * 8 * sizeof(float)));
  CU_TRY(cudaMemcpy(r->line_pos, positions, static_cast<size_t>(n_lines) * 6 * sizeof(float), cudaMemcpyHostToDevice));
  CU_TRY(cudaMemcpy(r->line_col, colors, static_cast<size_t>(n_lines) * 8 * sizeof(float), cudaMemcpyHostToDevice));
  static const float kIdentity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  std::memcpy(r->line_model, model ? model : kIdentity, sizeof(r->line_model));
  r->n_lines = n_lines;
  return VKGSB_OK;
}

int vkgsb_draw(vkgsb_renderer* r, void* dst, int dst_is_device, void* stream) {
  if (!r) return fail(VKGSB_ERR_INVALID, "renderer is null");
  if (set_device(r)) return VKGSB_ERR_CUDA;
  std::lock_guard<std::mutex> g(r->draw_mutex);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : r->stream;
  // a device destination is rendered into directly (no copy): the blend kernel's stores go to `dst`, wherever it lives
  if (int e = run_frame(r, s, dst && dst_is_device ? dst : nullptr)) return e;
  const size_t bytes = static_cast<size_t>(r->width) * r->height * 4;
  if (dst && !dst_is_device) {
    CU_TRY(cudaMemcpyAsync(dst, r->image[r->last_parity], bytes, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
  }
  return VKGSB_OK;
}

// The frames of a batch need nothing from each other, so consecutive frames run SIDE BY SIDE: frame i on the stream of
// work-buffer set i & 1, behind frame i - 2 only.  The memory-bound projection of one frame then shares the GPU with
// the latency-bound sort and the issue-bound blend of the other (the reference keeps two frames in flight for the
// same reason, engine.cc:1028-1035).  Everything the caller queued on `stream` before the call precedes the batch; the
// stream continues when the whole batch has finished.
int vkgsb_draw_batch(vkgsb_renderer* r, uint32_t n_views, const vkgsb_camera* cameras, void* dst, size_t dst_stride,
                     int dst_is_device, void* stream) {
  if (!r || !cameras) return fail(VKGSB_ERR_INVALID, "null argument");
  if (set_device(r)) return VKGSB_ERR_CUDA;
  std::lock_guard<std::mutex> g(r->draw_mutex);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : r->stream;
  const size_t bytes = static_cast<size_t>(r->width) * r->height * 4;
  if (dst && dst_stride < bytes) return fail(VKGSB_ERR_INVALID, "dst_stride smaller than one image");
  if (n_views == 0) return VKGSB_OK;
  const bool host = dst && !dst_is_device;
  // stage timing measures one frame at a time; a band group's members hand-shake frame by frame
  const bool side_by_side = !r->stage_timing && !r->grouped && n_views > 1;
  if (side_by_side) {
    CU_TRY(cudaEventRecord(r->batch_start, s));
    CU_TRY(cudaStreamWaitEvent(r->slot_stream[0], r->batch_start, 0));
    CU_TRY(cudaStreamWaitEvent(r->slot_stream[1], r->batch_start, 0));
  }
  bool copying[2] = {false, false};
  for (uint32_t i = 0; i < n_views; ++i) {
    r->cam = cameras[i];
    r->have_cam = true;
    const int p = static_cast<int>((r->grouped ? r->group_epoch + 1 : r->frame_counter + 1) & 1);  // the set run_frame will use
    cudaStream_t fs = side_by_side ? r->slot_stream[p] : s;
    // host destination: the frame leaves image[p] over PCIe on copy_stream while the next frames render; the set's next
    // frame waits for that copy before it overwrites the image
    if (host && copying[p]) CU_TRY(cudaStreamWaitEvent(fs, r->copy_done[p], 0));
    if (int e = run_frame(r, fs, dst && dst_is_device ? static_cast<uint8_t*>(dst) + i * dst_stride : nullptr)) return e;
    if (host) {
      CU_TRY(cudaEventRecord(r->frame_done[p], fs));
      CU_TRY(cudaStreamWaitEvent(r->copy_stream, r->frame_done[p], 0));
      CU_TRY(cudaMemcpyAsync(static_cast<uint8_t*>(dst) + i * dst_stride, r->image[p], bytes, cudaMemcpyDeviceToHost,
                             r->copy_stream));
      CU_TRY(cudaEventRecord(r->copy_done[p], r->copy_stream));
      copying[p] = true;
    }
  }
  if (side_by_side) {  // the caller's stream continues behind the last frame of either set
    CU_TRY(cudaStreamWaitEvent(s, r->flight[r->frame_counter & 3], 0));
    CU_TRY(cudaStreamWaitEvent(s, r->flight[(r->frame_counter - 1) & 3], 0));
  }
  if (host) {
    CU_TRY(cudaStreamSynchronize(r->copy_stream));  // returns when every image is in dst
    CU_TRY(cudaStreamSynchronize(s));
  }
  return VKGSB_OK;
}

int vkgsb_image_device_ptr(vkgsb_renderer* r, void** ptr) {
  if (!r || !ptr) return fail(VKGSB_ERR_INVALID, "null argument");
  *ptr = r->image[r->last_parity];  // the last frame drawn without a destination
  return VKGSB_OK;
}

int vkgsb_sync(vkgsb_renderer* r) {
  if (!r) return fail(VKGSB_ERR_INVALID, "renderer is null");
  if (set_device(r)) return VKGSB_ERR_CUDA;
  CU_TRY(drain_frames(r));
  if (r->grouped) {
    unsigned int t = 0;
    CU_TRY(cudaMemcpy(&t, &r->group_flags->timeout, sizeof(t), cudaMemcpyDeviceToHost));
    if (t) return fail(VKGSB_ERR_CUDA, "band group: a member did not deliver its share of the cull in time (were all members' frames issued?)");
  }
  return VKGSB_OK;
}

int vkgsb_wait_frame(vkgsb_renderer* r, uint32_t frames_back) {
  if (!r) return fail(VKGSB_ERR_INVALID, "renderer is null");
  if (frames_back == 0 || frames_back > 3) return fail(VKGSB_ERR_INVALID, "frames_back must be 1, 2 or 3");
  if (set_device(r)) return VKGSB_ERR_CUDA;
  uint64_t f;
  {
    std::lock_guard<std::mutex> g(r->draw_mutex);
    f = r->frame_counter + 1;  // the frame the caller is about to issue
  }
  if (f > frames_back) CU_TRY(cudaEventSynchronize(r->flight[(f - frames_back) & 3]));
  return VKGSB_OK;
}

int vkgsb_get_stats(vkgsb_renderer* r, vkgsb_stats* out) {
  if (!r || !out) return fail(VKGSB_ERR_INVALID, "null argument");
  if (set_device(r)) return VKGSB_ERR_CUDA;
  std::lock_guard<std::mutex> g(r->draw_mutex);
  CU_TRY(cudaDeviceSynchronize());
  std::memset(out, 0, sizeof(*out));
  out->total_point_count = r->total_points.load();
  out->loaded_point_count = r->loaded_points.load();
  out->visible_point_count = r->h_counts[r->last_parity][0];
  out->pair_count = r->h_counts[r->last_parity][1];
  out->pair_overflow = r->h_counts[r->last_parity][2];
  out->blend_retries = r->h_counts[r->last_parity][3];
  out->fragment_count = static_cast<uint64_t>(r->h_counts[r->last_parity][4]) | (static_cast<uint64_t>(r->h_counts[r->last_parity][5]) << 32);
  out->frame_counter = r->frame_counter;
  if (r->ev_recorded) {
    CU_TRY(cudaEventElapsedTime(&out->ms_project, r->ev[0], r->ev[1]));
    CU_TRY(cudaEventElapsedTime(&out->ms_cull, r->ev[0], r->ev[5]));
    CU_TRY(cudaEventElapsedTime(&out->ms_sort, r->ev[1], r->ev[2]));
    CU_TRY(cudaEventElapsedTime(&out->ms_bin, r->ev[2], r->ev[3]));
    CU_TRY(cudaEventElapsedTime(&out->ms_blend, r->ev[3], r->ev[4]));
    CU_TRY(cudaEventElapsedTime(&out->ms_total, r->ev[0], r->ev[4]));
  }
  return VKGSB_OK;
}

int vkgsb_read_sorted(vkgsb_renderer* r, uint32_t* keys, uint32_t* ids, uint32_t capacity, uint32_t* count) {
  if (!r || !count) return fail(VKGSB_ERR_INVALID, "null argument");
  if (set_device(r)) return VKGSB_ERR_CUDA;
  std::lock_guard<std::mutex> g(r->draw_mutex);
  CU_TRY(cudaDeviceSynchronize());
  uint32_t v = 0;
  CU_TRY(cudaMemcpy(&v, &r->ctrl[r->last_parity]->visible_count, 4, cudaMemcpyDeviceToHost));
  *count = v;
  if (v > capacity) return fail(VKGSB_ERR_CAPACITY, "capacity smaller than the visible count");
  if (v == 0) return VKGSB_OK;
  if (keys) {
    CU_TRY(cudaMemcpy(keys, r->keys[r->last_parity], v * 4ull, cudaMemcpyDeviceToHost));
    // the frame sorts the integer (1 - z) * 2^24; the tap returns the reference's key, floatBitsToUint(1 - z)
    // (rank.comp:40) - the conversion is exact both ways
    for (uint32_t i = 0; i < v; ++i) {
      const float f = static_cast<float>(keys[i]) * 5.9604644775390625e-08f;
      std::memcpy(&keys[i], &f, 4);
    }
  }
  if (ids) {
    // slots_alt is free between frames: gather ids there; the ids by slot come from the frame's cull index
    launch_expand_ids(r->cull[r->last_parity], r->scene_n.load(), r->vis_id, r->stream);
    launch_gather_sorted(r->ctrl[r->last_parity], r->slots[r->last_parity], r->vis_id, r->inst[r->last_parity], v, r->slots_alt[r->last_parity], nullptr, r->stream);
    CU_TRY(cudaStreamSynchronize(r->stream));
    CU_TRY(cudaMemcpy(ids, r->slots_alt[r->last_parity], v * 4ull, cudaMemcpyDeviceToHost));
  }
  return VKGSB_OK;
}

int vkgsb_row_histogram(vkgsb_renderer* r, uint32_t* rows, uint32_t capacity) {
  if (!r || !rows) return fail(VKGSB_ERR_INVALID, "null argument");
  if (r->height == 0 || capacity < r->height) return fail(VKGSB_ERR_CAPACITY, "rows[] is shorter than the viewport height");
  if (set_device(r)) return VKGSB_ERR_CUDA;
  std::lock_guard<std::mutex> g(r->draw_mutex);
  CU_TRY(cudaDeviceSynchronize());
  uint32_t* d_hist = nullptr;
  CU_TRY(cudaMalloc(&d_hist, r->height * sizeof(uint32_t)));
  launch_row_histogram(r->ctrl[r->last_parity], r->rrec[r->last_parity], r->scene_n.load(), r->height, d_hist, r->stream);
  cudaError_t e = cudaStreamSynchronize(r->stream);
  if (e == cudaSuccess) e = cudaMemcpy(rows, d_hist, r->height * sizeof(uint32_t), cudaMemcpyDeviceToHost);
  cudaFree(d_hist);
  CU_TRY(e);
  return VKGSB_OK;
}

int vkgsb_read_instances(vkgsb_renderer* r, float* inst, uint32_t capacity, uint32_t* count) {
  if (!r || !count) return fail(VKGSB_ERR_INVALID, "null argument");
  if (set_device(r)) return VKGSB_ERR_CUDA;
  std::lock_guard<std::mutex> g(r->draw_mutex);
  CU_TRY(cudaDeviceSynchronize());
  uint32_t v = 0;
  CU_TRY(cudaMemcpy(&v, &r->ctrl[r->last_parity]->visible_count, 4, cudaMemcpyDeviceToHost));
  *count = v;
  if (v > capacity) return fail(VKGSB_ERR_CAPACITY, "capacity smaller than the visible count");
  if (v == 0 || !inst) return VKGSB_OK;
  if (!r->last_frame_has_instances)
    return fail(VKGSB_ERR_INVALID, "instance records were not kept: set VKGSB_OPT_KEEP_INSTANCES before drawing");
  float* tmp = nullptr;
  CU_TRY(cudaMalloc(&tmp, v * 48ull));
  launch_gather_sorted(r->ctrl[r->last_parity], r->slots[r->last_parity], r->vis_id, r->inst[r->last_parity], v, nullptr, tmp, r->stream);
  cudaError_t e = cudaStreamSynchronize(r->stream);
  if (e == cudaSuccess) e = cudaMemcpy(inst, tmp, v * 48ull, cudaMemcpyDeviceToHost);
  cudaFree(tmp);
  CU_TRY(e);
  return VKGSB_OK;
}

int vkgsb_read_scene(vkgsb_renderer* r, float* pos, float* cov, float* opacity, uint16_t* sh, uint32_t capacity,
                     uint32_t* count) {
  if (!r || !count) return fail(VKGSB_ERR_INVALID, "null argument");
  if (set_device(r)) return VKGSB_ERR_CUDA;
  const uint32_t n = r->scene_n.load();
  *count = n;
  if (n > capacity) return fail(VKGSB_ERR_CAPACITY, "capacity smaller than the scene");
  if (n == 0) return VKGSB_OK;
  float *dp = nullptr, *dc = nullptr, *dop = nullptr;
  uint16_t* ds = nullptr;
  cudaError_t e = cudaSuccess;
  if (pos && e == cudaSuccess) e = cudaMalloc(&dp, n * 12ull);
  if (cov && e == cudaSuccess) e = cudaMalloc(&dc, n * 24ull);
  if (opacity && e == cudaSuccess) e = cudaMalloc(&dop, n * 4ull);
  if (sh && e == cudaSuccess) e = cudaMalloc(&ds, n * 96ull);
  if (e == cudaSuccess) {
    launch_export_scene(r->scene, n, dp, dc, dop, ds, r->stream);
    e = cudaStreamSynchronize(r->stream);
  }
  if (pos && e == cudaSuccess) e = cudaMemcpy(pos, dp, n * 12ull, cudaMemcpyDeviceToHost);
  if (cov && e == cudaSuccess) e = cudaMemcpy(cov, dc, n * 24ull, cudaMemcpyDeviceToHost);
  if (opacity && e == cudaSuccess) e = cudaMemcpy(opacity, dop, n * 4ull, cudaMemcpyDeviceToHost);
  if (sh && e == cudaSuccess) e = cudaMemcpy(sh, ds, n * 96ull, cudaMemcpyDeviceToHost);
  cudaFree(dp); cudaFree(dc); cudaFree(dop); cudaFree(ds);
  CU_TRY(e);
  return VKGSB_OK;
}

int vkgsb_read_order(vkgsb_renderer* r, uint32_t* order, uint32_t capacity, uint32_t* count) {
  if (!r || !count) return fail(VKGSB_ERR_INVALID, "null argument");
  if (set_device(r)) return VKGSB_ERR_CUDA;
  const uint32_t n = r->scene_n.load();
  *count = n;
  if (!order) return VKGSB_OK;
  if (n > capacity) return fail(VKGSB_ERR_CAPACITY, "capacity smaller than the scene");
  if (n) CU_TRY(cudaMemcpy(order, r->order, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost));
  return VKGSB_OK;
}

// ---- band group ----------------------------------------------------------------------------------------------------------
static int group_setup(vkgsb_renderer* r, uint32_t rank, uint32_t world, void* const* bases, const bool* ipc,
                       const uint32_t* edges) {
  if (world < 2 || world > static_cast<uint32_t>(kMaxGroup) || rank >= world) return fail(VKGSB_ERR_INVALID, "bad group size / rank");
  for (uint32_t g = 0; g < world; ++g)
    if (edges[g] > edges[g + 1]) return fail(VKGSB_ERR_INVALID, "band edges must be non-decreasing");
  std::lock_guard<std::mutex> g(r->draw_mutex);
  CU_TRY(cudaDeviceSynchronize());
  for (int p = 0; p < 2; ++p) {
    GroupParams& gp = r->gp[p];
    gp = GroupParams{};
    gp.rank = rank;
    gp.world = world;
    for (uint32_t i = 0; i <= world; ++i) gp.edges[i] = edges[i];
    for (uint32_t m = 0; m < world; ++m) {
      uint8_t* base = m == rank ? r->group_block : static_cast<uint8_t*>(bases[m]);
      gp.peer[m] = m == rank ? r->cull[p] : CullIndex{};
      gp.peer[m].mask = reinterpret_cast<uint32_t*>(base + r->group_off_mask[p]);       // every member was created with the
      gp.peer[m].tile_cnt = reinterpret_cast<uint32_t*>(base + r->group_off_cnt[p]);    // same max_splats: same layout
      gp.flags[m] = reinterpret_cast<GroupFlags*>(base + r->group_off_flags);
    }
  }
  for (uint32_t m = 0; m < world; ++m) {
    r->group_peer_base[m] = m == rank ? nullptr : bases[m];
    r->group_peer_ipc[m] = m != rank && ipc[m];
  }
  CU_TRY(cudaMemset(r->group_flags, 0, sizeof(GroupFlags)));
  r->group_epoch = 0;
  r->band_y0 = edges[rank];
  r->band_y1 = edges[rank + 1];
  r->grouped = true;
  invalidate_graph(r);
  return VKGSB_OK;
}

int vkgsb_group_export(vkgsb_renderer* r, uint8_t handle[64]) {
  if (!r || !handle) return fail(VKGSB_ERR_INVALID, "null argument");
  if (set_device(r)) return VKGSB_ERR_CUDA;
  cudaIpcMemHandle_t h;
  CU_TRY(cudaIpcGetMemHandle(&h, r->group_block));
  std::memcpy(handle, &h, 64);
  return VKGSB_OK;
}

int vkgsb_group_join(vkgsb_renderer* r, uint32_t rank, uint32_t world, const uint8_t* handles, const uint32_t* edges) {
  if (!r || !handles || !edges) return fail(VKGSB_ERR_INVALID, "null argument");
  if (world > static_cast<uint32_t>(kMaxGroup)) return fail(VKGSB_ERR_INVALID, "at most 16 members");
  if (set_device(r)) return VKGSB_ERR_CUDA;
  if (r->grouped) return fail(VKGSB_ERR_INVALID, "already a group member: vkgsb_group_leave first");
  void* bases[kMaxGroup] = {nullptr};
  bool ipc[kMaxGroup] = {false};
  for (uint32_t m = 0; m < world; ++m) {
    if (m == rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + 64 * m, 64);
    cudaError_t e = cudaIpcOpenMemHandle(&bases[m], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (uint32_t k = 0; k < m; ++k)
        if (bases[k]) cudaIpcCloseMemHandle(bases[k]);
      return fail(VKGSB_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    }
    ipc[m] = true;
  }
  return group_setup(r, rank, world, bases, ipc, edges);
}

int vkgsb_group_join_local(vkgsb_renderer* const* members, uint32_t world, const uint32_t* edges) {
  if (!members || !edges) return fail(VKGSB_ERR_INVALID, "null argument");
  if (world > static_cast<uint32_t>(kMaxGroup)) return fail(VKGSB_ERR_INVALID, "at most 16 members");
  for (uint32_t m = 0; m < world; ++m) {
    if (!members[m]) return fail(VKGSB_ERR_INVALID, "null member");
    if (members[m]->grouped) return fail(VKGSB_ERR_INVALID, "already a group member");
    if (members[m]->max_splats != members[0]->max_splats) return fail(VKGSB_ERR_INVALID, "members must share max_splats");
  }
  void* bases[kMaxGroup] = {nullptr};
  bool ipc[kMaxGroup] = {false};
  for (uint32_t m = 0; m < world; ++m) bases[m] = members[m]->group_block;
  for (uint32_t m = 0; m < world; ++m) {
    if (set_device(members[m])) return VKGSB_ERR_CUDA;
    if (int e = group_setup(members[m], m, world, bases, ipc, edges)) return e;
  }
  return VKGSB_OK;
}

int vkgsb_group_leave(vkgsb_renderer* r) {
  if (!r) return fail(VKGSB_ERR_INVALID, "renderer is null");
  if (!r->grouped) return VKGSB_OK;
  if (set_device(r)) return VKGSB_ERR_CUDA;
  std::lock_guard<std::mutex> g(r->draw_mutex);
  CU_TRY(cudaDeviceSynchronize());
  for (int m = 0; m < kMaxGroup; ++m) {
    if (r->group_peer_ipc[m] && r->group_peer_base[m]) cudaIpcCloseMemHandle(r->group_peer_base[m]);
    r->group_peer_base[m] = nullptr;
    r->group_peer_ipc[m] = false;
  }
  r->grouped = false;
  r->band_y0 = r->band_y1 = 0;
  invalidate_graph(r);
  return VKGSB_OK;
}

// ---- cross-process destinations on one node --------------------------------------------------------------
int vkgsb_shared_create(int device, size_t bytes, void** d_ptr, uint8_t handle[64]) {
  if (!d_ptr || !handle || bytes == 0) return fail(VKGSB_ERR_INVALID, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  CU_TRY(cudaSetDevice(device));
  void* p = nullptr;
  CU_TRY(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(VKGSB_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  std::memcpy(handle, &h, 64);
  *d_ptr = p;
  return VKGSB_OK;
}

int vkgsb_shared_open(int device, const uint8_t handle[64], void** d_ptr) {
  if (!d_ptr || !handle) return fail(VKGSB_ERR_INVALID, "null argument");
  CU_TRY(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  CU_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return VKGSB_OK;
}

int vkgsb_shared_close(int device, void* d_ptr) {
  if (!d_ptr) return VKGSB_OK;
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaIpcCloseMemHandle(d_ptr));
  return VKGSB_OK;
}

int vkgsb_shared_read(int device, const void* d_ptr, size_t offset, size_t bytes, void* host_dst) {
  if (!d_ptr || !host_dst) return fail(VKGSB_ERR_INVALID, "null argument");
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaDeviceSynchronize());
  CU_TRY(cudaMemcpy(host_dst, static_cast<const uint8_t*>(d_ptr) + offset, bytes, cudaMemcpyDeviceToHost));
  return VKGSB_OK;
}

int vkgsb_shared_destroy(int device, void* d_ptr) {
  if (!d_ptr) return VKGSB_OK;
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaFree(d_ptr));
  return VKGSB_OK;
}

// ---- stage-level sort plug-in (vrdx* surface) -----------------------------------------------------------------------
// storage layout: [hist 4x256 u32][tickets 4 u32 (+pad)][tree of partition aggregates + arrival counters,
// sort_lookback_bytes(max_n)][keys_alt max_n][vals_alt max_n]
static size_t sort_storage_layout(uint32_t max_n, size_t* off_lookback, size_t* off_keys, size_t* off_vals) {
  size_t o = 4 * 256 * 4 + 64;
  *off_lookback = o;
  o += sort_lookback_bytes(max_n);
  o = (o + 255) & ~size_t(255);
  *off_keys = o;
  o += (static_cast<size_t>(max_n) * 4 + 255) & ~size_t(255);
  *off_vals = o;
  o += (static_cast<size_t>(max_n) * 4 + 255) & ~size_t(255);
  return o;
}

int vkgsb_sort_storage_bytes(uint32_t max_element_count, size_t* bytes) {
  if (!bytes) return fail(VKGSB_ERR_INVALID, "bytes is null");
  size_t a, b, c;
  *bytes = sort_storage_layout(max_element_count, &a, &b, &c);
  return VKGSB_OK;
}

int vkgsb_sort_key_value_indirect(void* stream, uint32_t max_element_count, const uint32_t* d_count, uint32_t* d_keys,
                                  uint32_t* d_values, void* d_storage) {
  if (max_element_count == 0) return VKGSB_OK;  // nothing to sort; pointers are not inspected
  if (!d_count || !d_keys || !d_values || !d_storage) return fail(VKGSB_ERR_INVALID, "null device pointer");
  if (max_element_count >= (1u << 30)) return fail(VKGSB_ERR_CAPACITY, "fewer than 2^30 elements");
  if ((reinterpret_cast<uintptr_t>(d_keys) & 15u) || (reinterpret_cast<uintptr_t>(d_values) & 3u) ||
      (reinterpret_cast<uintptr_t>(d_storage) & 255u))
    return fail(VKGSB_ERR_INVALID, "d_keys must be 16-byte aligned (it is read as uint4), d_values 4-byte, d_storage 256-byte");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  size_t off_lb, off_k, off_v;
  sort_storage_layout(max_element_count, &off_lb, &off_k, &off_v);
  uint8_t* base = static_cast<uint8_t*>(d_storage);
  CU_TRY(cudaMemsetAsync(base, 0, 4 * 256 * 4 + 64, s));
  SortArgs a{};
  a.d_count = d_count;
  a.max_n = max_element_count;
  a.keys = d_keys; a.vals = d_values;
  a.keys_alt = reinterpret_cast<uint32_t*>(base + off_k);
  a.vals_alt = reinterpret_cast<uint32_t*>(base + off_v);
  a.hist = reinterpret_cast<uint32_t*>(base);
  a.tickets = reinterpret_cast<uint32_t*>(base + 4 * 256 * 4);
  a.lookback = reinterpret_cast<uint32_t*>(base + off_lb);
  a.begin_bit = 0;
  a.npass = 4;
  launch_sort(a, s);
  CU_TRY(cudaGetLastError());
  return VKGSB_OK;
}

}  // extern "C"
