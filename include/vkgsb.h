/*
 * vkgsb.h — C ABI of the B200-native vkgs renderer path (libvkgsb.so).
 *
 * Drop-in boundary for the reference's per-frame hot path: everything below is what a binding of
 * jaesung-cs/vkgs would call instead of recording Vulkan commands in Engine::Impl::Draw()
 * (src/vkgs/engine/engine.cc:742-1378).  Plain pointers and sizes only; no C++/torch types.
 * All functions return VKGSB_OK (0) or an error code; vkgsb_last_error() gives the text of the
 * calling thread's last failure.  No exception crosses this boundary.  There is no CPU fallback:
 * without a CUDA device vkgsb_create fails with VKGSB_ERR_CUDA.
 *
 * Threading (same contract as the reference): one renderer = one CUDA device + one
 * stream; draw/set_* are not re-entrant per renderer; load_ply_async / load_progress / cancel_load
 * may be called from any thread.  Frames share the renderer's work buffers: a frame drawn on another stream
 * than the previous one is ordered behind it, and loads / vkgsb_set_lines / vkgsb_destroy drain every frame
 * issued, whichever stream it ran on.
 *
 * Matrices are column-major float[16] (m[c*4+r]) exactly as glm / the reference's UBO
 * (src/vkgs/vulkan/shader/uniforms.h:10-15).
 */
#ifndef VKGSB_H_
#define VKGSB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VKGSB_API __declspec(dllexport)
#else
#define VKGSB_API __attribute__((visibility("default")))
#endif

typedef struct vkgsb_renderer vkgsb_renderer;

enum vkgsb_status {
  VKGSB_OK = 0,
  VKGSB_ERR_INVALID = 1,  /* bad argument */
  VKGSB_ERR_CUDA = 2,     /* CUDA runtime / no device */
  VKGSB_ERR_IO = 3,       /* file open / read / malformed PLY */
  VKGSB_ERR_CAPACITY = 4, /* more splats than max_splats, image larger than allocated */
  VKGSB_ERR_NO_SCENE = 5, /* draw before any splats were loaded */
  VKGSB_ERR_CANCELLED = 6 /* load cancelled */
};

/* How destination colour is accumulated (SURVEY.md §7 hard part 1). */
enum vkgsb_blend_mode {
  VKGSB_BLEND_FP32 = 0,  /* fp32 accumulation, one UNORM8 quantisation at the end; front-to-back with early exit */
  VKGSB_BLEND_UNORM8 = 1 /* the reference's render target (render_pass.cc:15): destination re-quantised to
                            UNORM8 after every splat, back-to-front; the walk starts where the splats in front already
                            hide everything behind, certified per pixel to within 1/255 of the full walk */
};

enum vkgsb_pixel_format {
  VKGSB_FORMAT_RGBA8 = 0,
  VKGSB_FORMAT_BGRA8 = 1 /* memory order of the reference's swapchain image (swapchain.cc:18) */
};

typedef struct vkgsb_config {
  uint32_t struct_size; /* = sizeof(vkgsb_config) */
  int32_t device;       /* CUDA device ordinal */
  uint32_t max_splats;  /* storage is pre-allocated once, like the reference's MAX_SPLAT_COUNT (engine.cc:1653);
                           0 => 1<<23 */
  uint32_t max_width;   /* 0 => 3840 */
  uint32_t max_height;  /* 0 => 2160 */
  uint64_t max_pairs;   /* capacity of the (bin, splat) binning list; 0 => 8 * max_splats */
} vkgsb_config;

/* The per-frame parameter block: shader::Camera (uniforms.h:10-15) + the `mat4 model` push constant
 * (engine.cc:1024-1025,1185-1186). */
typedef struct vkgsb_camera {
  float projection[16];
  float view[16];
  float camera_position[3];
  float pad0;
  float model[16];
} vkgsb_camera;

/* Mirrors FrameInfo (engine.cc:1620-1636).  Stage times are only filled when stage timing is enabled. */
typedef struct vkgsb_stats {
  uint32_t total_point_count;
  uint32_t loaded_point_count;
  uint32_t visible_point_count; /* V of the last drawn frame */
  uint32_t pair_overflow;       /* 1 if the binning list hit max_pairs (farthest splats were dropped) */
  uint64_t pair_count;          /* (bin, splat) pairs of the last frame */
  float ms_project;             /* rank.comp + projection.comp equivalent */
  float ms_sort;                /* vrdxCmdSortKeyValueIndirect equivalent */
  float ms_bin;                 /* tile binning (no reference equivalent: replaces the HW rasteriser's setup) */
  float ms_blend;               /* splat.vert/.frag + ROP equivalent */
  float ms_total;
  uint64_t frame_counter;
  uint32_t blend_retries;       /* VKGSB_BLEND_UNORM8: warp (16x8 pixels) attempts whose late-start bracket was still more
                                   than 1 level wide at the front and that started again from deeper (a cost indicator) */
  uint32_t pad0;
  uint64_t fragment_count;      /* fragments (pixel x splat pairs inside the +-3 sigma quad) the blend stage shaded in the
                                   last frame; only counted while VKGSB_OPT_COUNT_FRAGMENTS is on */
  float ms_cull;                /* the part of ms_project spent in the cull kernel (rank.comp equivalent) */
  uint32_t pad1;
} vkgsb_stats;

enum vkgsb_option {
  VKGSB_OPT_STAGE_TIMING = 0, /* 1: launch stages eagerly with CUDA events between them; 0: one CUDA graph per frame */
  VKGSB_OPT_BLEND_MODE = 1,   /* vkgsb_blend_mode */
  VKGSB_OPT_PIXEL_FORMAT = 2, /* vkgsb_pixel_format */
  VKGSB_OPT_BAND_Y0 = 3,      /* render image rows [y0,y1) only (tile-band sharding): the rows of the band are
                                 bit-identical to the same rows of the full frame */
  VKGSB_OPT_BAND_Y1 = 4,      /* 0 => full height */
  VKGSB_OPT_KEEP_INSTANCES = 5, /* 1: also store the reference-format instance records (projection.comp:177-179) so
                                  vkgsb_read_instances can return them; off by default (48 B/visible splat of writes) */
  VKGSB_OPT_BAND_CULL = 6,    /* default 1: while a band is set, splats whose footprint provably cannot reach it are
                                 dropped at the cull, so sort, projection and binning shrink with the band; the visible
                                 count and the parity taps then describe the band's subset (same relative order).
                                 0: cull on the centre alone, as the full frame does (rank.comp:37) */
  VKGSB_OPT_COUNT_FRAGMENTS = 7, /* 1: the blend stage also counts the fragments it shades (vkgsb_stats.fragment_count): the
                                 stage's work unit for fragments/s figures; costs a few percent, off by default */
  VKGSB_OPT_UNORM8_CUT_EXP = 8, /* k in [1, 18], default 4: VKGSB_BLEND_UNORM8 starts its back-to-front walk where the
                                 transmittance of the splats in front is below 10^-k (deeper = fewer retries, longer
                                 walks); the result is certified either way */
  VKGSB_OPT_L2_PIN_MB = 9,    /* default 72: the centres (12 B/splat) of the first that-many MB of splats are kept in the
                                 126 MB L2 across frames (evict_last), so the centre gathers of a scene of up to ~6 M
                                 splats are L2 hits from the second frame on; 0 disables */
  VKGSB_OPT_SPATIAL_ORDER = 10 /* default 1: a load stores the scene in Morton order of the splat centres (the reference
                                 keeps the file's order, splat_load_thread.cc:138-159), so that the cull decides whole tiles
                                 of 256 splats from their bounding box and the visible splats' payload is read in runs.
                                 Pixels are the same; equal depth keys - which the reference orders by a race,
                                 rank.comp:38 - are ordered by stored index.  vkgsb_read_order maps stored index -> index in
                                 the file.  0: keep the file's order.  Applies to the next load. */
};

VKGSB_API const char* vkgsb_last_error(void);
VKGSB_API int vkgsb_device_count(int* count);

/* Engine::Engine() (engine.cc:114-524): allocate every buffer once. */
VKGSB_API int vkgsb_create(int device, uint32_t max_splats, vkgsb_renderer** out);
VKGSB_API int vkgsb_create_ex(const vkgsb_config* cfg, vkgsb_renderer** out);
VKGSB_API void vkgsb_destroy(vkgsb_renderer* r);
VKGSB_API int vkgsb_set_option(vkgsb_renderer* r, int option, int64_t value);

/* Engine::LoadSplats (engine.cc:541-544) + SplatLoadThread (splat_load_thread.cc:55-207) + parse_ply.comp.
 * vkgsb_load_ply blocks until the scene is resident; _async returns at once, cancelling a load in flight. */
VKGSB_API int vkgsb_load_ply(vkgsb_renderer* r, const char* path);
VKGSB_API int vkgsb_load_ply_async(vkgsb_renderer* r, const char* path);
/* SplatLoadThread::GetProgress (splat_load_thread.h:35-39).  state: 0 idle, 1 loading, 2 done, <0 -vkgsb_status. */
VKGSB_API int vkgsb_load_progress(vkgsb_renderer* r, uint32_t* total, uint32_t* loaded, int* state);
VKGSB_API int vkgsb_cancel_load(vkgsb_renderer* r);
VKGSB_API int vkgsb_wait_load(vkgsb_renderer* r);

/* Same ingest without a file: `rows` = n PLY vertices as floats (host memory), `offsets` = the 60-entry
 * float-offset table of splat_load_thread.cc:114-135 (offsets[59] = row stride in floats). */
VKGSB_API int vkgsb_upload_splats(vkgsb_renderer* r, uint32_t n, const float* rows, const uint32_t offsets[60]);

/* Write the camera UBO + model push constant (engine.cc:1047-1050,1185-1186) and the viewport (engine.cc:1419-1431). */
VKGSB_API int vkgsb_set_camera(vkgsb_renderer* r, const vkgsb_camera* cam);
VKGSB_API int vkgsb_set_viewport(vkgsb_renderer* r, uint32_t width, uint32_t height);

/* Opaque line layer under the splats: the reference's axis and grid (engine.cc:618-680 geometry, drawn at
 * engine.cc:1440-1469 through color.vert/.frag with a LINE_LIST pipeline that tests AND writes depth,
 * engine.cc:398-415).  The lines are drawn first; the splats are then depth-tested LESS against them without writing
 * depth (engine.cc:298-299) - the reason the reference keeps a graphics pipeline at all (DETAILS.md:7).
 * positions: 2 * n_lines xyz; colors: 2 * n_lines straight-alpha rgba (color.frag premultiplies); model: the lines' own
 * push constant, column-major (engine.cc:1444-1448 uses diag(10, 10, 10, 1)); NULL = identity.  All host pointers,
 * copied.  n_lines = 0 removes the layer (the default: frames then cost exactly what they did without it). */
VKGSB_API int vkgsb_set_lines(vkgsb_renderer* r, uint32_t n_lines, const float* positions, const float* colors,
                              const float model[16]);

/* Splat centres per image row of the LAST frame drawn (rows[height], host memory): the load a screen-band partition
 * balances its band edges on (vkgs_b200.dist.balanced_band_edges).  No counterpart in the reference. */
VKGSB_API int vkgsb_row_histogram(vkgsb_renderer* r, uint32_t* rows, uint32_t capacity);

/* One frame: rank -> sort -> projection -> draw (engine.cc:1164-1290), into an RGBA8/BGRA8 image of
 * width*height*4 bytes.  dst may be NULL (image stays in the renderer, see vkgsb_image_device_ptr), a host pointer
 * (dst_is_device = 0: device->host copy, returns when the pixels are in dst) or a device pointer (dst_is_device = 1:
 * asynchronous on `stream`; the blend stage writes its pixels straight into dst - no copy - and while a band is set
 * only the band's rows of dst are written).  dst may live on another GPU of the node (vkgsb_shared_open): the pixel
 * stores then travel over NVLink.  stream = a cudaStream_t cast to void*, NULL = the renderer's own stream. */
VKGSB_API int vkgsb_draw(vkgsb_renderer* r, void* dst, int dst_is_device, void* stream);
/* n_views frames with one scene: cameras[i] -> dst + i*stride bytes (same dst rules). */
VKGSB_API int vkgsb_draw_batch(vkgsb_renderer* r, uint32_t n_views, const vkgsb_camera* cameras, void* dst,
                               size_t dst_stride, int dst_is_device, void* stream);
VKGSB_API int vkgsb_image_device_ptr(vkgsb_renderer* r, void** ptr);
VKGSB_API int vkgsb_sync(vkgsb_renderer* r);
/* Frame pacing, the reference's fence wait (engine.cc:1028-1035: frame i waits for frame i - 2): blocks until the frame
 * issued `frames_back` (1..3) frames before the NEXT one has finished on the device.  A render loop that calls this with
 * 2 before every vkgsb_draw(dst = NULL) keeps two frames in flight instead of filling the launch queue. */
VKGSB_API int vkgsb_wait_frame(vkgsb_renderer* r, uint32_t frames_back);
VKGSB_API int vkgsb_get_stats(vkgsb_renderer* r, vkgsb_stats* out);

/* Destinations shared between the processes of one node (one process per GPU): the consumer process
 * allocates a device buffer and exports a 64-byte handle (CUDA IPC); each producer process maps it and passes
 * mapped pointer + offset as vkgsb_draw's device destination, so a finished view - or one screen band's rows of a
 * frame - lands in the consumer GPU's memory without a copy kernel, a staging buffer or a collective.  The reference is
 * single-GPU (context.cc:94-132); its interop pattern for handing images to another API is external memory by file
 * descriptor (interop/cuda_image.cu:77-132).  Ordering is the caller's: a producer's frame is complete when its stream
 * is, and processes meet with whatever barrier they already have. */
VKGSB_API int vkgsb_shared_create(int device, size_t bytes, void** d_ptr, uint8_t handle[64]);
VKGSB_API int vkgsb_shared_open(int device, const uint8_t handle[64], void** d_ptr);
VKGSB_API int vkgsb_shared_close(int device, void* d_ptr);
/* consumer side without a CUDA runtime of its own: bytes [offset, offset + bytes) of the buffer -> host memory */
VKGSB_API int vkgsb_shared_read(int device, const void* d_ptr, size_t offset, size_t bytes, void* host_dst);
VKGSB_API int vkgsb_shared_destroy(int device, void* d_ptr);

/* Band group (BASELINE configs[4]): W renderers - one per GPU, normally one per process - draw the W screen
 * bands of the same frames; member g draws rows [edges[g], edges[g + 1]) (VKGSB_OPT_BAND_Y0 / Y1 are set by the join).
 * The cull, which every member would otherwise repeat over the whole scene, is shared out: member j tests the splats of
 * its 1/W share against the frustum and against every band's footprint bound and writes each band's visibility bits
 * straight into that band's member over NVLink (peer mappings of one allocation per renderer, CUDA IPC); members
 * hand-shake through frame-numbered flags in each other's memory, no collective and no host round trip.  Contract: all
 * members hold the same scene and were created with the same max_splats; every member issues the same sequence of frames
 * (same camera, viewport, options) after the join; a member that does not makes the others' frames give up after ~2 s,
 * reported by vkgsb_sync.  Frames are bit-identical to the same rows of the ungrouped full frame.
 * vkgsb_group_export: this renderer's 64-byte handle; vkgsb_group_join: handles = world x 64 bytes in rank order (the own
 * entry is ignored), edges = world + 1 rows; vkgsb_group_join_local: the members live in this process (tests, one
 * process driving several GPUs). */
VKGSB_API int vkgsb_group_export(vkgsb_renderer* r, uint8_t handle[64]);
VKGSB_API int vkgsb_group_join(vkgsb_renderer* r, uint32_t rank, uint32_t world, const uint8_t* handles,
                               const uint32_t* edges);
VKGSB_API int vkgsb_group_join_local(vkgsb_renderer* const* members, uint32_t world, const uint32_t* edges);
VKGSB_API int vkgsb_group_leave(vkgsb_renderer* r);

/* External memory and semaphores: the present path of a desktop viewer (the reference's interop pattern,
 * src/vkgs/engine/interop/cuda_image.cu:77-132 and cuda_semaphore.cu:56-86, device extensions at
 * src/vkgs/vulkan/context.cc:204-216,240-241).  The application creates its VkImage with exportable memory, takes the
 * memory's file descriptor (vkGetMemoryFdKHR) and hands it over; the returned device pointer is a frame destination for
 * vkgsb_draw(dst, dst_is_device = 1), so the blend kernel writes the frame straight into the image's memory (the
 * reference measured its copy-based variant at 1 ms per 1600x900 frame, DETAILS.md:43).  On success an OPAQUE_FD belongs
 * to CUDA and must not be closed by the caller.  VKGSB_EXTERNAL_CUDA_POSIX_FD is the same contract for an allocation
 * exported by CUDA's own virtual-memory API (cuMemExportToShareableHandle) - another process's CUDA allocation, or the
 * stand-in for the VkImage's memory where no Vulkan exists (tests); vkgsb_external_alloc creates such an allocation and
 * returns its fd (the caller owns and closes the fd). */
enum vkgsb_external_handle_type {
  VKGSB_EXTERNAL_OPAQUE_FD = 0,     /* VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT (cudaExternalMemoryHandleTypeOpaqueFd) */
  VKGSB_EXTERNAL_CUDA_POSIX_FD = 1  /* CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR */
};
typedef struct vkgsb_external vkgsb_external;
VKGSB_API int vkgsb_external_import(int device, int fd, size_t bytes, int handle_type, vkgsb_external** out, void** d_ptr);
VKGSB_API int vkgsb_external_alloc(int device, size_t bytes, vkgsb_external** out, int* fd, void** d_ptr);
VKGSB_API int vkgsb_external_release(vkgsb_external* ext);
/* A VkSemaphore exported as an opaque fd (vkGetSemaphoreFdKHR): signalled on `stream` behind a frame so that the queue
 * presenting the image waits for it; waited on before a frame overwrites an image the other API still reads. */
VKGSB_API int vkgsb_external_semaphore_import(int device, int fd, void** sem);
VKGSB_API int vkgsb_external_semaphore_signal(void* sem, void* stream);
VKGSB_API int vkgsb_external_semaphore_wait(void* sem, void* stream);
VKGSB_API int vkgsb_external_semaphore_release(void* sem);

/* Parity taps (test / debugging): state of the last drawn frame, copied to host.
 * read_sorted: keys/ids in sorted (far -> near) order = SplatStorage.key / .index after vrdx (engine.cc:1218-1219).
 * read_instances: 12 floats per visible splat in sorted order = SplatStorage.instance (projection.comp:177-179).
 * read_scene: the activated scene in the reference layout (engine.cc:1639-1651): pos[n*3], cov[n*6],
 *             opacity[n], sh[n*48] (IEEE half bits), in STORED order (the ids of read_sorted index it). Any pointer may
 *             be NULL.
 * read_order: order[i] = index in the loaded file / uploaded rows of stored splat i (VKGSB_OPT_SPATIAL_ORDER; the
 *             identity when that option is off). */
VKGSB_API int vkgsb_read_sorted(vkgsb_renderer* r, uint32_t* keys, uint32_t* ids, uint32_t capacity, uint32_t* count);
VKGSB_API int vkgsb_read_instances(vkgsb_renderer* r, float* inst, uint32_t capacity, uint32_t* count);
VKGSB_API int vkgsb_read_scene(vkgsb_renderer* r, float* pos, float* cov, float* opacity, uint16_t* sh,
                               uint32_t capacity, uint32_t* count);
VKGSB_API int vkgsb_read_order(vkgsb_renderer* r, uint32_t* order, uint32_t capacity, uint32_t* count);

/* Stage-level plug-in for the sort alone, the vrdx* surface of third_party/vulkan_radix_sort
 * (include/vk_radix_sort.h:19-76): ascending, stable, in place, element count read ON THE DEVICE from d_count
 * (vrdxCmdSortKeyValueIndirect), caller-owned storage from vkgsb_sort_storage_bytes
 * (vrdxGetSorterKeyValueStorageRequirements).  All pointers are device pointers; asynchronous on `stream`. */
VKGSB_API int vkgsb_sort_storage_bytes(uint32_t max_element_count, size_t* bytes);
/* d_keys must be 16-byte aligned, d_storage 256-byte (cudaMalloc alignment); fewer than 2^30 elements. */
VKGSB_API int vkgsb_sort_key_value_indirect(void* stream, uint32_t max_element_count, const uint32_t* d_count,
                                            uint32_t* d_keys, uint32_t* d_values, void* d_storage);

/* Host helper: vkgs::Camera (camera.cc:25-45) for an orbit pose; fills projection/view/camera_position, model = I. */
VKGSB_API int vkgsb_camera_orbit(uint32_t width, uint32_t height, float fovy, float r, float phi, float theta,
                                 const float center[3], vkgsb_camera* out);

/* Host helper: the viewer's mouse operations (camera.cc:47-70) applied to a default vkgs::Camera in the order Rotate(rot_x,
 * rot_y), Zoom(zoom), SetFov(fov) if fov > 0, Translate(tx, ty, tz), DollyZoom(dolly) if dolly != 0. */
VKGSB_API int vkgsb_camera_apply(uint32_t width, uint32_t height, float rot_x, float rot_y, float zoom, float fov,
                                 float tx, float ty, float tz, float dolly, vkgsb_camera* out);

#ifdef __cplusplus
}
#endif
#endif /* VKGSB_H_ */
