/*
 * woxel_b200.h -- C ABI of libwoxel_b200.so: woxel's per-pixel VDB345 HDDA+SDF raycast on B200.
 *
 * The reference (NemoInfo/woxel; citations are path:line relative to its root) has no FFI seam
 * for this path; its boundary is the wgpu compute pass.  Every entry point below replaces the
 * reference interface named beside it, so a maintainer swaps the wgpu plumbing for these calls
 * (see INTEGRATION.md for the Rust `extern "C"` block and build.rs).
 *
 *   reference interface                                                     replaced by
 *   ----------------------------------------------------------------------  ------------------------
 *   WgpuContext::new: instance/adapter/device   src/render/wgpu_context.rs:33-99      wx_init
 *   model upload: vdb.atlas() + write_texture x3, MaskUniform::from(&vdb).bind
 *       + write_buffer x6                        src/render/wgpu_context.rs:101-159,
 *                                                :506-573 ; src/vdb/vdb345.rs:108-264  wx_tree_upload
 *   ComputeState uniform (256 B)                 src/render/gpu_types/compute_state.rs:9-29
 *                                                src/shaders/raycast.comp.wgsl:1-23    WxState
 *   compute pass: set_bind_group x4 + dispatch_workgroups(W/8, H/4, 1) of cp_main
 *                                                src/render/wgpu_context.rs:270-282
 *                                                src/shaders/raycast.comp.wgsl:60-68   wx_render / wx_render_device
 *   frame capture copy_texture_to_buffer + map   src/render/wgpu_context.rs:374-405    wx_render (host rgba_out)
 *   drop(WgpuContext) / atlas_group, masks_group src/render/wgpu_context.rs:16-30      wx_tree_free / wx_shutdown
 *
 * Conventions: plain pointers and sizes only; every function returns WX_OK (0) or a negative
 * WxStatus, never unwinds, never aborts.  A context is used from one thread at a time (the
 * reference renders on the winit thread, src/runtime.rs:97-109).  There is NO CPU fallback:
 * without a CUDA device wx_init fails with WX_ERR_NO_DEVICE.
 */
#ifndef WOXEL_B200_H
#define WOXEL_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WX_ABI_VERSION 2

typedef enum WxStatus {
  WX_OK = 0,
  WX_ERR_INVALID_ARGUMENT = -1,
  WX_ERR_NO_DEVICE = -2,     /* no CUDA device / driver: the path does not run on the CPU */
  WX_ERR_CUDA = -3,          /* a CUDA runtime call failed; wx_last_error() has the text */
  WX_ERR_OUT_OF_MEMORY = -4,
  WX_ERR_BAD_TREE = -5,      /* child index out of range, masks inconsistent with counts */
  WX_ERR_UNSUPPORTED = -6,   /* e.g. an SDF distance >= 2^31 */
  WX_ERR_PEER_ACCESS = -7
} WxStatus;

typedef struct WxContext WxContext;
typedef struct WxTree WxTree;

/*
 * Host-side description of a VDB345 tree in the reference's node order
 * (N5s by ascending origin [x,y,z]; children by ascending offset; N4 / N3 indices are running
 * counters over that DFS -- src/vdb/vdb345.rs:134-158, :186-261).  It carries exactly the
 * information of vdb.origins() + vdb.masks() + vdb.atlas() after vdb.compute_sdf(), minus the
 * atlas padding.  The library copies and re-packs it during wx_tree_upload.
 *
 *   offsets: N5 ((x&4095)>>7)<<10 | ((y&4095)>>7)<<5 | (z&4095)>>7 ; N4 ((x&127)>>3)<<8 | ((y&127)>>3)<<4 | (z&127)>>3 ;
 *            N3 (x&7)<<6 | (y&7)<<3 | (z&7)                        (src/vdb/data_structure.rs:58-70)
 *   masks:   u64 word k holds offsets 64k..64k+63, bit (o & 63)   (src/vdb/data_structure.rs:102-108)
 *   tabN:    per slot the u32 the reference stores in the atlas texel: child index where the child
 *            bit is set, else the SDF distance in cells of that level (ignored where the value bit
 *            is set: raycast.comp.wgsl:431-433, :463-465, :490-492).
 */
typedef struct WxTreeDesc {
  uint32_t n5, n4, n3;
  const int32_t *origins;  /* n5 x 3 */
  const uint64_t *kids5;   /* n5 x 512 */
  const uint64_t *vals5;   /* n5 x 512 */
  const uint32_t *tab5;    /* n5 x 32768 */
  const uint64_t *kids4;   /* n4 x 64 */
  const uint64_t *vals4;   /* n4 x 64 */
  const uint32_t *tab4;    /* n4 x 4096 */
  const uint64_t *vals3;   /* n3 x 8 */
  const void *tab3;        /* n3 x 512 SDF distances, element size tab3_elem_bytes */
  uint32_t tab3_elem_bytes; /* 4 (the reference's u32 texels) or 1 (pre-narrowed, all < 256) */
  uint32_t reserved;
} WxTreeDesc;

/* == ComputeState (src/render/gpu_types/compute_state.rs:9-29), 256 bytes, std140-compatible. */
typedef struct WxState {
  float view_proj[16];       /*   0 */
  float camera_to_world[16]; /*  64 */
  float eye[4];              /* 128 */
  float u[4];                /* 144 */
  float mv[4];               /* 160 */
  float wp[4];               /* 176 */
  uint32_t render_mode[4];   /* 192  [0]: 0 Gray 1 Rgb 2 Ray 3 Diffuse 4 Glossy (src/render/egui_dev.rs:11-18) */
  uint32_t show_345[4];      /* 208 */
  float sun_dir[4];          /* 224 */
  float sun_color[4];        /* 240  rgb + intensity */
} WxState;

/*
 * Optional per-pixel outputs beside the RGBA8 frame ("AOVs").  The reference shader exposes only
 * colour; these expose the HDDAout of the primary ray (raycast.comp.wgsl:128-142) so that hit
 * voxel, leaf index and depth can be compared.  Any member may be NULL.  Arrays are
 * n_states x H x W, in the same memory space as the rgba buffer of the call.
 */
typedef struct WxAov {
  uint8_t *state;  /* 0 hit, 1 out of bounds, 2 max steps */
  int32_t *voxel;  /* x3: vec3<i32>(floor(hit.p)) */
  int32_t *leaf;   /* reference N3 index (parents[2].idx) when the last lookup ended in a leaf, else -1 */
  uint8_t *level;  /* num_parents of the last lookup */
  uint32_t *iters; /* hit.i */
  float *depth;    /* |hit.p - eye| */
  uint8_t *mask;   /* bit0 x, bit1 y, bit2 z */
  float *pos;      /* x3: hit.p */
} WxAov;

/* Which rows of every frame this call renders: bands of `band_rows` rows, band b belongs to shard
 * (b % count); index/count = 0/1 renders everything.  band_rows must be a multiple of 8. */
typedef struct WxShard {
  uint32_t index, count, band_rows, reserved;
} WxShard;

typedef struct WxTreeInfo {
  uint32_t n5, n4, n3;
  uint32_t leaf_bits;        /* 4, 8 or 32: width of a packed leaf distance */
  uint64_t device_bytes;     /* per device */
  uint32_t max_dist[3];      /* max SDF distance seen in tab5 / tab4 / tab3 */
  uint32_t n_devices;
} WxTreeInfo;

typedef struct WxRenderInfo {
  float kernel_ms;          /* device time of the raycast kernel(s) of the last wx_render*, CUDA events */
  float total_ms;           /* device time of the whole call incl. copies (wx_render only) */
  uint64_t rays;            /* pixels dispatched */
  uint32_t launches;        /* kernels launched by the call */
  uint32_t reserved;
} WxRenderInfo;

int wx_abi_version(void);
const char *wx_strerror(int status);
const char *wx_last_error(const WxContext *ctx); /* detail text of the last failure on this context */

/* n_devices == 0: use the current device only.  device_ids may be NULL (0..n_devices-1).  A device id may be listed more
 * than once: every entry gets its own streams, frame and tree replica and takes its share of the row bands, which lets the
 * multi-device code paths be exercised on a single GPU. */
int wx_init(int n_devices, const int *device_ids, WxContext **out);
int wx_shutdown(WxContext *ctx);
int wx_device_count(const WxContext *ctx);

/* Re-packs `desc` into the device layout and replicates it read-only on every device of ctx. */
int wx_tree_upload(WxContext *ctx, const WxTreeDesc *desc, WxTree **out);
int wx_tree_free(WxContext *ctx, WxTree *tree);
int wx_tree_info(const WxTree *tree, WxTreeInfo *info);

typedef struct WxSdfInfo {
  uint32_t max_dist[3]; /* largest distance written to tab5 / tab4 / tab3 */
  uint32_t rounds;      /* relaxation rounds run over all levels and both passes */
  float device_ms;      /* device time of the sweeps (CUDA events), without the host<->device copies */
  float total_ms;       /* wall time of the call */
} WxSdfInfo;

/*
 * VDB345::compute_sdf (src/vdb/vdb345.rs:290-628) on the GPU: the step the reference runs on the CPU
 * between reading a model and uploading it (src/render/wgpu_context.rs:104, :513).  `topo` carries the
 * topology exactly as for wx_tree_upload (masks; tab5 / tab4 hold the child index where the child bit
 * is set, their other entries and tab3 are ignored).  On return tab5_out / tab4_out hold what
 * vdb.atlas() holds after vdb.compute_sdf(): the child index where the child bit is set, else the
 * distance; tab3_out holds the distance of every inactive voxel (0 where active), as u8
 * (tab3_elem_bytes == 1) or u32 (== 4).  The results equal the reference's sequential, order-dependent
 * sweep value for value.  WX_ERR_UNSUPPORTED: a leaf distance does not fit (retry with
 * tab3_elem_bytes = 4; distances >= 32768 are not supported on this path).  Out arrays are host memory.
 */
int wx_compute_sdf(WxContext *ctx, const WxTreeDesc *topo, uint32_t *tab5_out, uint32_t *tab4_out, void *tab3_out,
                   uint32_t tab3_elem_bytes, WxSdfInfo *info);

/*
 * Model load in one call: compute_sdf on the GPU and the device tables built from its result in place, replicated on
 * every device of ctx -- what `vdb.compute_sdf(); vdb.atlas(); MaskUniform::from(&vdb)` + the uploads do in the
 * reference (src/render/wgpu_context.rs:101-159, :506-573), without the distances ever visiting the host.  `topo` as for
 * wx_compute_sdf (plus vals5 / vals4).  The tree renders exactly like wx_compute_sdf + wx_tree_upload.
 * WX_ERR_UNSUPPORTED: a leaf distance above 255 (take the two-call path, which has the u32 brick layout).
 */
int wx_tree_build(WxContext *ctx, const WxTreeDesc *topo, WxTree **out, WxSdfInfo *info);

/*
 * One frame per state (n_states > 1 = camera batch).  Blocking.  rgba_out is HOST memory,
 * n_states x height x width x 4 bytes (pinned memory -- wx_host_alloc_pinned -- lets the copies overlap the rendering; pageable memory
 * works: every kernel is then enqueued before the first copy, because a copy into pageable memory blocks the host until it is done) -- or device memory
 * of ANY GPU of the process's unified address space (another GPU's buffer, also one opened with wx_ipc_open): the frame
 * is then delivered there by DMA, chunk by chunk while later chunks render; that is how one process per GPU gathers its
 * frames on GPU 0 over NVLink (bench.py).  aov_out is host memory.
 * Pixels outside the reference's dispatch (x >= (W/8)*8 or y >= (H/4)*4, wgpu_context.rs:281) are
 * written as 0,0,0,0 like a freshly created texture.  With several devices the frames are split
 * by row bands / cameras, rendered concurrently and gathered on device 0 over NVLink.
 */
int wx_render(WxContext *ctx, const WxTree *tree, const WxState *states, uint32_t n_states, uint32_t width,
              uint32_t height, uint8_t *rgba_out, const WxAov *aov_out);

/*
 * Same kernel, device-resident output, asynchronous on `stream` (a cudaStream_t, NULL = default
 * stream) of device `device_index` of the context.  `states` is host memory (copied into the
 * launch).  rgba_dev / aov_dev are device pointers valid on that device (they may be peer-mapped
 * memory of another GPU: the kernel's final store then IS the gather).  Only the rows selected by
 * `shard` (NULL = all) are written.
 */
int wx_render_device(WxContext *ctx, int device_index, const WxTree *tree, const WxState *states, uint32_t n_states,
                     uint32_t width, uint32_t height, uint8_t *rgba_dev, const WxAov *aov_dev, const WxShard *shard,
                     void *stream);

/*
 * One shard of a frame that is split over several processes, one per GPU (BASELINE config 4's tile partition; wx_render on a
 * multi-device context does the same inside one process).  Renders the rows `shard` selects (bands of 8 rows, band b belongs to
 * shard b % count) of every frame into the context's own memory and delivers exactly those rows to rgba_out, which has the
 * layout of the whole frame stack: host memory, or device memory of any GPU -- e.g. a frame on GPU 0 opened with
 * wx_ipc_open, the rows then travel over NVLink by DMA.  Blocking.  shard->band_rows must be 8.
 */
int wx_render_shard(WxContext *ctx, const WxTree *tree, const WxState *states, uint32_t n_states, uint32_t width,
                    uint32_t height, const WxShard *shard, uint8_t *rgba_out);

int wx_last_render_info(const WxContext *ctx, WxRenderInfo *info);

/*
 * Per-context options (ABI 2).  They replace the getenv() knobs round 1 read inside the library: a release build reads no
 * environment variable.  None of them changes a result except WX_OPT_MARCH.
 */
typedef enum WxOption {
  WX_OPT_MARCH = 1,         /* 0 (default): exact -- frames and AOVs bit-identical to the strict-f32 restatement of the shader.
                               1: tolerance mode -- the north-star bar (hit voxel + leaf equal on >= 99.9 % of the pixels, depth
                               within 1e-4 relative, RGB within 1/255) instead of bit identity: p += t * dir is one fused
                               multiply-add per axis.
                               2: 1 + rays start at their entry into the bounding box of the active cells.  Meets the bar on
                               scenes whose rays mostly hit or miss the box (the benchmark scenes: 99.98 %), NOT on sparse scenes
                               where many rays cross the box and leave (94-98 %: their out-of-bounds colour depends on the axis
                               of the LAST step, which changes with the start point).  Measurement only.
                               Render mode 2 (Ray: colours the iteration count) always takes the exact march. */
  WX_OPT_KERNEL = 2,        /* 0 (default): tiled grid; 1: persistent kernel, warp-level tile queue; 2: persistent kernel,
                               CTA-level chunk queue.  Same results; measured slower (profiles/). */
  WX_OPT_RENDER_CHUNKS = 3, /* row chunks of a pipelined wx_render (0 = automatic) */
  WX_OPT_SMEM_PAD = 4,      /* bytes of unused dynamic shared memory per CTA (lowers the resident CTAs per SM; measurement only) */
  WX_OPT_NVTX = 5,          /* 1: NVTX ranges around upload / sweep / render / read-back (default 0) */
  WX_OPT_LONG_FIRST = 6,    /* 1 (default): the tiles that held the longest rays in the previous launch of the same frame geometry
                               start first (a small high-priority kernel ahead of the main grid): shortens the drain at the end of
                               a launch.  Same pixels; 0 = plain launches */
  WX_OPT_LONG_THRESHOLD = 7 /* iterations of a primary ray from which its tile counts as long (default 96) */
} WxOption;
int wx_set_option(WxContext *ctx, int option, int64_t value);
int wx_get_option(const WxContext *ctx, int option, int64_t *value_out);

/*
 * Capture step after the raycast (replaces the copy_texture_to_buffer + recorder conversion of
 * src/render/wgpu_context.rs:374-405 and src/render/recorder.rs:20-37, :132-140): converts the frame(s)
 * the last wx_render left on device 0 from RGBA8 to RGB8, every colour byte passed through the
 * reference's linear_to_srgb, and copies them to rgb_out (host, n_states x height x width x 3 bytes).
 * The arguments must be those of that wx_render call.
 */
int wx_capture_srgb(WxContext *ctx, uint32_t n_states, uint32_t width, uint32_t height, uint8_t *rgb_out);
/* The 256-entry transfer table wx_capture_srgb applies (host arithmetic, no device needed). */
int wx_srgb_table(uint8_t table_out[256]);

/* Pure host arithmetic, no device needed: row_mask_out[y] = 1 iff `shard` renders row y of a frame of
 * `height` rows (the same band dealing wx_render / wx_render_device launch with).  The shards
 * 0..count-1 partition the rows.  Replaces nothing in the reference (it has one adapter,
 * src/render/wgpu_context.rs:46-53); it exists so that the multi-GPU split can be tested without GPUs. */
int wx_shard_rows(uint32_t height, const WxShard *shard, uint8_t *row_mask_out);

/* Device memory helpers so that a host language needs no CUDA runtime binding of its own. */
int wx_device_alloc(WxContext *ctx, int device_index, size_t bytes, void **out);
int wx_device_free(WxContext *ctx, int device_index, void *ptr);
int wx_host_alloc_pinned(size_t bytes, void **out);
int wx_host_free_pinned(void *ptr);
int wx_memcpy_d2h(WxContext *ctx, int device_index, void *dst_host, const void *src_dev, size_t bytes, void *stream);
int wx_stream_synchronize(WxContext *ctx, int device_index, void *stream);
/* Cross-process gather target: export a wx_device_alloc'd buffer / open it in another process. */
int wx_ipc_export(WxContext *ctx, int device_index, void *ptr, uint8_t handle_out[64]);
int wx_ipc_open(WxContext *ctx, int device_index, const uint8_t handle[64], void **out);
int wx_ipc_close(WxContext *ctx, int device_index, void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* WOXEL_B200_H */
