/*
 * woxel_host.h -- C view of the host-side surface (libwoxel_host.so) that sits ABOVE the
 * device ABI of woxel_b200.h: the reference's `src/vdb` (VDB345 tree, .vdb reader, compute_sdf,
 * GPU serialisation), `src/scene` (camera) and `src/render` (ComputeState, frame entry point).
 * The reference's host is Rust; no Rust toolchain exists in this environment, so the host is C++
 * (woxel_b200/host/) and this header is how Python (ctypes) and other languages reach it.
 * Names mirror the reference: each function cites the Rust item it stands for.
 */
#ifndef WOXEL_HOST_H
#define WOXEL_HOST_H
#include <stddef.h>
#include <stdint.h>

#include "woxel_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct WxhVdb WxhVdb;       /* VDB345<u32>            src/vdb/vdb345.rs:12   */
typedef struct WxhFlat WxhFlat;     /* origins()+masks()+atlas() flattened, src/vdb/vdb345.rs:108-264 */
typedef struct WxhRenderer WxhRenderer; /* WgpuContext       src/render/wgpu_context.rs:16-30 */

enum { /* negative statuses of this library; WX_* statuses pass through unchanged */
  WXH_ERR_IO = -101, WXH_ERR_MAGIC = -102, WXH_ERR_VERSION = -103, WXH_ERR_COMPRESSION = -104,
  WXH_ERR_GRID_NAME = -105, WXH_ERR_NODE_METADATA = -106, WXH_ERR_BLOSC = -107, WXH_ERR_UNSUPPORTED = -108,
  WXH_ERR_INVALID_ARGUMENT = -109
};
const char *wxh_last_error(void); /* thread-local text of the last failure */

/* Node trait index maths (src/vdb/data_structure.rs:43-91); level = 3, 4 or 5 */
int wxh_global_to_node(int level, const int32_t g[3], int32_t out[3]);
int64_t wxh_global_to_offset(int level, const int32_t g[3]);
int wxh_offset_to_child(int level, uint32_t offset, uint32_t out[3]);
int64_t wxh_child_to_offset(int level, const uint32_t c[3]);

/* VDB345 (src/vdb/vdb345.rs) */
WxhVdb *wxh_vdb_new(void);                                                     /* VDB::new            */
void wxh_vdb_free(WxhVdb *v);
void wxh_vdb_set_voxel(WxhVdb *v, int32_t x, int32_t y, int32_t z, uint32_t value); /* set_voxel :26     */
void wxh_vdb_set_voxels(WxhVdb *v, const int32_t *xyz, size_t n, uint32_t value);
/* get_voxel :69 -> kind 0 Offs 1 Leaf 2 Innr 3 Root 4 Bkgr ; level is 5/4 for Innr */
int wxh_vdb_get_voxel(const WxhVdb *v, int32_t x, int32_t y, int32_t z, uint32_t *value, int *level);
void wxh_vdb_count_nodes(const WxhVdb *v, uint64_t out[3]);                    /* count_nodes :266    */
uint64_t wxh_vdb_count_leaf_values(const WxhVdb *v);
void wxh_vdb_compute_sdf(WxhVdb *v);                                           /* compute_sdf :290    */

/* VdbReader (src/vdb/read.rs:55-141) */
typedef struct WxhVdbInfo {
  uint32_t file_version, library_major, library_minor, grid_count;
  uint32_t grid_compression;
  int32_t is_half_float;
  int64_t file_voxel_count; /* -1 if the grid has no such metadata */
  uint64_t grid_pos, block_pos, end_pos;
} WxhVdbInfo;
int wxh_vdb_read(const char *path, const char *grid_name, WxhVdb **out, WxhVdbInfo *info);
/* One c-blosc 1.x frame as the reader meets them inside a .vdb (src/vdb/read.rs:514-533 hands them to the c-blosc library):
 * BloscLZ, LZ4, Snappy, zlib and Zstd codecs, byte or bit shuffle.  *out_len receives the decoded size; WXH_ERR_INVALID_ARGUMENT with
 * *out_len set when `cap` is too small; WXH_ERR_BLOSC for a corrupt or unsupported frame. */
int wxh_blosc_decompress(const uint8_t *frame, size_t n, uint8_t *out, size_t cap, size_t *out_len);

/* to_flat: what replaces atlas()/masks()/origins() */
WxhFlat *wxh_vdb_to_flat(const WxhVdb *v, int narrow_leaves);
void wxh_flat_free(WxhFlat *f);
void wxh_flat_desc(const WxhFlat *f, WxTreeDesc *out); /* pointers stay valid until wxh_flat_free */

/* Procedural scenes of the benchmark configs; not in the reference */
WxhVdb *wxh_build_sphere(int32_t half, double radius, double band);
WxhVdb *wxh_build_torus(int32_t half, double major, double minor, double band);
WxhVdb *wxh_build_fog(int32_t half, double tau, double *occupancy);

/* ComputeState::build (src/render/gpu_types/compute_state.rs:87-131) + SunSettings::default (egui_dev.rs:355-367) */
int wxh_compute_state_build(const float eye[3], const float target[3], const float up[3], float aspect, float fovy_deg,
                            float resolution_width, uint32_t render_mode, const uint32_t show_grid[3],
                            const float sun_dir3[3], const float sun_color3[3], float sun_intensity, WxState *out);
void wxh_default_sun(float dir3[3], float color3[3], float *intensity);

/* Renderer = WgpuContext::new / change_vdb_model / render (src/render/wgpu_context.rs:33, :506, :207) */
int wxh_renderer_new(uint32_t width, uint32_t height, int n_devices, WxhRenderer **out);
void wxh_renderer_free(WxhRenderer *r);
int wxh_renderer_change_vdb_model(WxhRenderer *r, WxhVdb *v, int run_compute_sdf);
int wxh_renderer_change_vdb_model_file(WxhRenderer *r, const char *path, const char *grid_name);
int wxh_renderer_set_options(WxhRenderer *r, uint32_t render_mode, const uint32_t show_grid[3], const float sun_dir3[3],
                             const float sun_color3[3], float sun_intensity);
/* Scene.camera in, rgba8 frame (height x width x 4, host) out */
int wxh_renderer_render(WxhRenderer *r, const float eye[3], const float target[3], const float up[3], float aspect,
                        float fovy_deg, uint8_t *rgba_out);
/* compute_sdf of change_vdb_model on the GPU (default 1; wx_compute_sdf, identical values) or on the host (0) */
int wxh_renderer_set_sdf_on_gpu(WxhRenderer *r, int on);
void wxh_renderer_last_sdf(const WxhRenderer *r, WxSdfInfo *out); /* device_ms == 0: the host sweep ran */
/* fills the distances of a flat tree (taken before compute_sdf) on the GPU; WX_ERR_UNSUPPORTED: use wxh_vdb_compute_sdf */
int wxh_flat_compute_sdf_gpu(WxhFlat *f, WxContext *ctx, WxSdfInfo *info);
/* Frame dump after wx_capture_srgb (the reference's recorder pipes frames to ffmpeg, src/render/recorder.rs:67-105):
 * rgb is height x width x 3 bytes.  Binary PPM (P6), or PNG (8-bit RGB, sRGB chunk). */
int wxh_write_ppm(const char *path, const uint8_t *rgb, uint32_t width, uint32_t height);
int wxh_write_png(const char *path, const uint8_t *rgb, uint32_t width, uint32_t height);
WxContext *wxh_renderer_context(WxhRenderer *r);
WxTree *wxh_renderer_tree(WxhRenderer *r);

#ifdef __cplusplus
}
#endif
#endif
