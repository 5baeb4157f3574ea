/*
 * wxo.h -- CPU ORACLE for the woxel raycast hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This directory holds a plain-C restatement of the reference's algorithm for the one
 * hot path this repository accelerates (reference = NemoInfo/woxel, cited as path:line
 * relative to the reference root).  It is the CHECKER the CUDA product path is compared
 * against and the CPU baseline `bench.py` times beside it.  Nothing under woxel_b200/
 * (the product) may include, link, import or call anything in here; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY STATUS: the index maths, the .vdb reader topology and set/get_voxel are pinned
 * by the reference's own unit-test vectors (src/vdb/data_structure.rs:424-485,
 * src/vdb/vdb345.rs:703-723, src/vdb/read.rs:796-806).  The RAYCAST ITSELF IS
 * "PARITY UNPINNED": the reference ships no golden image / known-answer vector for
 * src/shaders/raycast.comp.wgsl and its toolchain (rustc, wgpu, naga, a Vulkan ICD)
 * is absent from this image, so the shader cannot be executed here.  The raycast oracle
 * is a line-by-line restatement of the WGSL in strict IEEE f32 without contraction.
 *
 * What is restated, and from where:
 *   tree / index maths      src/vdb/data_structure.rs:43-91, src/vdb/vdb345.rs:26-106
 *   compute_sdf             src/vdb/vdb345.rs:290-628
 *   origins/masks/atlas     src/vdb/vdb345.rs:108-264, :673-694 ; src/render/gpu_types/mask.rs:95-119
 *   .vdb reader             src/vdb/read.rs:62-349, :378-629
 *   ComputeState::build     src/render/gpu_types/compute_state.rs:87-131 ; src/render/camera.rs:31-35
 *                           (+ cgmath 0.18.0 look_at_rh / Matrix4::invert, restated from its published algorithm)
 *   raycast                 src/shaders/raycast.comp.wgsl:60-519
 */
#ifndef WXO_H
#define WXO_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct WxoTree WxoTree;       /* the reference's VDB345<u32> (pointer tree)            */
typedef struct WxoGpuData WxoGpuData; /* what vdb.origins()/masks()/atlas() hand to the shader  */

/* VdbEndpoint discriminants (src/vdb/data_structure.rs:337-344) */
enum { WXO_EP_OFFS = 0, WXO_EP_LEAF = 1, WXO_EP_INNR5 = 2, WXO_EP_INNR4 = 3, WXO_EP_ROOT = 4, WXO_EP_BKGR = 5 };

/* error codes of the reader (src/vdb/read.rs:31-53) */
enum {
  WXO_OK = 0,
  WXO_ERR_IO = -1,
  WXO_ERR_MAGIC = -2,
  WXO_ERR_VERSION = -3,
  WXO_ERR_COMPRESSION = -4,
  WXO_ERR_GRID_NAME = -5,
  WXO_ERR_NODE_METADATA = -6,
  WXO_ERR_BLOSC = -7,
  WXO_ERR_UNSUPPORTED = -8
};

/* ---- index maths (data_structure.rs:43-91); level = 3,4,5 ------------------------------ */
void wxo_global_to_node(int level, const int32_t g[3], int32_t out[3]);
uint32_t wxo_global_to_offset(int level, const int32_t g[3]);
void wxo_offset_to_child(int level, uint32_t offset, uint32_t out[3]);
uint32_t wxo_child_to_offset(int level, const uint32_t c[3]);

/* ---- tree (vdb345.rs:26-106, :266-287) -------------------------------------------------- */
WxoTree *wxo_tree_new(void);
void wxo_tree_free(WxoTree *t);
void wxo_set_voxel(WxoTree *t, int32_t x, int32_t y, int32_t z, uint32_t v);
void wxo_set_voxels(WxoTree *t, const int32_t *xyz, size_t n, uint32_t v);
/* returns a WXO_EP_* discriminant; *value receives the payload (dist / value / background) */
int wxo_get_voxel(const WxoTree *t, int32_t x, int32_t y, int32_t z, uint64_t *value);
void wxo_count_nodes(const WxoTree *t, uint64_t out[3]);
uint64_t wxo_count_leaf_values(const WxoTree *t); /* the count read.rs:772-794 asserts on */
void wxo_compute_sdf(WxoTree *t);                 /* vdb345.rs:290-628 */

/* Build a tree from reference-layout topology (origins sorted + DFS-ordered masks as u64 words).
 * Used to hand the oracle a scene that was generated elsewhere.  Leaf values are set to 1. */
WxoTree *wxo_tree_from_topology(uint32_t n5, uint32_t n4, uint32_t n3, const int32_t *origins /* n5*3 */,
                                const uint64_t *kids5, const uint64_t *vals5, const uint64_t *kids4,
                                const uint64_t *vals4, const uint64_t *vals3);

/* ---- .vdb reader (read.rs) --------------------------------------------------------------- */
typedef struct WxoVdbInfo {
  uint32_t file_version, library_major, library_minor, grid_count;
  uint32_t grid_compression; /* per-grid flags of the grid that was read */
  int32_t is_half_float;
  int64_t file_voxel_count; /* grid metadata "file_voxel_count", -1 if absent */
  uint64_t grid_pos, block_pos, end_pos;
  uint64_t topology_end_pos; /* stream position after the topology pass (== block_pos in a sane file) */
  uint32_t root_tiles, root_nodes;
} WxoVdbInfo;
int wxo_vdb_read(const char *path, const char *grid_name, WxoTree **out, WxoVdbInfo *info);

/* ---- reference GPU serialisation (vdb345.rs:108-264) -------------------------------------- */
WxoGpuData *wxo_serialise(const WxoTree *t);
/* Build the same structure directly from flat per-node tables (slot = child index or SDF distance). */
WxoGpuData *wxo_gpudata_from_tables(uint32_t n5, uint32_t n4, uint32_t n3, const int32_t *origins /* n5*3 */,
                                    const uint64_t *kids5, const uint64_t *vals5, const uint32_t *tab5,
                                    const uint64_t *kids4, const uint64_t *vals4, const uint32_t *tab4,
                                    const uint64_t *vals3, const uint32_t *tab3);
void wxo_gpudata_free(WxoGpuData *g);
void wxo_gpudata_counts(const WxoGpuData *g, uint32_t n[3], uint32_t atlas_dim[3]);
const int32_t *wxo_gpudata_origins(const WxoGpuData *g);                 /* n5 x 4 (x,y,z,0) */
const uint32_t *wxo_gpudata_mask(const WxoGpuData *g, int which);        /* 0 kids5 1 vals5 2 kids4 3 vals4 4 vals3 */
const uint32_t *wxo_gpudata_atlas(const WxoGpuData *g, int level_543);   /* 0:node5s 1:node4s 2:node3s, [x][y][z] */
/* Per-node tables (slot order = node offset) gathered back out of the atlases: tab5 n5*32768 etc. */
void wxo_gpudata_tables(const WxoGpuData *g, uint32_t *tab5, uint32_t *tab4, uint32_t *tab3);

/* ---- uniform (compute_state.rs:9-29; raycast.comp.wgsl:1-23); exactly 256 bytes ----------- */
typedef struct WxoState {
  float view_proj[16];
  float camera_to_world[16];
  float eye[4];
  float u[4], mv[4], wp[4];
  uint32_t render_mode[4];
  uint32_t show_345[4];
  float sun_dir[4];
  float sun_color[4];
} WxoState;

void wxo_compute_state_build(const float eye[3], const float target[3], const float up[3], float aspect,
                             float fovy_deg, float resolution_width, uint32_t render_mode,
                             const uint32_t show_grid[3], const float sun_dir3[3], const float sun_color3[3],
                             float sun_intensity, WxoState *out);
void wxo_default_sun(float dir3[3], float color3[3], float *intensity); /* egui_dev.rs:355-367 */

/* ---- raycast (raycast.comp.wgsl) ----------------------------------------------------------- */
typedef struct WxoAov { /* any pointer may be NULL; all are W*H, row-major like the image */
  uint8_t *state;   /* 0 hit, 1 oob, 2 maxed */
  int32_t *voxel;   /* 3 per pixel: vec3<i32>(floor(hit.p)) */
  int32_t *leaf;    /* parents[2].idx when the terminating lookup is level 3, else -1 */
  uint8_t *level;   /* num_parents of the terminating lookup */
  uint32_t *iters;  /* hit.i */
  float *depth;     /* |hit.p - eye| */
  uint8_t *mask;    /* bit0 x, bit1 y, bit2 z */
  float *pos;       /* 3 per pixel: hit.p */
} WxoAov;

typedef struct WxoStats {
  uint64_t rays;          /* hdda_ray invocations (primary + secondary) */
  uint64_t primary_rays;
  uint64_t lookups[4];    /* by num_parents of the lookup result, all rays */
  uint64_t primary_lookups[4];
  uint64_t alg_bytes;         /* sum b(level) over lookups (12/12/8 B for level 1/2/3, 0 B for
                                 level-0 misses), all rays, + 4 B per pixel */
  uint64_t primary_alg_bytes; /* same, primary rays only, + 4 B per pixel */
  uint32_t max_iters;
  uint64_t hit, oob, maxed; /* primary ray terminal states */
} WxoStats;

/* Renders rows [y0, y1) of the W x H frame; pixels outside the reference's dispatch
 * (x >= (W/8)*8 or y >= (H/4)*4, wgpu_context.rs:281) are left zero.  rgba/aov are full-frame buffers. */
void wxo_render(const WxoGpuData *g, const WxoState *s, uint32_t width, uint32_t height, uint32_t y0, uint32_t y1,
                uint8_t *rgba, const WxoAov *aov, int threads, WxoStats *stats);

/* ---- capture (src/render/recorder.rs:20-37, :132-140) ----------------------------------------- */
uint8_t wxo_linear_to_srgb(uint8_t value);
void wxo_frame_to_srgb_rgb(const uint8_t* rgba, size_t n_pixels, uint8_t* rgb);

#ifdef __cplusplus
}
#endif
#endif
