/*
 * abi_smoke.c -- a plain C99 consumer of include/woxel_b200.h and include/woxel_host.h (TEST code).
 *
 * Proves that the headers are valid C (compiled with -std=c99 -pedantic -Werror), that WxState has the layout of the
 * reference's ComputeState (src/render/gpu_types/compute_state.rs:9-29), and drives the boundary the way a non-Python
 * host does: build a small model with set_voxel, flatten it, wx_tree_build, wx_render, wx_capture_srgb.
 *
 *   exit 0, prints "no-device"  : no GPU -- wx_init refused with WX_ERR_NO_DEVICE (there is no CPU fallback)
 *   exit 0, prints "rendered …" : a frame was rendered and passed the checks below
 *   exit 1                      : anything else
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "woxel_b200.h"
#include "woxel_host.h"

#define CHECK(cond, what)                                  \
  do {                                                     \
    if (!(cond)) {                                         \
      fprintf(stderr, "abi_smoke: FAILED: %s\n", (what));  \
      return 1;                                            \
    }                                                      \
  } while (0)

/* C99 has no static_assert: negative array size on failure */
typedef char assert_state_size[sizeof(WxState) == 256 ? 1 : -1];
typedef char assert_state_eye[offsetof(WxState, eye) == 128 ? 1 : -1];
typedef char assert_state_mode[offsetof(WxState, render_mode) == 192 ? 1 : -1];
typedef char assert_state_sun[offsetof(WxState, sun_color) == 240 ? 1 : -1];

int main(void) {
  WxContext *ctx = NULL;
  int rc, i, st;

  CHECK(wx_abi_version() == WX_ABI_VERSION, "wx_abi_version");
  for (st = 0; st >= -7; --st) CHECK(wx_strerror(st) != NULL && strlen(wx_strerror(st)) > 0, "wx_strerror");

  /* host-side model: a 24^3 block of voxels around the origin, straddling all eight N5s */
  {
    WxhVdb *v = wxh_vdb_new();
    WxhFlat *flat;
    WxTreeDesc desc;
    WxTree *tree = NULL;
    WxSdfInfo sdf;
    WxState state;
    WxRenderInfo info;
    uint64_t nodes[3];
    int x, y, z;
    const uint32_t w = 128, h = 64;
    const float eye[3] = {0.5f, 0.5f, -200.5f}, target[3] = {0.5f, 0.5f, 0.5f}, up[3] = {0.f, 1.f, 0.f};
    const uint32_t grid[3] = {0, 0, 0};
    float sun_dir[3], sun_col[3], sun_i;
    uint8_t *rgba, *rgba2, *rgb;
    static uint8_t hit_state[128 * 64];
    static float depth[128 * 64];
    WxAov aov;
    size_t hits = 0, lit = 0;

    CHECK(v != NULL, "wxh_vdb_new");
    for (x = -12; x < 12; ++x)
      for (y = -12; y < 12; ++y)
        for (z = -12; z < 12; ++z) wxh_vdb_set_voxel(v, x, y, z, 1u);
    wxh_vdb_count_nodes(v, nodes);
    CHECK(nodes[0] == 8 && nodes[1] == 8 && nodes[2] == 64, "count_nodes of a 24^3 block around the origin");
    CHECK(wxh_vdb_count_leaf_values(v) == 24u * 24u * 24u, "count_leaf_values");
    flat = wxh_vdb_to_flat(v, 0);
    CHECK(flat != NULL, "wxh_vdb_to_flat");
    wxh_flat_desc(flat, &desc);
    CHECK(desc.n5 == 8 && desc.n4 == 8 && desc.n3 == 64, "flat counts");

    wxh_default_sun(sun_dir, sun_col, &sun_i);
    rc = wxh_compute_state_build(eye, target, up, (float)w / (float)h, 45.f, (float)w, 3u, grid, sun_dir, sun_col, sun_i, &state);
    CHECK(rc == 0 && state.render_mode[0] == 3u && state.eye[2] == -200.5f, "wxh_compute_state_build");

    rc = wx_init(0, NULL, &ctx);
    if (rc == WX_ERR_NO_DEVICE) {
      CHECK(ctx == NULL, "wx_init must not return a context without a device");
      wxh_flat_free(flat);
      wxh_vdb_free(v);
      printf("no-device\n");
      return 0;
    }
    CHECK(rc == WX_OK && ctx != NULL, "wx_init");

    rc = wx_tree_build(ctx, &desc, &tree, &sdf);
    if (rc != WX_OK) fprintf(stderr, "wx_tree_build: %s\n", wx_last_error(ctx));
    CHECK(rc == WX_OK && tree != NULL && sdf.rounds > 0, "wx_tree_build");

    rgba = (uint8_t *)malloc((size_t)w * h * 4);
    rgb = (uint8_t *)malloc((size_t)w * h * 3);
    CHECK(rgba && rgb, "malloc");
    memset(rgba, 0x5A, (size_t)w * h * 4);
    memset(&aov, 0, sizeof(aov));
    aov.state = hit_state, aov.depth = depth;
    rc = wx_render(ctx, tree, &state, 1, w, h, rgba, &aov);
    if (rc != WX_OK) fprintf(stderr, "wx_render: %s\n", wx_last_error(ctx));
    CHECK(rc == WX_OK, "wx_render");
    CHECK(wx_last_render_info(ctx, &info) == WX_OK && info.rays == (uint64_t)w * h && info.launches >= 1, "wx_last_render_info");
    for (i = 0; i < (int)(w * h); ++i) {
      CHECK(rgba[4 * i + 3] == 255, "alpha of a dispatched pixel");
      CHECK(hit_state[i] <= 1, "every ray hits or leaves the world");
      if (hit_state[i] == 0) {
        ++hits;
        /* Only the -z face is visible from the axis.  Diffuse mode (raycast.comp.wgsl:182-194) with the default sun
         * normalize(1,-1,0.5), colour (1, 210/255, 160/255): ambient 0.3*(0.4,0.4,0.3)*(0.4,0.2,0.2) + diffuse
         * 0.7*sun*(0.4,0.2,0.2)*(0.5/1.5) = (0.1413, 0.0624, 0.0473) -> unorm8 (36, 16, 12); no occluder towards the sun. */
        CHECK(rgba[4 * i] == 36 && rgba[4 * i + 1] == 16 && rgba[4 * i + 2] == 12, "colour of the lit -z face");
        ++lit;
      }
    }
    CHECK(hits >= 64 && hits <= 144, "the 24-voxel face is about 10 x 10 pixels at this distance");
    /* the pixel at the image centre looks straight at the block's -z face, 188.5 voxels away */
    i = (int)((h / 2) * w + w / 2);
    CHECK(hit_state[i] == 0 && depth[i] > 188.0f && depth[i] < 189.5f, "centre pixel hits the near face of the block");
    /* the same frame without AOVs (the pipelined read-back path) is the same frame */
    rgba2 = (uint8_t *)malloc((size_t)w * h * 4);
    CHECK(rgba2 != NULL, "malloc");
    CHECK(wx_render(ctx, tree, &state, 1, w, h, rgba2, NULL) == WX_OK, "wx_render without AOVs");
    CHECK(memcmp(rgba, rgba2, (size_t)w * h * 4) == 0, "frames with and without AOVs are identical");
    free(rgba2);

    rc = wx_capture_srgb(ctx, 1, w, h, rgb);
    CHECK(rc == WX_OK, "wx_capture_srgb");
    {
      uint8_t lut[256];
      CHECK(wx_srgb_table(lut) == WX_OK, "wx_srgb_table");
      for (i = 0; i < (int)(w * h); ++i)
        CHECK(rgb[3 * i] == lut[rgba[4 * i]] && rgb[3 * i + 1] == lut[rgba[4 * i + 1]] && rgb[3 * i + 2] == lut[rgba[4 * i + 2]],
              "capture == table(frame)");
    }

    /* argument errors come back as statuses with a text, never as a crash */
    CHECK(wx_render(ctx, tree, &state, 1, 0, h, rgba, NULL) == WX_ERR_INVALID_ARGUMENT, "empty frame is refused");
    CHECK(strlen(wx_last_error(ctx)) > 0, "wx_last_error has a text");

    printf("rendered %ux%u: %lu hit pixels, %lu lit, sdf %u rounds, kernel %.3f ms\n", w, h,
           (unsigned long)hits, (unsigned long)lit, sdf.rounds, (double)info.kernel_ms);
    free(rgba);
    free(rgb);
    CHECK(wx_tree_free(ctx, tree) == WX_OK, "wx_tree_free");
    CHECK(wx_shutdown(ctx) == WX_OK, "wx_shutdown");
    wxh_flat_free(flat);
    wxh_vdb_free(v);
  }
  return 0;
}
