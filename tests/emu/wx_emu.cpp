// wx_emu.cpp -- TEST INFRASTRUCTURE (never linked into the product).  The device code of woxel_b200/csrc/wx_device.cuh --
// ray generation, march_fast / march_exact, descend, shading, the rgba8 store -- compiled for the host through
// cuda_shim.h and run once per pixel over the tables wx_pack.h builds (the ones wx_tree_upload sends to the GPU).
// Two uses:
//   * tests/test_device_emu.py compares it with the oracle bit for bit on the CPU (every value-exact rewrite of the fast
//     march is therefore checked without a GPU, for each of the reciprocal values MUFU.RCP may return);
//   * wxe_render's statistics replay the lanes of each 4x8-pixel warp in lockstep and count which table levels a warp
//     step touches: the divergence figures in DESIGN.md section 4 come from here.
#include "cuda_shim.h"

#include <vector>

#include "../../woxel_b200/csrc/wx_internal.h"
#include "../../woxel_b200/csrc/wx_pack.h"

thread_local int wx_emu_rcp_bump = 0;
thread_local WxEmuTrace* wx_emu_trace = nullptr;

namespace {
using namespace wx;

int g_march = 0;  // wxe_set_march: 0 exact, 1 tolerance mode, 2 tolerance mode + bounding-box clip (WX_OPT_MARCH); like the launcher, mode 2 always runs exact
template <int MODE>
void pixel(const RenderParams& P, const PixelRef& q) {
  if (g_march != kMarchExact && MODE != 2) {
    if (P.has_aov) render_pixel<MODE, true, kMarchTolerance>(P, q);
    else render_pixel<MODE, false, kMarchTolerance>(P, q);
    return;
  }
  if (P.has_aov) render_pixel<MODE, true>(P, q);
  else render_pixel<MODE, false>(P, q);
}
void pixel_mode(uint32_t mode, const RenderParams& P, const PixelRef& q) {
  switch (mode) {  // the switch of launch_raycast (wx_raycast.cu)
    case 1: return pixel<1>(P, q);
    case 2: return pixel<2>(P, q);
    case 3: return pixel<3>(P, q);
    case 4: return pixel<4>(P, q);
    default: return pixel<0>(P, q);
  }
}
}  // namespace

enum {
  WXE_RAYS = 0, WXE_WARPS, WXE_LANE_STEPS, WXE_WARP_STEPS, WXE_ROOT_BLOCKS, WXE_N5_BLOCKS, WXE_N4_BLOCKS, WXE_LEAF_BLOCKS,
  WXE_GENERIC_ITERS, WXE_COMBO0, /* 8 entries: bit0 N5 table, bit1 N4 table, bit2 leaf brick touched by the warp step */
  WXE_LANE_TABLE_READS = WXE_COMBO0 + 8, WXE_TRUNCATED, WXE_N_STATS
};

extern "C" int wxe_n_stats(void) { return WXE_N_STATS; }
extern "C" void wxe_set_march(int march) { g_march = march; }

// Renders states[0..n_states) (each in its own render mode) exactly as wx_render would.  warp_w x (32 / warp_w) is the
// pixel footprint of a warp for the lockstep statistics (stats may be null; they cover the primary rays of mode-0 style
// marches only: secondary rays are traced into the same per-lane buffer after the primary one and are ignored).
namespace {
// The tables wx_tree_upload would send to the GPU, and the launch parameters of a whole-frame, single-shard launch.
struct EmuScene {
  std::vector<uint32_t> e5, e4, grid, f4;
  std::vector<uint8_t> l3;
  std::vector<int4> origins;
  wx::RenderParams P;
  int init(const WxTreeDesc* d, const WxState* states, uint32_t n_states, uint32_t width, uint32_t height, uint8_t* rgba, const WxAov* aov) {
    using namespace wx;
    uint32_t max5 = 0, max4 = 0, max3v = 0;
    int rc = pack_internal(d->n5, 32768, 128.f, d->kids5, d->vals5, d->tab5, d->n4, e5, &max5);
    if (rc) return rc;
    rc = pack_internal(d->n4, 4096, 8.f, d->kids4, d->vals4, d->tab4, d->n3, e4, &max4);
    if (rc) return rc;
    const uint32_t leaf_bits = pack_leaves(d->n3, d->vals3, d->tab3, d->tab3_elem_bytes, l3, &max3v);
    bias_origins(d->n5, d->origins, origins);
    int16_t root_grid[64];
    build_root_grid(d->n5, d->origins, root_grid);
    // slack so that zero-sized levels have a base address
    e5.push_back(0), e4.push_back(0), l3.push_back(0), origins.push_back(make_int4(0, 0, 0, 0));
    memset(&P, 0, sizeof(P));
    const bool fast_ok = fast_march_ok(leaf_bits, max5, max4, max3v);
    const bool with_grid = world_grid_ok(fast_ok, d->n4, d->n3);
    int32_t bbox[6] = {1 << 30, 1 << 30, 1 << 30, -1, -1, -1};
    if (with_grid) build_grid_tables(d->n5, d->n4, origins, root_grid, e5.data(), e4.data(), grid, f4), grid_bbox_cells(grid, bbox);
    fill_dev_tree(P.tree, e5.data(), e4.data(), l3.data(), origins.data(), d->n5, d->n4, d->n3, leaf_bits == 8 ? 9 : 11, fast_ok, root_grid,
                  with_grid ? grid.data() : nullptr, with_grid ? f4.data() : nullptr, g_march == 2 ? bbox : nullptr);
    P.n_states = n_states;
    P.states = states;
    P.s0 = states[0];
    P.width = width, P.height = height;
    P.disp_w = (width / 8) * 8, P.disp_h = (height / 4) * 4;  // wgpu_context.rs:281
    P.row_end = height;
    P.rgba = reinterpret_cast<uchar4*>(rgba);
    if (aov) {
      P.aov.state = aov->state, P.aov.voxel = aov->voxel, P.aov.leaf = aov->leaf, P.aov.level = aov->level;
      P.aov.iters = aov->iters, P.aov.depth = aov->depth, P.aov.mask = aov->mask, P.aov.pos = aov->pos;
      P.has_aov = 1u;
    }
    return 0;
  }
};
}  // namespace

extern "C" int wxe_render(const WxTreeDesc* d, const WxState* states, uint32_t n_states, uint32_t width, uint32_t height, uint8_t* rgba,
                          const WxAov* aov, int rcp_bump, uint32_t warp_w, uint64_t* stats) {
  if (!d || !states || !rgba || n_states == 0) return WX_ERR_INVALID_ARGUMENT;
  EmuScene scene;
  const int rc = scene.init(d, states, n_states, width, height, rgba, aov);
  if (rc) return rc;
  const RenderParams& P = scene.P;
  if (warp_w == 0 || 32 % warp_w) warp_w = 4;
  const uint32_t warp_h = 32 / warp_w;
  const uint32_t tiles_x = (width + warp_w - 1) / warp_w, tiles_y = (height + warp_h - 1) / warp_h;
  std::vector<std::atomic<uint64_t>> acc(WXE_N_STATS);
  for (auto& a : acc) a.store(0);

  for (uint32_t cam = 0; cam < n_states; ++cam) {
    const uint32_t mode = states[cam].render_mode[0];
    parallel_for((size_t)tiles_x * tiles_y, [&](size_t tile) {
      wx_emu_rcp_bump = rcp_bump;
      const uint32_t ty = (uint32_t)(tile / tiles_x), tx = (uint32_t)(tile % tiles_x);
      constexpr uint32_t kCap = 1024;
      uint8_t trace[32][kCap];
      uint32_t len[32];
      uint64_t local[WXE_N_STATS] = {0};
      for (uint32_t lane = 0; lane < 32; ++lane) {
        PixelRef q;
        q.x = tx * warp_w + lane % warp_w, q.y = ty * warp_h + lane / warp_w, q.cam = cam;
        q.in_frame = q.x < width && q.y < height;
        q.dispatched = q.x < P.disp_w && q.y < P.disp_h;
        WxEmuTrace t{trace[lane], 0, kCap};
        wx_emu_trace = stats ? &t : nullptr;
        pixel_mode(mode, P, q);
        wx_emu_trace = nullptr;
        len[lane] = t.n;
      }
      if (!stats) return;
      // primary ray of a lane = the steps up to and including its first ending step; a mode-0/1/2 frame traces nothing else
      uint32_t max_len = 0, rays = 0;
      for (uint32_t lane = 0; lane < 32; ++lane) max_len = std::max(max_len, len[lane]), rays += len[lane] ? 1 : 0;
      if (!rays) return;
      local[WXE_RAYS] += rays, local[WXE_WARPS] += 1;
      for (uint32_t k = 0; k < max_len; ++k) {
        uint32_t root = 0, b5 = 0, b4 = 0, b3 = 0, deepest = 0;
        for (uint32_t lane = 0; lane < 32; ++lane) {
          if (len[lane] <= k) continue;
          const uint32_t start = trace[lane][k] >> 4, end = trace[lane][k] & 15u;
          local[WXE_LANE_STEPS] += 1;
          const uint32_t r5 = (start == 1u || (start == 0u && end >= 1u)) ? 1u : 0u;
          const uint32_t r4 = (start == 2u || (start <= 1u && end >= 2u)) ? 1u : 0u;
          const uint32_t r3 = end == 3u ? 1u : 0u;
          root |= start == 0u, b5 |= r5, b4 |= r4, b3 |= r3;
          deepest = std::max(deepest, r5 + r4 + r3);
          local[WXE_LANE_TABLE_READS] += r5 + r4 + r3;
        }
        local[WXE_WARP_STEPS] += 1;
        local[WXE_ROOT_BLOCKS] += root, local[WXE_N5_BLOCKS] += b5, local[WXE_N4_BLOCKS] += b4, local[WXE_LEAF_BLOCKS] += b3;
        local[WXE_GENERIC_ITERS] += deepest;
        local[WXE_COMBO0 + (b5 | (b4 << 1) | (b3 << 2))] += 1;
      }
      for (uint32_t lane = 0; lane < 32; ++lane) local[WXE_TRUNCATED] += len[lane] >= kCap;
      for (int i = 0; i < WXE_N_STATS; ++i)
        if (local[i]) acc[i].fetch_add(local[i]);
    });
  }
  if (stats)
    for (int i = 0; i < WXE_N_STATS; ++i) stats[i] = acc[i].load();
  return WX_OK;
}

// The ticket protocol of raycast_persistent_cta with real threads: n_ctas "CTAs" of warps_per_cta "warps" (one thread each, the
// role of lane 0) pull (chunk, tile) tickets until the queue says done.  counts[chunk * 16 + tile] receives how often each tile
// was handed out (must be exactly 1 everywhere); returns the number of chunks rendered privately after a lost publish race.
extern "C" int wxe_queue_sim(uint32_t n_chunks, uint32_t n_ctas, uint32_t warps_per_cta, uint32_t* counts, uint32_t seed) {
  using namespace wx;
  RenderParams P;
  memset(&P, 0, sizeof(P));
  uint32_t counter = 0;
  P.work_counter = &counter;
  P.n_chunks = n_chunks;
  std::vector<uint32_t> state(n_ctas, 16u);  // as the kernel initialises s_state
  std::atomic<int> privately{0}, bad{0};
  std::vector<std::thread> th;
  for (uint32_t c = 0; c < n_ctas; ++c)
    for (uint32_t w = 0; w < warps_per_cta; ++w)
      th.emplace_back([&, c, w]() {
        uint32_t rng = seed * 2654435761u + c * 97u + w * 7919u + 1u;
        uint32_t priv = kNoTile;
        for (;;) {
          const bool was_private = priv != kNoTile;
          uint32_t t = 0;
          const uint32_t chunk = next_ticket(P, &state[c], priv, t);
          if (chunk == kNoTile) break;
          if (!was_private && priv != kNoTile) privately.fetch_add(1);
          if (chunk >= n_chunks || t >= 16u) {
            bad.fetch_add(1);
            break;
          }
          __atomic_fetch_add(&counts[(size_t)chunk * 16u + t], 1u, __ATOMIC_RELAXED);
          rng = rng * 1664525u + 1013904223u;  // uneven "tile render times" to shake the interleavings
          for (volatile uint32_t spin = 0; spin < ((rng >> 24) & 63u) * 8u; ++spin) {
          }
          if ((rng >> 20 & 15u) == 0u) std::this_thread::yield();
        }
      });
  for (auto& t : th) t.join();
  return bad.load() ? -1 : privately.load();
}

// Pixel coverage of the chunk / tile / lane mapping of raycast_persistent_cta for one shard of a frame, with the launch geometry
// of launch_raycast (wx_raycast.cu): hits[y * width + x] += 1 for every in-frame pixel some (chunk, tile, lane) maps to.
extern "C" int wxe_chunk_coverage(uint32_t width, uint32_t height, uint32_t shard_index, uint32_t shard_count, uint32_t band_rows,
                                  uint32_t n_cams, uint32_t* hits) {
  using namespace wx;
  RenderParams P;
  memset(&P, 0, sizeof(P));
  P.width = width, P.height = height, P.row_base = 0, P.row_end = height;
  P.disp_w = (width / 8) * 8, P.disp_h = (height / 4) * 4;
  if (shard_count <= 1 || band_rows == 0) {
    P.shard_index = 0, P.shard_count = 1, P.band_rows = ((height + 7) / 8) * 8, P.own_bands = 1;
  } else {
    P.shard_index = shard_index, P.shard_count = shard_count, P.band_rows = band_rows;
    P.own_bands = shard_own_bands(height, shard_index, shard_count, band_rows);
  }
  const uint64_t own_rows = (uint64_t)P.own_bands * P.band_rows;
  P.chunks_x = (width + 31u) / 32u;
  P.chunks_y = (uint32_t)((own_rows + 15u) / 16u);
  P.n_chunks = P.chunks_x * P.chunks_y * n_cams;
  for (uint32_t chunk = 0; chunk < P.n_chunks; ++chunk)
    for (uint32_t t = 0; t < 16; ++t)
      for (uint32_t lane = 0; lane < 32; ++lane) {
        const PixelRef q = pixel_of_chunk_tile(P, chunk, t, lane);
        if (q.in_frame) hits[((size_t)q.cam * height + q.y) * width + q.x] += 1;
      }
  return 0;
}

// A whole frame through the work distribution of raycast_persistent_cta: n_ctas x warps_per_cta threads play the warps (each
// renders the 32 lanes of its tile one after the other), pulling (chunk, tile) tickets exactly as the kernel's loop does; camera
// batches render the states whose mode equals that of state 0 only if all modes are equal (one launch = one mode, as in wx_api).
extern "C" int wxe_render_cta_queue(const WxTreeDesc* d, const WxState* states, uint32_t n_states, uint32_t width, uint32_t height,
                                    uint8_t* rgba, const WxAov* aov, uint32_t n_ctas, uint32_t warps_per_cta) {
  using namespace wx;
  if (!d || !states || !rgba || n_states == 0 || n_ctas == 0 || warps_per_cta == 0) return WX_ERR_INVALID_ARGUMENT;
  for (uint32_t i = 1; i < n_states; ++i)
    if (states[i].render_mode[0] != states[0].render_mode[0]) return WX_ERR_INVALID_ARGUMENT;
  EmuScene scene;
  const int rc = scene.init(d, states, n_states, width, height, rgba, aov);
  if (rc) return rc;
  RenderParams& P = scene.P;
  // launch geometry of launch_raycast for a single shard (wx_raycast.cu)
  P.shard_index = 0, P.shard_count = 1, P.band_rows = ((height + 7) / 8) * 8, P.own_bands = 1;
  P.chunks_x = (width + 31u) / 32u;
  P.chunks_y = (P.band_rows + 15u) / 16u;
  P.n_chunks = P.chunks_x * P.chunks_y * n_states;
  uint32_t counter = 0;
  P.work_counter = &counter;
  const uint32_t mode = states[0].render_mode[0];
  std::vector<uint32_t> state(n_ctas, 16u);
  std::vector<std::thread> th;
  for (uint32_t c = 0; c < n_ctas; ++c)
    for (uint32_t w = 0; w < warps_per_cta; ++w)
      th.emplace_back([&, c]() {
        uint32_t priv = kNoTile;
        for (;;) {
          uint32_t t = 0;
          const uint32_t chunk = next_ticket(P, &state[c], priv, t);
          if (chunk == kNoTile) break;
          for (uint32_t lane = 0; lane < 32; ++lane) pixel_mode(mode, P, pixel_of_chunk_tile(P, chunk, t, lane));
        }
      });
  for (auto& t : th) t.join();
  return WX_OK;
}
