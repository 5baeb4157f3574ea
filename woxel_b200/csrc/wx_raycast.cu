// wx_raycast.cu -- the raycast kernels (cp_main of the reference, src/shaders/raycast.comp.wgsl:60-68)
// and their launcher.  One thread per primary ray; a warp is a 4x8-pixel tile (the reference's workgroup
// is 8x4; the taller shape measured faster, see WX_WARP_W), a CTA is four such tiles side by side (16x8
// pixels).  The per-pixel code (render_pixel, shade_and_store) lives in wx_device.cuh.
#include <algorithm>

#include "wx_device.cuh"
#include "wx_internal.h"

namespace wx {

#ifndef WX_CTA_WARPS
#define WX_CTA_WARPS 4  // 4: CTA = 2x2 warp tiles (16x8 px); 2: 2x1 (16x4 px); 1: one tile (8x4 px)
#endif
#ifndef WX_WARP_W
// Pixels per warp row: a warp covers WX_WARP_W x (32 / WX_WARP_W) pixels.  The reference's workgroup is 8 x 4; measured on
// the 4K sphere (profiles/r1_variants_g.txt): 1x32 1.038 ms, 2x16 0.963, 4x8 0.925, 8x4 0.948, 16x2 1.002, 32x1 1.108.
#define WX_WARP_W 4
#endif
constexpr int kWarpW = WX_WARP_W, kWarpH = 32 / WX_WARP_W;
// the warps of a CTA tile a 16-pixel-wide block (8 rows for 4 warps): row bands of 8 stay the sharding unit
#ifndef WX_CTA_WARPS_X
#define WX_CTA_WARPS_X ((16 / kWarpW) < WX_CTA_WARPS ? (16 / kWarpW) : WX_CTA_WARPS)  // A/B: 2 = the four warps as 2 x 2 (8 x 16 px)
#endif
constexpr int kCtaWarpsX = WX_CTA_WARPS_X;
constexpr int kCtaWarpsY = WX_CTA_WARPS / kCtaWarpsX;
constexpr int kTileW = kCtaWarpsX * kWarpW, kTileH = kCtaWarpsY * kWarpH;  // CTA footprint in pixels
#define WX_LANE_X(warp, lane) (((warp) % kCtaWarpsX) * kWarpW + ((lane) % kWarpW))
#define WX_LANE_Y(warp, lane) (((warp) / kCtaWarpsX) * kWarpH + ((lane) / kWarpW))
constexpr int kThreads = 32 * WX_CTA_WARPS;
#ifndef WX_MIN_BLOCKS
#define WX_MIN_BLOCKS (36 / WX_CTA_WARPS)  // resident CTAs per SM the register budget is capped for (36 warps -> 56 registers)
#endif

// Launch order.  The hardware starts the CTAs of a grid in linear order (x fastest) and a launch ends with a drain as long as
// its slowest remaining warp (a lone 355-step ray runs ~75 us).  The longest rays of a frame graze the
// silhouette of what is in view, which for a framed object lies towards the border of the image; the centre holds short, early
// hits.  So tiles are taken from the outside in -- columns 0, n-1, 1, n-2, ... and rows likewise: the slow rays start first and
// the grid ends on the cheap centre tiles.  Only the order changes, not which pixel a thread renders (WX_NO_OUTSIDE_IN: A/B).
__device__ __forceinline__ uint32_t outside_in(uint32_t i, uint32_t n) {
#ifdef WX_NO_OUTSIDE_IN
  return i;
#else
  return (i & 1u) ? n - 1u - (i >> 1) : (i >> 1);
#endif
}

// One CTA's tile: tile column tx, tile row trow among the rows this launch owns, camera cam_i of the launch.  Tile id =
// (cam_i * tile_rows + trow) * tiles_x + tx.  Records the tile for the next launch's long list when one of its rays was long.
template <int MODE, bool AOV, int MARCH>
__device__ __forceinline__ void render_tile(const RenderParams& P, uint32_t tx, uint32_t trow, uint32_t cam_i, uint32_t tile_id) {
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // band of this tile row and the row inside it; the two usual shapes (one band, or one tile row per band) need no division
  uint32_t own_band = 0, in_band = trow;
  if (P.tile_rows_per_band == 1u) own_band = trow, in_band = 0;
  else if (P.own_bands > 1u) own_band = trow / P.tile_rows_per_band, in_band = trow - own_band * P.tile_rows_per_band;
  const uint32_t band = own_band * P.shard_count + P.shard_index;
  PixelRef q;
  q.x = tx * kTileW + WX_LANE_X(warp, lane);
  q.y = P.row_base + band * P.band_rows + in_band * kTileH + WX_LANE_Y(warp, lane);
  q.cam = P.cam_base + cam_i;
  q.in_frame = q.x < P.width && q.y < P.row_end;
  q.dispatched = q.x < P.disp_w && q.y < P.disp_h;
  const uint32_t iters = render_pixel<MODE, AOV, MARCH>(P, q);
  if (P.next_flag != nullptr && iters >= P.sched_threshold) {  // a long ray: this tile starts first next time (few lanes get here)
    if (atomicExch(P.next_flag + tile_id, 1u) == 0u) {
      const uint32_t slot = atomicAdd(P.next_list, 1u);
      if (slot < P.sched_cap) P.next_list[1u + slot] = tile_id;
      else P.next_flag[tile_id] = 0u;  // list full: the tile stays in the main grid
    }
  }
}

// Tiled kernel: one thread per pixel of the grid, a warp per 4x8-pixel tile, a CTA per four tiles (16x8 pixels).
template <int MODE, bool AOV, int MARCH>
__device__ __forceinline__ void tiled_body(const RenderParams& P) {
  const uint32_t tx = outside_in(blockIdx.x, gridDim.x);
  const uint32_t trow = outside_in(blockIdx.y, gridDim.y);      // tile row among the rows this launch owns
  const uint32_t tile_id = (blockIdx.z * (P.frame_tile_rows ? P.frame_tile_rows : gridDim.y) + trow + P.tile_row_offset) * gridDim.x + tx;
  if (P.prev_flag != nullptr && __ldg(P.prev_flag + tile_id) != 0u) return;  // rendered by the long-tile kernel of this launch
  render_tile<MODE, AOV, MARCH>(P, tx, trow, blockIdx.z, tile_id);
}
// The tiles of the long list (one CTA each; CTAs beyond the list's length leave at once).  Launched before the main grid on a
// high-priority stream, so that the slowest rays of the frame are the first to start.
template <int MODE, bool AOV, int MARCH>
__device__ __forceinline__ void long_body(const RenderParams& P) {
  const uint32_t n = min(__ldg(P.prev_list), P.sched_cap);
  if (blockIdx.x >= n) return;
  const uint32_t tile_id = __ldg(P.prev_list + 1u + blockIdx.x);
  const uint32_t per_cam = P.tiles_x * (P.tile_rows_per_band * P.own_bands);
  const uint32_t cam_i = tile_id / per_cam, rem = tile_id - cam_i * per_cam;
  const uint32_t trow = rem / P.tiles_x;
  render_tile<MODE, AOV, MARCH>(P, rem - trow * P.tiles_x, trow, cam_i, tile_id);
}

template <int MODE, bool AOV>
__global__ void __launch_bounds__(kThreads, WX_MIN_BLOCKS) raycast_kernel(const __grid_constant__ RenderParams P) {
  tiled_body<MODE, AOV, kMarchExact>(P);
}
template <int MODE, bool AOV>
__global__ void __launch_bounds__(kThreads, WX_MIN_BLOCKS) raycast_long_tiles(const __grid_constant__ RenderParams P) {
  long_body<MODE, AOV, kMarchExact>(P);
}

// The same kernels in tolerance mode (WX_OPT_MARCH = 1, wx_device.cuh: fused p += t * dir, rays start at the bounding box of
// the active cells).  Separate entry points so that the exact kernel's symbol, registers and SASS do not depend on them.
template <int MODE, bool AOV>
__global__ void __launch_bounds__(kThreads, WX_MIN_BLOCKS) raycast_kernel_tol(const __grid_constant__ RenderParams P) {
  tiled_body<MODE, AOV, kMarchTolerance>(P);
}
template <int MODE, bool AOV>
__global__ void __launch_bounds__(kThreads, WX_MIN_BLOCKS) raycast_long_tiles_tol(const __grid_constant__ RenderParams P) {
  long_body<MODE, AOV, kMarchTolerance>(P);
}

// ---------------------------------------------------------------------------------------------
// Persistent kernel: the grid is the number of CTAs the device holds at once; every warp pulls 8x4-pixel
// tiles from a global counter until the frame is done.  Tiles are numbered so that 16 consecutive ones
// form a 32x16-pixel block: the tiles a warp (and an SM) works on one after the other are neighbours and
// find their nodes in L1.  Compared with the tiled grid this removes the CTA tail (a CTA slot is held
// until its longest warp ends: ncu shows 47 % warps active of the 56 % the registers allow).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ PixelRef pixel_of(const RenderParams& P, uint32_t tile, uint32_t lane) {
  const uint32_t chunk = tile >> 4, t = tile & 15u;
  const uint32_t per_cam = P.chunks_x * P.chunks_y;
  const uint32_t cam_i = chunk / per_cam, rem = chunk - cam_i * per_cam;
  const uint32_t cy = rem / P.chunks_x, cx = rem - cy * P.chunks_x;
  PixelRef q;
  q.x = (cx * 4u + (t & 3u)) * 8u + (lane & 7u);
  const uint32_t vrow = (cy * 4u + (t >> 2)) * 4u + (lane >> 3);  // row among the rows this launch owns
  const uint32_t band = (vrow / P.band_rows) * P.shard_count + P.shard_index;
  q.y = P.row_base + band * P.band_rows + vrow % P.band_rows;
  q.cam = P.cam_base + cam_i;
  q.in_frame = q.x < P.width && q.y < P.row_end && vrow < P.own_bands * P.band_rows;
  q.dispatched = q.x < P.disp_w && q.y < P.disp_h;
  return q;
}

#ifndef WX_TILES_PER_GRAB
#define WX_TILES_PER_GRAB 1  // (measured: 1 -> 0.976 ms, 4 -> 1.20 ms, 16 -> 2.01 ms on the 4K sphere) consecutive tiles (of the 16 of a 32x16-pixel chunk) a warp renders per queue access
#endif
template <int MODE, bool AOV>
__global__ void __launch_bounds__(kThreads, WX_MIN_BLOCKS) raycast_persistent(const __grid_constant__ RenderParams P) {
  const uint32_t lane = threadIdx.x & 31;
  for (;;) {
    uint32_t tile = 0;
    if (lane == 0) tile = atomicAdd(P.work_counter, (uint32_t)WX_TILES_PER_GRAB);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= P.n_chunks * 16u) break;
#pragma unroll 1
    for (uint32_t k = 0; k < (uint32_t)WX_TILES_PER_GRAB; ++k) {
      (void)render_pixel<MODE, AOV>(P, pixel_of(P, tile + k, lane));
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Persistent kernel with a CTA-level chunk queue (WX_OPT_KERNEL = 2).  Measured in round 2 (profiles/r2_cta_queue.txt):
// bit-identical (tests/test_parity_gpu.py::test_cta_queue_kernel_is_bit_identical) and 2.1x SLOWER than the tiled grid --
// lane 0's ticket loop (a shared-memory atomic read and a CAS per tile) sits in front of every tile.  Kept selectable, not the default.
// Why: in the tiled grid a CTA slot is held until its slowest warp ends (warp slots are 92.5 % used inside a CTA on the
// bench frame, tools/warp_stats.py), and the warp-level queue above cures that but scatters the warps of a CTA over the
// frame (their tiles no longer share nodes in L1).  Here the CTA owns a 32x16-pixel chunk; its warps take the chunk's 16
// tiles (4x8 pixels, row-major: four consecutive tickets are a 16x8 strip, the tiled kernel's CTA footprint) from a
// shared-memory ticket, and the first warp to find the chunk exhausted fetches the next chunk from the global counter
// for everybody -- no barrier, nobody waits for the stragglers of the previous chunk.
// s_state = chunk << 5 | tiles handed out (0..16); kQueueDone once the global counter ran out.
// ---------------------------------------------------------------------------------------------
template <int MODE, bool AOV>
__global__ void __launch_bounds__(kThreads, WX_MIN_BLOCKS) raycast_persistent_cta(const __grid_constant__ RenderParams P) {
  static_assert(kWarpW == 4, "the chunk layout below is written for 4x8-pixel warps");
  __shared__ uint32_t s_state;
  if (threadIdx.x == 0) s_state = 16u;  // "chunk 0 exhausted": the first warp to ask fetches a real one
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  uint32_t priv = kNoTile;
  for (;;) {
    uint32_t chunk = kNoTile, t = 0;
    if (lane == 0) chunk = next_ticket(P, &s_state, priv, t);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (chunk == kNoTile) break;
    (void)render_pixel<MODE, AOV>(P, pixel_of_chunk_tile(P, chunk, t, lane));
    __syncwarp();
  }
}

template <int MODE>
static cudaError_t launch_mode_persistent(const RenderParams& P, unsigned ctas, cudaStream_t stream, bool cta_queue) {
  if (cta_queue) {
    if (P.has_aov) raycast_persistent_cta<MODE, true><<<ctas, kThreads, 0, stream>>>(P);
    else raycast_persistent_cta<MODE, false><<<ctas, kThreads, 0, stream>>>(P);
  } else {
    if (P.has_aov) raycast_persistent<MODE, true><<<ctas, kThreads, 0, stream>>>(P);
    else raycast_persistent<MODE, false><<<ctas, kThreads, 0, stream>>>(P);
  }
  return cudaGetLastError();
}

// pad: bytes of (unused) dynamic shared memory per CTA (WX_OPT_SMEM_PAD) -- a measurement knob that lowers the number of
// resident CTAs per SM.
// long_stream != nullptr (and P.prev_list set): the long-tile kernel goes first, on that (high-priority) stream; the caller has
// ordered it after `stream`'s earlier work and joins it afterwards.
// main_grid == false: only the long-tile kernel (the first phase of a frame whose main grid is launched in row chunks).
template <int MODE>
static cudaError_t launch_mode(const RenderParams& P, dim3 grid, cudaStream_t stream, size_t pad, cudaStream_t long_stream, bool main_grid) {
  if (pad > 48 * 1024) {
    (void)cudaFuncSetAttribute(raycast_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
    (void)cudaFuncSetAttribute(raycast_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
  }
  if (long_stream != nullptr && P.prev_list != nullptr) {
    if (P.has_aov) raycast_long_tiles<MODE, true><<<P.sched_cap, kThreads, 0, long_stream>>>(P);
    else raycast_long_tiles<MODE, false><<<P.sched_cap, kThreads, 0, long_stream>>>(P);
  }
  if (!main_grid) return cudaGetLastError();
  if (P.has_aov) raycast_kernel<MODE, true><<<grid, kThreads, pad, stream>>>(P);
  else raycast_kernel<MODE, false><<<grid, kThreads, pad, stream>>>(P);
  return cudaGetLastError();
}
template <int MODE>
static cudaError_t launch_mode_tol(const RenderParams& P, dim3 grid, cudaStream_t stream, cudaStream_t long_stream, bool main_grid) {
  if (long_stream != nullptr && P.prev_list != nullptr) {
    if (P.has_aov) raycast_long_tiles_tol<MODE, true><<<P.sched_cap, kThreads, 0, long_stream>>>(P);
    else raycast_long_tiles_tol<MODE, false><<<P.sched_cap, kThreads, 0, long_stream>>>(P);
  }
  if (!main_grid) return cudaGetLastError();
  if (P.has_aov) raycast_kernel_tol<MODE, true><<<grid, kThreads, 0, stream>>>(P);
  else raycast_kernel_tol<MODE, false><<<grid, kThreads, 0, stream>>>(P);
  return cudaGetLastError();
}

// WX_OPT_KERNEL selects a work-queue kernel.  Measured on the 4K sphere frame: tiled 0.924 ms, warp-level queue 0.964 ms,
// CTA-level queue 1.985 ms (profiles/r2_cta_queue.txt) -- the CTA tail the queues remove is not what limits the tiled grid,
// so the simpler kernel is the default.

uint32_t raycast_tile_height() { return (uint32_t)kTileH; }
// The list costs one memset, a fork / join and a second (mostly empty) launch per launch, ~8 us: worth it on a 4K frame (0.76 ms:
// -5 %) and on one rank's share of a 4K frame (0.18 -> 0.155 ms), not on a 1080p frame (0.10 ms: +5...10 %,
// profiles/r2_longfirst.txt).  The criterion is therefore the size of the FRAME (times the cameras of a batch: config 5's 64 x 1080p
// launch gains 7 %), not of the launch.
constexpr uint64_t kLongFirstMinPixels = 1ull << 22;

// Fills the launch geometry of P (shard -> bands -> tile rows) and launches frames
// [P.cam_base, P.cam_base + n_cams) in render mode `render_mode` (the caller groups a camera batch
// by mode).  P.n_states is the total number of states behind P.states / P.s0.
// sched (optional): prepares the long-tiles-first list for this launch geometry (sched->prepare fills P.prev_* / P.next_*
// and returns the stream of the long-tile kernel, already ordered after `stream`; sched->finish joins it back).
cudaError_t launch_raycast(RenderParams& P, uint32_t n_cams, uint32_t render_mode, cudaStream_t stream, uint32_t* launches,
                           uint32_t* work_counter, uint32_t resident_ctas, const LaunchOptions& opt, TileSched* sched, const SchedCall* call) {
  *launches = 0;
  if (P.shard_count == 0) P.shard_count = 1, P.shard_index = 0;
  if (P.row_end == 0 || P.row_end > P.height) P.row_end = P.height;
  if (P.band_rows == 0 || P.shard_count == 1) {
    // a single shard owns every row of [row_base, row_end): one band of that height (rounded up to the tile height)
    if (P.row_base >= P.row_end) return cudaSuccess;
    P.band_rows = ((P.row_end - P.row_base + kTileH - 1) / kTileH) * kTileH;
    P.shard_count = 1, P.shard_index = 0;
    P.own_bands = 1;
  } else {
    // Band dealing over a row range: only from a row where the deal starts over (a multiple of one round of bands),
    // so that a band belongs to the same shard as in the whole frame.
    if (P.row_base % (P.shard_count * P.band_rows) != 0u) return cudaErrorInvalidValue;
    if (P.row_base >= P.row_end) return cudaSuccess;
    P.own_bands = shard_own_bands(P.row_end - P.row_base, P.shard_index, P.shard_count, P.band_rows);
  }
  P.tiles_x = (P.width + kTileW - 1) / kTileW;
  P.tile_rows_per_band = P.band_rows / kTileH;
  P.disp_w = (P.width / 8) * 8;
  P.disp_h = (P.height / 4) * 4;
  if (P.own_bands == 0 || n_cams == 0 || P.tiles_x == 0) return cudaSuccess;
  const uint64_t tile_rows = (uint64_t)P.tile_rows_per_band * P.own_bands;
  if (P.tiles_x > 0x7fffffffu || tile_rows > 65535u || n_cams > 65535u) return cudaErrorInvalidConfiguration;
  if (work_counter && opt.kernel != 0 && opt.march == kMarchExact) {
    const uint64_t own_rows = (uint64_t)P.own_bands * P.band_rows;
    P.chunks_x = (P.width + 31u) / 32u;
    P.chunks_y = (uint32_t)((own_rows + 15u) / 16u);
    const uint64_t n_chunks = (uint64_t)P.chunks_x * P.chunks_y * n_cams;
    if (n_chunks < (1ull << 27)) {  // the tile counter (16 per chunk) must fit 32 bits
      P.n_chunks = (uint32_t)n_chunks;
      P.work_counter = work_counter;
      cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(uint32_t), stream);
      if (e != cudaSuccess) return e;
      // one warp per tile at most; otherwise every resident CTA slot of the device
      const unsigned ctas = (unsigned)std::min<uint64_t>((n_chunks * 16 + WX_CTA_WARPS - 1) / WX_CTA_WARPS, resident_ctas ? resident_ctas : 148u * (unsigned)(WX_MIN_BLOCKS));
      *launches = 1;
      const bool cta_queue = opt.kernel == 2;
      switch (render_mode) {
        case 1: return launch_mode_persistent<1>(P, ctas, stream, cta_queue);
        case 2: return launch_mode_persistent<2>(P, ctas, stream, cta_queue);
        case 3: return launch_mode_persistent<3>(P, ctas, stream, cta_queue);
        case 4: return launch_mode_persistent<4>(P, ctas, stream, cta_queue);
        default: return launch_mode_persistent<0>(P, ctas, stream, cta_queue);
      }
    }
  }
  dim3 grid(P.tiles_x, (unsigned)tile_rows, n_cams);
  *launches = 1;
  cudaStream_t long_stream = nullptr;
  const uint64_t n_tiles = (uint64_t)P.tiles_x * tile_rows * n_cams;
  const int phase = call ? call->phase : kSchedWhole;
  P.frame_tile_rows = 0, P.tile_row_offset = 0;
  bool main_grid = true;
  // (modes 3 and 4 spend most of their time in secondary rays, which the primary ray's iteration count does not predict:
  // measured no gain there, profiles/r2_longfirst.txt)
  if (phase == kSchedChunk) {  // a row chunk of a frame whose lists kSchedLongOnly prepared (or declined to)
    if (sched != nullptr) {
      P.sched_threshold = opt.long_threshold;
      P.frame_tile_rows = call->frame_tile_rows, P.tile_row_offset = call->tile_row_offset;
      sched->attach(P);
    }
  } else if (sched != nullptr && opt.long_first && render_mode <= 2u && (uint64_t)P.width * P.height * n_cams >= kLongFirstMinPixels && n_tiles >= 512 &&
             n_tiles < (1ull << 31)) {
    P.sched_threshold = opt.long_threshold;
    cudaError_t e = sched->prepare(P, n_cams, (uint32_t)n_tiles, stream, &long_stream);
    if (e != cudaSuccess) return e;
    if (long_stream != nullptr) *launches = 2;
  }
  if (phase == kSchedLongOnly) {
    main_grid = false;
    *launches = long_stream != nullptr ? 1 : 0;
    if (long_stream == nullptr) return cudaSuccess;  // no list yet (or not applicable): the chunks render everything
  }
  cudaError_t le;
  if (opt.march != kMarchExact && render_mode != 2u) {  // mode 2 colours the iteration count: always exact
    switch (render_mode) {
      case 1: le = launch_mode_tol<1>(P, grid, stream, long_stream, main_grid); break;
      case 3: le = launch_mode_tol<3>(P, grid, stream, long_stream, main_grid); break;
      case 4: le = launch_mode_tol<4>(P, grid, stream, long_stream, main_grid); break;
      default: le = launch_mode_tol<0>(P, grid, stream, long_stream, main_grid); break;
    }
  } else {
    switch (render_mode) {
      case 1: le = launch_mode<1>(P, grid, stream, opt.smem_pad, long_stream, main_grid); break;
      case 2: le = launch_mode<2>(P, grid, stream, opt.smem_pad, long_stream, main_grid); break;
      case 3: le = launch_mode<3>(P, grid, stream, opt.smem_pad, long_stream, main_grid); break;
      case 4: le = launch_mode<4>(P, grid, stream, opt.smem_pad, long_stream, main_grid); break;
      default: le = launch_mode<0>(P, grid, stream, opt.smem_pad, long_stream, main_grid); break;  // Gray and the shader's `default:` arms
    }
  }
  if (le == cudaSuccess && long_stream != nullptr && phase == kSchedWhole) le = sched->finish(stream);  // (kSchedLongOnly: the caller joins)
  return le;
}

}  // namespace wx
