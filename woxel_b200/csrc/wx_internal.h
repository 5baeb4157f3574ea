// wx_internal.h -- host-side declarations shared by the CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wx {
struct RenderParams;
constexpr int kBandRowsMultiple = 8;  // band_rows of a WxShard must be a multiple of the CTA tile height
// Band b (rows [b*band_rows, (b+1)*band_rows)) belongs to shard b % count.
inline uint32_t shard_own_bands(uint32_t height, uint32_t index, uint32_t count, uint32_t band_rows) {
  const uint32_t total = (height + band_rows - 1) / band_rows;
  return total > index ? (total - index + count - 1) / count : 0;
}
// work_counter: one zero-initialisable u32 of device memory private to this launch (nullptr = tiled kernel);
// resident_ctas: CTAs the device holds at once (SM count x CTAs per SM), the persistent grid.
// Per-context options of wx_set_option (include/woxel_b200.h) that reach the launcher.
struct LaunchOptions {
  int kernel = 0;        // WX_OPT_KERNEL: 0 tiled grid, 1 warp-level tile queue, 2 CTA-level chunk queue
  size_t smem_pad = 0;   // WX_OPT_SMEM_PAD
  int march = 0;         // WX_OPT_MARCH: 0 exact, 1 tolerance mode
  int long_first = 1;    // WX_OPT_LONG_FIRST: tiles that held long rays in the previous launch of the same geometry start first
  uint32_t long_threshold = 96;  // WX_OPT_LONG_THRESHOLD: ~3x the mean primary ray of the benchmark scenes (20-50 iterations)
};

// Long-tiles-first state of one launch geometry on one stream (wx_api.cu owns the objects; wx_raycast.cu drives them).
struct TileSched {
  virtual cudaError_t prepare(RenderParams& P, uint32_t n_cams, uint32_t n_tiles, cudaStream_t stream, cudaStream_t* long_stream) = 0;
  virtual void attach(RenderParams& P) = 0;  // the lists prepare() set up for this frame, for a launch that renders part of it
  virtual cudaError_t finish(cudaStream_t stream) = 0;
  virtual ~TileSched() {}
};
// How a launch takes part in the long-tiles-first scheme: the whole frame in one launch (list prepared, long-tile kernel, main
// grid, joined); or a frame whose main grid is launched in row chunks (wx_render's pipelined read-back): first kSchedLongOnly
// with the whole frame's geometry (list prepared, long-tile kernel only), then one kSchedChunk launch per row chunk.
enum { kSchedWhole = 0, kSchedLongOnly = 1, kSchedChunk = 2 };
struct SchedCall {
  int phase = kSchedWhole;
  uint32_t frame_tile_rows = 0, tile_row_offset = 0;  // kSchedChunk
};
cudaError_t launch_raycast(RenderParams& P, uint32_t n_cams, uint32_t render_mode, cudaStream_t stream, uint32_t* launches,
                           uint32_t* work_counter, uint32_t resident_ctas, const LaunchOptions& opt, TileSched* sched = nullptr,
                           const SchedCall* call = nullptr);
constexpr uint32_t kCtasPerSm = 9;
uint32_t raycast_tile_height();  // rows of a CTA's footprint (8 in the default build)
// RGBA8 (linear) -> RGB8 through recorder.rs' linear_to_srgb; rgba_dev must be 16-byte aligned, rgb_dev 4-byte aligned.
cudaError_t launch_srgb_rgb8(const uint8_t* rgba_dev, uint8_t* rgb_dev, size_t n_pixels, cudaStream_t stream);
void build_srgb_lut(uint8_t lut[256]);
}  // namespace wx
struct WxTreeDesc;
namespace wx {
// wx_sdf.cu: compute_sdf on the current device; info = max distance per level [0..2], values that did not fit [3], relaxation rounds [4]
// dev != nullptr: instead of the host tables, fill the raycast kernel's own tables (caller-allocated device memory of the
// current device: e5 n5*32768 u32, e4 n4*4096 u32, l3 n3*512 u8); info[3] then counts leaf distances above 255.
struct SdfDeviceTargets {
  uint32_t* e5;
  uint32_t* e4;
  uint8_t* l3;
};
cudaError_t compute_sdf_device(const WxTreeDesc& d, uint32_t* tab5_out, uint32_t* tab4_out, void* tab3_out, uint32_t tab3_elem_bytes,
                               uint32_t info[5], float* device_ms, cudaStream_t stream, const SdfDeviceTargets* dev = nullptr);
}  // namespace wx
