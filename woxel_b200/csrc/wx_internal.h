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
cudaError_t launch_raycast(RenderParams& P, uint32_t n_cams, uint32_t render_mode, cudaStream_t stream, uint32_t* launches,
                           uint32_t* work_counter, uint32_t resident_ctas);
constexpr uint32_t kCtasPerSm = 9;
}  // namespace wx
