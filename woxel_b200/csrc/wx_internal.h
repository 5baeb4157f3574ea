// wx_internal.h -- host-side declarations shared by the CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wx {
struct RenderParams;
constexpr int kBandRowsMultiple = 8;  // band_rows of a WxShard must be a multiple of the CTA tile height
cudaError_t launch_raycast(RenderParams& P, uint32_t n_cams, uint32_t render_mode, cudaStream_t stream, uint32_t* launches);
}  // namespace wx
