// wx_capture.cu -- the step after the raycast in the reference's frame: capture for recording.
// src/render/wgpu_context.rs:374-405 copies the frame texture to a buffer and src/render/recorder.rs:20-37
// turns the RGBA8 frame into RGB8, passing every colour byte through linear_to_srgb (:132-140) before the
// encoder sees it.  Here that is one kernel over the frame that wx_render left on device 0: 4 bytes in,
// 3 bytes out per pixel, the transfer function as a 256-entry table in constant memory.
#include <cmath>

#include "wx_device.cuh"
#include "wx_internal.h"

namespace wx {

__constant__ uint8_t c_srgb_lut[256];

// recorder.rs:132-140, evaluated in binary32 like the reference (f32::powf, then round half away from zero)
void build_srgb_lut(uint8_t lut[256]) {
  for (int v = 0; v < 256; ++v) {
    const float c = (float)v / 255.0f;
    const float s = c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
    lut[v] = (uint8_t)roundf(s * 255.0f);
  }
}

// One thread per 4 pixels: 16 bytes in (one uint4), 12 bytes out (three u32), both coalesced.
__global__ void __launch_bounds__(256) srgb_rgb8_kernel(const uint4* __restrict__ rgba4, uint32_t* __restrict__ rgb3, size_t n_quads,
                                                         const uchar4* __restrict__ rgba_tail, uint8_t* __restrict__ rgb_tail,
                                                         uint32_t n_tail) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n_quads) {
    const uint4 p = __ldg(rgba4 + q);
    const uint32_t w[4] = {p.x, p.y, p.z, p.w};
    uint8_t o[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[3 * k + 0] = c_srgb_lut[w[k] & 255u];
      o[3 * k + 1] = c_srgb_lut[(w[k] >> 8) & 255u];
      o[3 * k + 2] = c_srgb_lut[(w[k] >> 16) & 255u];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
      rgb3[3 * q + k] = (uint32_t)o[4 * k] | ((uint32_t)o[4 * k + 1] << 8) | ((uint32_t)o[4 * k + 2] << 16) | ((uint32_t)o[4 * k + 3] << 24);
  }
  if (q < n_tail) {  // the last n % 4 pixels
    const uchar4 p = rgba_tail[q];
    rgb_tail[3 * q + 0] = c_srgb_lut[p.x], rgb_tail[3 * q + 1] = c_srgb_lut[p.y], rgb_tail[3 * q + 2] = c_srgb_lut[p.z];
  }
}

cudaError_t launch_srgb_rgb8(const uint8_t* rgba_dev, uint8_t* rgb_dev, size_t n_pixels, cudaStream_t stream) {
  static bool lut_ready[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 64 && !lut_ready[dev]) {
    uint8_t lut[256];
    build_srgb_lut(lut);
    e = cudaMemcpyToSymbolAsync(c_srgb_lut, lut, 256, 0, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(stream);  // `lut` is a stack buffer
    if (e != cudaSuccess) return e;
    lut_ready[dev] = true;
  }
  if (n_pixels == 0) return cudaSuccess;
  const size_t n_quads = n_pixels / 4;
  const uint32_t n_tail = (uint32_t)(n_pixels % 4);
  const size_t threads = n_quads > n_tail ? n_quads : n_tail;
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  srgb_rgb8_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(rgba_dev), reinterpret_cast<uint32_t*>(rgb_dev), n_quads,
                                               reinterpret_cast<const uchar4*>(rgba_dev) + 4 * n_quads, rgb_dev + 12 * n_quads, n_tail);
  return cudaGetLastError();
}

}  // namespace wx
