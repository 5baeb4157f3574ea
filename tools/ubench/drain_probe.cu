// Micro-benchmark for the streamed read-back's copy kernel: device -> mapped pinned host memory with SM stores, no render kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o drain_probe drain_probe.cu && ./drain_probe
// Prints GB/s for CTA shapes (threads per CTA x CTAs), with and without a 32-register cap, and for cudaMemcpyAsync.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int UNROLL>
__device__ __forceinline__ void copy_range(const uint4* s4, uint4* d4, uint32_t n16) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x; base < n16; base += UNROLL * stride) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint32_t k = base + u * stride + threadIdx.x;
      if (k < n16) v[u] = __ldcg(s4 + k);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint32_t k = base + u * stride + threadIdx.x;
      if (k < n16) d4[k] = v[u];
    }
  }
}
template <int UNROLL>
__global__ void copy_plain(const uint4* s, uint4* d, uint32_t n16, uint32_t pieces) {
  const uint32_t per = n16 / pieces;
  for (uint32_t p = 0; p < pieces; ++p) copy_range<UNROLL>(s + (size_t)p * per, d + (size_t)p * per, per);
}
template <int UNROLL>
__global__ void __maxnreg__(32) copy_capped(const uint4* s, uint4* d, uint32_t n16, uint32_t pieces) {
  const uint32_t per = n16 / pieces;
  for (uint32_t p = 0; p < pieces; ++p) copy_range<UNROLL>(s + (size_t)p * per, d + (size_t)p * per, per);
}
int main() {
  const size_t bytes = 3840ull * 2160 * 4;
  const uint32_t n16 = (uint32_t)(bytes / 16);
  uint4 *src, *dsth, *dstd;
  cudaMalloc(&src, bytes); cudaMalloc(&dstd, bytes); cudaHostAlloc(&dsth, bytes, cudaHostAllocDefault);
  cudaMemset(src, 7, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](const char* what, auto launch) {
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("%-44s %8.3f ms  %7.1f GB/s  (%s)\n", what, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  time("cudaMemcpyAsync D2H", [&] { cudaMemcpyAsync(dsth, src, bytes, cudaMemcpyDeviceToHost, 0); });
  for (uint32_t pieces : {1u, 64u}) {
    printf("pieces %u (each CTA walks every piece)\n", pieces);
    time("host  16 x 256 unroll 4", [&] { copy_plain<4><<<16, 256>>>(src, dsth, n16, pieces); });
    time("host   4 x 256 unroll 4", [&] { copy_plain<4><<<4, 256>>>(src, dsth, n16, pieces); });
    time("host 128 x  32 unroll 2", [&] { copy_plain<2><<<128, 32>>>(src, dsth, n16, pieces); });
    time("host 128 x  32 unroll 2, 32 regs", [&] { copy_capped<2><<<128, 32>>>(src, dsth, n16, pieces); });
    time("host 128 x  32 unroll 4", [&] { copy_plain<4><<<128, 32>>>(src, dsth, n16, pieces); });
    time("host 296 x  32 unroll 2", [&] { copy_plain<2><<<296, 32>>>(src, dsth, n16, pieces); });
    time("host 592 x  32 unroll 4", [&] { copy_plain<4><<<592, 32>>>(src, dsth, n16, pieces); });
    time("host 128 x  64 unroll 4", [&] { copy_plain<4><<<128, 64>>>(src, dsth, n16, pieces); });
    time("device 128 x 32 unroll 2", [&] { copy_plain<2><<<128, 32>>>(src, dstd, n16, pieces); });
  }
  return 0;
}
