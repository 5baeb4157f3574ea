// l2_peak.cu -- the L2 read bandwidth of one B200, the second roofline denominator SURVEY 8(d) asks for ("a measured L2
// peak") beside MEASURED_PEAKS.json's HBM copy figure.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_peak l2_peak.cu && ./l2_peak
// Every thread streams 16-byte loads over a buffer that stays resident in the 126 MB L2 (16 ... 96 MB), 148 x 8 CTAs of 256
// threads, grid-stride, many passes inside one launch; a 2 GB buffer gives the DRAM figure of the same kernel for comparison.
// ld.global.nc with L1::no_allocate so that the 256 KB L1s do not answer instead of the L2.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) stream_read(const uint4* __restrict__ p, size_t n16, int passes, unsigned* sink) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int k = 0; k < passes; ++k) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {  // four independent loads in flight per thread
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(p + i + u * stride));
#pragma unroll
      for (int u = 0; u < 4; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    for (; i < n16; i += stride) {
      uint4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

int main() {
  const size_t MB = 1 << 20;
  const size_t sizes[] = {16 * MB, 32 * MB, 48 * MB, 64 * MB, 96 * MB, 2048 * MB};
  uint4* buf;
  unsigned* sink;
  if (cudaMalloc(&buf, 2048 * MB) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) return 1;
  cudaMemset(buf, 1, 2048 * MB);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  for (size_t bytes : sizes) {
    const int passes = bytes > 1024 * MB ? 2 : (int)(4096 * MB / bytes);
    const size_t n16 = bytes / 16;
    stream_read<<<148 * 8, 256>>>(buf, n16, 2, sink);  // warm the L2
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      stream_read<<<148 * 8, 256>>>(buf, n16, passes, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    printf("read %5zu MB x %4d passes: %8.3f ms  %8.1f GB/s%s\n", bytes / MB, passes, best, (double)bytes * passes / best / 1e6,
           bytes > 1024 * MB ? "  (DRAM)" : "  (L2-resident)");
  }
  return cudaGetLastError() != cudaSuccess;
}
