// fog_gen.cu -- BENCH / TEST INFRASTRUCTURE, not part of the product library and not part of the reference.
//
// Evaluates the value-noise fog of BASELINE config 4 (active iff fbm(c / 256) > tau, 5 octaves,
// lacunarity 2, gain 0.5, integer-hash lattice seeded 0x9E3779B9, evaluated in f64) on the GPU, so that the 2048^3
// volume (8.6 G voxels, up to 16.7 M leaves) can be built on the box in seconds; the host builder
// (woxel_b200/host/procedural.cpp::fbm_fog, the same arithmetic in the same order) needs minutes for it.
// Compiled with -fmad=false: every f64 operation is an IEEE operation, so the masks equal the host builder's
// bit for bit (tests/test_scenegen_gpu.py).
//
// Output: the value masks of ALL (2*half/8)^3 leaf cells of [-half, half)^3 in the reference's DFS node order
// (N5s sorted by origin x,y,z; N4s by ascending offset; leaves by ascending offset -- vdb345.rs:134-158), 8 u64 words per
// leaf, bit o = (x<<6)|(y<<3)|z of word o>>6.  tests/scenes.py drops the empty leaves and derives the upper levels.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ uint32_t hash3(int32_t x, int32_t y, int32_t z) {
  uint32_t h = 0x9E3779B9u;
  h ^= (uint32_t)x * 0x85EBCA6Bu, h = (h << 13) | (h >> 19), h *= 0xC2B2AE35u;
  h ^= (uint32_t)y * 0x27D4EB2Fu, h = (h << 13) | (h >> 19), h *= 0xC2B2AE35u;
  h ^= (uint32_t)z * 0x165667B1u, h = (h << 13) | (h >> 19), h *= 0xC2B2AE35u;
  h ^= h >> 16, h *= 0x85EBCA6Bu, h ^= h >> 13, h *= 0xC2B2AE35u, h ^= h >> 16;
  return h;
}
__device__ __forceinline__ double lattice(int32_t x, int32_t y, int32_t z) { return (double)hash3(x, y, z) * (1.0 / 4294967296.0); }
__device__ __forceinline__ double smooth(double t) { return t * t * (3.0 - 2.0 * t); }
__device__ __forceinline__ double lerp(double a, double b, double t) { return a + (b - a) * t; }

__device__ double value_noise(double x, double y, double z) {
  const double fx = floor(x), fy = floor(y), fz = floor(z);
  const int32_t ix = (int32_t)fx, iy = (int32_t)fy, iz = (int32_t)fz;
  const double tx = smooth(x - fx), ty = smooth(y - fy), tz = smooth(z - fz);
  const double x00 = lerp(lattice(ix, iy, iz), lattice(ix + 1, iy, iz), tx), x10 = lerp(lattice(ix, iy + 1, iz), lattice(ix + 1, iy + 1, iz), tx);
  const double x01 = lerp(lattice(ix, iy, iz + 1), lattice(ix + 1, iy, iz + 1), tx), x11 = lerp(lattice(ix, iy + 1, iz + 1), lattice(ix + 1, iy + 1, iz + 1), tx);
  return lerp(lerp(x00, x10, ty), lerp(x01, x11, ty), tz);
}

__device__ double fbm(double x, double y, double z) {
  double sum = 0.0, amp = 0.5, norm = 0.0;
  for (int o = 0; o < 5; ++o) {
    sum += amp * value_noise(x, y, z);
    norm += amp;
    x *= 2.0, y *= 2.0, z *= 2.0;
    amp *= 0.5;
  }
  return sum / norm;
}

// One thread per (leaf, mask word): word w of a leaf holds the 64 voxels of its x-slice w.
__global__ void fog_masks_kernel(int32_t half, uint32_t hn, double tau, uint64_t first, uint64_t count, uint64_t* masks, unsigned long long* n_active) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t word = 0;
  if (i < count) {
    const uint64_t g = first + i, t = g >> 3;
    const uint32_t w = (uint32_t)(g & 7u);
    const uint64_t per5 = (uint64_t)hn * hn * hn * 4096u;
    const uint32_t i5 = (uint32_t)(t / per5);
    const uint64_t r = t % per5;
    const uint32_t n4 = (uint32_t)(r >> 12), l = (uint32_t)(r & 4095u);
    const uint32_t ax = n4 / (hn * hn), ay = (n4 / hn) % hn, az = n4 % hn;
    const int32_t x = (((i5 >> 2) & 1u) ? 0 : -half) + (int32_t)(ax * 128u + (l >> 8) * 8u + w);
    const int32_t y0 = (((i5 >> 1) & 1u) ? 0 : -half) + (int32_t)(ay * 128u + ((l >> 4) & 15u) * 8u);
    const int32_t z0 = ((i5 & 1u) ? 0 : -half) + (int32_t)(az * 128u + (l & 15u) * 8u);
    for (uint32_t b = 0; b < 64u; ++b) {
      const int32_t y = y0 + (int32_t)(b >> 3), z = z0 + (int32_t)(b & 7u);
      if (fbm((x + 0.5) / 256.0, (y + 0.5) / 256.0, (z + 0.5) / 256.0) > tau) word |= 1ull << b;
    }
    masks[i] = word;
  }
  unsigned long long c = (unsigned long long)__popcll(word);
  for (int s = 16; s; s >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, s);
  if ((threadIdx.x & 31u) == 0u && c) atomicAdd(n_active, c);
}

}  // namespace

// masks_out: host memory, (2*half/8)^3 x 8 words.  half: a multiple of 128, at most 4096.  Returns 0 or a negative cudaError_t.
extern "C" int wxs_fog_masks(int device, int32_t half, double tau, uint64_t* masks_out, uint64_t* n_active_out) {
  if (half <= 0 || half > 4096 || (half & 127) || !masks_out) return -1;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return -(int)e;
  const uint32_t hn = (uint32_t)half / 128u;
  const uint64_t words = 8ull * hn * hn * hn * 4096ull * 8ull;
  const uint64_t slab = 1ull << 27;  // 1 GiB of masks per launch: bounded device memory and launch time
  uint64_t* d = nullptr;
  unsigned long long* d_n = nullptr;
  if ((e = cudaMalloc(&d, (size_t)(words < slab ? words : slab) * 8)) != cudaSuccess) return -(int)e;
  if ((e = cudaMalloc(&d_n, 8)) != cudaSuccess) { cudaFree(d); return -(int)e; }
  cudaMemset(d_n, 0, 8);
  for (uint64_t first = 0; first < words && e == cudaSuccess; first += slab) {
    const uint64_t count = words - first < slab ? words - first : slab;
    fog_masks_kernel<<<(unsigned)((count + 127) / 128), 128>>>(half, hn, tau, first, count, d, d_n);
    e = cudaMemcpy(masks_out + first, d, (size_t)count * 8, cudaMemcpyDeviceToHost);
  }
  unsigned long long n = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&n, d_n, 8, cudaMemcpyDeviceToHost);
  if (n_active_out) *n_active_out = n;
  cudaFree(d), cudaFree(d_n);
  return e == cudaSuccess ? 0 : -(int)e;
}

// Active-voxel fraction only (tau calibration): no masks leave the device.
extern "C" int wxs_fog_occupancy(int device, int32_t half, double tau, double* occupancy) {
  if (half <= 0 || half > 4096 || (half & 127) || !occupancy) return -1;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return -(int)e;
  const uint32_t hn = (uint32_t)half / 128u;
  const uint64_t words = 8ull * hn * hn * hn * 4096ull * 8ull;
  const uint64_t slab = 1ull << 27;
  uint64_t* d = nullptr;
  unsigned long long* d_n = nullptr;
  if ((e = cudaMalloc(&d, (size_t)(words < slab ? words : slab) * 8)) != cudaSuccess) return -(int)e;
  if ((e = cudaMalloc(&d_n, 8)) != cudaSuccess) { cudaFree(d); return -(int)e; }
  cudaMemset(d_n, 0, 8);
  for (uint64_t first = 0; first < words; first += slab) {
    const uint64_t count = words - first < slab ? words - first : slab;
    fog_masks_kernel<<<(unsigned)((count + 127) / 128), 128>>>(half, hn, tau, first, count, d, d_n);
  }
  unsigned long long n = 0;
  e = cudaMemcpy(&n, d_n, 8, cudaMemcpyDeviceToHost);
  *occupancy = (double)n / (8.0 * (double)half * (double)half * (double)half);
  cudaFree(d), cudaFree(d_n);
  return e == cudaSuccess ? 0 : -(int)e;
}
