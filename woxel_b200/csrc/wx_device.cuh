// wx_device.cuh -- device-side data layout and the per-ray march for sm_100a.
//
// What this replaces: the body of src/shaders/raycast.comp.wgsl of the reference (cp_main :60-68,
// hdda_ray :84-126, ray_trace :152-265, reflect_ray2/1 :267-342, get_vdb_leaf_* :360-494).  It is
// not a translation of that file: the reference reads 2 mask words + 1 atlas texel per tree level
// (three 3-D textures + five mask buffers); here every level is ONE table of 32-bit entries (or a
// 4/8-bit brick at the leaf) so a lookup is one load per level, and the bottom-up parent-origin
// cache (:360-396) is three XORs against the previously visited voxel.
//
// Arithmetic contract: IEEE binary32, one rounding per source operation, in the operation order of
// the WGSL (this translation unit is compiled with -fmad=false; division and sqrt are the IEEE
// ones).  Results are bit-identical to the CPU oracle, which is how parity is proven.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/woxel_b200.h"

namespace wx {

// ---------------------------------------------------------------------------------------------
// Device tree.  All arrays are 256-B aligned (cudaMalloc); node records are 128 KB / 16 KB /
// 256 B (4-bit leaves) so every node starts on a 128-B line.
//
//   e5[n5][32768], e4[n4][4096] : u32 entry per slot
//        bit 31 set  -> child; low 31 bits = child node index (reference DFS index)
//        bit 31 clear-> tile;  value = SDF distance in cells of the level, 0 = active tile (hit)
//        (value-mask is tested before child-mask in the reference, raycast.comp.wgsl:431-437; the
//         packer applies that priority once, at upload)
//   l3[n3] : one brick per leaf, LEAF_BITS per voxel in offset order (x<<6 | y<<3 | z),
//        0 = active voxel (hit), else SDF distance in voxels.  LEAF_BITS = 4 when every distance
//        is <= 15 (always true for leaves that contain an active voxel), else 8, else 32.
// ---------------------------------------------------------------------------------------------
struct DevTree {
  const uint32_t* __restrict__ e5;
  const uint32_t* __restrict__ e4;
  const uint32_t* __restrict__ l3;
  const int4* __restrict__ origins_g;  // n5 entries (x,y,z,0); used when n5 > kInlineOrigins
  uint32_t n5, n4, n3;
  uint32_t leaf_bits;
  int4 origins_c[8];  // first 8 origins, read from the constant bank
};
constexpr uint32_t kInlineOrigins = 8;
constexpr uint32_t kChildFlag = 0x80000000u;

struct AovPtrs {
  uint8_t* state;
  int32_t* voxel;
  int32_t* leaf;
  uint8_t* level;
  uint32_t* iters;
  float* depth;
  uint8_t* mask;
  float* pos;
};

struct RenderParams {
  DevTree tree;
  WxState s0;              // the state when n_states == 1 (lives in the constant bank)
  const WxState* states;   // device array when n_states > 1
  uint32_t n_states;
  uint32_t cam_base;        // first frame of this launch
  uint32_t width, height;
  uint32_t disp_w, disp_h;  // (W/8)*8, (H/4)*4 : what the reference dispatches (wgpu_context.rs:281)
  uint32_t shard_index, shard_count, band_rows;
  uint32_t own_bands;       // number of row bands this launch renders
  uint32_t tiles_x;         // blocks per row of tiles
  uint32_t tile_rows_per_band;
  uchar4* rgba;
  AovPtrs aov;
  uint32_t has_aov;
};

// ---------------------------------------------------------------------------------------------
// small vector helpers (plain IEEE ops; no contraction in this TU)
// ---------------------------------------------------------------------------------------------
struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 splat(float s) { return V3{s, s, s}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 operator+(V3 a, float s) { return V3{a.x + s, a.y + s, a.z + s}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 normalize3(V3 a) { return a / sqrtf(dot3(a, a)); }
__device__ __forceinline__ V3 max3(V3 a, V3 b) { return V3{fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }
__device__ __forceinline__ V3 min3(V3 a, V3 b) { return V3{fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
__device__ __forceinline__ V3 mix3(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }  // e1*(1-e3)+e2*e3
__device__ __forceinline__ V3 sign11(V3 d) {
  return V3{d.x < 0.f ? -1.f : 1.f, d.y < 0.f ? -1.f : 1.f, d.z < 0.f ? -1.f : 1.f};
}

// ---------------------------------------------------------------------------------------------
// Tree cursor: the path to the voxel visited last.  depth = level of the last result
// (the reference's VdbLeaf.num_parents): 0 nothing cached, 1 n5, 2 n5+n4, 3 n5+n4+n3.
// ---------------------------------------------------------------------------------------------
struct Cursor {
  int lx, ly, lz;
  uint32_t n5, n4, n3;
  uint32_t depth;
};

struct Lookup {
  uint32_t dist;   // 0 => hit
  float cell;      // edge of one cell of the level the lookup ended on: 4096, 128, 8, 1
};

__device__ __forceinline__ uint32_t leaf_dist(const DevTree& T, uint32_t n3, int x, int y, int z) {
  const uint32_t o3 = ((uint32_t)(x & 7) << 6) | ((uint32_t)(y & 7) << 3) | (uint32_t)(z & 7);
  if (T.leaf_bits == 4) {
    const uint32_t w = __ldg(T.l3 + (size_t)n3 * 64 + (o3 >> 3));
    return (w >> ((o3 & 7) * 4)) & 15u;
  } else if (T.leaf_bits == 8) {
    const uint32_t w = __ldg(T.l3 + (size_t)n3 * 128 + (o3 >> 2));
    return (w >> ((o3 & 3) * 8)) & 255u;
  }
  return __ldg(T.l3 + (size_t)n3 * 512 + o3);
}

// L(pos): pure function of pos; the cursor only shortens the walk.
__device__ __forceinline__ Lookup lookup(const DevTree& T, Cursor& c, int x, int y, int z) {
  const uint32_t diff = (uint32_t)((x ^ c.lx) | (y ^ c.ly) | (z ^ c.lz));
  c.lx = x, c.ly = y, c.lz = z;
  uint32_t e;
  if (c.depth == 3 && diff < 8u) goto leaf;
  if (c.depth >= 2 && diff < 128u) goto node4;
  if (c.depth >= 1 && diff < 4096u) goto node5;
  {  // root: first origin equal to (pos >> 12) << 12   (raycast.comp.wgsl:398-413)
    const int gx = (x >> 12) << 12, gy = (y >> 12) << 12, gz = (z >> 12) << 12;
    uint32_t found = 0xffffffffu;
    const uint32_t nc = T.n5 < kInlineOrigins ? T.n5 : kInlineOrigins;
#pragma unroll
    for (uint32_t i = 0; i < kInlineOrigins; ++i) {
      const int4 o = T.origins_c[i];
      if (i < nc && found == 0xffffffffu && o.x == gx && o.y == gy && o.z == gz) found = i;
    }
    if (found == 0xffffffffu) {
      for (uint32_t i = kInlineOrigins; i < T.n5; ++i) {
        const int4 o = __ldg(T.origins_g + i);
        if (o.x == gx && o.y == gy && o.z == gz) {
          found = i;
          break;
        }
      }
    }
    if (found == 0xffffffffu) {
      c.depth = 0;
      return Lookup{1u, 4096.f};
    }
    c.n5 = found;
  }
node5:
  e = __ldg(T.e5 + (size_t)c.n5 * 32768u +
            ((((uint32_t)(x & 4095) >> 7) << 10) | (((uint32_t)(y & 4095) >> 7) << 5) | ((uint32_t)(z & 4095) >> 7)));
  if (!(e & kChildFlag)) {
    c.depth = 1;
    return Lookup{e, 128.f};
  }
  c.n4 = e & ~kChildFlag;
node4:
  e = __ldg(T.e4 + (size_t)c.n4 * 4096u +
            ((((uint32_t)(x & 127) >> 3) << 8) | (((uint32_t)(y & 127) >> 3) << 4) | ((uint32_t)(z & 127) >> 3)));
  if (!(e & kChildFlag)) {
    c.depth = 2;
    return Lookup{e, 8.f};
  }
  c.n3 = e & ~kChildFlag;
leaf:
  c.depth = 3;
  return Lookup{leaf_dist(T, c.n3, x, y, z), 1.f};
}

// ---------------------------------------------------------------------------------------------
// hdda_ray (raycast.comp.wgsl:84-126)
// ---------------------------------------------------------------------------------------------
struct HitOut {
  uint32_t state;  // 0 hit, 1 out of bounds, 2 max steps
  V3 p;
  uint32_t mask;   // bit0 x, bit1 y, bit2 z
  uint32_t i;
  uint32_t level;  // num_parents of the last lookup
  uint32_t n3;     // leaf index of the last lookup when level == 3
};

constexpr uint32_t kMaxRaySteps = 1000u;

__device__ __forceinline__ HitOut hdda_ray(const DevTree& T, V3 src, V3 dir) {
  V3 p = src;
  const V3 step = sign11(dir);
  const V3 step01 = max3(splat(0.f), step);
  const V3 idir = V3{1.f / dir.x, 1.f / dir.y, 1.f / dir.z};
  const V3 nudge = 4e-4f * step;
  uint32_t mask = 0;
  Cursor c{0, 0, 0, 0, 0, 0, 0};
  HitOut out;
  uint32_t i = 0;
  for (; i < kMaxRaySteps; ++i) {
    const int x = __float2int_rd(p.x), y = __float2int_rd(p.y), z = __float2int_rd(p.z);
    const Lookup l = lookup(T, c, x, y, z);
    if (l.dist == 0u) {
      out.state = 0u;
      break;
    }
    if (4096.f < fabsf(p.x) || 4096.f < fabsf(p.y) || 4096.f < fabsf(p.z)) {
      out.state = 1u;
      break;
    }
    const float size = (float)l.dist * l.cell;
    // modulo_vec3f(p, size) = p - size * floor(p / size)
    const V3 m = V3{p.x - size * floorf(p.x / size), p.y - size * floorf(p.y / size), p.z - size * floorf(p.z / size)};
    const V3 tmax = idir * (size * step01 - m);
    const float t = fminf(fminf(tmax.x, tmax.y), tmax.z);
    p = p + t * dir;
    const bool bx = (tmax.x <= tmax.y) && (tmax.x <= tmax.z);
    const bool by = (tmax.y <= tmax.z) && (tmax.y <= tmax.x);
    const bool bz = (tmax.z <= tmax.x) && (tmax.z <= tmax.y);
    mask = (uint32_t)bx | ((uint32_t)by << 1) | ((uint32_t)bz << 2);
    p = p + nudge * V3{bx ? 1.f : 0.f, by ? 1.f : 0.f, bz ? 1.f : 0.f};
  }
  if (i == kMaxRaySteps) out.state = 2u;
  out.p = p;
  out.mask = mask;
  out.i = i;
  out.level = c.depth;
  out.n3 = c.n3;
  return out;
}

// secondary rays share one out-of-line copy of the march
static __device__ __noinline__ HitOut hdda_ray_secondary(const DevTree& T, V3 src, V3 dir) { return hdda_ray(T, src, dir); }

// ---------------------------------------------------------------------------------------------
// shading (raycast.comp.wgsl:144-342)
// ---------------------------------------------------------------------------------------------
#define WX_K_D 0.7f
#define WX_K_A 0.3f
#define WX_REFLECTIVITY 0.9f
#define WX_WALL_I 0.1f
#define WX_BASE_COLOR (V3{0.4f, 0.2f, 0.2f})
#define WX_AMBIENT_COLOR (V3{0.4f, 0.4f, 0.3f})

__device__ __forceinline__ V3 maskf(uint32_t m) { return V3{(float)(m & 1u), (float)((m >> 1) & 1u), (float)((m >> 2) & 1u)}; }
__device__ __forceinline__ V3 sun_rgb(const WxState& s) { return V3{s.sun_color[0], s.sun_color[1], s.sun_color[2]}; }
__device__ __forceinline__ V3 sun_dir(const WxState& s) { return V3{s.sun_dir[0], s.sun_dir[1], s.sun_dir[2]}; }

// colour of an out-of-bounds secondary ray (:296-306, :328-338)
__device__ __forceinline__ V3 wall_flat(V3 N) {
  const V3 Np = max3(splat(0.f), N);
  const V3 Nn = -min3(splat(0.f), N);
  V3 r = V3{WX_WALL_I, 0.f, 0.f} * Np.x;
  r = r + V3{0.f, WX_WALL_I, 0.f} * Np.y;
  r = r + V3{0.f, 0.f, WX_WALL_I} * Np.z;
  r = r + V3{WX_WALL_I, WX_WALL_I, 0.f} * Nn.x;
  r = r + V3{0.f, WX_WALL_I, WX_WALL_I} * Nn.y;
  r = r + V3{WX_WALL_I, 0.f, WX_WALL_I} * Nn.z;
  return r;
}

// BASE + I * sun, with the sun term x0.05 when a shadow ray finds an occluder (:199-208, :280-290, :317-325)
__device__ __forceinline__ V3 sun_lit(const DevTree& T, const WxState& s, const HitOut& hit, V3 step, V3 N) {
  float I = s.sun_color[3] * WX_K_D * dot3(-sun_dir(s), N);
  I = fmaxf(0.0f, I);
  if (I != 0.0f && hdda_ray_secondary(T, hit.p - (4e-2f * step) * maskf(hit.mask), -sun_dir(s)).state == 0u)
    return WX_BASE_COLOR + (I * sun_rgb(s)) * 0.05f;
  return WX_BASE_COLOR + I * sun_rgb(s);
}

__device__ __forceinline__ V3 reflect_ray1(const DevTree& T, const WxState& s, V3 src, V3 dir) {
  const HitOut hit = hdda_ray_secondary(T, src, dir);
  const V3 step = sign11(dir);
  if (hit.state == 0u) return sun_lit(T, s, hit, step, normalize3((-step) * maskf(hit.mask)));
  if (hit.state == 1u) return wall_flat(normalize3((-step) * maskf(hit.mask)));
  return dir;
}

__device__ __forceinline__ V3 reflect_ray2(const DevTree& T, const WxState& s, V3 src, V3 dir) {
  const HitOut hit = hdda_ray_secondary(T, src, dir);
  const V3 step = sign11(dir);
  if (hit.state == 0u) {
    const V3 N = normalize3((-step) * maskf(hit.mask));
    const V3 rdir = normalize3(dir - (2.0f * N) * dot3(dir, N));
    const V3 rsrc = hit.p - (4e-2f * step) * maskf(hit.mask);
    const V3 rcol = reflect_ray1(T, s, rsrc, rdir);
    const V3 mcol = sun_lit(T, s, hit, step, N);
    return mix3(mcol, rcol, WX_REFLECTIVITY);
  }
  if (hit.state == 1u) return wall_flat(normalize3((-step) * maskf(hit.mask)));
  return dir;
}

__device__ __forceinline__ float fmod_trunc(float x, float y) { return x - y * truncf(x / y); }  // WGSL `%`
__device__ __forceinline__ bool any_mod0(V3 fp, float m) {
  return fmod_trunc(fp.x, m) == 0.f || fmod_trunc(fp.y, m) == 0.f || fmod_trunc(fp.z, m) == 0.f;
}

// ray_trace (:152-265) given the primary HitOut.  MODE is the warp-uniform render mode.
template <int MODE>
__device__ __forceinline__ V3 shade(const DevTree& T, const WxState& s, const HitOut& hit, V3 dir) {
  const V3 step = sign11(dir);
  if (hit.state == 0u) {
    V3 grid = splat(0.0f);
    if (MODE <= 2 || MODE > 4) {  // modes 3 and 4 never read `grid`
      const V3 fp = V3{floorf(hit.p.x), floorf(hit.p.y), floorf(hit.p.z)};
      if (s.show_345[2] == 1u && any_mod0(fp, 4096.f)) grid = V3{-0.3f, -0.3f, 1.0f};
      else if (s.show_345[1] == 1u && any_mod0(fp, 128.f)) grid = V3{0.6f, -0.2f, -0.2f};
      else if (s.show_345[0] == 1u && any_mod0(fp, 8.f)) grid = V3{-0.1f, 0.5f, 0.3f};
    }
    if (MODE == 1) return (grid + splat(0.1f)) + maskf(hit.mask) * V3{0.4f, 0.4f, 0.4f};
    if (MODE == 2) {
      const float t = (float)hit.i / (float)200u;
      return grid + mix3(V3{0.72f, 1.0f, 0.99f}, V3{1.0f, 0.0f, 0.0f}, t);
    }
    if (MODE == 3) {
      const V3 N = normalize3((-step) * maskf(hit.mask));
      const float LN = fmaxf(0.0f, s.sun_color[3] * dot3(-sun_dir(s), N));
      const V3 I_d = ((WX_K_D * sun_rgb(s)) * WX_BASE_COLOR) * LN;
      const V3 I_a = (WX_K_A * WX_AMBIENT_COLOR) * WX_BASE_COLOR;
      if (LN != 0.0f && hdda_ray_secondary(T, hit.p - (4e-2f * step) * maskf(hit.mask), -sun_dir(s)).state == 0u) return I_a;
      return I_a + I_d;
    }
    if (MODE == 4) {
      const V3 N = normalize3((-step) * maskf(hit.mask));
      const V3 mcol = sun_lit(T, s, hit, step, N);
      const V3 rdir = normalize3(dir - (2.0f * N) * dot3(dir, N));
      const V3 rsrc = hit.p - (4e-2f * step) * maskf(hit.mask);
      const V3 rcol = reflect_ray2(T, s, rsrc, rdir);
      return mix3(mcol, rcol, WX_REFLECTIVITY);
    }
    return grid + dot3(maskf(hit.mask) * V3{0.2f, 0.2f, 0.3f}, splat(1.0f));  // Gray and any other mode
  }
  if (hit.state == 1u) {
    if (MODE == 2) {
      const float t = (float)hit.i / (float)200u;
      return mix3(V3{0.72f, 1.0f, 0.99f}, V3{1.0f, 0.0f, 0.0f}, t) + dot3(maskf(hit.mask) * V3{0.04f, 0.08f, 0.12f}, splat(1.0f));
    }
    if (MODE == 4) {
      const V3 N = normalize3((-step) * maskf(hit.mask));
      const V3 Np = max3(splat(0.f), N);
      const V3 Nn = -min3(splat(0.f), N);
      const float t = hit.p.y / 4096.f;
      V3 r = mix3(V3{WX_WALL_I, 0.f, 0.f}, V3{WX_WALL_I * 0.1f, 0.f, 0.f}, t) * Np.x;
      r = r + V3{0.f, WX_WALL_I, 0.f} * Np.y;
      r = r + mix3(V3{0.f, 0.f, WX_WALL_I}, V3{0.f, 0.f, WX_WALL_I * 0.1f}, t) * Np.z;
      r = r + mix3(V3{WX_WALL_I, WX_WALL_I, 0.f}, V3{WX_WALL_I * 0.1f, WX_WALL_I * 0.1f, 0.f}, t) * Nn.x;
      r = r + V3{0.f, WX_WALL_I, WX_WALL_I} * Nn.y;
      r = r + mix3(V3{WX_WALL_I, 0.f, WX_WALL_I}, V3{WX_WALL_I * 0.1f, 0.f, WX_WALL_I * 0.1f}, t) * Nn.z;
      return r;
    }
    return splat(0.0f) + dot3(maskf(hit.mask) * V3{0.01f, 0.02f, 0.03f}, splat(1.0f));
  }
  return dir;  // max steps exceeded
}

// rgba8unorm store conversion: clamp, x255, round half to even; NaN -> 0
__device__ __forceinline__ uint32_t unorm8(float c) {
  if (c != c) return 0u;
  c = fminf(fmaxf(c, 0.f), 1.f);
  return (uint32_t)__float2int_rn(c * 255.0f);
}

}  // namespace wx
