// wx_device.cuh -- device-side data layout and the per-ray march for sm_100a.
//
// What this replaces: the body of src/shaders/raycast.comp.wgsl of the reference (cp_main :60-68,
// hdda_ray :84-126, ray_trace :152-265, reflect_ray2/1 :267-342, get_vdb_leaf_* :360-494).  It is
// not a translation of that file:
//   * the reference reads 2 mask words + 1 atlas texel per tree level (three 3-D textures + five
//     mask buffers); here every internal level is ONE table of 32-bit entries and a leaf is one
//     512-byte brick, so a lookup is one load per level;
//   * the bottom-up parent-origin cache (:360-396) is three XORs against the voxel visited last;
//   * voxel coordinates are never converted: floor(p) is one round-down add of 1.5*2^23 whose
//     float bit pattern IS the (biased) integer coordinate;
//   * modulo_vec3f (:78-80; three IEEE divisions + three floors per step) is replaced by an exact
//     floor((x+1/2)/size) built from one MUFU.RCP and fused multiply-adds (proof below), packed two
//     lanes per instruction with the Blackwell f32x2 forms (FADD2 / FFMA2).
//
// Arithmetic contract: IEEE binary32, one rounding per source operation of the WGSL, in its
// operation order (this translation unit is compiled with -fmad=false; fused operations appear
// only where they are provably equal to the separately rounded sequence).  Results are
// bit-identical to the CPU oracle, which is how parity is proven (tests/test_parity_gpu.py).
//
// WX_HOST_EMU: defined ONLY by tests/emu (test infrastructure), which compiles this header with g++ to check the code below
// against the oracle on the CPU.  No product build defines it (woxel_b200/csrc/Makefile, tests/test_abi_symbols.py): the
// library has no CPU path, and wx_init fails without a CUDA device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/woxel_b200.h"

namespace wx {

// ---------------------------------------------------------------------------------------------
// Device tree.  All arrays are 256-B aligned (cudaMalloc); node records are 128 KB / 16 KB / 512 B
// (2 KB for wide leaves), so every node starts on a 128-B line.
//
//   e5[n5][32768], e4[n4][4096] : u32 entry per slot, offset order of the reference
//        bit 31 set  -> child; low 31 bits = child node index (reference DFS index)
//        bit 31 clear-> tile; the entry is the f32 bit pattern of `size` = SDF distance * cell edge
//                       (128 for an N5 slot, 8 for an N4 slot), i.e. exactly the value hdda_ray
//                       computes at raycast.comp.wgsl:104; 0.0f = active tile (hit).
//        (value-mask is tested before child-mask in the reference, :431-437; the packer applies
//         that priority once, at upload)
//   l3[n3] : one brick per leaf, one byte per voxel in offset order (x<<6 | y<<3 | z): 0 = active
//        voxel (hit), else SDF distance in voxels.  When a distance exceeds 255 the whole level is
//        stored as u32 per voxel instead (leaf_shift 11 instead of 9).
//   origins: (x, y, z) + kBias per N5, in the reference's sorted order.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kBias = 0x4B400000u;      // bit pattern of 1.5*2^23: float(kMagic + x) has bits kBias + x
constexpr float kMagic = 12582912.0f;        // 1.5 * 2^23
constexpr uint32_t kChildFlag = 0x80000000u;
constexpr uint32_t kNoCache = 0x80000000u;   // cursor.dbits: nothing cached, every lookup starts at the root
constexpr float kFastMaxSize = 1048576.0f;   // 2^20: largest step size the fast march accepts
constexpr int kRootNone = -1, kRootScan = -2, kRootBeyond = 0x4000, kRootIndexMask = 0x3fff;

struct DevTree {
  const uint32_t* __restrict__ e5;
  const uint32_t* __restrict__ e4;
  const uint8_t* __restrict__ l3;
  // child entry e (flag included) -> node address: adj + e * node_bytes  (adj = base - 2^31 * node_bytes)
  const char* __restrict__ e4_adj;
  const char* __restrict__ l3_adj;
  const int4* __restrict__ origins_g;  // all n5 biased origins (scanned only outside the root grid)
  uint32_t n5, n4, n3;
  uint32_t leaf_shift;                 // log2(bytes per leaf): 9 (u8 voxels) or 11 (u32 voxels)
  uint32_t fast_ok;                    // every tile/leaf size < kFastMaxSize
  // The 4x4x4 root cells covering [-8192, 8192)^3, cell = ((x>>12)+2)*16 + ((y>>12)+2)*4 + (z>>12)+2: everything a ray
  // can reach from inside the +-4096 world with one level-0 step.  Entry: kRootNone (no N5 there), kRootScan (index does
  // not fit: scan origins), else the N5 index, | kRootBeyond when the cell reaches outside the world (its origin has a
  // component outside [-4096, 0]), where the march must test the bounds.  Outside the grid: scan origins.
  int16_t root_grid[64];
  // World grid (the march's usual top level, see "World grid" below); grid == nullptr when the tree does not qualify.
  const uint32_t* __restrict__ grid;  // kGridCells entries
  const uint32_t* __restrict__ f4;    // re-encoded N4 tables: the child entries of e4 as index bases (grid_word3)
  // Bounding box (voxel coordinates, 128-voxel cell granularity) of everything a ray can hit: the in-world grid cells that are
  // children or active tiles.  Used only by the tolerance-mode march (WX_OPT_MARCH = 1); bb_lo[0] > bb_hi[0] = empty / unknown.
  float bb_lo[3], bb_hi[3];
#ifdef WX_ROOT_PTRS
  // A/B variant (measured 0.7 % SLOWER than the int16 cells + address arithmetic, profiles/r1_variants_h.txt; not the default):
  // the same cells as ready-made N5 table addresses of THIS replica: kRootPtrNone = no N5, kRootPtrScan = scan the
  // origins (both have bit 1 set), else the address of e5[n5] with bit 0 set when the cell reaches outside the world (tables
  // are 128-B aligned, so bits 0..6 of an address are free).
  uint64_t root_ptr[64];
#endif
};
constexpr uint64_t kRootPtrNone = 2ull, kRootPtrScan = 6ull;

// ---------------------------------------------------------------------------------------------
// World grid: the top level the march normally reads (march_grid).  The reference finds the N5 of a position by scanning
// origins[] and then indexes that N5's table (raycast.comp.wgsl:398-444).  Every position a ray can visit while it is inside
// the +-4096 world lies in [-8192 - 1, 8192 + 1)^3 as long as no step is longer than 4096, so ONE dense table of 128-voxel
// cells over that cube replaces the root lookup and all N5 tables: 128^3 cells + a pad of one cell row/plane on either end,
// 8.5 MB, of which only the 64^3 in-world cells (1 MB, the size of 8 N5 tables) are ever hot.
//   entry, in-world cell of an existing N5 : the N5's slot -- tile: f32 bits of `size` (0.0f = active tile) as in e5;
//                                             child: word4 (below)
//          in-world cell without an N5     : f32 bits of 4096.0f (dist 1 at level 0, :411)
//          cell outside the world, no N5   : kEntryVoid -- the lookup there is "dist 1 at level 0" and the bounds test of
//                                             :100-103 follows: the ray ends out of bounds unless it sits exactly on +4096
//          anything else                   : kEntrySlow -- a cell outside the world inside some N5 (a hit is possible before
//                                             the bounds test), a tile larger than 4096, the pads.  The march then hands the
//                                             ray to the generic loop (march_fast from the current state), which knows all that.
// Tile sizes are positive, so bit 31 marks a child (as in e5 / e4); kEntrySlow is a float below 1 that no tile can be.
// Child words are INDEX BASES, not node numbers: with biased voxel coordinates (x, y, z) the N4 slot of a position is
//   f4[(x >> 3) * 256 + (y >> 3) * 16 + (z >> 3) + (word4 << 4)],   word4 = kChildFlag | (n4 * 4096 - C4(N4 origin)) >> 4   (mod 2^32)
// and the voxel of a leaf l3[x * 64 + y * 8 + z + (word3 << 3)], word3 = kChildFlag | (n3 * 512 - C3(leaf origin)) >> 3: no masking of
// the coordinates, no 64-bit node pointer in the cursor.  C4 is a multiple of 16 and C3 of 8, and the shift that undoes the
// division also drops the flag bit (it is part of the index add: one LEA).
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kGS = 128u, kGS2 = kGS * kGS;
constexpr uint32_t kGridPad = kGS2 + kGS + 1u;
constexpr size_t kGridCells = (size_t)kGS * kGS2 + 2u * (size_t)kGridPad;
// cell index of biased coordinates: (x >> 7) * kGS2 + (y >> 7) * kGS + (z >> 7) + kGridK   (mod 2^32)
constexpr uint32_t kGridK = kGridPad - ((kBias >> 7) - 64u) * (kGS2 + kGS + 1u);
constexpr uint32_t kEntrySlow = 0x02020202u;  // what cudaMemset(.., 0x02, ..) writes: 9.6e-38f
constexpr uint32_t kEntryVoid = 0x01010101u;  // 2.4e-38f
constexpr float kGridMaxSize = 4096.f;
constexpr uint32_t kGridMaxN3 = (1u << 23) - 1u, kGridMaxN4 = (1u << 20) - 1u;  // index bases must fit 32 bits

__host__ __device__ inline uint32_t grid_word4(uint32_t n4, uint32_t ox, uint32_t oy, uint32_t oz) {  // biased N4 origin
  return kChildFlag | ((n4 * 4096u - ((ox >> 3) * 256u + (oy >> 3) * 16u + (oz >> 3))) >> 4);
}
__host__ __device__ inline uint32_t grid_word3(uint32_t n3, uint32_t ox, uint32_t oy, uint32_t oz) {  // biased leaf origin
  return kChildFlag | ((n3 * 512u - (ox * 64u + oy * 8u + oz)) >> 3);
}
// Entry of cell (cx, cy, cz) of the 128^3 cube: cell coordinate = true coordinate / 128 + 64; in-world cells are [32, 96)^3.
__host__ __device__ inline uint32_t grid_cell_entry(uint32_t cx, uint32_t cy, uint32_t cz, const int16_t* root_grid, const uint32_t* e5) {
  const int v = (int)root_grid[(cx >> 5) * 16u + (cy >> 5) * 4u + (cz >> 5)];  // (true >> 12) + 2 == cell >> 5
  if (((cx - 32u) | (cy - 32u) | (cz - 32u)) >= 64u) return v == kRootNone ? kEntryVoid : kEntrySlow;  // outside the world
  if (v == kRootNone) return 0x45800000u;                                       // 4096.0f
  if (v < 0 || (v & kRootBeyond)) return kEntrySlow;
  const uint32_t e = e5[(size_t)(v & kRootIndexMask) * 32768u + (((cx & 31u) << 10) | ((cy & 31u) << 5) | (cz & 31u))];
  if (e & kChildFlag)
    return grid_word4(e & ~kChildFlag, (cx - 64u) * 128u + kBias, (cy - 64u) * 128u + kBias, (cz - 64u) * 128u + kBias);
  return e > 0x45800000u ? kEntrySlow : e;  // a tile larger than 4096 (positive floats order like their bit patterns)
}

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
  uint32_t row_base, row_end;  // single-shard launches may cover rows [row_base, row_end) only (pipelined read-back)
  // persistent kernel: work queue of 32x16-pixel chunks over the rows this launch owns
  uint32_t* work_counter;      // zeroed before the launch
  uint32_t chunks_x, chunks_y, n_chunks;  // per camera: chunks_x * chunks_y; n_chunks = that * cameras of the launch
  uchar4* rgba;
  AovPtrs aov;
  uint32_t has_aov;
  // Long tiles first (wx_raycast.cu): the tiles (CTA footprints) of the PREVIOUS launch with this geometry that held a ray of at
  // least sched_threshold iterations -- prev_list[0] = how many, prev_list[1..] their ids, prev_flag[tile] = 1 for each -- are
  // rendered by a small kernel that starts first; the main grid skips them.  Every launch records the same for the next one
  // (next_list / next_flag, zeroed beforehand).  All null / 0: plain launch.
  const uint32_t* prev_list;
  const uint32_t* prev_flag;
  uint32_t* next_list;
  uint32_t* next_flag;
  uint32_t sched_cap, sched_threshold;
  // A launch that renders only a row chunk of the frame the lists describe (wx_render's pipelined read-back): tile ids are those of
  // the whole frame -- frame_tile_rows tile rows, this launch's first one being tile_row_offset.  0 / 0: the launch is the frame.
  uint32_t frame_tile_rows, tile_row_offset;
};
// MARCH template parameter of the kernels: how hdda_ray is evaluated (WX_OPT_MARCH)
constexpr int kMarchExact = 0, kMarchTolerance = 1;

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
// Packed binary32 pairs (Blackwell f32x2: one issue slot, two IEEE results).  ptxas turns a pair
// built from one scalar into a broadcast operand, so bc() costs nothing.
// NOTE: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false (and folds
// fma2(a, b, -0.0) back into it); where a product must be rounded on its own and then added, the
// add is done with scalar FADDs, which ptxas leaves alone.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
#ifndef WX_HOST_EMU
__device__ __forceinline__ f32x2 pk(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
#else  // tests/emu: this header compiled by the host compiler; the PTX-only primitives come from tests/emu/cuda_shim.h
__device__ __forceinline__ f32x2 pk(float lo, float hi) { return (f32x2)__float_as_uint(lo) | ((f32x2)__float_as_uint(hi) << 32); }
#endif
__device__ __forceinline__ f32x2 bc(float v) { return pk(v, v); }
__device__ __forceinline__ float lo(f32x2 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi(f32x2 v) { return __uint_as_float((uint32_t)(v >> 32)); }
#ifndef WX_NO_F32X2
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2_rd(f32x2 a, f32x2 b) {  // round toward -inf
  f32x2 d;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {  // never feed this to add2/sub2 (see NOTE)
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
#else  // A/B variant: the same operations as scalar instructions (more issue slots, more independent work per stage)
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { return pk(__fadd_rn(lo(a), lo(b)), __fadd_rn(hi(a), hi(b))); }
__device__ __forceinline__ f32x2 add2_rd(f32x2 a, f32x2 b) { return pk(__fadd_rd(lo(a), lo(b)), __fadd_rd(hi(a), hi(b))); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { return pk(__fsub_rn(lo(a), lo(b)), __fsub_rn(hi(a), hi(b))); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { return pk(__fmaf_rn(lo(a), lo(b), lo(c)), __fmaf_rn(hi(a), hi(b), hi(c))); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { return pk(__fmul_rn(lo(a), lo(b)), __fmul_rn(hi(a), hi(b))); }
#endif
#ifndef WX_HOST_EMU
__device__ __forceinline__ float rcp_approx(float v) {  // MUFU.RCP, <= 1 ulp
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
#else
__device__ __forceinline__ float rcp_approx(float v) { return wx_emu_rcp(v); }  // 1/v, bumped by the ulps the test asks for
#endif

// ---------------------------------------------------------------------------------------------
// Tree cursor: the path to the voxel visited last.
//   dbits = 0            : q5, q4, q3 valid (last lookup ended in a leaf,       num_parents 3)
//           8            : q5, q4 valid     (last lookup ended on an N4 slot,   num_parents 2)
//           128          : q5 valid         (last lookup ended on an N5 slot,   num_parents 1)
//           kNoCache|L<<28: nothing reusable (no N5 here, or an N5 that reaches beyond the +-4096 world,
//                          where the march must test the bounds every step); L = num_parents
// OR-ing dbits into the coordinate difference makes "same node AND cached that deep" one compare.
// ---------------------------------------------------------------------------------------------
struct Cursor {
  uint32_t lx, ly, lz;  // biased voxel coordinates of the last lookup
  uint32_t dbits;
  const uint32_t* q5;
  const uint32_t* q4;
  const uint8_t* q3;
};

__device__ __forceinline__ uint32_t cursor_level(uint32_t dbits) {
  if (dbits & kNoCache) return (dbits >> 28) & 7u;
  return dbits == 0u ? 3u : (dbits == 8u ? 2u : 1u);
}

__device__ __forceinline__ float u32_to_float(uint32_t v) {  // I2FP (alu pipe); a u8/u16 source would go to the XU pipe
#ifndef WX_HOST_EMU
  float f;
  asm("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(v));
  return f;
#else
  return (float)v;
#endif
}

// Walk down from the level the cursor is still valid for: dv >= 128 starts at the N5 table c.q5,
// dv >= 8 at the N4 table c.q4, else at the leaf brick c.q3.  Each level's code exists once and the
// lanes of a warp fall through it together, whatever level they started on.  Returns `size`
// (0 = hit) and leaves the cursor on the node the walk ended in.
// WIDE: the leaf level may be stored as u32 per voxel (only the exact march handles that).
#ifdef WX_DESCEND_GOTO  // A/B variant: one goto chain; measured 2.5 % slower than the structured walk below
template <bool WIDE>
__device__ __forceinline__ float descend(const DevTree& T, Cursor& c, uint32_t dv, uint32_t x, uint32_t y, uint32_t z) {
  uint32_t e;
  if (dv < 8u) goto leaf;
  if (dv < 128u) goto node4;
  e = __ldg(c.q5 + (((x << 3) & 0x7C00u) | ((y >> 2) & 0x3E0u) | ((z >> 7) & 31u)));
  if ((int32_t)e >= 0) {
    c.dbits = 128u;
    return __uint_as_float(e);
  }
  c.q4 = reinterpret_cast<const uint32_t*>(T.e4_adj + (uint64_t)e * 16384ull);
node4:
  e = __ldg(c.q4 + (((x << 5) & 0xF00u) | ((y << 1) & 0xF0u) | ((z >> 3) & 15u)));
  if ((int32_t)e >= 0) {
    c.dbits = 8u;
    return __uint_as_float(e);
  }
  c.q3 = reinterpret_cast<const uint8_t*>(T.l3_adj + (WIDE ? ((uint64_t)e << T.leaf_shift) : (uint64_t)e * 512ull));
leaf:
  c.dbits = 0u;
  const uint32_t o3 = ((x & 7u) << 6) | ((y & 7u) << 3) | (z & 7u);
  if (!WIDE || T.leaf_shift == 9u) return u32_to_float(__ldg(c.q3 + o3));
  return u32_to_float(__ldg(reinterpret_cast<const uint32_t*>(c.q3) + o3));
}
#else
// One structured `if` per level: the lanes of a warp reconverge before the leaf level, whatever mix of start
// levels they have.  ptxas threads the jump from "the N5 entry is a child" straight into the N4 block, so lanes that
// came down from the N5 and lanes that started at the N4 run that block one group after the other (ncu: 6.1 M
// executions at 16 lanes where 4.8 M at 20 would do); WX_JOIN makes the level opaque at the join so that the groups
// meet there first (-DWX_DESCEND_JOIN, A/B).
#ifdef WX_DESCEND_JOIN
#ifdef WX_HOST_EMU
#define WX_JOIN(lvl) ((void)0)
#else
#define WX_JOIN(lvl) asm volatile("" : "+r"(lvl))
#endif
#else
#define WX_JOIN(lvl) ((void)0)
#endif
template <bool WIDE>
__device__ __forceinline__ float descend(const DevTree& T, Cursor& c, uint32_t dv, uint32_t x, uint32_t y, uint32_t z) {
  float size = 0.f;
  uint32_t lvl = dv >= 128u ? 5u : (dv >= 8u ? 4u : 3u);  // level whose table is read next; 0 = walk finished
  if (lvl == 5u) {
    const uint32_t e = __ldg(c.q5 + (((x << 3) & 0x7C00u) | ((y >> 2) & 0x3E0u) | ((z >> 7) & 31u)));
    if ((int32_t)e >= 0) {
      c.dbits = 128u, size = __uint_as_float(e), lvl = 0u;
    } else {
      c.q4 = reinterpret_cast<const uint32_t*>(T.e4_adj + (uint64_t)e * 16384ull), lvl = 4u;
    }
  }
  WX_JOIN(lvl);
  if (lvl == 4u) {
    const uint32_t e = __ldg(c.q4 + (((x << 5) & 0xF00u) | ((y << 1) & 0xF0u) | ((z >> 3) & 15u)));
    if ((int32_t)e >= 0) {
      c.dbits = 8u, size = __uint_as_float(e), lvl = 0u;
    } else {
      c.q3 = reinterpret_cast<const uint8_t*>(T.l3_adj + (WIDE ? ((uint64_t)e << T.leaf_shift) : (uint64_t)e * 512ull)), lvl = 3u;
    }
  }
  if (lvl == 3u) {
    c.dbits = 0u;
    const uint32_t o3 = ((x & 7u) << 6) | ((y & 7u) << 3) | (z & 7u);
    if (!WIDE || T.leaf_shift == 9u) size = u32_to_float(__ldg(c.q3 + o3));
    else size = u32_to_float(__ldg(reinterpret_cast<const uint32_t*>(c.q3) + o3));
  }
  return size;
}
#endif

// Root: first origin equal to (pos >> 12) << 12 (raycast.comp.wgsl:398-413).  Returns the N5 index or -1.
static __device__ __noinline__ int scan_roots(const DevTree& T, uint32_t x, uint32_t y, uint32_t z) {
  const int gx = (int)(x & ~4095u), gy = (int)(y & ~4095u), gz = (int)(z & ~4095u);
  for (uint32_t i = 0; i < T.n5; ++i) {
    const int4 o = __ldg(T.origins_g + i);
    if (o.x == gx && o.y == gy && o.z == gz) return (int)i;
  }
  return -1;
}
// The N5 at (pos >> 12) << 12 contains positions outside the +-4096 world (origin component not in [-4096, 0]).
__device__ __forceinline__ bool reaches_beyond(uint32_t x, uint32_t y, uint32_t z) {
  const uint32_t span = 4096u;  // biased origin - (kBias - 4096) must be 0 or 4096
  return ((x & ~4095u) - (kBias - 4096u)) > span || ((y & ~4095u) - (kBias - 4096u)) > span || ((z & ~4095u) - (kBias - 4096u)) > span;
}
// Root entry for the N5 cell of (x, y, z): kRootNone, or the N5 index with kRootBeyond or-ed in.
__device__ __forceinline__ int find_root(const DevTree& T, uint32_t x, uint32_t y, uint32_t z) {
  const uint32_t c0 = (kBias >> 12) - 2u;
  const uint32_t cx = (x >> 12) - c0, cy = (y >> 12) - c0, cz = (z >> 12) - c0;
  if ((cx | cy | cz) < 4u) {
    const int v = (int)T.root_grid[cx * 16u + cy * 4u + cz];
    if (v != kRootScan) return v;
  }
  const int n5 = scan_roots(T, x, y, z);
  return n5 < 0 ? kRootNone : (n5 | (reaches_beyond(x, y, z) ? kRootBeyond : 0));
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
#ifndef WX_EMU_STEP
#define WX_EMU_STEP(dv0, dbits) ((void)(dv0))  // tests/emu records the level walk of every step here; nothing on the device
#endif
#ifndef WX_STAGE_MIN
#define WX_STAGE_MIN 8  // lanes that must share a leaf before its brick is staged (WX_STAGE_LEAF experiment)
#endif
#ifndef WX_UNROLL
#define WX_UNROLL 2  // two steps per loop trip: the cursor's last-voxel registers alternate instead of being copied
#endif
#if WX_UNROLL > 1
#define WX_STR2(x) #x
#define WX_STR(x) WX_STR2(x)
#define WX_UNROLL_PRAGMA _Pragma(WX_STR(unroll WX_UNROLL))
#else
#define WX_UNROLL_PRAGMA
#endif

__device__ __forceinline__ bool out_of_bounds(float x, float y, float z) {
  return 4096.f < fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z));
}

// The N5 changed (or nothing is cached): find the root entry.  On success c.q5 is set and dv becomes 128
// (walk down from the N5 table); without an N5 dv stays >= 4096 and the cursor caches nothing.  Returns
// `beyond`: the bounds test of :100-103 can succeed at this position (it cannot inside an N5 whose
// origin lies in [-4096, 0]^3).
__device__ __forceinline__ bool enter_root(const DevTree& T, Cursor& c, uint32_t& dv, uint32_t x, uint32_t y, uint32_t z) {
#ifdef WX_ROOT_PTRS
  const uint32_t c0 = (kBias >> 12) - 2u;
  const uint32_t cx = (x >> 12) - c0, cy = (y >> 12) - c0, cz = (z >> 12) - c0;
  uint64_t v = kRootPtrScan;
  if ((cx | cy | cz) < 4u) v = T.root_ptr[cx * 16u + cy * 4u + cz];
  if ((uint32_t)v & 2u) {  // no N5 here, or the cell's N5 must be found by the scan
    int n5 = -1;
    if ((uint32_t)v & 4u) n5 = scan_roots(T, x, y, z);
    if (n5 < 0) {
      c.dbits = kNoCache;
      return true;
    }
    v = (uint64_t)(T.e5 + (size_t)n5 * 32768u) | (reaches_beyond(x, y, z) ? 1ull : 0ull);
  }
  c.q5 = reinterpret_cast<const uint32_t*>(v & ~1ull);
  dv = 128u;
  return ((uint32_t)v & 1u) != 0u;
#else
  const int r = find_root(T, x, y, z);
  if (r < 0) {
    c.dbits = kNoCache;
    return true;
  }
  c.q5 = T.e5 + (size_t)(r & ~kRootBeyond) * 32768u;
  dv = 128u;
  return (r & kRootBeyond) != 0;
#endif
}

// One lookup L(pos) through the cursor, for the exact march (bounds tested every step).
template <bool WIDE>
__device__ __forceinline__ float lookup(const DevTree& T, Cursor& c, uint32_t x, uint32_t y, uint32_t z) {
  uint32_t dv = ((x ^ c.lx) | c.dbits) | (y ^ c.ly) | (z ^ c.lz);
  c.lx = x, c.ly = y, c.lz = z;
  if (dv >= 4096u) (void)enter_root(T, c, dv, x, y, z);
  if (dv >= 4096u) return 4096.f;  // no N5 here: dist 1 at level 0 (:411)
  return descend<WIDE>(T, c, dv, x, y, z);
}

__device__ __forceinline__ void finish(const DevTree& T, const Cursor& c, HitOut& out) {
  out.level = cursor_level(c.dbits);
  out.n3 = out.level == 3u ? (uint32_t)((size_t)(c.q3 - T.l3) >> T.leaf_shift) : 0u;
}

// Exact march: the WGSL's operations one by one (IEEE division, floor).  Used for rays the fast
// march excludes (a direction component so small that 1/dir overflows) and for trees with a step
// size >= 2^20.
static __device__ __noinline__ HitOut march_exact(const DevTree& T, V3 src, V3 dir) {
  V3 p = src;
  const V3 step = sign11(dir);
  const V3 step01 = max3(splat(0.f), step);
  const V3 idir = V3{1.f / dir.x, 1.f / dir.y, 1.f / dir.z};
  const V3 nudge = 4e-4f * step;
  uint32_t mask = 0;
  Cursor c{0u, 0u, 0u, kNoCache, nullptr, nullptr, nullptr};
  HitOut out;
  out.state = 2u;
  uint32_t i = 0;
  for (; i < kMaxRaySteps; ++i) {
    const uint32_t x = (uint32_t)__float2int_rd(p.x) + kBias, y = (uint32_t)__float2int_rd(p.y) + kBias,
                   z = (uint32_t)__float2int_rd(p.z) + kBias;
    const float size = lookup<true>(T, c, x, y, z);
    if (size == 0.f) {
      out.state = 0u;
      break;
    }
    if (out_of_bounds(p.x, p.y, p.z)) {
      out.state = 1u;
      break;
    }
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
  out.p = p;
  out.mask = mask;
  out.i = i;
  finish(T, c, out);
  return out;
}

// Fast march.  Preconditions (checked by hdda_ray): every |1/dir| < 1e30 (so no tMax is NaN and
// none overflows) and every size < 2^20.  Differences from the text of :84-126, each value-exact:
//
//  (1) floor(p): t = p (+, round down) 1.5*2^23 has the integer floor(p) in its low mantissa bits for
//      |p| < 2^22; its bit pattern is kBias + floor(p), which is what the tree is indexed with.
//  (2) size * floor(p / size): for finite p and integer size >= 1, RN(p/size) can only reach an
//      integer k from the inside of [k, k+1) -- p is at least one ulp away from the lattice plane
//      k*size, and one ulp of p is more than half an ulp of k times size -- so
//      floor(RN(p/size)) == floor(p/size) == floor((x + 1/2)/size) with x = floor(p).  (x + 1/2)/size
//      is at least 1/(2 size) away from every integer, and fma(x, r, r/2) with r = MUFU.RCP(size)
//      is within 4097.5 * 2^-21.4 / size < 1/(2 size) of it (|x| <= 4096 when the bounds test has
//      passed), so its floor is exact.  size * that floor is an integer below 2^24: the FMA that
//      subtracts p from it rounds once, exactly like p - size * floor(p / size).  (The one input this changes is a denormal
//      negative p, where RN(p/size) underflows to -0: not reachable from a camera.)
//  (3) size * step01 is exactly size or 0, so fma(size, step01, -m) rounds once, like the WGSL.
//  (4) mask: without NaNs, (tx <= ty && tx <= tz) == (tx == min(tx, ty, tz)).
//  (5) p += 4e-4 * step * mask adds exactly +-4e-4 or +-0: a predicated add.
//  (6) the bounds test can only succeed where lookup() says so (see Cursor).
__device__ __forceinline__ float keep(float v) {  // the value stays in its register (no rematerialisation in the loop)
#ifndef WX_HOST_EMU
  asm volatile("" : "+f"(v));
#endif
  return v;
}

struct FastRay {
  f32x2 pxy, dxy, ixy, s01xy;
  float pz, dz, iz, s01z;
  float ndx, ndy, ndz;
  float ltx, lty, ltz, lt;  // tMax and its minimum of the last step (the mask is derived on exit)
  float size;               // result of the last lookup: 0 = hit
  Cursor c;
  uint32_t i;

  __device__ __forceinline__ void init(V3 src, V3 dir, V3 idir) {
    pxy = pk(src.x, src.y), pz = src.z;
    dxy = pk(dir.x, dir.y), dz = dir.z;
    ixy = pk(idir.x, idir.y), iz = idir.z;
    s01xy = pk(keep(dir.x < 0.f ? 0.f : 1.f), keep(dir.y < 0.f ? 0.f : 1.f));
    s01z = keep(dir.z < 0.f ? 0.f : 1.f);
    ndx = keep(dir.x < 0.f ? -4e-4f : 4e-4f), ndy = keep(dir.y < 0.f ? -4e-4f : 4e-4f), ndz = keep(dir.z < 0.f ? -4e-4f : 4e-4f);
    ltx = 1.f, lty = 1.f, ltz = 1.f, lt = 0.f;
    size = 1.f;
    c = Cursor{0u, 0u, 0u, kNoCache, nullptr, nullptr, nullptr};
    i = 0;
  }

  // One iteration of hdda_ray's loop body (:90-122) without the counter.  Returns true when the ray ended:
  // it hit (size == 0, state 0) or left the world (size != 0, state 1).
  __device__ __forceinline__ bool step(const DevTree& T) {
    const f32x2 txy = add2_rd(pxy, bc(kMagic));
    const float tz = __fadd_rd(pz, kMagic);
    const uint32_t x = (uint32_t)txy, y = (uint32_t)(txy >> 32), z = __float_as_uint(tz);
    uint32_t dv = ((x ^ c.lx) | c.dbits) | (y ^ c.ly) | (z ^ c.lz);
    c.lx = x, c.ly = y, c.lz = z;
    // lookup L(pos): from the root only when the N5 changed, else from the deepest cached node
    bool beyond = false;
    const uint32_t dv0 = dv;
    if (dv >= 4096u) beyond = enter_root(T, c, dv, x, y, z);
    if (dv < 4096u) size = descend<false>(T, c, dv, x, y, z);
    else size = 4096.f;  // no N5 here: dist 1 at level 0 (:411)
    WX_EMU_STEP(dv0, c.dbits);
    if (size == 0.f) return true;
    if (beyond) {  // the only places where the bounds test of :100-103 can succeed (see Cursor)
      if (out_of_bounds(lo(pxy), hi(pxy), pz)) return true;
      c.dbits = kNoCache | (cursor_level(c.dbits) << 28);
    }
    const float r = rcp_approx(size);
    const float hr = 0.5f * r;
    const f32x2 xfxy = add2(txy, bc(-kMagic));                         // float(floor(p)), exact
    const float xfz = tz - kMagic;
    const f32x2 qxy = add2_rd(fma2(xfxy, bc(r), bc(hr)), bc(kMagic));  // kMagic + floor(p / size)
    const float qz = __fadd_rd(fmaf(xfz, r, hr), kMagic);
#ifndef WX_NM_SEPARATE
    // k = floor(p / size) = q - kMagic (exact), then -modulo_vec3f(p, size) = fma(k, size, -p): k * size is an exact integer
    // below 2^24, so this rounds once, like g - p.  (One instruction fewer than forming g first: measured 0.7 % faster,
    // profiles/r1_variants_h.txt.)
    const f32x2 kxy = add2(qxy, bc(-kMagic));
    const float kz = qz - kMagic;
    const f32x2 nmxy = fma2(kxy, bc(size), pk(-lo(pxy), -hi(pxy)));
    const float nmz = fmaf(kz, size, -pz);
#else  // A/B: g = size * floor(p / size) first, then g - p
    const float nms = -kMagic * size;
    const f32x2 gxy = fma2(qxy, bc(size), bc(nms));                    // size * floor(p / size), exact
    const float gz = fmaf(qz, size, nms);
    const f32x2 nmxy = sub2(gxy, pxy);                                 // -modulo_vec3f(p, size)
    const float nmz = gz - pz;
#endif
    const f32x2 tmxy = mul2(ixy, fma2(bc(size), s01xy, nmxy));         // tMax
    const float tmz = iz * fmaf(size, s01z, nmz);
    ltx = lo(tmxy), lty = hi(tmxy), ltz = tmz;
    lt = fminf(fminf(ltx, lty), ltz);
    // p += t * dir: the product is rounded on its own (scalar adds: ptxas would fuse a packed pair)
    const f32x2 axy = mul2(bc(lt), dxy);
    float px = lo(pxy) + lo(axy), py = hi(pxy) + hi(axy);
    pz = pz + lt * dz;
    if (ltx == lt) px += ndx;
    if (lty == lt) py += ndy;
    if (ltz == lt) pz += ndz;
    pxy = pk(px, py);
    return false;
  }

  // HDDAout (:128-142) once step() returned true (`ended`) or the step budget ran out
  __device__ __forceinline__ HitOut result(const DevTree& T, bool ended) const {
    HitOut out;
    out.state = !ended ? 2u : (size == 0.f ? 0u : 1u);
    out.p = V3{lo(pxy), hi(pxy), pz};
    out.mask = (uint32_t)(ltx == lt) | ((uint32_t)(lty == lt) << 1) | ((uint32_t)(ltz == lt) << 2);
    out.i = i;
    finish(T, c, out);
    return out;
  }
};

// The generic fast march: any tree with byte leaves and sizes below 2^20, any start position below 2^21, N5s anywhere.  Since
// march_grid took over the usual rays this is the out-of-line fallback: rays that start beyond the world grid, trees that do
// not qualify for it, and rays march_grid hands over when they reach a cell it does not decide itself (`resume`: iteration
// index, and tMax / its minimum of the step before, from which the mask is derived if the ray ends at once).
static __device__ __noinline__ HitOut march_fast(const DevTree& T, V3 src, V3 dir, V3 idir, uint32_t i0 = 0u, float ltx = 1.f,
                                                 float lty = 1.f, float ltz = 1.f, float lt = 0.f) {
  FastRay r;
  r.init(src, dir, idir);
  r.i = i0, r.ltx = ltx, r.lty = lty, r.ltz = ltz, r.lt = lt;
  if (r.i & 1u) {  // the loop below counts in pairs (kMaxRaySteps is even)
    if (r.step(T)) return r.result(T, true);
    r.i += 1u;
  }
  // Two steps per trip: the cursor's last-voxel registers alternate instead of being copied and the step budget
  // is tested once per trip.
  for (; r.i < kMaxRaySteps; r.i += 2u) {
    if (r.step(T)) break;
    if (r.step(T)) {
      r.i += 1u;
      break;
    }
  }
  return r.result(T, r.i < kMaxRaySteps);  // a break leaves i below the budget
}

// ---------------------------------------------------------------------------------------------
// march_grid: hdda_ray over the world grid (see "World grid" at the top).  The arithmetic of a step is FastRay's, value for
// value; what differs is the lookup: no root level, one 32-bit index base per level instead of a node pointer, coordinates
// never masked.  Preconditions (grid_ray_ok): those of the fast march, a tree with a world grid, and a start position inside
// (-8192, 8192)^3.  From an in-world position no step is longer than 4096 (larger tiles and everything outside the world are
// kEntrySlow cells), so every lookup stays inside the grid or its pads; a ray that meets a slow cell leaves for march_fast.
// ---------------------------------------------------------------------------------------------
struct VoxelId {
  uint32_t x, y, z;  // biased voxel coordinates
};

// Warp votes of march_grid's top-level runs.  They only ever select between code paths that give each lane the same result;
// tests/emu runs the lanes one by one (a "warp" of one lane).
__device__ __forceinline__ uint32_t warp_live() {  // the lanes executing this together with the caller
#ifndef WX_HOST_EMU
  return __activemask();
#else
  return 1u;
#endif
}
__device__ __forceinline__ bool warp_all(uint32_t live, bool p) {  // p holds on every lane of `live` (all of them must be here)
#ifndef WX_HOST_EMU
  return __all_sync(live, p) != 0;
#else
  return p;
#endif
}

__device__ __forceinline__ VoxelId opaque_copy(const VoxelId& s) {
  VoxelId d;
#ifndef WX_HOST_EMU
  asm volatile("mov.b32 %0, %3;\n\tmov.b32 %1, %4;\n\tmov.b32 %2, %5;" : "=r"(d.x), "=r"(d.y), "=r"(d.z) : "r"(s.x), "r"(s.y), "r"(s.z));
#else
  d = s;
#endif
  return d;
}

struct GridRay {
  f32x2 pxy, dxy, ixy, s01xy;
  float pz, dz, iz, s01z;
  float ndx, ndy, ndz;
  float ltx, lty, ltz, lt;
  float size;                // last lookup: 0 = hit, a value in (0, 1) = slow cell, else the step size
  uint32_t dbits;            // 0: w4 and w3 valid (last lookup ended in a leaf), 8: w4 valid, 128: neither
  uint32_t w4, w3;           // index bases of the N4 table / leaf brick the cursor is in (grid_word4 << 4 / grid_word3 << 3)
  uint32_t i;
#if defined(WX_STAGE_LEAF) && !defined(WX_HOST_EMU)
  uint32_t stage_tag, stage_phase;  // A/B experiment: brick in this warp's shared-memory slot; parity of its mbarrier (2 = not initialised)
#endif

  __device__ __forceinline__ void init(V3 src, V3 dir, V3 idir) {
    pxy = pk(src.x, src.y), pz = src.z;
    dxy = pk(dir.x, dir.y), dz = dir.z;
    ixy = pk(idir.x, idir.y), iz = idir.z;
    s01xy = pk(keep(dir.x < 0.f ? 0.f : 1.f), keep(dir.y < 0.f ? 0.f : 1.f));
    s01z = keep(dir.z < 0.f ? 0.f : 1.f);
    ndx = keep(dir.x < 0.f ? -4e-4f : 4e-4f), ndy = keep(dir.y < 0.f ? -4e-4f : 4e-4f), ndz = keep(dir.z < 0.f ? -4e-4f : 4e-4f);
    ltx = 1.f, lty = 1.f, ltz = 1.f, lt = 0.f;
    size = 1.f;
    dbits = 128u, w4 = 0u, w3 = 0u;
    i = 0;
#if defined(WX_STAGE_LEAF) && !defined(WX_HOST_EMU)
    stage_tag = 0xfffffffeu, stage_phase = 2u;
#endif
  }

  // One iteration of hdda_ray's loop body (:90-122) without the counter: `last` is the voxel of the lookup before, `cur`
  // receives this one's (the caller alternates two VoxelIds, so nothing is copied).  Returns true when the ray left the
  // loop: a hit (size == 0) or a slow cell (0 < size < 1).
  // TOL (tolerance mode, WX_OPT_MARCH = 1): p += t * dir is one fused multiply-add per axis -- one rounding instead of two,
  // three instructions fewer; no longer bit-identical to the strict restatement, but within the north-star bar against it.
  template <bool TOL>
  __device__ __forceinline__ bool step(const DevTree& T, const VoxelId& last, VoxelId& cur) {
    const f32x2 txy = add2_rd(pxy, bc(kMagic));
    const float tz = __fadd_rd(pz, kMagic);
    const uint32_t x = (uint32_t)txy, y = (uint32_t)(txy >> 32), z = __float_as_uint(tz);
    const uint32_t dv = ((x ^ last.x) | dbits) | (y ^ last.y) | (z ^ last.z);
    cur.x = x, cur.y = y, cur.z = z;
    // lookup L(pos): from the deepest node the cursor still holds
    uint32_t lvl = dv >= 128u ? 5u : (dv >= 8u ? 4u : 3u);  // level whose table is read next; 0 = walk finished
    if (lvl == 5u) {
      const uint32_t e = __ldg(T.grid + (uint32_t)((x >> 7) * kGS2 + (y >> 7) * kGS + (z >> 7) + kGridK));
      if ((int32_t)e >= 0) dbits = 128u, size = __uint_as_float(e), lvl = 0u;
      else w4 = e << 4, lvl = 4u;  // the shift drops the flag; an operation here (not a copy) lets the load target `size` directly
    }
    if (lvl == 4u) {
      const uint32_t e = __ldg(T.f4 + (uint32_t)((x >> 3) * 256u + (y >> 3) * 16u + (z >> 3) + w4));
      if ((int32_t)e >= 0) dbits = 8u, size = __uint_as_float(e), lvl = 0u;
      else {
        w3 = e << 3, lvl = 3u;
#if defined(WX_PREFETCH_LEAF) && !defined(WX_HOST_EMU)
        // A/B experiment: on entering a leaf, pull its whole 512-byte brick (four 128-byte lines) towards L1 -- a ray that grazes the
        // shell otherwise takes one L2 round trip per voxel it advances in x (a 32-byte sector holds 4 y x 8 z voxels of one x).
        {
          const uint8_t* brick = T.l3 + (uint32_t)(w3 + ((x & ~7u) * 64u + (y & ~7u) * 8u + (z & ~7u)));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(brick));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(brick + 128));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(brick + 256));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(brick + 384));
        }
#endif
      }
    }
#if defined(WX_STAGE_LEAF) && !defined(WX_HOST_EMU)
    // A/B EXPERIMENT (never the default; VERDICT r1 item 6, north_star "leaf bricks staged in shared memory, TMA where the layout
    // allows"): when at least WX_STAGE_MIN lanes of the warp are about to read the same leaf, its 512-byte brick is copied into a
    // per-warp slot of shared memory -- WX_STAGE_LEAF=1: cooperatively with 16-byte loads (LDG.128 + STS.128), =2: one
    // cp.async.bulk (TMA, UBLKCP) completing on an mbarrier -- and those lanes read their voxel from there; the slot stays valid
    // while the warp keeps returning to that leaf.  The decision is taken where all lanes still marching are converged (after
    // the per-level walk), so the slot is only ever touched by all live lanes together.  Results are unchanged.
    {
      __shared__ __align__(128) uint8_t s_brick[8][512];
      __shared__ __align__(8) unsigned long long s_bar[8];
      const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
      const bool leaf = lvl == 3u;
      const uint32_t brick = leaf ? w3 + ((x & ~7u) * 64u + (y & ~7u) * 8u + (z & ~7u)) : 0xffffffffu;  // byte offset of the leaf's brick
      const uint32_t live = __activemask();
      const uint32_t want = __ballot_sync(live, leaf);
      if (want != 0u) {
        const uint32_t lead_brick = __shfl_sync(live, brick, __ffs(want) - 1);
        const uint32_t same = __ballot_sync(live, brick == lead_brick);
        if (stage_tag != lead_brick && __popc(same) >= WX_STAGE_MIN) {
#if WX_STAGE_LEAF == 1
          const uint32_t rank = __popc(live & ((1u << lane) - 1u)), n = __popc(live);
          for (uint32_t k = rank; k < 32u; k += n)
            reinterpret_cast<uint4*>(s_brick[warp])[k] = __ldg(reinterpret_cast<const uint4*>(T.l3 + lead_brick) + k);
#else
          const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[warp]), dst = (uint32_t)__cvta_generic_to_shared(s_brick[warp]);
          if (stage_phase == 2u) {  // first use by this warp: initialise its barrier
            if (lane == (uint32_t)(__ffs(live) - 1)) {
              asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
              asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            stage_phase = 0u;
            __syncwarp(live);
          }
          if (lane == (uint32_t)(__ffs(live) - 1)) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 512;" ::"r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 512, [%2];" ::"r"(dst),
                         "l"(T.l3 + lead_brick), "r"(bar)
                         : "memory");
          }
          uint32_t done = 0u;
          while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(stage_phase) : "memory");
          stage_phase ^= 1u;
#endif
          stage_tag = lead_brick;
          __syncwarp(live);
        }
        if (leaf) {
          dbits = 0u;
          if (brick == stage_tag) size = u32_to_float(s_brick[warp][((x & 7u) << 6) | ((y & 7u) << 3) | (z & 7u)]);
          else size = u32_to_float(__ldg(T.l3 + (uint32_t)(x * 64u + y * 8u + z + w3)));
        }
        __syncwarp(live);  // nobody restages while a lane still reads
      }
    }
#else
    if (lvl == 3u) {
      dbits = 0u;
      size = u32_to_float(__ldg(T.l3 + (uint32_t)(x * 64u + y * 8u + z + w3)));
    }
#endif
    if (size == 0.f || size >= 1.f) WX_EMU_STEP(dv >= 128u ? 128u : dv, dbits);  // (tests/emu only; a slow cell's lookup is redone, and traced, by march_fast)
    if (size < 1.f) return true;  // hit, or a slow cell
    advance<TOL>(txy, tz);
    return false;
  }

  // The step itself (:104-121) from the position whose floor is (txy, tz) (kMagic-biased) through a cell of pitch `size`.
  template <bool TOL>
  __device__ __forceinline__ void advance(const f32x2 txy, const float tz) {
    const float r = rcp_approx(size);
    const float hr = 0.5f * r;
    const f32x2 xfxy = add2(txy, bc(-kMagic));                         // float(floor(p)), exact
    const float xfz = tz - kMagic;
    const f32x2 qxy = add2_rd(fma2(xfxy, bc(r), bc(hr)), bc(kMagic));  // kMagic + floor(p / size)
    const float qz = __fadd_rd(fmaf(xfz, r, hr), kMagic);
    const f32x2 kxy = add2(qxy, bc(-kMagic));                          // see FastRay::step for the argument
    const float kz = qz - kMagic;
    const f32x2 nmxy = fma2(kxy, bc(size), pk(-lo(pxy), -hi(pxy)));    // -modulo_vec3f(p, size)
    const float nmz = fmaf(kz, size, -pz);
    const f32x2 tmxy = mul2(ixy, fma2(bc(size), s01xy, nmxy));         // tMax
    const float tmz = iz * fmaf(size, s01z, nmz);
    ltx = lo(tmxy), lty = hi(tmxy), ltz = tmz;
    lt = fminf(fminf(ltx, lty), ltz);
    float px, py;
    if (TOL) {
      const f32x2 nxy = fma2(bc(lt), dxy, pxy);
      px = lo(nxy), py = hi(nxy);
      pz = fmaf(lt, dz, pz);
    } else {
      const f32x2 axy = mul2(bc(lt), dxy);
      px = lo(pxy) + lo(axy), py = hi(pxy) + hi(axy);
      pz = pz + lt * dz;
    }
    if (ltx == lt) px += ndx;
    if (lty == lt) py += ndy;
    if (ltz == lt) pz += ndz;
    pxy = pk(px, py);
  }

  // A run of top-level steps (march_grid): while EVERY lane of `live` (the lanes that entered together) finds a plain tile in
  // the world grid -- size >= 128, i.e. every top-level tile that is not active -- all of them step, with no cursor test, no
  // level dispatch and no divergence inside the loop.  The first lookup that is something else for some lane ends the run for
  // everybody; each lane then finishes ITS lookup of that round: a plain tile is stepped through; a child leaves the cursor
  // saying "inside that N4, at this voxel" (`pending`, dbits = 8), so that the generic step continues the lookup from the N4
  // table; a hit or a slow cell ends the march (returns true, i not counted, as in step()).  `limit`: a lane only steps here
  // while i < limit.
  template <bool TOL>
  __device__ __forceinline__ bool top_run(const DevTree& T, VoxelId& pending, uint32_t live, uint32_t limit) {
    for (;;) {
      const f32x2 txy = add2_rd(pxy, bc(kMagic));
      const float tz = __fadd_rd(pz, kMagic);
      const uint32_t x = (uint32_t)txy, y = (uint32_t)(txy >> 32), z = __float_as_uint(tz);
      const uint32_t e = __ldg(T.grid + (uint32_t)((x >> 7) * kGS2 + (y >> 7) * kGS + (z >> 7) + kGridK));
      const bool plain = (int32_t)e >= 0x43000000 && i < limit;  // positive floats order like ints
      if (warp_all(live, plain)) {
        size = __uint_as_float(e);
        WX_EMU_STEP(128u, 128u);
        advance<TOL>(txy, tz);
        i += 1u;
        continue;
      }
      // the run is over; this lane's own lookup:
      if (plain) {
        size = __uint_as_float(e);
        WX_EMU_STEP(128u, 128u);
        advance<TOL>(txy, tz);
        i += 1u;
        return false;
      }
      if ((int32_t)e >= 0x43000000) return false;  // out of budget for this loop: the generic steps take over (and redo the lookup)
      if ((int32_t)e >= 0) {  // 0.0f = active tile (hit), kEntryVoid / kEntrySlow
        size = __uint_as_float(e);
        if (e == 0u) WX_EMU_STEP(128u, 128u);
        return true;
      }
      w4 = e << 4, dbits = 8u;
      pending.x = x, pending.y = y, pending.z = z;
      return false;
    }
  }
};

// Tolerance mode only: a ray that starts outside the bounding box of everything it could hit and crosses it starts two voxels
// before its entry point instead (the steps through the empty tiles in front of the box are skipped; the iteration count
// changes, which is why render mode 2 never runs in this mode).  Rays that miss the box are left alone: the colour of an
// out-of-bounds pixel depends on the axis of the ray's LAST step, which only the full march knows.
__device__ __forceinline__ V3 clip_to_bbox(const DevTree& T, V3 src, V3 dir, V3 idir) {
  if (T.bb_lo[0] > T.bb_hi[0] || out_of_bounds(src.x, src.y, src.z)) return src;  // (a start outside the world ends at once, :100-103)
  const float ax = (T.bb_lo[0] - src.x) * idir.x, bx = (T.bb_hi[0] - src.x) * idir.x;
  const float ay = (T.bb_lo[1] - src.y) * idir.y, by = (T.bb_hi[1] - src.y) * idir.y;
  const float az = (T.bb_lo[2] - src.z) * idir.z, bz = (T.bb_hi[2] - src.z) * idir.z;
  const float t_near = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
  const float t_far = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
  if (!(t_near < t_far) || !(t_near > 2.f)) return src;  // misses the box, or starts inside / right in front of it (NaNs: no clip)
  const float t = t_near - 2.f;
  return V3{fmaf(t, dir.x, src.x), fmaf(t, dir.y, src.y), fmaf(t, dir.z, src.z)};
}

template <bool TOL>
__device__ __forceinline__ HitOut march_grid(const DevTree& T, V3 src, V3 dir, V3 idir) {
  GridRay r;
  r.init(TOL ? clip_to_bbox(T, src, dir, idir) : src, dir, idir);
  // Two steps per trip over two alternating VoxelIds: nothing is copied inside the loop, and the step budget is tested once
  // per trip.  The copies at the exits are opaque so that the compiler does not merge a and b.
  // Top-level runs (-DWX_TOP_RUN, an A/B variant that is NOT the default): 40 % of the warp steps of the bench frame read the
  // world grid only (rays crossing empty tiles), and after such a step the next lookup starts at the grid again.  While that
  // holds for every lane of the warp still marching, the lanes can run top_run(), a loop of their own without cursor test,
  // level dispatch and convergence barriers (45 instead of 53 instructions per step); a lane that meets a child there hands
  // its lookup to the generic step below, so lanes may leave a run with different iteration counts (hence the single last
  // step after the loop).  Measured: the warp vote that decides whether to enter a run costs 8 instructions per trip of the
  // generic loop (VOTE.ANY + R2UR + BRA.DIV + VOTE.ALL + ...), which eats the gain: 0.756 ms with runs, 0.740 ms without.
  VoxelId a{0u, 0u, 0u}, b{0u, 0u, 0u}, v{0u, 0u, 0u};
  bool ended = false;
  while (r.i < kMaxRaySteps - 1u) {  // two lookups or more to go
#ifdef WX_TOP_RUN  // A/B: measured 2 % SLOWER than without (profiles/r2_top_run_ab.txt) -- not the default
    const uint32_t live = warp_live();
    if (warp_all(live, r.dbits == 128u && r.i < kMaxRaySteps - 2u) && r.template top_run<TOL>(T, a, live, kMaxRaySteps - 2u)) {
      ended = true;  // (the voxel is not needed for an end on the top level)
      break;
    }
#endif
    if (r.template step<TOL>(T, a, b)) {
      v = opaque_copy(b), ended = true;
      break;
    }
    if (r.template step<TOL>(T, b, a)) {
      v = opaque_copy(a), r.i += 1u, ended = true;
      break;
    }
    r.i += 2u;
  }
  if (!ended) {
    if (r.i == kMaxRaySteps - 1u) {  // one lookup left (the count became odd in a top-level run)
      if (r.template step<TOL>(T, a, b)) ended = true;
      else r.i += 1u;
      v = opaque_copy(b);
    } else {
      v = opaque_copy(a);  // out of steps: the voxel of the last lookup
    }
  }
  const V3 p = V3{lo(r.pxy), hi(r.pxy), r.pz};
  HitOut out;
  out.p = p;
  out.mask = (uint32_t)(r.ltx == r.lt) | ((uint32_t)(r.lty == r.lt) << 1) | ((uint32_t)(r.ltz == r.lt) << 2);
  out.i = r.i;
  if (ended && r.size != 0.f) {
    // the usual end of a ray that misses: a cell outside the world without an N5 -- dist 1 at level 0 (:411), then out of bounds (:100-103)
    if (__float_as_uint(r.size) == kEntryVoid && out_of_bounds(p.x, p.y, p.z)) {
      out.state = 1u, out.level = 0u, out.n3 = 0u;
      WX_EMU_STEP(128u, kNoCache);  // (tests/emu only: this lookup read the grid and ended at level 0)
      return out;
    }
    return march_fast(T, p, dir, idir, r.i, r.ltx, r.lty, r.ltz, r.lt);  // slow cell: redo this lookup there
  }
  out.state = ended ? 0u : 2u;
  out.level = r.dbits == 0u ? 3u : (r.dbits == 8u ? 2u : 1u);
  // a ray that ran out of steps on an in-world cell without an N5 ended at level 0 (the grid stores that cell as a 4096 tile)
  if (!ended && r.dbits == 128u && find_root(T, v.x, v.y, v.z) < 0) out.level = 0u;
  out.n3 = out.level == 3u ? (r.w3 + ((v.x & ~7u) * 64u + (v.y & ~7u) * 8u + (v.z & ~7u))) >> 9 : 0u;
  return out;
}

// march_grid applies to this ray.
__device__ __forceinline__ bool grid_ray_ok(const DevTree& T, V3 src, V3 idir) {
  return T.grid != nullptr && fmaxf(fmaxf(fabsf(idir.x), fabsf(idir.y)), fabsf(idir.z)) < 1e30f &&
         fmaxf(fmaxf(fabsf(src.x), fabsf(src.y)), fabsf(src.z)) < 8192.f;
}

// The fast march applies to this ray (see its preconditions).
__device__ __forceinline__ bool fast_ray_ok(const DevTree& T, V3 src, V3 idir) {
  return T.fast_ok && fmaxf(fmaxf(fabsf(idir.x), fabsf(idir.y)), fabsf(idir.z)) < 1e30f &&
         fmaxf(fmaxf(fabsf(src.x), fabsf(src.y)), fabsf(src.z)) < 2097152.f;
}

template <int MARCH>
__device__ __forceinline__ HitOut hdda_ray(const DevTree& T, V3 src, V3 dir) {
  const V3 idir = V3{1.f / dir.x, 1.f / dir.y, 1.f / dir.z};
#ifndef WX_NO_GRID
  if (grid_ray_ok(T, src, idir)) return march_grid<MARCH == kMarchTolerance>(T, src, dir, idir);
#endif
  if (fast_ray_ok(T, src, idir)) return march_fast(T, src, dir, idir);
  return march_exact(T, src, dir);
}

// secondary rays share one out-of-line copy of the march
template <int MARCH>
static __device__ __noinline__ HitOut hdda_ray_secondary(const DevTree& T, V3 src, V3 dir) { return hdda_ray<MARCH>(T, src, dir); }

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
template <int MARCH>
__device__ __forceinline__ V3 sun_lit(const DevTree& T, const WxState& s, const HitOut& hit, V3 step, V3 N) {
  float I = s.sun_color[3] * WX_K_D * dot3(-sun_dir(s), N);
  I = fmaxf(0.0f, I);
  if (I != 0.0f && hdda_ray_secondary<MARCH>(T, hit.p - (4e-2f * step) * maskf(hit.mask), -sun_dir(s)).state == 0u)
    return WX_BASE_COLOR + (I * sun_rgb(s)) * 0.05f;
  return WX_BASE_COLOR + I * sun_rgb(s);
}

template <int MARCH>
__device__ __forceinline__ V3 reflect_ray1(const DevTree& T, const WxState& s, V3 src, V3 dir) {
  const HitOut hit = hdda_ray_secondary<MARCH>(T, src, dir);
  const V3 step = sign11(dir);
  if (hit.state == 0u) return sun_lit<MARCH>(T, s, hit, step, normalize3((-step) * maskf(hit.mask)));
  if (hit.state == 1u) return wall_flat(normalize3((-step) * maskf(hit.mask)));
  return dir;
}

template <int MARCH>
__device__ __forceinline__ V3 reflect_ray2(const DevTree& T, const WxState& s, V3 src, V3 dir) {
  const HitOut hit = hdda_ray_secondary<MARCH>(T, src, dir);
  const V3 step = sign11(dir);
  if (hit.state == 0u) {
    const V3 N = normalize3((-step) * maskf(hit.mask));
    const V3 rdir = normalize3(dir - (2.0f * N) * dot3(dir, N));
    const V3 rsrc = hit.p - (4e-2f * step) * maskf(hit.mask);
    const V3 rcol = reflect_ray1<MARCH>(T, s, rsrc, rdir);
    const V3 mcol = sun_lit<MARCH>(T, s, hit, step, N);
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
template <int MODE, int MARCH>
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
      if (LN != 0.0f && hdda_ray_secondary<MARCH>(T, hit.p - (4e-2f * step) * maskf(hit.mask), -sun_dir(s)).state == 0u) return I_a;
      return I_a + I_d;
    }
    if (MODE == 4) {
      const V3 N = normalize3((-step) * maskf(hit.mask));
      const V3 mcol = sun_lit<MARCH>(T, s, hit, step, N);
      const V3 rdir = normalize3(dir - (2.0f * N) * dot3(dir, N));
      const V3 rsrc = hit.p - (4e-2f * step) * maskf(hit.mask);
      const V3 rcol = reflect_ray2<MARCH>(T, s, rsrc, rdir);
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

// ---------------------------------------------------------------------------------------------
// cp_main (raycast.comp.wgsl:60-68) for one pixel.  Lives here (not next to the kernels) so that the host emulation of
// tests/emu compiles exactly the code the kernels run.
// ---------------------------------------------------------------------------------------------
struct PixelRef {
  uint32_t x, y, cam;
  bool in_frame;    // a pixel of the frame this launch owns
  bool dispatched;  // inside the reference's dispatch (wgpu_context.rs:281)
};

template <int MODE, bool AOV, int MARCH>
__device__ __forceinline__ void shade_and_store(const RenderParams& P, const PixelRef& q, const HitOut& hit, V3 dir) {
  const WxState& s = (P.n_states == 1) ? P.s0 : P.states[q.cam];
  const size_t pix = ((size_t)q.cam * P.height + q.y) * P.width + q.x;
  const V3 col = shade<MODE, MARCH>(P.tree, s, hit, dir);
  P.rgba[pix] = make_uchar4((unsigned char)unorm8(col.x), (unsigned char)unorm8(col.y), (unsigned char)unorm8(col.z), 255);
  if (AOV) {
    const AovPtrs& a = P.aov;
    if (a.state) a.state[pix] = (uint8_t)hit.state;
    if (a.voxel) {
      a.voxel[3 * pix + 0] = __float2int_rd(hit.p.x);
      a.voxel[3 * pix + 1] = __float2int_rd(hit.p.y);
      a.voxel[3 * pix + 2] = __float2int_rd(hit.p.z);
    }
    if (a.leaf) a.leaf[pix] = hit.level == 3u ? (int32_t)hit.n3 : -1;
    if (a.level) a.level[pix] = (uint8_t)hit.level;
    if (a.iters) a.iters[pix] = hit.i;
    if (a.depth) {
      const V3 d = hit.p - V3{s.eye[0], s.eye[1], s.eye[2]};
      a.depth[pix] = sqrtf(dot3(d, d));
    }
    if (a.mask) a.mask[pix] = (uint8_t)hit.mask;
    if (a.pos) a.pos[3 * pix + 0] = hit.p.x, a.pos[3 * pix + 1] = hit.p.y, a.pos[3 * pix + 2] = hit.p.z;
  }
}

// cp_main (:60-68) for one pixel: ray generation, hdda_ray, ray_trace, store.
// Returns the iteration count of the primary ray (0 for pixels outside the frame / the dispatch): the cost the launcher's
// long-tiles-first list is built from.
template <int MODE, bool AOV, int MARCH = kMarchExact>
__device__ __forceinline__ uint32_t render_pixel(const RenderParams& P, const PixelRef& q) {
  if (!q.in_frame) return 0u;
  if (!q.dispatched) {  // never dispatched by the reference: zero-initialised texel (and zeroed AOVs)
    const size_t pix = ((size_t)q.cam * P.height + q.y) * P.width + q.x;
    P.rgba[pix] = make_uchar4(0, 0, 0, 0);
    if (AOV) {
      const AovPtrs& a = P.aov;
      if (a.state) a.state[pix] = 0;
      if (a.voxel) a.voxel[3 * pix + 0] = a.voxel[3 * pix + 1] = a.voxel[3 * pix + 2] = 0;
      if (a.leaf) a.leaf[pix] = 0;
      if (a.level) a.level[pix] = 0;
      if (a.iters) a.iters[pix] = 0;
      if (a.depth) a.depth[pix] = 0.f;
      if (a.mask) a.mask[pix] = 0;
      if (a.pos) a.pos[3 * pix + 0] = a.pos[3 * pix + 1] = a.pos[3 * pix + 2] = 0.f;
    }
    return 0u;
  }
  // the ray basis: from the constant bank for a single state (the usual frame), else from the batch in global memory
  V3 u, mv, wp, eye;
  if (P.n_states == 1) {
    u = V3{P.s0.u[0], P.s0.u[1], P.s0.u[2]}, mv = V3{P.s0.mv[0], P.s0.mv[1], P.s0.mv[2]};
    wp = V3{P.s0.wp[0], P.s0.wp[1], P.s0.wp[2]}, eye = V3{P.s0.eye[0], P.s0.eye[1], P.s0.eye[2]};
  } else {
    const float4* s4 = reinterpret_cast<const float4*>(P.states + q.cam);  // eye, u, mv, wp are the float4s 8..11 of the state
    const float4 e = __ldg(s4 + 8), a = __ldg(s4 + 9), b = __ldg(s4 + 10), c = __ldg(s4 + 11);
    eye = V3{e.x, e.y, e.z}, u = V3{a.x, a.y, a.z}, mv = V3{b.x, b.y, b.z}, wp = V3{c.x, c.y, c.z};
  }
  const float px = (float)q.x + 0.001f, py = (float)q.y + 0.001f;
  const V3 dir = normalize3((px * u + py * mv) + wp);
  const HitOut hit = hdda_ray<MARCH>(P.tree, eye, dir);
  shade_and_store<MODE, AOV, MARCH>(P, q, hit, dir);
  return hit.i;
}

// ---------------------------------------------------------------------------------------------
// Work distribution of the persistent kernel with a CTA-level chunk queue (raycast_persistent_cta, wx_raycast.cu).  Lives here
// so that tests/emu can run the ticket protocol with real threads on the CPU (tests/test_device_emu.py).
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kQueueDone = 0xffffffffu, kNoTile = 0xffffffffu;

__device__ __forceinline__ PixelRef pixel_of_chunk_tile(const RenderParams& P, uint32_t chunk, uint32_t t, uint32_t lane) {
  const uint32_t per_cam = P.chunks_x * P.chunks_y;
  const uint32_t cam_i = chunk / per_cam, rem = chunk - cam_i * per_cam;
  const uint32_t cy = rem / P.chunks_x, cx = rem - cy * P.chunks_x;
  PixelRef q;
  q.x = cx * 32u + (t & 7u) * 4u + (lane & 3u);
  const uint32_t vrow = cy * 16u + (t >> 3) * 8u + (lane >> 2);  // row among the rows this launch owns
  const uint32_t band = (vrow / P.band_rows) * P.shard_count + P.shard_index;
  q.y = P.row_base + band * P.band_rows + vrow % P.band_rows;
  q.cam = P.cam_base + cam_i;
  q.in_frame = q.x < P.width && q.y < P.row_end && vrow < P.own_bands * P.band_rows;
  q.dispatched = q.x < P.disp_w && q.y < P.disp_h;
  return q;
}

// Lane 0 of a warp: the next (chunk, tile) for this warp, or kNoTile when the frame is done.  `priv` is a chunk this warp
// fetched but could not publish (another warp published one at the same moment): it renders that one by itself.
__device__ __forceinline__ uint32_t next_ticket(const RenderParams& P, uint32_t* s_state, uint32_t& priv, uint32_t& tile_in_chunk) {
  if (priv != kNoTile) {  // private chunk: tiles 1..15 (tile 0 was rendered when it was fetched)
    const uint32_t chunk = priv >> 5, t = priv & 31u;
    priv = t + 1u < 16u ? priv + 1u : kNoTile;
    tile_in_chunk = t;
    return chunk;
  }
  for (;;) {
    const uint32_t old = atomicAdd(s_state, 0u);
    if (old != kQueueDone && (old & 31u) < 16u) {
      if (atomicCAS(s_state, old, old + 1u) != old) continue;
      tile_in_chunk = old & 31u;
      return old >> 5;
    }
    if (old == kQueueDone) return kNoTile;
    const uint32_t g = atomicAdd(P.work_counter, 1u);
    if (g >= P.n_chunks) {  // nothing left globally; a chunk published meanwhile by a CTA-mate is still served above
      (void)atomicCAS(s_state, old, kQueueDone);
      continue;
    }
    tile_in_chunk = 0u;
    if (atomicCAS(s_state, old, (g << 5) | 1u) != old) priv = (g << 5) | 1u;  // lost the race: keep g for ourselves
    return g;
  }
}

}  // namespace wx
