// wx_pack.h -- host-side packing of a WxTreeDesc (reference order: masks + one u32 per slot, what
// src/vdb/vdb345.rs:108-264 origins()/masks()/atlas() produce) into the raycast kernel's own layout
// (wx_device.cuh: entry tables, leaf bricks, biased origins, root cells).  Plain host C++: used by
// wx_tree_upload (wx_api.cu) and by the host emulation of the device code (tests/emu), so that the
// emulation walks exactly the tables the GPU walks.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "wx_device.cuh"

namespace wx {

template <class F>
static void parallel_for(size_t n, F&& f) {
  unsigned hw = std::thread::hardware_concurrency();
  size_t nt = std::min<size_t>(hw ? hw : 1, std::max<size_t>(1, n / 64));
  if (nt <= 1) {
    for (size_t i = 0; i < n; ++i) f(i);
    return;
  }
  std::atomic<size_t> next{0};
  std::vector<std::thread> th;
  auto body = [&]() {
    for (;;) {
      size_t b = next.fetch_add(256);
      if (b >= n) break;
      size_t e = std::min(n, b + 256);
      for (size_t i = b; i < e; ++i) f(i);
    }
  };
  for (size_t t = 1; t < nt; ++t) th.emplace_back(body);
  body();
  for (auto& t : th) t.join();
}

static inline bool bit(const uint64_t* m, size_t i) { return (m[i >> 6] >> (i & 63)) & 1ull; }

static inline uint32_t float_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

// One internal level.  Returns 0, or a negative status; *max_dist receives the largest tile distance.
// A tile entry is the f32 bit pattern of f32(dist) * cell, the `size` of raycast.comp.wgsl:104.
static inline int pack_internal(uint32_t n_nodes, uint32_t slots, float cell, const uint64_t* kids, const uint64_t* vals,
                                const uint32_t* tab, uint32_t n_children, std::vector<uint32_t>& out, uint32_t* max_dist) {
  out.assign((size_t)n_nodes * slots, 0u);
  std::atomic<int> status{0};
  std::atomic<uint32_t> mx{0};
  parallel_for(n_nodes, [&](size_t node) {
    const uint64_t* k = kids + node * (slots / 64);
    const uint64_t* v = vals + node * (slots / 64);
    const uint32_t* t = tab + node * slots;
    uint32_t* o = out.data() + node * slots;
    uint32_t local_max = 0;
    for (uint32_t s = 0; s < slots; ++s) {
      if (bit(v, s)) {
        o[s] = 0u;  // active tile: a hit, whatever the child bit says (raycast.comp.wgsl:431-433)
      } else if (bit(k, s)) {
        if (t[s] >= n_children) {
          status.store(WX_ERR_BAD_TREE);
          return;
        }
        o[s] = kChildFlag | t[s];
      } else {
        if (t[s] & kChildFlag) {
          status.store(WX_ERR_UNSUPPORTED);
          return;
        }
        o[s] = float_bits((float)t[s] * cell);
        local_max = std::max(local_max, t[s]);
      }
    }
    uint32_t cur = mx.load();
    while (local_max > cur && !mx.compare_exchange_weak(cur, local_max)) {
    }
  });
  *max_dist = mx.load();
  return status.load();
}

// The leaf level: one brick per leaf, 0 = active voxel, else the SDF distance.  The largest inactive-voxel distance
// decides the width (one byte per voxel, else u32).  Returns the bits per voxel (8 or 32).
static inline uint32_t pack_leaves(uint32_t n3, const uint64_t* vals3, const void* tab3, uint32_t tab3_elem_bytes,
                                   std::vector<uint8_t>& l3, uint32_t* max_dist) {
  const uint8_t* t8 = (const uint8_t*)tab3;
  const uint32_t* t32 = (const uint32_t*)tab3;
  const bool narrow = tab3_elem_bytes == 1;
  std::atomic<uint32_t> mx3{0};
  parallel_for(n3, [&](size_t leaf) {
    const uint64_t* v = vals3 + leaf * 8;
    uint32_t local_max = 0;
    for (uint32_t s = 0; s < 512; ++s)
      if (!bit(v, s)) local_max = std::max(local_max, narrow ? (uint32_t)t8[leaf * 512 + s] : t32[leaf * 512 + s]);
    uint32_t cur = mx3.load();
    while (local_max > cur && !mx3.compare_exchange_weak(cur, local_max)) {
    }
  });
  const uint32_t max3v = mx3.load();
  const uint32_t leaf_bits = max3v <= 255 ? 8 : 32;
  const uint32_t leaf_shift = leaf_bits == 8 ? 9 : 11;
  l3.assign((size_t)n3 << leaf_shift, 0u);
  parallel_for(n3, [&](size_t leaf) {
    const uint64_t* v = vals3 + leaf * 8;
    uint8_t* o8 = l3.data() + (leaf << leaf_shift);
    uint32_t* o32 = reinterpret_cast<uint32_t*>(o8);
    for (uint32_t s = 0; s < 512; ++s) {
      const uint32_t dist = bit(v, s) ? 0u : (narrow ? (uint32_t)t8[leaf * 512 + s] : t32[leaf * 512 + s]);
      if (leaf_bits == 8) o8[s] = (uint8_t)dist;
      else o32[s] = dist;
    }
  });
  *max_dist = max3v;
  return leaf_bits;
}

// Origins biased like the voxel coordinates the kernel derives from float bits (modular arithmetic).
static inline void bias_origins(uint32_t n5, const int32_t* origins, std::vector<int4>& out) {
  out.resize(n5);
  for (uint32_t i = 0; i < n5; ++i)
    out[i] = make_int4((int)((uint32_t)origins[3 * i] + kBias), (int)((uint32_t)origins[3 * i + 1] + kBias),
                       (int)((uint32_t)origins[3 * i + 2] + kBias), 0);
}

// DevTree::root_grid: the N5 of each 4096^3 cell of [-8192, 8192)^3.
static inline void build_root_grid(uint32_t n5, const int32_t* origins, int16_t root_grid[64]) {
  for (int c = 0; c < 64; ++c) root_grid[c] = (int16_t)kRootNone;
  for (uint32_t i = n5; i-- > 0;) {  // descending: the first of equal origins wins, as in the reference's scan
    const int32_t* o = origins + 3 * i;
    const int64_t cx = ((int64_t)o[0] >> 12) + 2, cy = ((int64_t)o[1] >> 12) + 2, cz = ((int64_t)o[2] >> 12) + 2;
    if ((o[0] & 4095) || (o[1] & 4095) || (o[2] & 4095)) continue;  // an unaligned origin never equals (pos >> 12) << 12
    if (cx < 0 || cx > 3 || cy < 0 || cy > 3 || cz < 0 || cz > 3) continue;
    const bool beyond = cx < 1 || cx > 2 || cy < 1 || cy > 2 || cz < 1 || cz > 2;  // origin component outside [-4096, 0]
    root_grid[cx * 16 + cy * 4 + cz] = i <= (uint32_t)kRootIndexMask ? (int16_t)(i | (beyond ? kRootBeyond : 0)) : (int16_t)kRootScan;
  }
}

// The fast march needs byte leaves and every step size below 2^20 (wx_device.cuh); anything else takes the exact march.
static inline bool fast_march_ok(uint32_t leaf_bits, uint32_t max5, uint32_t max4, uint32_t max3v) {
  return leaf_bits == 8 && (double)max5 * 128.0 < (double)kFastMaxSize && (double)max4 * 8.0 < (double)kFastMaxSize &&
         (double)max3v < (double)kFastMaxSize;
}

// The world grid applies (wx_device.cuh): the fast march's conditions, and index bases that fit 32 bits.
static inline bool world_grid_ok(bool fast_ok, uint32_t n4, uint32_t n3) { return fast_ok && n4 <= kGridMaxN4 && n3 <= kGridMaxN3; }

// Host version of the world grid and the re-encoded N4 tables (the library derives the same on the device from the tables it
// has just uploaded or swept, build_grid_kernel / build_f4_kernel in wx_api.cu, through the same helpers of wx_device.cuh).
// o4 (scratch): biased origin of every N4, from the N5 slot that points to it.
static inline void build_grid_tables(uint32_t n5, uint32_t n4, const std::vector<int4>& origins, const int16_t root_grid[64],
                                     const uint32_t* e5, const uint32_t* e4, std::vector<uint32_t>& grid, std::vector<uint32_t>& f4) {
  std::vector<uint32_t> o4((size_t)n4 * 3u, 0u);
  for (uint32_t i = 0; i < n5; ++i)
    for (uint32_t s = 0; s < 32768u; ++s) {
      const uint32_t e = e5[(size_t)i * 32768u + s];
      if (!(e & kChildFlag)) continue;
      uint32_t* o = o4.data() + (size_t)(e & ~kChildFlag) * 3u;
      o[0] = (uint32_t)origins[i].x + (s >> 10) * 128u, o[1] = (uint32_t)origins[i].y + ((s >> 5) & 31u) * 128u, o[2] = (uint32_t)origins[i].z + (s & 31u) * 128u;
    }
  grid.assign(kGridCells, kEntrySlow);  // the pads stay slow
  parallel_for((size_t)kGS * kGS2, [&](size_t c) {
    grid[(size_t)kGridPad + c] = grid_cell_entry((uint32_t)(c >> 14), (uint32_t)((c >> 7) & 127u), (uint32_t)(c & 127u), root_grid, e5);
  });
  f4.assign((size_t)n4 * 4096u + 1u, 0u);
  parallel_for(n4, [&](size_t node) {
    const uint32_t* o = o4.data() + node * 3u;
    for (uint32_t s = 0; s < 4096u; ++s) {
      const uint32_t e = e4[node * 4096u + s];
      f4[node * 4096u + s] = (e & kChildFlag) ? grid_word3(e & ~kChildFlag, o[0] + (s >> 8) * 8u, o[1] + ((s >> 4) & 15u) * 8u, o[2] + (s & 15u) * 8u) : e;
    }
  });
}

// Bounding box of what a ray can hit, in grid cells (lo xyz, hi xyz inclusive; lo > hi = empty): the in-world cells that are
// children or active tiles (host version; wx_api.cu reduces the same on the device).
static inline void grid_bbox_cells(const std::vector<uint32_t>& grid, int32_t bbox[6]) {
  bbox[0] = bbox[1] = bbox[2] = 1 << 30, bbox[3] = bbox[4] = bbox[5] = -1;
  for (uint32_t cx = 32; cx < 96; ++cx)
    for (uint32_t cy = 32; cy < 96; ++cy)
      for (uint32_t cz = 32; cz < 96; ++cz) {
        const uint32_t e = grid[(size_t)kGridPad + (size_t)cx * kGS2 + (size_t)cy * kGS + cz];
        if (!((e & kChildFlag) || e == 0u)) continue;
        const int32_t c[3] = {(int32_t)cx, (int32_t)cy, (int32_t)cz};
        for (int k = 0; k < 3; ++k) bbox[k] = std::min(bbox[k], c[k]), bbox[3 + k] = std::max(bbox[3 + k], c[k]);
      }
}

// The kernel's view of one replica of the tree.  grid / f4: nullptr when the tree has no world grid.
static inline void fill_dev_tree(DevTree& T, const uint32_t* e5, const uint32_t* e4, const uint8_t* l3, const int4* origins, uint32_t n5,
                                 uint32_t n4, uint32_t n3, uint32_t leaf_shift, bool fast_ok, const int16_t root_grid[64],
                                 const uint32_t* grid = nullptr, const uint32_t* f4 = nullptr, const int32_t* bbox_cells = nullptr) {
  T.e5 = e5, T.e4 = e4, T.l3 = l3, T.origins_g = origins;
  T.grid = (grid && f4) ? grid : nullptr, T.f4 = f4;
  T.bb_lo[0] = 1.f, T.bb_hi[0] = 0.f;  // empty
  if (bbox_cells && bbox_cells[0] <= bbox_cells[3]) {  // grid cells (lo xyz, hi xyz inclusive) -> voxel coordinates
    for (int k = 0; k < 3; ++k) T.bb_lo[k] = (float)((bbox_cells[k] - 64) * 128), T.bb_hi[k] = (float)((bbox_cells[3 + k] - 64 + 1) * 128);
  }
  // a child entry keeps its flag bit: node = adj + entry * node_bytes, adj = base - 2^31 * node_bytes
  T.e4_adj = reinterpret_cast<const char*>(e4) - ((uint64_t)kChildFlag << 14);
  T.l3_adj = reinterpret_cast<const char*>(l3) - ((uint64_t)kChildFlag << leaf_shift);
  T.n5 = n5, T.n4 = n4, T.n3 = n3;
  T.leaf_shift = leaf_shift;
  T.fast_ok = fast_ok ? 1u : 0u;
  memcpy(T.root_grid, root_grid, sizeof(T.root_grid));
#ifdef WX_ROOT_PTRS
  for (int c = 0; c < 64; ++c) {
    const int v = root_grid[c];
    T.root_ptr[c] = v == kRootNone ? kRootPtrNone : v == kRootScan ? kRootPtrScan
                    : ((uint64_t)(e5 + (size_t)(v & kRootIndexMask) * 32768u) | ((v & kRootBeyond) ? 1ull : 0ull));
  }
#endif
}

}  // namespace wx
