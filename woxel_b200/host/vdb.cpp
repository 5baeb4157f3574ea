// vdb.cpp -- VDB345 over flat node arenas: set/get_voxel, compute_sdf, to_flat.
// Reference behaviour: src/vdb/vdb345.rs (cited per function).
#include "vdb.hpp"

#include <algorithm>
#include <cstring>

namespace woxel::vdb {

namespace {
inline void set_bit(uint64_t* m, Offset o) { m[o >> 6] |= 1ull << (o & 63); }
inline bool get_bit(const uint64_t* m, Offset o) { return (m[o >> 6] >> (o & 63)) & 1ull; }
constexpr uint32_t kInf = 0xFFFFFFFEu;  // "MAX - 1 so adding 1 doesn't wrap around" (vdb345.rs:300-319)
}  // namespace

// ---------------------------------------------------------------------------------------------
// set_voxel / get_voxel (vdb345.rs:26-106)
// ---------------------------------------------------------------------------------------------
uint32_t VDB345::descend_create(GlobalCoordinates p) {
  const GlobalCoordinates key = N5::global_to_node(p);
  RootData& rd = root[key];
  if (!rd.is_node) {  // vacant entry or a root tile: becomes a node (vdb345.rs:32-40)
    rd.is_node = true;
    rd.node = (uint32_t)n5.size();
    rd.tile_value = 0, rd.tile_active = false;
    n5.emplace_back();
    n5.back().origin = p;  // N5::new(p): the reference stores the voxel, the map key is the aligned origin
  }
  const uint32_t i5 = rd.node;
  const Offset o5 = N5::global_to_offset(p);
  if (!n5[i5].child(o5)) {
    n5[i5].slot[o5] = (uint32_t)n4.size();
    set_bit(n5[i5].child_mask, o5);
    n4.emplace_back();
  }
  const uint32_t i4 = n5[i5].slot[o5];
  const Offset o4 = N4::global_to_offset(p);
  if (!n4[i4].child(o4)) {
    n4[i4].slot[o4] = (uint32_t)n3.size();
    set_bit(n4[i4].child_mask, o4);
    n3.emplace_back();
  }
  return n4[i4].slot[o4];
}

void VDB345::set_voxel(GlobalCoordinates p, uint32_t v) {
  Node3& leaf = n3[descend_create(p)];
  const Offset o = N3::global_to_offset(p);
  set_bit(leaf.value_mask, o);
  leaf.slot[o] = v;
}

void VDB345::add_leaf(GlobalCoordinates leaf_origin, const uint64_t value_mask[8]) {
  Node3& leaf = n3[descend_create(leaf_origin)];
  for (Offset o = 0; o < N3::SIZE; ++o)
    if (get_bit(value_mask, o)) {
      set_bit(leaf.value_mask, o);
      leaf.slot[o] = 1;
    }
}

VdbEndpoint VDB345::get_voxel(GlobalCoordinates p) const {
  auto it = root.find(N5::global_to_node(p));
  if (it == root.end()) return {VdbEndpoint::Bkgr, background, 0};
  if (!it->second.is_node) return {VdbEndpoint::Root, it->second.tile_value, 0};
  const Node5& a = n5[it->second.node];
  const Offset o5 = N5::global_to_offset(p);
  if (!a.child(o5)) return {VdbEndpoint::Innr, a.slot[o5], 5};
  const Node4& b = n4[a.slot[o5]];
  const Offset o4 = N4::global_to_offset(p);
  if (!b.child(o4)) return {VdbEndpoint::Innr, b.slot[o4], 4};
  const Node3& c = n3[b.slot[o4]];
  const Offset o3 = N3::global_to_offset(p);
  return {c.active(o3) ? VdbEndpoint::Leaf : VdbEndpoint::Offs, c.slot[o3], 0};
}

std::vector<GlobalCoordinates> VDB345::origins() const {  // vdb345.rs:108-117
  std::vector<GlobalCoordinates> out;
  for (const auto& [key, rd] : root)
    if (rd.is_node) out.push_back(key);
  return out;
}

std::array<size_t, 3> VDB345::count_nodes() const {  // vdb345.rs:266-287
  std::array<size_t, 3> c{0, 0, 0};
  for (const auto& [key, rd] : root) {
    if (!rd.is_node) continue;
    c[0]++;
    const Node5& a = n5[rd.node];
    for (Offset o5 = 0; o5 < N5::SIZE; ++o5) {
      if (!a.child(o5)) continue;
      c[1]++;
      const Node4& b = n4[a.slot[o5]];
      for (uint32_t w = 0; w < N4::MASK_WORDS; ++w) c[2] += (size_t)__builtin_popcountll(b.child_mask[w]);
    }
  }
  return c;
}

uint64_t VDB345::count_leaf_values() const {  // what read.rs:772-806 asserts against file_voxel_count
  uint64_t c = 0;
  for (const auto& [key, rd] : root) {
    if (!rd.is_node) continue;
    const Node5& a = n5[rd.node];
    for (Offset o5 = 0; o5 < N5::SIZE; ++o5) {
      if (!a.child(o5)) continue;
      const Node4& b = n4[a.slot[o5]];
      for (Offset o4 = 0; o4 < N4::SIZE; ++o4) {
        if (!b.child(o4)) continue;
        const Node3& l = n3[b.slot[o4]];
        for (uint32_t w = 0; w < N3::MASK_WORDS; ++w) c += (uint64_t)__builtin_popcountll(l.value_mask[w]);
      }
    }
  }
  return c;
}

// ---------------------------------------------------------------------------------------------
// compute_sdf (vdb345.rs:290-628)
//
// The reference runs one forward and one backward sweep over the tree in DFS order (N5 by origin,
// then slots ascending, recursively).  Every inactive slot takes  min(self, neighbour + 1)  over
// the 13 already-visited neighbours of its own level; a neighbour that is a child / active voxel
// / of another level / missing contributes 1.  Two facts let the sweep be restructured without
// changing a single value:
//   (1) slots of the three levels never read each other's distances (a neighbour of another level
//       always contributes exactly 1), so each level is swept on its own, in the same DFS order;
//   (2) within one slot the order of the 13 neighbours is irrelevant (the value is >= 1 throughout,
//       so "= 1" and "min(., 1)" coincide).
// Leaves are swept through a 10^3 halo copy: halo cells hold the neighbour's current distance, or 0
// for anything that contributes 1.
// ---------------------------------------------------------------------------------------------
namespace {

struct Sweep {
  int32_t nb[13][3];
};

Sweep make_sweep(bool backward) {  // vdb345.rs:327-343
  Sweep s;
  int n = 0;
  const int sg = backward ? 1 : -1;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dz = -1; dz <= 1; ++dz) s.nb[n][0] = sg, s.nb[n][1] = dy, s.nb[n][2] = dz, ++n;
  for (int dz = -1; dz <= 1; ++dz) s.nb[n][0] = 0, s.nb[n][1] = sg, s.nb[n][2] = dz, ++n;
  s.nb[n][0] = 0, s.nb[n][1] = 0, s.nb[n][2] = sg;
  return s;
}

// Sweep of the tile slots of one internal node.  `outside(global)` returns the contribution source
// for a neighbour that lies in another node: the neighbour's current distance when it is a tile
// of the same level, or 0 ("contributes 1").
template <class NM, class Node, class Outside>
void sweep_internal(Node& node, GlobalCoordinates origin, const Sweep& sw, bool backward, Outside&& outside) {
  constexpr int32_t cell = 1 << NM::CHILD_TOTAL_LOG2_D;
  for (uint32_t k = 0; k < NM::SIZE; ++k) {
    const Offset o = backward ? NM::SIZE - 1 - k : k;
    if (node.child(o)) continue;
    const LocalCoordinates c = NM::offset_to_child(o);
    uint32_t cur = node.slot[o];
    for (const auto& d : sw.nb) {
      const int32_t nc[3] = {(int32_t)c[0] + d[0], (int32_t)c[1] + d[1], (int32_t)c[2] + d[2]};
      uint32_t src;
      if ((uint32_t)nc[0] < NM::DIM && (uint32_t)nc[1] < NM::DIM && (uint32_t)nc[2] < NM::DIM) {
        const Offset no = NM::child_to_offset({(uint32_t)nc[0], (uint32_t)nc[1], (uint32_t)nc[2]});
        src = node.child(no) ? 0u : node.slot[no];
      } else {
        src = outside(GlobalCoordinates{origin[0] + nc[0] * cell, origin[1] + nc[1] * cell, origin[2] + nc[2] * cell});
      }
      cur = std::min(cur, src + 1);
    }
    node.slot[o] = cur;
  }
}

}  // namespace

void VDB345::compute_sdf() {
  // ---- initialise (vdb345.rs:292-325) ----
  std::vector<uint32_t> order5;  // arena indices of N5s in sorted-origin order
  std::vector<GlobalCoordinates> origin5;
  for (const auto& [key, rd] : root)
    if (rd.is_node) order5.push_back(rd.node), origin5.push_back(key);

  // DFS lists of N4s and leaves with their global origins
  struct Ref {
    uint32_t idx;
    GlobalCoordinates origin;
  };
  std::vector<Ref> dfs4, dfs3;
  for (size_t r = 0; r < order5.size(); ++r) {
    Node5& a = n5[order5[r]];
    for (Offset o5 = 0; o5 < N5::SIZE; ++o5) {
      if (!a.child(o5)) {
        a.slot[o5] = kInf;
        continue;
      }
      const LocalCoordinates c5 = N5::offset_to_child(o5);
      const GlobalCoordinates g4 = {origin5[r][0] + (int32_t)c5[0] * 128, origin5[r][1] + (int32_t)c5[1] * 128, origin5[r][2] + (int32_t)c5[2] * 128};
      dfs4.push_back({a.slot[o5], g4});
      Node4& b = n4[a.slot[o5]];
      for (Offset o4 = 0; o4 < N4::SIZE; ++o4) {
        if (!b.child(o4)) {
          b.slot[o4] = kInf;
          continue;
        }
        const LocalCoordinates c4 = N4::offset_to_child(o4);
        dfs3.push_back({b.slot[o4], {g4[0] + (int32_t)c4[0] * 8, g4[1] + (int32_t)c4[1] * 8, g4[2] + (int32_t)c4[2] * 8}});
        Node3& l = n3[b.slot[o4]];
        for (Offset o3 = 0; o3 < N3::SIZE; ++o3)
          if (!l.active(o3)) l.slot[o3] = kInf;
      }
    }
  }

  // neighbour classification for slots that lie in another node
  auto outside5 = [this](GlobalCoordinates g) -> uint32_t {
    const VdbEndpoint e = get_voxel(g);
    return (e.kind == VdbEndpoint::Innr && e.level == 5) ? e.value : 0u;
  };
  auto outside4 = [this](GlobalCoordinates g) -> uint32_t {
    const VdbEndpoint e = get_voxel(g);
    return (e.kind == VdbEndpoint::Innr && e.level == 4) ? e.value : 0u;
  };
  // leaf containing voxel g, or nullptr when g does not resolve to a leaf
  auto find_leaf = [this](GlobalCoordinates g) -> const Node3* {
    auto it = root.find(N5::global_to_node(g));
    if (it == root.end() || !it->second.is_node) return nullptr;
    const Node5& a = n5[it->second.node];
    const Offset o5 = N5::global_to_offset(g);
    if (!a.child(o5)) return nullptr;
    const Node4& b = n4[a.slot[o5]];
    const Offset o4 = N4::global_to_offset(g);
    if (!b.child(o4)) return nullptr;
    return &n3[b.slot[o4]];
  };

  for (int pass = 0; pass < 2; ++pass) {
    const bool backward = pass == 1;
    const Sweep sw = make_sweep(backward);

    // ---- N5-slot tiles (forward :368-396 / backward :508-536) ----
    for (size_t k = 0; k < order5.size(); ++k) {
      const size_t r = backward ? order5.size() - 1 - k : k;
      sweep_internal<N5>(n5[order5[r]], origin5[r], sw, backward, outside5);
    }
    // ---- N4-slot tiles (forward :405-436 / backward :545-576) ----
    for (size_t k = 0; k < dfs4.size(); ++k) {
      const Ref& ref = dfs4[backward ? dfs4.size() - 1 - k : k];
      sweep_internal<N4>(n4[ref.idx], ref.origin, sw, backward, outside4);
    }
    // ---- leaf voxels (forward :444-477 / backward :586-619) ----
    uint32_t halo[10][10][10];
    for (size_t k = 0; k < dfs3.size(); ++k) {
      const Ref& ref = dfs3[backward ? dfs3.size() - 1 - k : k];
      Node3& leaf = n3[ref.idx];
      // halo: only the half-space this sweep reads is needed, but filling all 26 neighbours keeps it simple
      for (int bx = -1; bx <= 1; ++bx)
        for (int by = -1; by <= 1; ++by)
          for (int bz = -1; bz <= 1; ++bz) {
            const int x0 = bx < 0 ? 0 : (bx == 0 ? 1 : 9), x1 = bx < 0 ? 1 : (bx == 0 ? 9 : 10);
            const int y0 = by < 0 ? 0 : (by == 0 ? 1 : 9), y1 = by < 0 ? 1 : (by == 0 ? 9 : 10);
            const int z0 = bz < 0 ? 0 : (bz == 0 ? 1 : 9), z1 = bz < 0 ? 1 : (bz == 0 ? 9 : 10);
            const Node3* src = (bx | by | bz) == 0 ? &leaf : find_leaf({ref.origin[0] + bx * 8, ref.origin[1] + by * 8, ref.origin[2] + bz * 8});
            for (int x = x0; x < x1; ++x)
              for (int y = y0; y < y1; ++y)
                for (int z = z0; z < z1; ++z) {
                  if (!src) {
                    halo[x][y][z] = 0u;
                    continue;
                  }
                  const Offset o = N3::child_to_offset({(uint32_t)((x - 1) & 7), (uint32_t)((y - 1) & 7), (uint32_t)((z - 1) & 7)});
                  halo[x][y][z] = src->active(o) ? 0u : src->slot[o];
                }
          }
      for (uint32_t kk = 0; kk < N3::SIZE; ++kk) {
        const Offset o = backward ? N3::SIZE - 1 - kk : kk;
        if (leaf.active(o)) continue;
        const LocalCoordinates c = N3::offset_to_child(o);
        const int x = (int)c[0] + 1, y = (int)c[1] + 1, z = (int)c[2] + 1;
        uint32_t cur = halo[x][y][z];
        for (const auto& d : sw.nb) cur = std::min(cur, halo[x + d[0]][y + d[1]][z + d[2]] + 1);
        halo[x][y][z] = cur;
        leaf.slot[o] = cur;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// to_flat: origins() + masks() + atlas() of the reference without the atlas (vdb345.rs:108-264)
// ---------------------------------------------------------------------------------------------
FlatTree VDB345::to_flat(bool narrow_leaves) const {
  FlatTree f;
  const auto cnt = count_nodes();
  f.n5 = (uint32_t)cnt[0], f.n4 = (uint32_t)cnt[1], f.n3 = (uint32_t)cnt[2];
  f.origins.reserve((size_t)f.n5 * 3);
  f.kids5.reserve((size_t)f.n5 * 512), f.vals5.reserve((size_t)f.n5 * 512), f.tab5.resize((size_t)f.n5 * N5::SIZE);
  f.kids4.reserve((size_t)f.n4 * 64), f.vals4.reserve((size_t)f.n4 * 64), f.tab4.resize((size_t)f.n4 * N4::SIZE);
  f.vals3.reserve((size_t)f.n3 * 8);

  bool fits_u8 = narrow_leaves;
  if (fits_u8)
    for (const Node3& l : n3) {
      for (Offset o = 0; o < N3::SIZE && fits_u8; ++o)
        if (!l.active(o) && l.slot[o] > 255u) fits_u8 = false;
      if (!fits_u8) break;
    }
  f.narrow = fits_u8;
  if (fits_u8) f.tab3_u8.resize((size_t)f.n3 * N3::SIZE);
  else f.tab3.resize((size_t)f.n3 * N3::SIZE);

  uint32_t i5 = 0, i4 = 0, i3 = 0;
  for (const auto& [key, rd] : root) {
    if (!rd.is_node) continue;  // root tiles are skipped (vdb345.rs:191-194)
    const Node5& a = n5[rd.node];
    f.origins.insert(f.origins.end(), key.begin(), key.end());
    f.kids5.insert(f.kids5.end(), a.child_mask, a.child_mask + N5::MASK_WORDS);
    f.vals5.insert(f.vals5.end(), a.value_mask, a.value_mask + N5::MASK_WORDS);
    uint32_t* t5 = &f.tab5[(size_t)i5 * N5::SIZE];
    for (Offset o5 = 0; o5 < N5::SIZE; ++o5) {
      if (!a.child(o5)) {
        t5[o5] = a.slot[o5];
        continue;
      }
      const Node4& b = n4[a.slot[o5]];
      f.kids4.insert(f.kids4.end(), b.child_mask, b.child_mask + N4::MASK_WORDS);
      f.vals4.insert(f.vals4.end(), b.value_mask, b.value_mask + N4::MASK_WORDS);
      uint32_t* t4 = &f.tab4[(size_t)i4 * N4::SIZE];
      for (Offset o4 = 0; o4 < N4::SIZE; ++o4) {
        if (!b.child(o4)) {
          t4[o4] = b.slot[o4];
          continue;
        }
        const Node3& l = n3[b.slot[o4]];
        f.vals3.insert(f.vals3.end(), l.value_mask, l.value_mask + N3::MASK_WORDS);
        if (fits_u8) {
          uint8_t* t3 = &f.tab3_u8[(size_t)i3 * N3::SIZE];
          for (Offset o3 = 0; o3 < N3::SIZE; ++o3) t3[o3] = l.active(o3) ? 0 : (uint8_t)l.slot[o3];
        } else {
          memcpy(&f.tab3[(size_t)i3 * N3::SIZE], l.slot, sizeof(l.slot));
        }
        t4[o4] = i3++;
      }
      t5[o5] = i4++;
    }
    ++i5;
  }
  return f;
}

WxTreeDesc FlatTree::desc() const {
  WxTreeDesc d;
  memset(&d, 0, sizeof(d));
  d.n5 = n5, d.n4 = n4, d.n3 = n3;
  d.origins = origins.data();
  d.kids5 = kids5.data(), d.vals5 = vals5.data(), d.tab5 = tab5.data();
  d.kids4 = kids4.data(), d.vals4 = vals4.data(), d.tab4 = tab4.data();
  d.vals3 = vals3.data();
  d.tab3 = narrow ? (const void*)tab3_u8.data() : (const void*)tab3.data();
  d.tab3_elem_bytes = narrow ? 1 : 4;
  return d;
}

}  // namespace woxel::vdb
