// vdb.hpp -- host-side VDB345 data model: the surface of the reference's `src/vdb` module
// (data_structure.rs, vdb345.rs, read.rs) that the raycast path needs, re-designed around flat
// node arenas instead of boxed pointer nodes.
//
//   reference                                            here
//   ---------------------------------------------------  -----------------------------------------
//   trait Node (data_structure.rs:17-92)                 NodeMath<LOG2, TOTAL_LOG2>
//   LeafNode / InternalNode / RootNode (:95-258)         Node3 / Node4 / Node5 arenas + std::map root
//   VDB345<u32>::set_voxel / get_voxel (vdb345.rs:26-106) VDB345::set_voxel / get_voxel
//   VDB345::origins / count_nodes (:108-117, :266-287)   same names
//   VDB345::compute_sdf (:290-628)                       VDB345::compute_sdf (same result, flat arrays)
//   VDB345::masks + atlas (:119-264)                     VDB345::to_flat -> FlatTree (no atlas padding;
//                                                        this is what wx_tree_upload consumes)
//   VdbReader (read.rs:55-141)                           VdbReader
#pragma once
#include <array>
#include <cstdint>
#include <istream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/woxel_b200.h"

namespace woxel::vdb {

using GlobalCoordinates = std::array<int32_t, 3>;
using LocalCoordinates = std::array<uint32_t, 3>;
using Offset = uint32_t;

// The reference's `Node` trait: compile-time index maths of one tree level.
template <int LOG2_D_, int TOTAL_LOG2_D_>
struct NodeMath {
  static constexpr int LOG2_D = LOG2_D_;
  static constexpr int LOG2_DD = 2 * LOG2_D_;
  static constexpr int TOTAL_LOG2_D = TOTAL_LOG2_D_;
  static constexpr int CHILD_TOTAL_LOG2_D = TOTAL_LOG2_D_ - LOG2_D_;
  static constexpr uint32_t DIM = 1u << LOG2_D_;
  static constexpr uint32_t TOTAL_DIM = 1u << TOTAL_LOG2_D_;
  static constexpr uint32_t SIZE = 1u << (3 * LOG2_D_);
  static constexpr uint32_t MASK_WORDS = SIZE / 64;

  static GlobalCoordinates global_to_node(GlobalCoordinates g) {
    for (auto& c : g) c = (int32_t)((uint32_t)(c >> TOTAL_LOG2_D) << TOTAL_LOG2_D);
    return g;
  }
  static LocalCoordinates global_to_relative(GlobalCoordinates g) {
    return {(uint32_t)(g[0] & (int32_t)(TOTAL_DIM - 1)), (uint32_t)(g[1] & (int32_t)(TOTAL_DIM - 1)), (uint32_t)(g[2] & (int32_t)(TOTAL_DIM - 1))};
  }
  static LocalCoordinates relative_to_child(LocalCoordinates l) {
    return {l[0] >> CHILD_TOTAL_LOG2_D, l[1] >> CHILD_TOTAL_LOG2_D, l[2] >> CHILD_TOTAL_LOG2_D};
  }
  static Offset child_to_offset(LocalCoordinates c) { return (c[0] << LOG2_DD) | (c[1] << LOG2_D) | c[2]; }
  static Offset global_to_offset(GlobalCoordinates g) { return child_to_offset(relative_to_child(global_to_relative(g))); }
  static LocalCoordinates offset_to_child(Offset o) { return {o >> LOG2_DD, (o >> LOG2_D) & (DIM - 1), o & (DIM - 1)}; }
};
using N3 = NodeMath<3, 3>;
using N4 = NodeMath<4, 7>;
using N5 = NodeMath<5, 12>;

// What get_voxel can end on (data_structure.rs:337-344).
struct VdbEndpoint {
  enum Kind : int { Offs = 0, Leaf = 1, Innr = 2, Root = 3, Bkgr = 4 } kind;
  uint32_t value;  // distance, voxel value, tile value or background
  int level;       // Innr only: 5 = a slot of an N5 (an "N4 tile"), 4 = a slot of an N4
};

struct Node3 {
  uint64_t value_mask[N3::MASK_WORDS] = {};
  uint32_t slot[N3::SIZE] = {};  // voxel value where active, else SDF distance (LeafData::{Value,Tile})
  bool active(Offset o) const { return (value_mask[o >> 6] >> (o & 63)) & 1ull; }
};
struct Node4 {
  uint64_t child_mask[N4::MASK_WORDS] = {};
  uint64_t value_mask[N4::MASK_WORDS] = {};
  uint32_t slot[N4::SIZE] = {};  // arena index of the child leaf, else tile value / SDF distance
  bool child(Offset o) const { return (child_mask[o >> 6] >> (o & 63)) & 1ull; }
};
struct Node5 {
  uint64_t child_mask[N5::MASK_WORDS] = {};
  uint64_t value_mask[N5::MASK_WORDS] = {};
  std::vector<uint32_t> slot = std::vector<uint32_t>(N5::SIZE, 0u);
  GlobalCoordinates origin{};
  bool child(Offset o) const { return (child_mask[o >> 6] >> (o & 63)) & 1ull; }
};

struct RootData {
  bool is_node = false;
  uint32_t node = 0;       // arena index when is_node
  uint32_t tile_value = 0;
  bool tile_active = false;
};

// Flat, DFS-ordered serialisation consumed by wx_tree_upload (the reference's origins()+masks()+atlas()).
struct FlatTree {
  uint32_t n5 = 0, n4 = 0, n3 = 0;
  std::vector<int32_t> origins;                 // n5 x 3
  std::vector<uint64_t> kids5, vals5, kids4, vals4, vals3;
  std::vector<uint32_t> tab5, tab4;
  std::vector<uint32_t> tab3;                   // filled unless narrow
  std::vector<uint8_t> tab3_u8;                 // filled when narrow (all distances < 256)
  bool narrow = false;
  WxTreeDesc desc() const;
};

enum Compression : uint32_t { NONE = 0, ZIP = 0x1, ACTIVE_MASK = 0x2, BLOSC = 0x4, DEFAULT_COMPRESSION = 0x6 };

struct Metadata {
  std::map<std::string, std::string> strings;
  std::map<std::string, int64_t> ints;
  std::map<std::string, bool> bools;
  std::map<std::string, float> floats;
  bool is_half_float() const {
    auto it = bools.find("is_saved_as_half_float");
    return it != bools.end() && it->second;
  }
};

struct GridDescriptor {
  std::string name = "Demo", instance_parent, grid_type;
  uint64_t grid_pos = 0, block_pos = 0, end_pos = 0;
  uint32_t compression = NONE;
  Metadata meta_data;
};

class VDB345 {
 public:
  void set_voxel(GlobalCoordinates p, uint32_t v);
  VdbEndpoint get_voxel(GlobalCoordinates p) const;
  std::vector<GlobalCoordinates> origins() const;
  std::array<size_t, 3> count_nodes() const;
  uint64_t count_leaf_values() const;
  // Hierarchical two-pass chamfer distance into every inactive slot (vdb345.rs:290-628).
  void compute_sdf();
  // narrow_leaves: emit u8 leaf distances when all of them fit (4x smaller upload).
  FlatTree to_flat(bool narrow_leaves = true) const;

  // Bulk topology construction for procedural scenes: adds a leaf with the given value mask
  // (voxel values = 1).  Equivalent to set_voxel on every set bit.
  void add_leaf(GlobalCoordinates leaf_origin, const uint64_t value_mask[8]);

  std::map<GlobalCoordinates, RootData> root;  // ordered by [x,y,z]: the reference sorts on every walk
  uint32_t background = 0;
  GridDescriptor grid_descriptor;
  std::vector<Node5> n5;
  std::vector<Node4> n4;
  std::vector<Node3> n3;

 private:
  friend class VdbReader;
  uint32_t descend_create(GlobalCoordinates p);  // arena index of the leaf containing p
};

class VdbError : public std::runtime_error {
 public:
  enum Kind { MagicMismatch, UnsupportedVersion, IoError, InvalidCompression, InvalidGridName, InvalidNodeMetadata,
              UnsupportedBloscFormat, InvalidBloscData, UnexpectedMaskLength, Unsupported };
  VdbError(Kind k, const std::string& what) : std::runtime_error(what), kind(k) {}
  Kind kind;
};

struct ArchiveHeader {
  uint32_t file_version = 0, library_major = 0, library_minor = 0;
  std::string uuid;
  bool has_grid_offsets = false;
  uint32_t compression = NONE;
  uint32_t grid_number = 0;
  Metadata meta_data;
};

// .vdb reader (read.rs).  Supports what the reference supports: file versions >= 218, no / zlib /
// Blosc(BloscLZ, LZ4, Snappy, zlib, Zstd codecs; byte or bit shuffle) block compression, active-mask compression,
// half-float storage.
// Decodes one c-blosc 1.x frame (BloscLZ / LZ4 / Snappy / zlib / Zstd codec, byte or bit shuffle); throws VdbError.
std::vector<uint8_t> decompress_blosc_frame(const uint8_t* frame, size_t n, size_t max_bytes = (size_t)-1);

class VdbReader {
 public:
  explicit VdbReader(const std::string& path);
  explicit VdbReader(std::vector<uint8_t> bytes);
  VDB345 read_vdb345_grid(const std::string& name);

  ArchiveHeader header;
  std::map<std::string, GridDescriptor> grid_descriptors;
  struct Cursor;  // byte cursor over the file image (implementation detail of vdb_read.cpp)

 private:
  void parse_header();
  std::vector<uint8_t> buf_;
};

}  // namespace woxel::vdb
