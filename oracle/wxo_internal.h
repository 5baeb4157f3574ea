/* wxo_internal.h -- ORACLE (test infrastructure, see wxo.h): shared private structures. */
#ifndef WXO_INTERNAL_H
#define WXO_INTERNAL_H
#include "wxo.h"

#define WXO_N3_SIZE 512
#define WXO_N4_SIZE 4096
#define WXO_N5_SIZE 32768

/* LeafNode<u32, 3> (data_structure.rs:95-110).  LeafData::{Tile(usize), Value(u32)} is kept as a
 * 64-bit payload plus an `is_value` discriminant bit per slot. */
typedef struct WxoN3 {
  uint64_t data[WXO_N3_SIZE];
  uint64_t is_value[8];
  uint64_t value_mask[8];
} WxoN3;

/* InternalNode<u32, N3, 4> (data_structure.rs:163-173): InternalData::{Node(Box), Tile(u32)}.
 * child[i] != NULL <=> InternalData::Node. */
typedef struct WxoN4 {
  WxoN3 *child[WXO_N4_SIZE];
  uint32_t tile[WXO_N4_SIZE];
  uint64_t value_mask[64];
  uint64_t child_mask[64];
} WxoN4;

typedef struct WxoN5 {
  WxoN4 *child[WXO_N5_SIZE];
  uint32_t tile[WXO_N5_SIZE];
  uint64_t value_mask[512];
  uint64_t child_mask[512];
  int32_t origin[3];
} WxoN5;

/* RootData::{Node(Box<N5>), Tile(u32,bool)} keyed by [i32;3] (data_structure.rs:234-247). */
typedef struct WxoRootEntry {
  int32_t key[3];
  WxoN5 *node; /* NULL => Tile */
  uint32_t tile_value;
  int tile_active;
} WxoRootEntry;

struct WxoTree {
  WxoRootEntry *root; /* insertion order */
  size_t n_root, cap_root;
  uint32_t background;
};

struct WxoGpuData {
  uint32_t n[3];    /* n5, n4, n3 */
  uint32_t dim[3];  /* atlas nodes per side */
  uint32_t side[3]; /* atlas texels per side */
  uint32_t *atlas[3];
  uint32_t *mask[5];
  int32_t *origins;
};

static inline int wxo_n3_is_value(const WxoN3 *n, uint32_t i) { return (int)((n->is_value[i >> 6] >> (i & 63)) & 1); }

WxoN3 *wxo_n3_new(void);
WxoN4 *wxo_n4_new(void);
WxoN5 *wxo_n5_new(const int32_t origin[3]);
WxoRootEntry *wxo_root_find(const WxoTree *t, const int32_t key[3]);
WxoRootEntry *wxo_root_insert(WxoTree *t, const int32_t key[3]);
WxoRootEntry **wxo_root_sorted(const WxoTree *t);

#endif
