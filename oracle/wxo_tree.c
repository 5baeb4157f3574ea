/*
 * wxo_tree.c -- ORACLE (test infrastructure, see wxo.h): the reference's VDB345 pointer tree,
 * its index maths, set/get_voxel, compute_sdf and the origins()/masks()/atlas() serialisation.
 *
 * Follows src/vdb/data_structure.rs and src/vdb/vdb345.rs of the reference; each function
 * cites the lines it restates.  Written for fidelity, not speed.
 */
#include "wxo_internal.h"

#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------
 * Node constants (data_structure.rs:17-35).  level 3 = leaf, 4 = N4, 5 = N5.
 * ------------------------------------------------------------------------------------------ */
static int log2d(int level) { return level; }                                   /* LOG2_D        */
static int total_log2d(int level) { return level == 3 ? 3 : level == 4 ? 7 : 12; } /* TOTAL_LOG2_D  */
static int child_total_log2d(int level) { return total_log2d(level) - log2d(level); }

/* data_structure.rs:73-75 */
void wxo_global_to_node(int level, const int32_t g[3], int32_t out[3]) {
  int t = total_log2d(level);
  for (int i = 0; i < 3; i++) out[i] = (int32_t)((uint32_t)(g[i] >> t) << t); /* arithmetic >> then << */
}

/* data_structure.rs:58-70 */
uint32_t wxo_global_to_offset(int level, const int32_t g[3]) {
  int t = total_log2d(level), c = child_total_log2d(level), l = log2d(level);
  int32_t m = (1 << t) - 1;
  return (uint32_t)((((g[0] & m) >> c) << (2 * l)) | (((g[1] & m) >> c) << l) | ((g[2] & m) >> c));
}

/* data_structure.rs:80-87 */
void wxo_offset_to_child(int level, uint32_t offset, uint32_t out[3]) {
  int l = log2d(level);
  uint32_t dim = 1u << l;
  out[0] = offset >> (2 * l);
  out[1] = (offset >> l) & (dim - 1);
  out[2] = offset & (dim - 1);
}

/* data_structure.rs:89-91 */
uint32_t wxo_child_to_offset(int level, const uint32_t c[3]) {
  int l = log2d(level);
  return (c[0] << (2 * l)) | (c[1] << l) | c[2];
}

/* ---------------------------------------------------------------------------------------------
 * Allocation (data_structure.rs:128-140, :182-197)
 * ------------------------------------------------------------------------------------------ */
WxoN3 *wxo_n3_new(void) {
  WxoN3 *n = (WxoN3 *)calloc(1, sizeof(WxoN3)); /* data = Tile(0), masks 0 */
  return n;
}
WxoN4 *wxo_n4_new(void) { return (WxoN4 *)calloc(1, sizeof(WxoN4)); }
WxoN5 *wxo_n5_new(const int32_t origin[3]) {
  WxoN5 *n = (WxoN5 *)calloc(1, sizeof(WxoN5));
  memcpy(n->origin, origin, sizeof(n->origin));
  return n;
}

static void n4_free(WxoN4 *n4) {
  if (!n4) return;
  for (int i = 0; i < WXO_N4_SIZE; i++) free(n4->child[i]);
  free(n4);
}
static void n5_free(WxoN5 *n5) {
  if (!n5) return;
  for (int i = 0; i < WXO_N5_SIZE; i++) n4_free(n5->child[i]);
  free(n5);
}

WxoTree *wxo_tree_new(void) { return (WxoTree *)calloc(1, sizeof(WxoTree)); }

void wxo_tree_free(WxoTree *t) {
  if (!t) return;
  for (size_t i = 0; i < t->n_root; i++) n5_free(t->root[i].node);
  free(t->root);
  free(t);
}

/* HashMap<[i32;3], RootData> lookup (data_structure.rs:239). */
WxoRootEntry *wxo_root_find(const WxoTree *t, const int32_t key[3]) {
  for (size_t i = 0; i < t->n_root; i++) {
    WxoRootEntry *e = &t->root[i];
    if (e->key[0] == key[0] && e->key[1] == key[1] && e->key[2] == key[2]) return e;
  }
  return NULL;
}

/* HashMap insert: a later insert of an equal key replaces the value (read.rs:340-343 from_iter). */
WxoRootEntry *wxo_root_insert(WxoTree *t, const int32_t key[3]) {
  WxoRootEntry *e = wxo_root_find(t, key);
  if (e) {
    n5_free(e->node);
    e->node = NULL;
    return e;
  }
  if (t->n_root == t->cap_root) {
    t->cap_root = t->cap_root ? 2 * t->cap_root : 8;
    t->root = (WxoRootEntry *)realloc(t->root, t->cap_root * sizeof(WxoRootEntry));
  }
  e = &t->root[t->n_root++];
  memset(e, 0, sizeof(*e));
  memcpy(e->key, key, sizeof(e->key));
  return e;
}

static int key_cmp(const void *a, const void *b) {
  const WxoRootEntry *x = *(WxoRootEntry *const *)a, *y = *(WxoRootEntry *const *)b;
  for (int i = 0; i < 3; i++) {
    if (x->key[i] < y->key[i]) return -1;
    if (x->key[i] > y->key[i]) return 1;
  }
  return 0;
}

/* `.iter().sorted_by_key(|(key, _)| *key)` (vdb345.rs:110,134,190,353): lexicographic on [x,y,z]. */
WxoRootEntry **wxo_root_sorted(const WxoTree *t) {
  WxoRootEntry **v = (WxoRootEntry **)malloc((t->n_root + 1) * sizeof(*v));
  for (size_t i = 0; i < t->n_root; i++) v[i] = &t->root[i];
  qsort(v, t->n_root, sizeof(*v), key_cmp);
  return v;
}

/* ---------------------------------------------------------------------------------------------
 * set_voxel (vdb345.rs:26-66)
 * ------------------------------------------------------------------------------------------ */
void wxo_set_voxel(WxoTree *t, int32_t x, int32_t y, int32_t z, uint32_t v) {
  int32_t p[3] = {x, y, z}, key[3];
  wxo_global_to_node(5, p, key); /* root_key_from_coords, data_structure.rs:264-268 */
  uint32_t bit_index_4 = wxo_global_to_offset(5, p);
  uint32_t bit_index_3 = wxo_global_to_offset(4, p);
  uint32_t bit_index_0 = wxo_global_to_offset(3, p);

  WxoRootEntry *e = wxo_root_find(t, key);
  if (!e) { /* .or_insert(RootData::Node(N5::new(p))) */
    e = wxo_root_insert(t, key);
    e->node = wxo_n5_new(p);
  }
  if (!e->node) { /* RootData::Tile -> replaced by a node */
    e->node = wxo_n5_new(p);
    e->tile_value = 0;
    e->tile_active = 0;
  }
  WxoN5 *n5 = e->node;
  if (!n5->child[bit_index_4]) n5->child[bit_index_4] = wxo_n4_new(); /* Tile -> Node */
  n5->child_mask[bit_index_4 >> 6] |= 1ull << (bit_index_4 & 63);
  WxoN4 *n4 = n5->child[bit_index_4];
  if (!n4->child[bit_index_3]) n4->child[bit_index_3] = wxo_n3_new();
  n4->child_mask[bit_index_3 >> 6] |= 1ull << (bit_index_3 & 63);
  WxoN3 *n3 = n4->child[bit_index_3];
  n3->value_mask[bit_index_0 >> 6] |= 1ull << (bit_index_0 & 63);
  n3->is_value[bit_index_0 >> 6] |= 1ull << (bit_index_0 & 63); /* data = LeafData::Value(v) */
  n3->data[bit_index_0] = v;
}

void wxo_set_voxels(WxoTree *t, const int32_t *xyz, size_t n, uint32_t v) {
  for (size_t i = 0; i < n; i++) wxo_set_voxel(t, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], v);
}

/* ---------------------------------------------------------------------------------------------
 * get_voxel (vdb345.rs:69-106)
 * ------------------------------------------------------------------------------------------ */
int wxo_get_voxel(const WxoTree *t, int32_t x, int32_t y, int32_t z, uint64_t *value) {
  int32_t p[3] = {x, y, z}, key[3];
  uint64_t dummy;
  if (!value) value = &dummy;
  wxo_global_to_node(5, p, key);
  const WxoRootEntry *e = wxo_root_find(t, key);
  if (!e) {
    *value = t->background;
    return WXO_EP_BKGR;
  }
  if (!e->node) {
    *value = e->tile_value;
    return WXO_EP_ROOT;
  }
  const WxoN5 *n5 = e->node;
  uint32_t b4 = wxo_global_to_offset(5, p);
  const WxoN4 *n4 = n5->child[b4];
  if (!n4) {
    *value = n5->tile[b4];
    return WXO_EP_INNR5; /* Innr(v, 5) */
  }
  uint32_t b3 = wxo_global_to_offset(4, p);
  const WxoN3 *n3 = n4->child[b3];
  if (!n3) {
    *value = n4->tile[b3];
    return WXO_EP_INNR4; /* Innr(v, 4) */
  }
  uint32_t b0 = wxo_global_to_offset(3, p);
  *value = n3->data[b0];
  return wxo_n3_is_value(n3, b0) ? WXO_EP_LEAF : WXO_EP_OFFS;
}

/* ---------------------------------------------------------------------------------------------
 * count_nodes (vdb345.rs:266-287) and the voxel count the reader tests assert (read.rs:772-794)
 * ------------------------------------------------------------------------------------------ */
void wxo_count_nodes(const WxoTree *t, uint64_t out[3]) {
  out[0] = out[1] = out[2] = 0;
  for (size_t r = 0; r < t->n_root; r++) {
    const WxoN5 *n5 = t->root[r].node;
    if (!n5) continue;
    out[0]++;
    for (int i = 0; i < WXO_N5_SIZE; i++) {
      const WxoN4 *n4 = n5->child[i];
      if (!n4) continue;
      out[1]++;
      for (int j = 0; j < WXO_N4_SIZE; j++)
        if (n4->child[j]) out[2]++;
    }
  }
}

uint64_t wxo_count_leaf_values(const WxoTree *t) {
  uint64_t c = 0;
  for (size_t r = 0; r < t->n_root; r++) {
    const WxoN5 *n5 = t->root[r].node;
    if (!n5) continue;
    for (int i = 0; i < WXO_N5_SIZE; i++) {
      const WxoN4 *n4 = n5->child[i];
      if (!n4) continue;
      for (int j = 0; j < WXO_N4_SIZE; j++) {
        const WxoN3 *n3 = n4->child[j];
        if (!n3) continue;
        for (int k = 0; k < WXO_N3_SIZE; k++) c += wxo_n3_is_value(n3, (uint32_t)k);
      }
    }
  }
  return c;
}

/* ---------------------------------------------------------------------------------------------
 * compute_sdf (vdb345.rs:290-628)
 * ------------------------------------------------------------------------------------------ */
static uint64_t min_u64(uint64_t a, uint64_t b) { return a < b ? a : b; }
static uint32_t min_u32(uint32_t a, uint32_t b) { return a < b ? a : b; }

static int same_node(int level, const int32_t a[3], const int32_t b[3]) {
  int32_t na[3], nb[3];
  wxo_global_to_node(level, a, na);
  wxo_global_to_node(level, b, nb);
  return na[0] == nb[0] && na[1] == nb[1] && na[2] == nb[2];
}

/* One N5-slot tile (forward :368-396, backward :508-536). */
static void sdf_tile5(const WxoTree *t, WxoN5 *n5, uint32_t n4i, const int32_t global[3], const int32_t (*nbrs)[3]) {
  uint32_t child5[3];
  wxo_offset_to_child(5, n4i, child5);
  uint32_t *tile_value = &n5->tile[n4i];
  for (int k = 0; k < 13; k++) {
    const int32_t *dn = nbrs[k];
    int32_t nchild5[3], nglobal[3];
    for (int a = 0; a < 3; a++) {
      nchild5[a] = (int32_t)child5[a] + dn[a];
      nglobal[a] = global[a] + dn[a] * 128; /* N4::TOTAL_DIM */
    }
    if (same_node(5, nglobal, global)) {
      uint32_t nc[3] = {(uint32_t)nchild5[0], (uint32_t)nchild5[1], (uint32_t)nchild5[2]};
      uint32_t nid = wxo_child_to_offset(5, nc);
      if (n5->child[nid]) {
        *tile_value = 1;
        break;
      }
      *tile_value = min_u32(*tile_value, n5->tile[nid] + 1);
      continue;
    }
    uint64_t v;
    int ep = wxo_get_voxel(t, nglobal[0], nglobal[1], nglobal[2], &v);
    *tile_value = (ep == WXO_EP_INNR5) ? min_u32(*tile_value, (uint32_t)v + 1) : 1;
  }
}

/* One N4-slot tile (forward :405-436, backward :545-576). */
static void sdf_tile4(const WxoTree *t, WxoN4 *n4, uint32_t n3i, const int32_t global[3], const int32_t (*nbrs)[3]) {
  uint32_t child4[3];
  wxo_offset_to_child(4, n3i, child4);
  uint32_t *tile_value = &n4->tile[n3i];
  for (int k = 0; k < 13; k++) {
    const int32_t *dn = nbrs[k];
    int32_t nchild4[3], nglobal[3];
    for (int a = 0; a < 3; a++) {
      nchild4[a] = (int32_t)child4[a] + dn[a];
      nglobal[a] = global[a] + dn[a] * 8; /* N3::TOTAL_DIM */
    }
    if (same_node(4, nglobal, global)) {
      uint32_t nc[3] = {(uint32_t)nchild4[0], (uint32_t)nchild4[1], (uint32_t)nchild4[2]};
      uint32_t nid = wxo_child_to_offset(4, nc);
      if (n4->child[nid]) {
        *tile_value = 1;
        break;
      }
      *tile_value = min_u32(*tile_value, n4->tile[nid] + 1);
      continue;
    }
    uint64_t v;
    int ep = wxo_get_voxel(t, nglobal[0], nglobal[1], nglobal[2], &v);
    *tile_value = (ep == WXO_EP_INNR4) ? min_u32(*tile_value, (uint32_t)v + 1) : 1;
  }
}

/* One inactive leaf voxel (forward :444-477, backward :586-619). */
static void sdf_voxel(const WxoTree *t, WxoN3 *n3, uint32_t vi, const int32_t global[3], const int32_t (*nbrs)[3]) {
  uint32_t child3[3];
  wxo_offset_to_child(3, vi, child3);
  uint64_t *tile_value = &n3->data[vi];
  for (int k = 0; k < 13; k++) {
    const int32_t *dn = nbrs[k];
    int32_t nchild3[3], nglobal[3];
    for (int a = 0; a < 3; a++) {
      nchild3[a] = (int32_t)child3[a] + dn[a];
      nglobal[a] = global[a] + dn[a];
    }
    if (same_node(3, nglobal, global)) {
      uint32_t nc[3] = {(uint32_t)nchild3[0], (uint32_t)nchild3[1], (uint32_t)nchild3[2]};
      uint32_t nid = wxo_child_to_offset(3, nc);
      if (wxo_n3_is_value(n3, nid)) {
        *tile_value = 1;
        break;
      }
      *tile_value = min_u64(*tile_value, n3->data[nid] + 1);
      continue;
    }
    uint64_t v;
    int ep = wxo_get_voxel(t, nglobal[0], nglobal[1], nglobal[2], &v);
    *tile_value = (ep == WXO_EP_OFFS) ? min_u64(*tile_value, v + 1) : 1;
  }
}

static void sdf_pass_n5(const WxoTree *t, WxoRootEntry *e, const int32_t (*nbrs)[3], int backward) {
  WxoN5 *n5 = e->node;
  const int32_t *origin5 = e->key;
  for (int ii = 0; ii < WXO_N5_SIZE; ii++) {
    uint32_t n4i = (uint32_t)(backward ? WXO_N5_SIZE - 1 - ii : ii);
    uint32_t c5[3];
    int32_t g5[3];
    wxo_offset_to_child(5, n4i, c5);
    for (int a = 0; a < 3; a++) g5[a] = origin5[a] + (int32_t)c5[a] * 128;
    WxoN4 *n4 = n5->child[n4i];
    if (!n4) {
      sdf_tile5(t, n5, n4i, g5, nbrs);
      continue;
    }
    for (int jj = 0; jj < WXO_N4_SIZE; jj++) {
      uint32_t n3i = (uint32_t)(backward ? WXO_N4_SIZE - 1 - jj : jj);
      uint32_t c4[3];
      int32_t g4[3];
      wxo_offset_to_child(4, n3i, c4);
      for (int a = 0; a < 3; a++) g4[a] = g5[a] + (int32_t)c4[a] * 8;
      WxoN3 *n3 = n4->child[n3i];
      if (!n3) {
        sdf_tile4(t, n4, n3i, g4, nbrs);
        continue;
      }
      for (int kk = 0; kk < WXO_N3_SIZE; kk++) {
        uint32_t vi = (uint32_t)(backward ? WXO_N3_SIZE - 1 - kk : kk);
        if (wxo_n3_is_value(n3, vi)) continue;
        uint32_t c3[3];
        int32_t g3[3];
        wxo_offset_to_child(3, vi, c3);
        for (int a = 0; a < 3; a++) g3[a] = g4[a] + (int32_t)c3[a];
        sdf_voxel(t, n3, vi, g3, nbrs);
      }
    }
  }
}

void wxo_compute_sdf(WxoTree *t) {
  /* :292-325 initialise every tile with "infinity" (MAX - 1 so that +1 does not wrap) */
  for (size_t r = 0; r < t->n_root; r++) {
    WxoN5 *n5 = t->root[r].node;
    if (!n5) continue;
    for (int i = 0; i < WXO_N5_SIZE; i++) {
      WxoN4 *n4 = n5->child[i];
      if (!n4) {
        n5->tile[i] = UINT32_MAX - 1;
        continue;
      }
      for (int j = 0; j < WXO_N4_SIZE; j++) {
        WxoN3 *n3 = n4->child[j];
        if (!n3) {
          n4->tile[j] = UINT32_MAX - 1;
          continue;
        }
        for (int k = 0; k < WXO_N3_SIZE; k++)
          if (!wxo_n3_is_value(n3, (uint32_t)k)) n3->data[k] = UINT64_MAX - 1; /* usize::MAX - 1 */
      }
    }
  }

  /* :327-343 neighbour sets */
  int32_t f[13][3], b[13][3];
  int n = 0;
  static const int32_t d3[3] = {-1, 0, 1};
  for (int iy = 0; iy < 3; iy++)
    for (int iz = 0; iz < 3; iz++) {
      f[n][0] = -1, f[n][1] = d3[iy], f[n][2] = d3[iz];
      b[n][0] = 1, b[n][1] = d3[iy], b[n][2] = d3[iz];
      n++;
    }
  for (int iz = 0; iz < 3; iz++) {
    f[n][0] = 0, f[n][1] = -1, f[n][2] = d3[iz];
    b[n][0] = 0, b[n][1] = 1, b[n][2] = d3[iz];
    n++;
  }
  f[n][0] = 0, f[n][1] = 0, f[n][2] = -1;
  b[n][0] = 0, b[n][1] = 0, b[n][2] = 1;

  WxoRootEntry **sorted = wxo_root_sorted(t);
  /* forward pass :353-485 */
  for (size_t r = 0; r < t->n_root; r++)
    if (sorted[r]->node) sdf_pass_n5(t, sorted[r], f, 0);
  /* backward pass :488-627 */
  for (size_t r = t->n_root; r-- > 0;)
    if (sorted[r]->node) sdf_pass_n5(t, sorted[r], b, 1);
  free(sorted);
}

/* ---------------------------------------------------------------------------------------------
 * origins() / masks() / atlas()  (vdb345.rs:108-264, :673-694; mask.rs:95-119)
 * ------------------------------------------------------------------------------------------ */
static uint32_t closest_power_of_3(uint64_t n) { /* vdb345.rs:684-690 (a cube root, despite the name) */
  uint64_t i = 0;
  while (i * i * i < n) i++;
  return (uint32_t)i;
}

static void arr32_from_arr64(const uint64_t *in, int words64, uint32_t *out) { /* vdb345.rs:673-682 */
  for (int i = 0; i < words64; i++) {
    out[2 * i] = (uint32_t)in[i];
    out[2 * i + 1] = (uint32_t)(in[i] >> 32);
  }
}

/* atlas[x][y][z] with cubic side `side` (vdb345.rs:170-184) */
static inline size_t atlas_at(uint32_t side, uint32_t x, uint32_t y, uint32_t z) {
  return ((size_t)x * side + y) * side + z;
}

static WxoGpuData *gpudata_alloc(uint32_t n5, uint32_t n4, uint32_t n3) {
  WxoGpuData *g = (WxoGpuData *)calloc(1, sizeof(WxoGpuData));
  g->n[0] = n5, g->n[1] = n4, g->n[2] = n3;
  g->dim[0] = closest_power_of_3(n5);
  g->dim[1] = closest_power_of_3(n4);
  g->dim[2] = closest_power_of_3(n3);
  static const uint32_t nd[3] = {32, 16, 8};
  for (int l = 0; l < 3; l++) {
    g->side[l] = nd[l] * g->dim[l];
    size_t cnt = (size_t)g->side[l] * g->side[l] * g->side[l];
    g->atlas[l] = (uint32_t *)calloc(cnt ? cnt : 1, sizeof(uint32_t)); /* ValueType::zeroed() */
  }
  g->origins = (int32_t *)calloc((size_t)n5 * 4 + 4, sizeof(int32_t));
  g->mask[0] = (uint32_t *)calloc((size_t)n5 * 1024 + 1, 4);
  g->mask[1] = (uint32_t *)calloc((size_t)n5 * 1024 + 1, 4);
  g->mask[2] = (uint32_t *)calloc((size_t)n4 * 128 + 1, 4);
  g->mask[3] = (uint32_t *)calloc((size_t)n4 * 128 + 1, 4);
  g->mask[4] = (uint32_t *)calloc((size_t)n3 * 16 + 1, 4);
  return g;
}

WxoGpuData *wxo_serialise(const WxoTree *t) {
  uint64_t cnt[3];
  wxo_count_nodes(t, cnt);
  WxoGpuData *g = gpudata_alloc((uint32_t)cnt[0], (uint32_t)cnt[1], (uint32_t)cnt[2]);
  WxoRootEntry **sorted = wxo_root_sorted(t);
  size_t n5_idx = 0, n4_idx = 0, n3_idx = 0;
  for (size_t r = 0; r < t->n_root; r++) {
    const WxoN5 *n5 = sorted[r]->node;
    if (!n5) continue; /* root tiles are skipped (:111,:135,:191-194) */
    /* origins(): the map key, padded to [x,y,z,0] by mask.rs:104-109 */
    memcpy(&g->origins[4 * n5_idx], sorted[r]->key, 3 * sizeof(int32_t));
    /* origin_from_idx (:692-694) * DIM */
    uint32_t d5 = g->dim[0], d4 = g->dim[1], d3 = g->dim[2];
    uint32_t o5[3] = {(uint32_t)(n5_idx % d5) * 32, (uint32_t)((n5_idx / d5) % d5) * 32, (uint32_t)(n5_idx / (d5 * d5)) * 32};
    for (uint32_t off5 = 0; off5 < WXO_N5_SIZE; off5++) {
      uint32_t c5[3];
      wxo_offset_to_child(5, off5, c5);
      size_t a5 = atlas_at(g->side[0], o5[0] + c5[0], o5[1] + c5[1], o5[2] + c5[2]);
      const WxoN4 *n4 = n5->child[off5];
      if (!n4) {
        g->atlas[0][a5] = n5->tile[off5];
        continue;
      }
      uint32_t o4[3] = {(uint32_t)(n4_idx % d4) * 16, (uint32_t)((n4_idx / d4) % d4) * 16, (uint32_t)(n4_idx / ((size_t)d4 * d4)) * 16};
      for (uint32_t off4 = 0; off4 < WXO_N4_SIZE; off4++) {
        uint32_t c4[3];
        wxo_offset_to_child(4, off4, c4);
        size_t a4 = atlas_at(g->side[1], o4[0] + c4[0], o4[1] + c4[1], o4[2] + c4[2]);
        const WxoN3 *n3 = n4->child[off4];
        if (!n3) {
          g->atlas[1][a4] = n4->tile[off4];
          continue;
        }
        uint32_t o3[3] = {(uint32_t)(n3_idx % d3) * 8, (uint32_t)((n3_idx / d3) % d3) * 8, (uint32_t)(n3_idx / ((size_t)d3 * d3)) * 8};
        for (uint32_t off3 = 0; off3 < WXO_N3_SIZE; off3++) {
          uint32_t c3[3];
          wxo_offset_to_child(3, off3, c3);
          size_t a3 = atlas_at(g->side[2], o3[0] + c3[0], o3[1] + c3[1], o3[2] + c3[2]);
          /* Value(value) => value ; Tile(offset) => offset as u32  (:241-247) */
          g->atlas[2][a3] = (uint32_t)n3->data[off3];
        }
        /* masks(): n3_vals pushed in the same DFS order (:144-150) */
        arr32_from_arr64(n3->value_mask, 8, &g->mask[4][n3_idx * 16]);
        g->atlas[1][a4] = (uint32_t)n3_idx;
        n3_idx++;
      }
      arr32_from_arr64(n4->value_mask, 64, &g->mask[3][n4_idx * 128]);
      arr32_from_arr64(n4->child_mask, 64, &g->mask[2][n4_idx * 128]);
      g->atlas[0][a5] = (uint32_t)n4_idx;
      n4_idx++;
    }
    arr32_from_arr64(n5->value_mask, 512, &g->mask[1][n5_idx * 1024]);
    arr32_from_arr64(n5->child_mask, 512, &g->mask[0][n5_idx * 1024]);
    n5_idx++;
  }
  free(sorted);
  return g;
}

WxoGpuData *wxo_gpudata_from_tables(uint32_t n5, uint32_t n4, uint32_t n3, const int32_t *origins,
                                    const uint64_t *kids5, const uint64_t *vals5, const uint32_t *tab5,
                                    const uint64_t *kids4, const uint64_t *vals4, const uint32_t *tab4,
                                    const uint64_t *vals3, const uint32_t *tab3) {
  WxoGpuData *g = gpudata_alloc(n5, n4, n3);
  for (uint32_t i = 0; i < n5; i++) memcpy(&g->origins[4 * i], &origins[3 * i], 3 * sizeof(int32_t));
  arr32_from_arr64(kids5, (int)(n5 * 512), g->mask[0]);
  arr32_from_arr64(vals5, (int)(n5 * 512), g->mask[1]);
  arr32_from_arr64(kids4, (int)(n4 * 64), g->mask[2]);
  arr32_from_arr64(vals4, (int)(n4 * 64), g->mask[3]);
  arr32_from_arr64(vals3, (int)(n3 * 8), g->mask[4]);
  const uint32_t *tabs[3] = {tab5, tab4, tab3};
  const uint32_t cnt[3] = {n5, n4, n3};
  static const int lvl[3] = {5, 4, 3};
  static const uint32_t sz[3] = {WXO_N5_SIZE, WXO_N4_SIZE, WXO_N3_SIZE};
  static const uint32_t nd[3] = {32, 16, 8};
  for (int l = 0; l < 3; l++) {
    uint32_t d = g->dim[l];
    for (size_t i = 0; i < cnt[l]; i++) {
      uint32_t o[3] = {(uint32_t)(i % d) * nd[l], (uint32_t)((i / d) % d) * nd[l], (uint32_t)(i / ((size_t)d * d)) * nd[l]};
      for (uint32_t off = 0; off < sz[l]; off++) {
        uint32_t c[3];
        wxo_offset_to_child(lvl[l], off, c);
        g->atlas[l][atlas_at(g->side[l], o[0] + c[0], o[1] + c[1], o[2] + c[2])] = tabs[l][i * sz[l] + off];
      }
    }
  }
  return g;
}

void wxo_gpudata_free(WxoGpuData *g) {
  if (!g) return;
  for (int l = 0; l < 3; l++) free(g->atlas[l]);
  for (int m = 0; m < 5; m++) free(g->mask[m]);
  free(g->origins);
  free(g);
}

void wxo_gpudata_counts(const WxoGpuData *g, uint32_t n[3], uint32_t atlas_dim[3]) {
  for (int l = 0; l < 3; l++) {
    n[l] = g->n[l];
    atlas_dim[l] = g->dim[l];
  }
}
const int32_t *wxo_gpudata_origins(const WxoGpuData *g) { return g->origins; }
const uint32_t *wxo_gpudata_mask(const WxoGpuData *g, int which) { return g->mask[which]; }
const uint32_t *wxo_gpudata_atlas(const WxoGpuData *g, int l) { return g->atlas[l]; }

void wxo_gpudata_tables(const WxoGpuData *g, uint32_t *tab5, uint32_t *tab4, uint32_t *tab3) {
  uint32_t *tabs[3] = {tab5, tab4, tab3};
  static const int lvl[3] = {5, 4, 3};
  static const uint32_t sz[3] = {WXO_N5_SIZE, WXO_N4_SIZE, WXO_N3_SIZE};
  static const uint32_t nd[3] = {32, 16, 8};
  for (int l = 0; l < 3; l++) {
    if (!tabs[l]) continue;
    uint32_t d = g->dim[l];
    for (size_t i = 0; i < g->n[l]; i++) {
      uint32_t o[3] = {(uint32_t)(i % d) * nd[l], (uint32_t)((i / d) % d) * nd[l], (uint32_t)(i / ((size_t)d * d)) * nd[l]};
      for (uint32_t off = 0; off < sz[l]; off++) {
        uint32_t c[3];
        wxo_offset_to_child(lvl[l], off, c);
        tabs[l][i * sz[l] + off] = g->atlas[l][atlas_at(g->side[l], o[0] + c[0], o[1] + c[1], o[2] + c[2])];
      }
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * Tree from reference-layout topology (not in the reference: lets tests hand the oracle a
 * procedurally generated scene).  Node order as in masks(): N5 by sorted origin, children by
 * ascending offset (vdb345.rs:134-158).
 * ------------------------------------------------------------------------------------------ */
WxoTree *wxo_tree_from_topology(uint32_t n5c, uint32_t n4c, uint32_t n3c, const int32_t *origins,
                                const uint64_t *kids5, const uint64_t *vals5, const uint64_t *kids4,
                                const uint64_t *vals4, const uint64_t *vals3) {
  WxoTree *t = wxo_tree_new();
  size_t i4 = 0, i3 = 0;
  for (uint32_t i5 = 0; i5 < n5c; i5++) {
    WxoRootEntry *e = wxo_root_insert(t, &origins[3 * i5]);
    WxoN5 *n5 = e->node = wxo_n5_new(&origins[3 * i5]);
    memcpy(n5->child_mask, &kids5[(size_t)i5 * 512], sizeof(n5->child_mask));
    memcpy(n5->value_mask, &vals5[(size_t)i5 * 512], sizeof(n5->value_mask));
    for (uint32_t o5 = 0; o5 < WXO_N5_SIZE; o5++) {
      if (!((n5->child_mask[o5 >> 6] >> (o5 & 63)) & 1)) continue;
      if (i4 >= n4c) goto fail;
      WxoN4 *n4 = n5->child[o5] = wxo_n4_new();
      memcpy(n4->child_mask, &kids4[i4 * 64], sizeof(n4->child_mask));
      memcpy(n4->value_mask, &vals4[i4 * 64], sizeof(n4->value_mask));
      i4++;
      for (uint32_t o4 = 0; o4 < WXO_N4_SIZE; o4++) {
        if (!((n4->child_mask[o4 >> 6] >> (o4 & 63)) & 1)) continue;
        if (i3 >= n3c) goto fail;
        WxoN3 *n3 = n4->child[o4] = wxo_n3_new();
        memcpy(n3->value_mask, &vals3[i3 * 8], sizeof(n3->value_mask));
        memcpy(n3->is_value, n3->value_mask, sizeof(n3->is_value));
        for (uint32_t k = 0; k < WXO_N3_SIZE; k++)
          if (wxo_n3_is_value(n3, k)) n3->data[k] = 1;
        i3++;
      }
    }
  }
  if (i4 != n4c || i3 != n3c) goto fail;
  return t;
fail:
  wxo_tree_free(t);
  return NULL;
}
