/*
 * wxo_read.c -- ORACLE (test infrastructure, see wxo.h): the reference's .vdb reader.
 *
 * Restates src/vdb/read.rs (VdbReader::new :62-121, read_vdb345_grid :123-141, read_transform
 * :143-164, read_grid_descriptors :166-212, read_metadata :214-268, read_tree_topology :270-349,
 * read_internal_node_header :351-376, read_compressed :378-488, read_compressed_data :490-574,
 * read_tree_data :576-629) for T = u32, the instantiation the application uses
 * (src/render/wgpu_context.rs:103).
 *
 * Deviations, all deliberate:
 *  - HashMap iteration order (read.rs:583) is arbitrary in the reference; the oracle walks root
 *    nodes in file order.  Only leaf *values* depend on it and no pixel does (raycast.comp.wgsl:485-493).
 *  - Blosc-compressed blocks (read.rs:507-537) need c-blosc (blosc-src 0.2.1, a Cargo dependency that
 *    is not in the reference tree): reported as WXO_ERR_BLOSC.  Raw (non-positive length) Blosc
 *    blocks are handled.  zlib blocks go through the system zlib (the reference uses flate2 1.0.27).
 *  - Rust panics/todo!() become error returns.
 */
#include "wxo_internal.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define COMP_ZIP 0x1u
#define COMP_ACTIVE_MASK 0x2u
#define COMP_BLOSC 0x4u

#define VER_BOOST_UUID 218u
#define VER_SELECTIVE_COMPRESSION 220u
#define VER_NODE_MASK_COMPRESSION 222u
#define VER_PER_GRID_COMPRESSION 223u

typedef struct {
  const uint8_t *buf;
  size_t len, pos;
  int err;
} Rd;

static void rd_bytes(Rd *r, void *dst, size_t n) {
  if (r->err) {
    memset(dst, 0, n);
    return;
  }
  if (n > r->len - r->pos) {
    r->err = WXO_ERR_IO;
    memset(dst, 0, n);
    return;
  }
  memcpy(dst, r->buf + r->pos, n);
  r->pos += n;
}
static uint8_t rd_u8(Rd *r) { uint8_t v; rd_bytes(r, &v, 1); return v; }
static uint32_t rd_u32(Rd *r) { uint32_t v; rd_bytes(r, &v, 4); return v; }
static int32_t rd_i32(Rd *r) { int32_t v; rd_bytes(r, &v, 4); return v; }
static uint64_t rd_u64(Rd *r) { uint64_t v; rd_bytes(r, &v, 8); return v; }
static int64_t rd_i64(Rd *r) { int64_t v; rd_bytes(r, &v, 8); return v; }

static char *rd_string(Rd *r, size_t len) { /* read.rs:658-663 */
  if (r->err || len > r->len - r->pos) {
    r->err = r->err ? r->err : WXO_ERR_IO;
    return (char *)calloc(1, 1);
  }
  char *s = (char *)malloc(len + 1);
  rd_bytes(r, s, len);
  s[len] = 0;
  return s;
}
static char *rd_len_string(Rd *r) { return rd_string(r, rd_u32(r)); } /* read.rs:653-656 */

typedef struct {
  int is_half_float;        /* "is_saved_as_half_float" == Bool(true), data_structure.rs:350-352 */
  int64_t file_voxel_count; /* -1 if absent */
} Meta;

/* read.rs:214-268 */
static void read_metadata(Rd *r, Meta *m) {
  m->is_half_float = 0;
  m->file_voxel_count = -1;
  uint32_t n = rd_u32(r);
  for (uint32_t i = 0; i < n && !r->err; i++) {
    char *name = rd_len_string(r);
    char *type = rd_len_string(r);
    uint32_t meta_len = rd_u32(r);
    if (!strcmp(type, "string")) {
      free(rd_string(r, meta_len));
    } else if (!strcmp(type, "bool")) {
      uint8_t v = rd_u8(r);
      if (!strcmp(name, "is_saved_as_half_float")) m->is_half_float = (v == 1);
    } else if (!strcmp(type, "int32")) {
      (void)rd_i32(r);
    } else if (!strcmp(type, "int64")) {
      int64_t v = rd_i64(r);
      if (!strcmp(name, "file_voxel_count")) m->file_voxel_count = v;
    } else if (!strcmp(type, "float")) {
      (void)rd_u32(r);
    } else if (!strcmp(type, "vec3i")) {
      (void)rd_i32(r), (void)rd_i32(r), (void)rd_i32(r);
    } else { /* Unknown { name, data } */
      free(rd_string(r, meta_len));
    }
    free(name);
    free(type);
  }
}

typedef struct {
  char *name;
  uint64_t grid_pos, block_pos, end_pos;
  uint32_t compression;
  Meta meta;
} GridDesc;

typedef struct {
  Rd *r;
  uint32_t version;
  const GridDesc *gd;
} Ctx;

/* read.rs:490-574 ; `elem` = size_of::<T>() of the stored type (2 for f16, 4 for f32).
 * Returns a malloc'd byte buffer holding *out_count elements. */
static uint8_t *read_compressed_data(Ctx *c, size_t count, size_t elem, size_t *out_count) {
  Rd *r = c->r;
  uint32_t comp = c->gd->compression;
  uint8_t *data = NULL;
  *out_count = 0;
  if (comp & COMP_BLOSC) {
    int64_t nbytes = rd_i64(r);
    int64_t ccount = nbytes / (int64_t)elem;
    if (nbytes <= 0) {
      size_t n = (size_t)(-ccount);
      data = (uint8_t *)calloc(n * elem + 1, 1);
      rd_bytes(r, data, n * elem);
      if (n != count && !r->err) r->err = WXO_ERR_IO; /* assert_eq!(-compressed_count, count) */
      *out_count = n;
    } else {
      if ((uint64_t)nbytes > r->len - r->pos) {
        r->err = WXO_ERR_IO;
        return (uint8_t *)calloc(1, 1);
      }
      r->pos += (size_t)nbytes;
      if (count > 0) {
        r->err = WXO_ERR_BLOSC; /* blosc_decompress_ctx: c-blosc not available to the oracle */
      }
      data = (uint8_t *)calloc(1, 1);
    }
  } else if (comp & COMP_ZIP) {
    int64_t nbytes = rd_i64(r);
    int64_t ccount = nbytes / (int64_t)elem;
    if (nbytes <= 0) {
      size_t n = (size_t)(-ccount);
      data = (uint8_t *)calloc(n * elem + 1, 1);
      rd_bytes(r, data, n * elem);
      *out_count = n;
    } else {
      if ((uint64_t)nbytes > r->len - r->pos) {
        r->err = WXO_ERR_IO;
        return (uint8_t *)calloc(1, 1);
      }
      data = (uint8_t *)calloc(count * elem + 1, 1);
      uLongf dlen = (uLongf)(count * elem);
      /* ZlibDecoder + read_exact(count elements): fewer bytes than requested is an error */
      int zr = uncompress(data, &dlen, r->buf + r->pos, (uLong)nbytes);
      if ((zr != Z_OK && zr != Z_BUF_ERROR) || dlen != count * elem) r->err = WXO_ERR_IO;
      r->pos += (size_t)nbytes;
      *out_count = count;
    }
  } else {
    data = (uint8_t *)calloc(count * elem + 1, 1);
    rd_bytes(r, data, count * elem);
    *out_count = count;
  }
  return data;
}

static inline int bit64(const uint64_t *m, size_t i) { return (int)((m[i >> 6] >> (i & 63)) & 1); }

/* read.rs:378-488 for T = u32.  Returns malloc'd u32[*out_len]. */
static uint32_t *read_compressed(Ctx *c, size_t size, const uint64_t *value_mask, size_t value_mask_bits,
                                 size_t *out_len) {
  Rd *r = c->r;
  uint8_t md = 6; /* NoMaskAndAllVals */
  if (c->version >= VER_NODE_MASK_COMPRESSION) {
    md = rd_u8(r);
    if (md > 6 && !r->err) r->err = WXO_ERR_NODE_METADATA;
  }
  uint32_t inactive0 = 0, inactive1 = 0; /* T::zeroed(); read as size_of::<T>() = 4 bytes (Q6) */
  if (md == 4 || md == 2) {
    inactive0 = rd_u32(r);
  } else if (md == 5) {
    inactive0 = rd_u32(r);
    inactive1 = rd_u32(r);
  }
  size_t sel_words = (size + 63) / 64;
  uint64_t *selection = (uint64_t *)calloc(sel_words + 1, 8);
  if (md == 3 || md == 4 || md == 5) rd_bytes(r, selection, sel_words * 8);

  size_t count = size;
  if ((c->gd->compression & COMP_ACTIVE_MASK) && md != 6 && c->version >= VER_NODE_MASK_COMPRESSION) {
    count = 0;
    for (size_t i = 0; i < value_mask_bits; i++) count += (size_t)bit64(value_mask, i);
  }

  size_t elem = c->gd->meta.is_half_float ? 2 : 4;
  size_t got = 0;
  uint8_t *raw = read_compressed_data(c, count, elem, &got);
  uint32_t *data = (uint32_t *)calloc(got + 1, 4);
  for (size_t i = 0; i < got; i++) {
    if (elem == 2) { /* from_f16_bites, read.rs:635-642: [0,0,b0,b1] with [b1,b0] = f.to_le_bytes() */
      uint8_t b1 = raw[2 * i], b0 = raw[2 * i + 1];
      data[i] = ((uint32_t)b0 << 16) | ((uint32_t)b1 << 24);
    } else {
      memcpy(&data[i], raw + 4 * i, 4);
    }
  }
  free(raw);

  if ((c->gd->compression & COMP_ACTIVE_MASK) && got != size) { /* :462-484 */
    uint32_t *expanded = (uint32_t *)calloc(size + 1, 4);
    size_t read_idx = 0;
    for (size_t d = 0; d < size; d++) {
      if (d < value_mask_bits && bit64(value_mask, d)) {
        if (read_idx < got) expanded[d] = data[read_idx];
        else if (!r->err) r->err = WXO_ERR_IO; /* index panic in the reference */
        read_idx++;
      } else if (bit64(selection, d)) {
        expanded[d] = inactive1;
      } else {
        expanded[d] = inactive0;
      }
    }
    free(data);
    data = expanded;
    got = size;
  }
  free(selection);
  *out_len = got;
  return data;
}

/* read.rs:351-376.  Tile values are read and dropped (the reference never stores NodeHeader.data). */
static void read_internal_node_header(Ctx *c, int words, uint64_t *child_mask, uint64_t *value_mask) {
  rd_bytes(c->r, child_mask, (size_t)words * 8);
  rd_bytes(c->r, value_mask, (size_t)words * 8);
  size_t node_size = (size_t)words * 64, size = node_size;
  if (c->version < VER_NODE_MASK_COMPRESSION) {
    size = 0;
    for (size_t i = 0; i < node_size; i++) size += (size_t)!bit64(child_mask, i); /* count_zeros */
  }
  size_t n;
  free(read_compressed(c, size, value_mask, node_size, &n));
}

static int valid_compression(uint32_t v) { return (v & ~7u) == 0; } /* Compression::from_bits */

int wxo_vdb_read(const char *path, const char *grid_name, WxoTree **out, WxoVdbInfo *info) {
  *out = NULL;
  WxoVdbInfo dummy;
  if (!info) info = &dummy;
  memset(info, 0, sizeof(*info));
  info->file_voxel_count = -1;

  FILE *f = fopen(path, "rb");
  if (!f) return WXO_ERR_IO;
  fseek(f, 0, SEEK_END);
  long flen = ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t *buf = (uint8_t *)malloc((size_t)flen + 1);
  if (fread(buf, 1, (size_t)flen, f) != (size_t)flen) {
    fclose(f);
    free(buf);
    return WXO_ERR_IO;
  }
  fclose(f);
  Rd rd = {buf, (size_t)flen, 0, 0}, *r = &rd;
  int rc = WXO_OK;
  GridDesc *grids = NULL;
  uint32_t grid_number = 0;
  WxoTree *t = NULL;

  /* ---- VdbReader::new (:62-121) ---- */
  if (rd_u64(r) != 0x56444220ull) { rc = r->err ? r->err : WXO_ERR_MAGIC; goto done; }
  uint32_t version = rd_u32(r);
  if (version < VER_BOOST_UUID) { rc = WXO_ERR_VERSION; goto done; }
  info->file_version = version;
  info->library_major = rd_u32(r);
  info->library_minor = rd_u32(r);
  int has_grid_offsets = rd_u8(r) != 0;
  uint32_t compression = version < VER_PER_GRID_COMPRESSION ? (COMP_ZIP | COMP_ACTIVE_MASK) : (COMP_BLOSC | COMP_ACTIVE_MASK);
  if (version >= VER_SELECTIVE_COMPRESSION && version < VER_NODE_MASK_COMPRESSION)
    compression = (rd_u8(r) == 1) ? COMP_ZIP : 0;
  free(rd_string(r, 36)); /* uuid */
  Meta file_meta;
  read_metadata(r, &file_meta);
  grid_number = rd_u32(r);
  info->grid_count = grid_number;
  if (r->err) { rc = r->err; goto done; }

  /* ---- read_grid_descriptors (:166-212) ---- */
  if (!has_grid_offsets) { rc = WXO_ERR_UNSUPPORTED; goto done; } /* assert!(header.has_grid_offsets) */
  if (grid_number > 4096) { rc = WXO_ERR_IO; goto done; }
  grids = (GridDesc *)calloc(grid_number + 1, sizeof(GridDesc));
  for (uint32_t g = 0; g < grid_number; g++) {
    GridDesc *gd = &grids[g];
    gd->name = rd_len_string(r);
    free(rd_len_string(r)); /* grid_type */
    free(rd_len_string(r)); /* instance_parent */
    gd->grid_pos = rd_u64(r);
    gd->block_pos = rd_u64(r);
    gd->end_pos = rd_u64(r);
    gd->compression = compression;
    if (version >= VER_NODE_MASK_COMPRESSION) {
      gd->compression = rd_u32(r);
      if (!valid_compression(gd->compression) && !r->err) r->err = WXO_ERR_COMPRESSION;
    }
    read_metadata(r, &gd->meta);
    if (r->err) { rc = r->err; goto done; }
    if (gd->end_pos > r->len) { rc = WXO_ERR_IO; goto done; }
    r->pos = (size_t)gd->end_pos;
  }

  /* ---- read_vdb345_grid (:123-141) ---- */
  const GridDesc *gd = NULL;
  for (uint32_t g = 0; g < grid_number; g++)
    if (!strcmp(grids[g].name, grid_name)) gd = &grids[g]; /* HashMap: last insert of a name wins */
  if (!gd) { rc = WXO_ERR_GRID_NAME; goto done; }
  info->grid_compression = gd->compression;
  info->is_half_float = gd->meta.is_half_float;
  info->file_voxel_count = gd->meta.file_voxel_count;
  info->grid_pos = gd->grid_pos, info->block_pos = gd->block_pos, info->end_pos = gd->end_pos;
  if (gd->grid_pos > r->len) { rc = WXO_ERR_IO; goto done; }
  r->pos = (size_t)gd->grid_pos;
  if (version >= VER_NODE_MASK_COMPRESSION) {
    uint32_t cflags = rd_u32(r);
    if (!valid_compression(cflags)) { rc = WXO_ERR_COMPRESSION; goto done; }
  }
  Meta grid_meta;
  read_metadata(r, &grid_meta);
  { /* read_transform (:143-164) */
    char *tn = rd_len_string(r);
    int nvec = !strcmp(tn, "UniformScaleMap") ? 5 : (!strcmp(tn, "UniformScaleTranslateMap") || !strcmp(tn, "ScaleTranslateMap")) ? 6 : -1;
    free(tn);
    if (r->err) { rc = r->err; goto done; }
    if (nvec < 0) { rc = WXO_ERR_UNSUPPORTED; goto done; }
    for (int i = 0; i < 3 * nvec; i++) (void)rd_u64(r);
  }

  Ctx ctx = {r, version, gd};
  /* ---- read_tree_topology (:270-349) ---- */
  t = wxo_tree_new();
  if (rd_u32(r) != 1) { rc = r->err ? r->err : WXO_ERR_UNSUPPORTED; goto done; } /* buffer_count */
  t->background = rd_u32(r);
  uint32_t number_of_tiles = rd_u32(r);
  uint32_t number_of_node5s = rd_u32(r);
  info->root_tiles = number_of_tiles, info->root_nodes = number_of_node5s;
  for (uint32_t i = 0; i < number_of_tiles && !r->err; i++) {
    int32_t origin[3] = {rd_i32(r), rd_i32(r), rd_i32(r)}, key[3];
    wxo_global_to_node(5, origin, key);
    uint32_t value = rd_u32(r);
    int active = rd_u8(r) == 1;
    WxoRootEntry *e = wxo_root_insert(t, key);
    e->tile_value = value, e->tile_active = active;
  }
  for (uint32_t i = 0; i < number_of_node5s && !r->err; i++) {
    int32_t origin[3] = {rd_i32(r), rd_i32(r), rd_i32(r)}, key[3];
    wxo_global_to_node(5, origin, key);
    WxoN5 *n5 = wxo_n5_new(origin);
    read_internal_node_header(&ctx, 512, n5->child_mask, n5->value_mask);
    for (uint32_t o5 = 0; o5 < WXO_N5_SIZE && !r->err; o5++) {
      if (!bit64(n5->child_mask, o5)) continue;
      WxoN4 *n4 = wxo_n4_new();
      read_internal_node_header(&ctx, 64, n4->child_mask, n4->value_mask);
      for (uint32_t o4 = 0; o4 < WXO_N4_SIZE && !r->err; o4++) {
        if (!bit64(n4->child_mask, o4)) continue;
        WxoN3 *n3 = wxo_n3_new();
        rd_bytes(r, n3->value_mask, 64);
        n4->child[o4] = n3;
      }
      n5->child[o5] = n4;
    }
    WxoRootEntry *e = wxo_root_insert(t, key);
    e->node = n5;
  }
  if (r->err) { rc = r->err; goto done; }
  info->topology_end_pos = r->pos;

  /* ---- read_tree_data (:576-629) ---- */
  if (gd->block_pos > r->len) { rc = WXO_ERR_IO; goto done; }
  r->pos = (size_t)gd->block_pos;
  for (size_t ri = 0; ri < t->n_root && !r->err; ri++) {
    WxoN5 *n5 = t->root[ri].node;
    if (!n5) continue;
    for (uint32_t o5 = 0; o5 < WXO_N5_SIZE && !r->err; o5++) {
      WxoN4 *n4 = n5->child[o5];
      if (!n4) continue;
      for (uint32_t o4 = 0; o4 < WXO_N4_SIZE && !r->err; o4++) {
        WxoN3 *n3 = n4->child[o4];
        if (!n3) continue;
        uint64_t value_mask[8];
        rd_bytes(r, value_mask, 64);
        if (version < VER_NODE_MASK_COMPRESSION) {
          (void)rd_i32(r), (void)rd_i32(r), (void)rd_i32(r);
          if (rd_u8(r) != 1 && !r->err) r->err = WXO_ERR_UNSUPPORTED; /* assert_eq!(num_buffers, 1) */
        }
        size_t n;
        uint32_t *data = read_compressed(&ctx, WXO_N3_SIZE, value_mask, WXO_N3_SIZE, &n);
        for (size_t idx = 0; idx < n && idx < WXO_N3_SIZE; idx++) {
          if (bit64(n3->value_mask, idx)) { /* the TOPOLOGY mask decides (:614-623) */
            n3->is_value[idx >> 6] |= 1ull << (idx & 63);
            n3->data[idx] = data[idx];
          }
        }
        free(data);
      }
    }
  }
  if (r->err) { rc = r->err; goto done; }

done:
  if (grids) {
    for (uint32_t g = 0; g < grid_number; g++) free(grids[g].name);
    free(grids);
  }
  free(buf);
  if (rc != WXO_OK) {
    wxo_tree_free(t);
    return rc;
  }
  *out = t;
  return WXO_OK;
}
