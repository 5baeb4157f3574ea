/*
 * wxo_render.c -- ORACLE (test infrastructure, see wxo.h): the reference's per-pixel raycast and
 * the uniform that feeds it.
 *
 * Restates src/shaders/raycast.comp.wgsl line by line over the reference's own GPU data model
 * (3 cubic R32Uint atlases + 5 mask buffers + origins, src/vdb/vdb345.rs:108-264) and
 * src/render/gpu_types/compute_state.rs:87-131.  Arithmetic is IEEE binary32, one rounding per
 * WGSL operation, evaluated left to right, no contraction (build with -ffp-contract=off; the
 * Makefile does).  PARITY UNPINNED against a real wgpu run -- see wxo.h.
 *
 * Choices where WGSL leaves the result to the implementation (documented, mirrored by the CUDA path):
 *   normalize(v)        = v / sqrt(dot(v,v))                 (three divides)
 *   vec3<i32>(f)        = truncation of floor(f), saturating, NaN -> 0
 *   min/max with NaN    = fminf/fmaxf (the non-NaN operand)
 *   rgba8unorm store    = round-half-even(clamp(c,0,1) * 255), NaN -> 0
 *   x % y (f32)         = x - y * trunc(x / y)
 */
#include "wxo_internal.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;
typedef struct { int32_t x, y, z; } i3;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3s(float s) { return V3(s, s, s); }
static inline v3 add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 muls(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 smul(float s, v3 a) { return V3(s * a.x, s * a.y, s * a.z); }
static inline v3 adds(v3 a, float s) { return V3(a.x + s, a.y + s, a.z + s); }
static inline v3 divs(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline v3 neg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 normalize(v3 a) { return divs(a, sqrtf(dot(a, a))); }
static inline v3 vmax(v3 a, v3 b) { return V3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
static inline v3 vmin(v3 a, v3 b) { return V3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
static inline v3 vfloor(v3 a) { return V3(floorf(a.x), floorf(a.y), floorf(a.z)); }
/* mix(e1,e2,e3) = e1*(1-e3) + e2*e3 (WGSL spec) */
static inline v3 mix(v3 a, v3 b, float t) { return add(muls(a, 1.0f - t), muls(b, t)); }

static inline int32_t f2i(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.0f) return INT32_MAX;
  if (f <= -2147483648.0f) return INT32_MIN;
  return (int32_t)f;
}
static inline float fmod_wgsl(float x, float y) { return x - y * truncf(x / y); }

/* ---------------------------------------------------------------------------------------------
 * Lookup cache (raycast.comp.wgsl:344-354)
 * ------------------------------------------------------------------------------------------ */
typedef struct { i3 origin; uint32_t idx; } Parent;
typedef struct {
  v3 color;
  uint32_t dist;
  uint32_t num_parents;
  Parent parents[3];
} VdbLeaf;

typedef struct {
  uint64_t lookups[4];
  uint64_t bytes;
  uint64_t rays;
  uint32_t max_iters;
} RayStats;

typedef struct {
  const WxoGpuData *g;
  const WxoState *s;
  RayStats *st;         /* counters of the ray being traced */
  RayStats *prim, *sec; /* primary-ray counters / secondary-ray counters of this pixel */
} Rt;

/* :497-519 */
static inline i3 global_to_node(i3 p, uint32_t t) {
  i3 r = {(int32_t)((uint32_t)(p.x >> t) << t), (int32_t)((uint32_t)(p.y >> t) << t), (int32_t)((uint32_t)(p.z >> t) << t)};
  return r;
}
typedef struct { uint32_t x, y, z; } u3;
static inline u3 global_to_local(i3 p, uint32_t t) {
  int32_t m = (1 << t) - 1;
  u3 r = {(uint32_t)(p.x & m), (uint32_t)(p.y & m), (uint32_t)(p.z & m)};
  return r;
}
static inline u3 local_to_child_node(u3 p, uint32_t t) { u3 r = {p.x >> t, p.y >> t, p.z >> t}; return r; }
static inline uint32_t child_to_offset(u3 p, uint32_t log_d, uint32_t log_dd) { return (p.x << log_dd) | (p.y << log_d) | p.z; }
static inline u3 atlas_origin_from_idx(uint32_t idx, uint32_t dim) { u3 r = {idx % dim, (idx / dim) % dim, idx / (dim * dim)}; return r; }
static inline int i3_eq(i3 a, i3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

/* textureLoad(atlas, coord, 0).r : texel (x,y,z) = atlas[x][y][z] (wgpu_context.rs:119-150) */
static inline uint32_t texel(const WxoGpuData *g, int l, u3 c) {
  uint32_t side = g->side[l];
  return g->atlas[l][((size_t)c.x * side + c.y) * side + c.z];
}

static inline void count_lookup(Rt *rt, uint32_t level) {
  /* level1 12 B (2 mask words + slot), level2 12 B, level3 8 B (mask word + slot).
   * A level-0 result (origin scan that found no N5) is counted in lookups[0] but contributes 0 B:
   * that is the convention that reproduces SURVEY appendix D's bytes/ray (166.1 / 177.9 / 233.1)
   * and it keeps the roofline numerator conservative (the <=8 origins live in registers/constant
   * memory in any implementation). */
  static const uint32_t b[4] = {0, 12, 12, 8};
  rt->st->lookups[level]++;
  rt->st->bytes += b[level];
}

/* :477-494 */
static VdbLeaf get_vdb_leaf_from_node3(Rt *rt, i3 pos, VdbLeaf leaf) {
  const WxoGpuData *g = rt->g;
  u3 node3_local = global_to_local(pos, 3);
  uint32_t node3_offset = child_to_offset(node3_local, 3, 6);
  uint32_t node3_idx = leaf.parents[2].idx;
  uint32_t mi = node3_offset >> 5, mp = node3_offset & 31;
  int in_val3 = (g->mask[4][(size_t)node3_idx * 16 + mi] & (1u << mp)) != 0;
  uint32_t dim = g->side[2] >> 3; /* textureDimensions(node3s).x >> 3 */
  u3 ao = atlas_origin_from_idx(node3_idx, dim);
  u3 c = {node3_local.x + 8 * ao.x, node3_local.y + 8 * ao.y, node3_local.z + 8 * ao.z};
  uint32_t voxel = texel(g, 2, c);
  count_lookup(rt, 3);
  VdbLeaf r = leaf;
  r.num_parents = 3;
  if (in_val3) {
    r.color = v3s(0.1f);
    r.dist = 0;
    return r;
  }
  r.color = v3s(0.0f);
  r.dist = voxel;
  return r;
}

/* :446-475 */
static VdbLeaf get_vdb_leaf_from_node4(Rt *rt, i3 pos, VdbLeaf leaf) {
  const WxoGpuData *g = rt->g;
  u3 node4_local = global_to_local(pos, 7);
  u3 node4_child = local_to_child_node(node4_local, 3);
  uint32_t node4_offset = child_to_offset(node4_child, 4, 8);
  uint32_t node4_idx = leaf.parents[1].idx;
  uint32_t mi = node4_offset >> 5, mp = node4_offset & 31;
  int in_kid4 = (g->mask[2][(size_t)node4_idx * 128 + mi] & (1u << mp)) != 0;
  int in_val4 = (g->mask[3][(size_t)node4_idx * 128 + mi] & (1u << mp)) != 0;
  uint32_t dim = g->side[1] >> 4;
  u3 ao = atlas_origin_from_idx(node4_idx, dim);
  u3 c = {node4_child.x + 16 * ao.x, node4_child.y + 16 * ao.y, node4_child.z + 16 * ao.z};
  uint32_t node3_idx = texel(g, 1, c);
  if (in_val4) {
    count_lookup(rt, 2);
    VdbLeaf r = leaf;
    r.color = v3s(0.2f), r.dist = 0, r.num_parents = 2;
    return r;
  }
  if (!in_kid4) {
    count_lookup(rt, 2);
    VdbLeaf r = leaf;
    r.color = v3s(0.0f), r.dist = node3_idx, r.num_parents = 2;
    return r;
  }
  leaf.parents[2].origin = global_to_node(pos, 3);
  leaf.parents[2].idx = node3_idx;
  leaf.num_parents = 3;
  return get_vdb_leaf_from_node3(rt, pos, leaf);
}

/* :415-444 */
static VdbLeaf get_vdb_leaf_from_node5(Rt *rt, i3 pos, VdbLeaf leaf) {
  const WxoGpuData *g = rt->g;
  u3 node5_local = global_to_local(pos, 12);
  u3 node5_child = local_to_child_node(node5_local, 7);
  uint32_t node5_offset = child_to_offset(node5_child, 5, 10);
  uint32_t node5_idx = leaf.parents[0].idx;
  uint32_t mi = node5_offset >> 5, mp = node5_offset & 31;
  int in_kid5 = (g->mask[0][(size_t)node5_idx * 1024 + mi] & (1u << mp)) != 0;
  int in_val5 = (g->mask[1][(size_t)node5_idx * 1024 + mi] & (1u << mp)) != 0;
  uint32_t dim = g->side[0] >> 5; /* textureDimensions(node5s).y >> 5 */
  u3 ao = atlas_origin_from_idx(node5_idx, dim);
  u3 c = {node5_child.x + 32 * ao.x, node5_child.y + 32 * ao.y, node5_child.z + 32 * ao.z};
  uint32_t node4_idx = texel(g, 0, c);
  if (in_val5) {
    count_lookup(rt, 1);
    VdbLeaf r = leaf;
    r.color = v3s(0.2f), r.dist = 0, r.num_parents = 1;
    return r;
  }
  if (!in_kid5) {
    count_lookup(rt, 1);
    VdbLeaf r = leaf;
    r.color = v3s(0.0f), r.dist = node4_idx, r.num_parents = 1;
    return r;
  }
  leaf.parents[1].origin = global_to_node(pos, 7);
  leaf.parents[1].idx = node4_idx;
  leaf.num_parents = 2;
  return get_vdb_leaf_from_node4(rt, pos, leaf);
}

/* :398-413 */
static VdbLeaf get_vdb_leaf_from_nothing(Rt *rt, i3 pos, VdbLeaf leaf) {
  const WxoGpuData *g = rt->g;
  i3 node5_global = global_to_node(pos, 12);
  for (uint32_t node5_idx = 0; node5_idx < g->n[0]; node5_idx++) { /* arrayLength(&origins) */
    i3 o = {g->origins[4 * node5_idx], g->origins[4 * node5_idx + 1], g->origins[4 * node5_idx + 2]};
    if (i3_eq(node5_global, o)) {
      leaf.parents[0].origin = node5_global;
      leaf.parents[0].idx = node5_idx;
      leaf.num_parents = 1;
      return get_vdb_leaf_from_node5(rt, pos, leaf);
    }
  }
  count_lookup(rt, 0);
  VdbLeaf r = leaf;
  r.color = v3s(0.0f), r.dist = 1, r.num_parents = 0;
  return r;
}

/* :360-396 */
static VdbLeaf get_vdb_leaf_from_leaf(Rt *rt, i3 pos, VdbLeaf leaf) {
  if (leaf.num_parents == 3) {
    if (i3_eq(leaf.parents[2].origin, global_to_node(pos, 3))) return get_vdb_leaf_from_node3(rt, pos, leaf);
    if (i3_eq(leaf.parents[1].origin, global_to_node(pos, 7))) return get_vdb_leaf_from_node4(rt, pos, leaf);
    if (i3_eq(leaf.parents[0].origin, global_to_node(pos, 12))) return get_vdb_leaf_from_node5(rt, pos, leaf);
    return get_vdb_leaf_from_nothing(rt, pos, leaf);
  }
  if (leaf.num_parents == 2) {
    if (i3_eq(leaf.parents[1].origin, global_to_node(pos, 7))) return get_vdb_leaf_from_node4(rt, pos, leaf);
    if (i3_eq(leaf.parents[0].origin, global_to_node(pos, 12))) return get_vdb_leaf_from_node5(rt, pos, leaf);
    return get_vdb_leaf_from_nothing(rt, pos, leaf);
  }
  if (leaf.num_parents == 1) {
    if (i3_eq(leaf.parents[0].origin, global_to_node(pos, 12))) return get_vdb_leaf_from_node5(rt, pos, leaf);
    return get_vdb_leaf_from_nothing(rt, pos, leaf);
  }
  return get_vdb_leaf_from_nothing(rt, pos, leaf);
}

/* ---------------------------------------------------------------------------------------------
 * hdda_ray (:70-142)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  uint32_t state; /* 0 hit, 1 out of bounds, 2 max steps */
  VdbLeaf leaf;
  v3 p;
  int mx, my, mz; /* vec3<bool> mask */
  uint32_t i;
} HDDAout;

static inline v3 sign11(v3 p) { return V3(p.x < 0.f ? -1.f : 1.f, p.y < 0.f ? -1.f : 1.f, p.z < 0.f ? -1.f : 1.f); }
static inline v3 modulo_vec3f(v3 x, float y) { return sub(x, smul(y, vfloor(divs(x, y)))); }

#define HDDA_MAX_RAY_STEPS 1000u

static HDDAout hdda_ray(Rt *rt, v3 src, v3 dir) {
  static const float scale[4] = {1.f, 8.f, 128.f, 4096.f};
  v3 p = src;
  v3 step = sign11(dir);
  v3 step01 = vmax(v3s(0.f), step);
  v3 idir = V3(1.f / dir.x, 1.f / dir.y, 1.f / dir.z);
  int mx = 0, my = 0, mz = 0;
  VdbLeaf leaf;
  memset(&leaf, 0, sizeof(leaf));
  rt->st->rays++;
  for (uint32_t i = 0; i < HDDA_MAX_RAY_STEPS; i++) {
    v3 fp = vfloor(p);
    i3 ip = {f2i(fp.x), f2i(fp.y), f2i(fp.z)};
    leaf = get_vdb_leaf_from_leaf(rt, ip, leaf);
    if (leaf.dist == 0u) {
      HDDAout o = {0u, leaf, p, mx, my, mz, i};
      if (i > rt->st->max_iters) rt->st->max_iters = i;
      return o;
    }
    if (4096.f < fabsf(p.x) || 4096.f < fabsf(p.y) || 4096.f < fabsf(p.z)) {
      HDDAout o = {1u, leaf, p, mx, my, mz, i};
      if (i > rt->st->max_iters) rt->st->max_iters = i;
      return o;
    }
    float size = (float)leaf.dist;
    switch (leaf.num_parents) {
      case 3u: size *= scale[0]; break;
      case 2u: size *= scale[1]; break;
      case 1u: size *= scale[2]; break;
      case 0u: size *= scale[3]; break;
      default: size = scale[0]; break;
    }
    v3 tMax = mul(idir, sub(smul(size, step01), modulo_vec3f(p, size)));
    p = add(p, smul(fminf(fminf(tMax.x, tMax.y), tMax.z), dir));
    /* b1 = tMax.xyz <= tMax.yzx ; b2 = tMax.xyz <= tMax.zxy ; mask = b1 & b2 */
    mx = (tMax.x <= tMax.y) && (tMax.x <= tMax.z);
    my = (tMax.y <= tMax.z) && (tMax.y <= tMax.x);
    mz = (tMax.z <= tMax.x) && (tMax.z <= tMax.y);
    p = add(p, mul(smul(4e-4f, step), V3((float)mx, (float)my, (float)mz)));
  }
  HDDAout o = {2u, leaf, p, mx, my, mz, HDDA_MAX_RAY_STEPS};
  rt->st->max_iters = HDDA_MAX_RAY_STEPS;
  return o;
}

/* ---------------------------------------------------------------------------------------------
 * shading (:144-342)
 * ------------------------------------------------------------------------------------------ */
static const float k_d = 0.7f, k_a = 0.3f, REFLECTIVITY = 0.9f, WALL_I = 0.1f;
#define BASE_COLOR V3(0.4f, 0.2f, 0.2f)
#define AMBIENT_COLOR V3(0.4f, 0.4f, 0.3f)

static inline v3 maskf(const HDDAout *h) { return V3((float)h->mx, (float)h->my, (float)h->mz); }
static inline v3 sun_rgb(const WxoState *s) { return V3(s->sun_color[0], s->sun_color[1], s->sun_color[2]); }
static inline v3 sun_dir(const WxoState *s) { return V3(s->sun_dir[0], s->sun_dir[1], s->sun_dir[2]); }

/* the colour table shared by reflect_ray2/1 for out-of-bounds rays (:296-306, :328-338) */
static v3 wall_flat(v3 N) {
  v3 Np = vmax(v3s(0.f), N);
  v3 Nn = neg(vmin(v3s(0.f), N));
  v3 r = muls(V3(WALL_I, 0.f, 0.f), Np.x);
  r = add(r, muls(V3(0.f, WALL_I, 0.f), Np.y));
  r = add(r, muls(V3(0.f, 0.f, WALL_I), Np.z));
  r = add(r, muls(V3(WALL_I, WALL_I, 0.f), Nn.x));
  r = add(r, muls(V3(0.f, WALL_I, WALL_I), Nn.y));
  r = add(r, muls(V3(WALL_I, 0.f, WALL_I), Nn.z));
  return r;
}

/* mcol of the glossy modes: BASE + I*sun (x0.05 when the sun is occluded) (:199-208, :280-290, :317-325) */
static v3 sun_lit(Rt *rt, const HDDAout *hit, v3 step, v3 N) {
  const WxoState *s = rt->s;
  float I = s->sun_color[3] * k_d * dot(neg(sun_dir(s)), N);
  I = fmaxf(0.0f, I);
  if (I != 0.0f && hdda_ray(rt, sub(hit->p, mul(smul(4e-2f, step), maskf(hit))), neg(sun_dir(s))).state == 0u)
    return add(BASE_COLOR, muls(smul(I, sun_rgb(s)), 0.05f));
  return add(BASE_COLOR, smul(I, sun_rgb(s)));
}

/* :311-342 */
static v3 reflect_ray1(Rt *rt, v3 src, v3 dir) {
  HDDAout hit = hdda_ray(rt, src, dir);
  v3 step = sign11(dir);
  if (hit.state == 0u) {
    v3 N = normalize(mul(neg(step), maskf(&hit)));
    return sun_lit(rt, &hit, step, N);
  }
  if (hit.state == 1u) {
    v3 N = normalize(mul(neg(step), maskf(&hit)));
    return wall_flat(N);
  }
  return dir;
}

/* :267-309 */
static v3 reflect_ray2(Rt *rt, v3 src, v3 dir) {
  HDDAout hit = hdda_ray(rt, src, dir);
  v3 step = sign11(dir);
  if (hit.state == 0u) {
    v3 N = normalize(mul(neg(step), maskf(&hit)));
    v3 rdir = normalize(sub(dir, muls(smul(2.0f, N), dot(dir, N))));
    v3 rsrc = sub(hit.p, mul(smul(4e-2f, step), maskf(&hit)));
    v3 rcol = reflect_ray1(rt, rsrc, rdir);
    v3 mcol = sun_lit(rt, &hit, step, N);
    return mix(mcol, rcol, REFLECTIVITY);
  }
  if (hit.state == 1u) {
    v3 N = normalize(mul(neg(step), maskf(&hit)));
    return wall_flat(N);
  }
  return dir;
}

static inline int any_mod0(v3 fp, float m) {
  return fmod_wgsl(fp.x, m) == 0.f || fmod_wgsl(fp.y, m) == 0.f || fmod_wgsl(fp.z, m) == 0.f;
}

/* :152-265.  `primary` receives the primary ray's HDDAout (the oracle's AOVs). */
static v3 ray_trace(Rt *rt, v3 src, v3 dir, HDDAout *primary) {
  const WxoState *s = rt->s;
  rt->st = rt->prim;
  HDDAout hit = hdda_ray(rt, src, dir);
  rt->st = rt->sec;
  *primary = hit;
  v3 step = sign11(dir);
  uint32_t mode = s->render_mode[0];

  if (hit.state == 0u) {
    v3 grid = v3s(0.0f);
    v3 fp = vfloor(hit.p);
    if (s->show_345[2] == 1u && any_mod0(fp, 4096.f)) grid = V3(-0.3f, -0.3f, 1.0f);
    else if (s->show_345[1] == 1u && any_mod0(fp, 128.f)) grid = V3(0.6f, -0.2f, -0.2f);
    else if (s->show_345[0] == 1u && any_mod0(fp, 8.f)) grid = V3(-0.1f, 0.5f, 0.3f);

    switch (mode) {
      case 1u: /* Rgb */
        return add(add(grid, v3s(0.1f)), mul(maskf(&hit), V3(0.4f, 0.4f, 0.4f)));
      case 2u: { /* Ray length */
        float t = (float)hit.i / (float)200u;
        return add(grid, mix(V3(0.72f, 1.0f, 0.99f), V3(1.0f, 0.0f, 0.0f), t));
      }
      case 3u: { /* Diffuse */
        v3 N = normalize(mul(neg(step), maskf(&hit)));
        float LN = fmaxf(0.0f, s->sun_color[3] * dot(neg(sun_dir(s)), N));
        v3 I_d = muls(mul(smul(k_d, sun_rgb(s)), BASE_COLOR), LN);
        v3 I_a = mul(smul(k_a, AMBIENT_COLOR), BASE_COLOR);
        if (LN != 0.0f && hdda_ray(rt, sub(hit.p, mul(smul(4e-2f, step), maskf(&hit))), neg(sun_dir(s))).state == 0u)
          return I_a;
        return add(I_a, I_d);
      }
      case 4u: { /* Glossy */
        v3 N = normalize(mul(neg(step), maskf(&hit)));
        v3 mcol = sun_lit(rt, &hit, step, N);
        v3 rdir = normalize(sub(dir, muls(smul(2.0f, N), dot(dir, N))));
        v3 rsrc = sub(hit.p, mul(smul(4e-2f, step), maskf(&hit)));
        v3 rcol = reflect_ray2(rt, rsrc, rdir);
        return mix(mcol, rcol, REFLECTIVITY);
      }
      case 0u:
      default: /* Gray */
        return adds(grid, dot(mul(maskf(&hit), V3(0.2f, 0.2f, 0.3f)), v3s(1.0f)));
    }
  }

  if (hit.state == 1u) {
    switch (mode) {
      case 2u: {
        float t = (float)hit.i / (float)200u;
        return adds(mix(V3(0.72f, 1.0f, 0.99f), V3(1.0f, 0.0f, 0.0f), t),
                    dot(mul(maskf(&hit), V3(0.04f, 0.08f, 0.12f)), v3s(1.0f)));
      }
      case 4u: {
        v3 N = normalize(mul(neg(step), maskf(&hit)));
        v3 Np = vmax(v3s(0.f), N);
        v3 Nn = neg(vmin(v3s(0.f), N));
        float t = hit.p.y / 4096.f;
        v3 r = muls(mix(V3(WALL_I, 0.f, 0.f), V3(WALL_I * 0.1f, 0.f, 0.f), t), Np.x);
        r = add(r, muls(V3(0.f, WALL_I, 0.f), Np.y));
        r = add(r, muls(mix(V3(0.f, 0.f, WALL_I), V3(0.f, 0.f, WALL_I * 0.1f), t), Np.z));
        r = add(r, muls(mix(V3(WALL_I, WALL_I, 0.f), V3(WALL_I * 0.1f, WALL_I * 0.1f, 0.f), t), Nn.x));
        r = add(r, muls(V3(0.f, WALL_I, WALL_I), Nn.y));
        r = add(r, muls(mix(V3(WALL_I, 0.f, WALL_I), V3(WALL_I * 0.1f, 0.f, WALL_I * 0.1f), t), Nn.z));
        return r;
      }
      case 0u:
      case 1u:
      default:
        return adds(v3s(0.0f), dot(mul(maskf(&hit), V3(0.01f, 0.02f, 0.03f)), v3s(1.0f)));
    }
  }
  return dir; /* max steps exceeded (:264) */
}

static inline uint8_t unorm8(float c) {
  if (c != c) return 0;
  c = c < 0.f ? 0.f : (c > 1.f ? 1.f : c);
  return (uint8_t)lrintf(c * 255.0f); /* default rounding mode: half to even */
}

/* ---------------------------------------------------------------------------------------------
 * cp_main over a band of rows (:60-68 ; dispatch wgpu_context.rs:281)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const WxoGpuData *g;
  const WxoState *s;
  uint32_t width, height, y0, y1;
  uint8_t *rgba;
  const WxoAov *aov;
  volatile uint32_t *next_row;
  WxoStats stats;
} Job;

static void render_rows(Job *j) {
  const WxoState *s = j->s;
  const uint32_t disp_w = (j->width / 8) * 8, disp_h = (j->height / 4) * 4;
  v3 eye = V3(s->eye[0], s->eye[1], s->eye[2]);
  v3 u = V3(s->u[0], s->u[1], s->u[2]), mv = V3(s->mv[0], s->mv[1], s->mv[2]), wp = V3(s->wp[0], s->wp[1], s->wp[2]);
  WxoStats *S = &j->stats;
  for (;;) {
    uint32_t y = __atomic_fetch_add(j->next_row, 1u, __ATOMIC_RELAXED);
    if (y >= j->y1) break;
    for (uint32_t x = 0; x < j->width; x++) {
      size_t pix = (size_t)y * j->width + x;
      if (x >= disp_w || y >= disp_h) { /* never dispatched: texture keeps its zero initialisation */
        if (j->rgba) memset(j->rgba + 4 * pix, 0, 4);
        continue;
      }
      RayStats prim = {{0, 0, 0, 0}, 0, 0, 0}, sec = {{0, 0, 0, 0}, 0, 0, 0};
      Rt rt = {j->g, s, &prim, &prim, &sec};
      float px = (float)x + 0.001f, py = (float)y + 0.001f;
      v3 ray_dir = normalize(add(add(smul(px, u), smul(py, mv)), wp));
      HDDAout h;
      v3 color = ray_trace(&rt, eye, ray_dir, &h);
      if (j->rgba) {
        uint8_t *o = j->rgba + 4 * pix;
        o[0] = unorm8(color.x), o[1] = unorm8(color.y), o[2] = unorm8(color.z), o[3] = 255;
      }
      const WxoAov *a = j->aov;
      if (a) {
        v3 fp = vfloor(h.p);
        if (a->state) a->state[pix] = (uint8_t)h.state;
        if (a->voxel) a->voxel[3 * pix] = f2i(fp.x), a->voxel[3 * pix + 1] = f2i(fp.y), a->voxel[3 * pix + 2] = f2i(fp.z);
        if (a->leaf) a->leaf[pix] = h.leaf.num_parents == 3 ? (int32_t)h.leaf.parents[2].idx : -1;
        if (a->level) a->level[pix] = (uint8_t)h.leaf.num_parents;
        if (a->iters) a->iters[pix] = h.i;
        if (a->depth) a->depth[pix] = sqrtf(dot(sub(h.p, eye), sub(h.p, eye)));
        if (a->mask) a->mask[pix] = (uint8_t)(h.mx | (h.my << 1) | (h.mz << 2));
        if (a->pos) a->pos[3 * pix] = h.p.x, a->pos[3 * pix + 1] = h.p.y, a->pos[3 * pix + 2] = h.p.z;
      }
      S->rays += prim.rays + sec.rays;
      S->primary_rays += 1;
      for (int l = 0; l < 4; l++) S->lookups[l] += prim.lookups[l] + sec.lookups[l], S->primary_lookups[l] += prim.lookups[l];
      S->alg_bytes += prim.bytes + sec.bytes + 4;
      S->primary_alg_bytes += prim.bytes + 4;
      if (prim.max_iters > S->max_iters) S->max_iters = prim.max_iters; /* primary rays only */
      if (h.state == 0) S->hit++;
      else if (h.state == 1) S->oob++;
      else S->maxed++;
    }
  }
}

static void *render_thread(void *arg) {
  render_rows((Job *)arg);
  return NULL;
}

void wxo_render(const WxoGpuData *g, const WxoState *s, uint32_t width, uint32_t height, uint32_t y0, uint32_t y1,
                uint8_t *rgba, const WxoAov *aov, int threads, WxoStats *stats) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if (y1 > height) y1 = height;
  volatile uint32_t next_row = y0;
  Job *jobs = (Job *)calloc((size_t)threads, sizeof(Job));
  pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
  for (int t = 0; t < threads; t++) {
    Job j = {g, s, width, height, y0, y1, rgba, aov, &next_row, {0}};
    jobs[t] = j;
  }
  for (int t = 1; t < threads; t++) pthread_create(&tid[t], NULL, render_thread, &jobs[t]);
  render_rows(&jobs[0]);
  for (int t = 1; t < threads; t++) pthread_join(tid[t], NULL);
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    for (int t = 0; t < threads; t++) {
      const WxoStats *a = &jobs[t].stats;
      stats->rays += a->rays, stats->primary_rays += a->primary_rays;
      for (int l = 0; l < 4; l++) stats->lookups[l] += a->lookups[l], stats->primary_lookups[l] += a->primary_lookups[l];
      stats->alg_bytes += a->alg_bytes, stats->primary_alg_bytes += a->primary_alg_bytes;
      if (a->max_iters > stats->max_iters) stats->max_iters = a->max_iters;
      stats->hit += a->hit, stats->oob += a->oob, stats->maxed += a->maxed;
    }
  }
  free(jobs);
  free(tid);
}

/* ---------------------------------------------------------------------------------------------
 * ComputeState::build (compute_state.rs:87-131) over cgmath 0.18.0 f32 maths.
 *
 * cgmath is a Cargo dependency (Cargo.lock: cgmath 0.18.0), not in the reference tree; the three
 * routines used are restated from its published algorithm:
 *   Matrix4::look_at_rh(eye, center, up) = look_to_rh(eye, center - eye, up):
 *       f = normalize(dir); s = normalize(cross(f, up)); u = cross(s, f);
 *       columns (s.x,u.x,-f.x,0) (s.y,u.y,-f.y,0) (s.z,u.z,-f.z,0) (-dot(eye,s), -dot(eye,u), dot(eye,f), 1)
 *   InnerSpace::normalize(v) = v * (1 / magnitude(v))
 *   SquareMatrix::invert (Matrix4) = transposed cofactors * (1 / det), det by first-row expansion.
 * The rounding order inside these is therefore a best effort; the resulting 256-byte block is an
 * INPUT of the raycast and both oracle and CUDA path consume the same bytes.
 * ------------------------------------------------------------------------------------------ */
static v3 cross(v3 a, v3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static v3 cg_normalize(v3 a) { return muls(a, 1.0f / sqrtf(dot(a, a))); }

static float det3(float a, float b, float c, float d, float e, float f, float g, float h, float i) {
  /* Matrix3::determinant, columns (a,b,c) (d,e,f) (g,h,i) */
  return a * (e * i - h * f) - d * (b * i - h * c) + g * (b * f - e * c);
}

/* m is column-major: m[4*col + row] */
static int invert4(const float *m, float *inv) {
  /* determinant by cofactor expansion along the first row of columns (cgmath Matrix4::determinant) */
  float c0 = det3(m[5], m[6], m[7], m[9], m[10], m[11], m[13], m[14], m[15]);
  float c1 = det3(m[1], m[2], m[3], m[9], m[10], m[11], m[13], m[14], m[15]);
  float c2 = det3(m[1], m[2], m[3], m[5], m[6], m[7], m[13], m[14], m[15]);
  float c3 = det3(m[1], m[2], m[3], m[5], m[6], m[7], m[9], m[10], m[11]);
  float det = m[0] * c0 - m[4] * c1 + m[8] * c2 - m[12] * c3;
  if (det == 0.0f) return 0;
  float inv_det = 1.0f / det;
  /* t = transpose(m); cf(i,j) = det(t without column i and row j) * sign * inv_det; inv column i = cf(i,0..3) */
  float t[16];
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) t[4 * c + r] = m[4 * r + c];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      float sub3[9];
      int k = 0;
      for (int c = 0; c < 4; c++) {
        if (c == i) continue;
        for (int r = 0; r < 4; r++) {
          if (r == j) continue;
          sub3[k++] = t[4 * c + r];
        }
      }
      float d = det3(sub3[0], sub3[1], sub3[2], sub3[3], sub3[4], sub3[5], sub3[6], sub3[7], sub3[8]);
      float sign = ((i + j) & 1) ? -1.0f : 1.0f;
      inv[4 * i + j] = d * sign * inv_det;
    }
  return 1;
}

void wxo_compute_state_build(const float eye[3], const float target[3], const float up[3], float aspect,
                             float fovy_deg, float resolution_width, uint32_t render_mode,
                             const uint32_t show_grid[3], const float sun_dir3[3], const float sun_color3[3],
                             float sun_intensity, WxoState *out) {
  memset(out, 0, sizeof(*out));
  v3 e = V3(eye[0], eye[1], eye[2]);
  v3 dir = sub(V3(target[0], target[1], target[2]), e);
  v3 f = cg_normalize(dir);
  v3 s = cg_normalize(cross(f, V3(up[0], up[1], up[2])));
  v3 uu = cross(s, f);
  float view[16] = {s.x, uu.x, -f.x, 0.f, s.y, uu.y, -f.y, 0.f, s.z, uu.z, -f.z, 0.f, -dot(e, s), -dot(e, uu), dot(e, f), 1.f};
  float c2w[16];
  if (!invert4(view, c2w)) memset(c2w, 0, sizeof(c2w)); /* reference panics */
  memcpy(out->view_proj, view, sizeof(view));
  memcpy(out->camera_to_world, c2w, sizeof(c2w));
  out->eye[0] = eye[0], out->eye[1] = eye[1], out->eye[2] = eye[2], out->eye[3] = 0.0f;
  float height = resolution_width / aspect;
  const float *u = &c2w[0], *v = &c2w[4], *w = &c2w[8];
  /* f32::to_radians = deg * (PI / 180) ; tan = libm tanf */
  float tan_half = tanf((fovy_deg * (3.14159265358979323846f / 180.0f)) * 0.5f);
  for (int k = 0; k < 4; k++) {
    /* wp = (-W/2)*u + (height/2)*v - w*(height/2)/tan(fovy/2) */
    out->wp[k] = ((-resolution_width / 2.0f) * u[k] + (height / 2.0f) * v[k]) - (w[k] * (height / 2.0f)) / tan_half;
    out->u[k] = u[k];
    out->mv[k] = -v[k];
  }
  out->render_mode[0] = render_mode;
  for (int k = 0; k < 3; k++) {
    out->show_345[k] = show_grid[k];
    out->sun_dir[k] = sun_dir3[k];
    out->sun_color[k] = sun_color3[k];
  }
  out->sun_color[3] = sun_intensity;
}

/* SunSettings::default (egui_dev.rs:355-367): glam Vec3::normalize = v * (1/length) */
void wxo_default_sun(float dir3[3], float color3[3], float *intensity) {
  v3 d = cg_normalize(V3(1.0f, -1.0f, 0.5f));
  dir3[0] = d.x, dir3[1] = d.y, dir3[2] = d.z;
  color3[0] = 255.f / 255.f, color3[1] = 210.f / 255.f, color3[2] = 160.f / 255.f;
  *intensity = 1.0f;
}

/* linear_to_srgb (src/render/recorder.rs:132-140): the byte-wise transfer function the recorder applies to
 * every colour channel of a captured frame before encoding (Frame::ndarray_frame, :20-37, drops alpha). */
uint8_t wxo_linear_to_srgb(uint8_t value) {
  const float c = (float)value / 255.0f;
  float srgb;
  if (c <= 0.0031308f) srgb = 12.92f * c;
  else srgb = 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
  return (uint8_t)roundf(srgb * 255.0f); /* f32::round: half away from zero; `as u8` saturates, in range here */
}

/* Frame::ndarray_frame: RGBA8 [n] -> RGB8 [n] through wxo_linear_to_srgb */
void wxo_frame_to_srgb_rgb(const uint8_t* rgba, size_t n_pixels, uint8_t* rgb) {
  for (size_t i = 0; i < n_pixels; ++i)
    for (int k = 0; k < 3; ++k) rgb[3 * i + k] = wxo_linear_to_srgb(rgba[4 * i + k]);
}
