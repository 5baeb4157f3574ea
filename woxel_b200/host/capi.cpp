// capi.cpp -- extern "C" view of the C++ host (include/woxel_host.h).
#include <cstring>
#include <new>
#include <string>

#include "../../include/woxel_host.h"
#include "procedural.hpp"
#include "render.hpp"
#include "scene.hpp"
#include "vdb.hpp"

using namespace woxel;

struct WxhVdb {
  vdb::VDB345 v;
};
struct WxhFlat {
  vdb::FlatTree f;
};
struct WxhRenderer {
  render::Renderer r;
  WxhRenderer(uint32_t w, uint32_t h, int n) : r(w, h, n) {}
};

static thread_local std::string g_err;
extern "C" const char* wxh_last_error(void) { return g_err.c_str(); }

static int status_of(const vdb::VdbError& e) {
  switch (e.kind) {
    case vdb::VdbError::MagicMismatch: return WXH_ERR_MAGIC;
    case vdb::VdbError::UnsupportedVersion: return WXH_ERR_VERSION;
    case vdb::VdbError::InvalidCompression: return WXH_ERR_COMPRESSION;
    case vdb::VdbError::InvalidGridName: return WXH_ERR_GRID_NAME;
    case vdb::VdbError::InvalidNodeMetadata: return WXH_ERR_NODE_METADATA;
    case vdb::VdbError::UnsupportedBloscFormat:
    case vdb::VdbError::InvalidBloscData: return WXH_ERR_BLOSC;
    case vdb::VdbError::Unsupported:
    case vdb::VdbError::UnexpectedMaskLength: return WXH_ERR_UNSUPPORTED;
    default: return WXH_ERR_IO;
  }
}

template <class F>
static int guarded(F&& f) {
  try {
    return f();
  } catch (const vdb::VdbError& e) {
    g_err = e.what();
    return status_of(e);
  } catch (const render::RenderError& e) {
    g_err = e.what();
    return e.status;
  } catch (const std::bad_alloc&) {
    g_err = "out of host memory";
    return WX_ERR_OUT_OF_MEMORY;
  } catch (const std::exception& e) {
    g_err = e.what();
    return WXH_ERR_UNSUPPORTED;
  }
}

template <class F>
static auto by_level(int level, F&& f) {
  switch (level) {
    case 3: return f(vdb::N3{});
    case 4: return f(vdb::N4{});
    default: return f(vdb::N5{});
  }
}
static bool bad_level(int level) { return level < 3 || level > 5; }

extern "C" int wxh_global_to_node(int level, const int32_t g[3], int32_t out[3]) {
  if (bad_level(level)) return WXH_ERR_INVALID_ARGUMENT;
  by_level(level, [&](auto nm) {
    auto r = decltype(nm)::global_to_node({g[0], g[1], g[2]});
    out[0] = r[0], out[1] = r[1], out[2] = r[2];
    return 0;
  });
  return 0;
}
extern "C" int64_t wxh_global_to_offset(int level, const int32_t g[3]) {
  if (bad_level(level)) return WXH_ERR_INVALID_ARGUMENT;
  return by_level(level, [&](auto nm) { return (int64_t) decltype(nm)::global_to_offset({g[0], g[1], g[2]}); });
}
extern "C" int wxh_offset_to_child(int level, uint32_t offset, uint32_t out[3]) {
  if (bad_level(level)) return WXH_ERR_INVALID_ARGUMENT;
  by_level(level, [&](auto nm) {
    auto r = decltype(nm)::offset_to_child(offset);
    out[0] = r[0], out[1] = r[1], out[2] = r[2];
    return 0;
  });
  return 0;
}
extern "C" int64_t wxh_child_to_offset(int level, const uint32_t c[3]) {
  if (bad_level(level)) return WXH_ERR_INVALID_ARGUMENT;
  return by_level(level, [&](auto nm) { return (int64_t) decltype(nm)::child_to_offset({c[0], c[1], c[2]}); });
}

extern "C" WxhVdb* wxh_vdb_new(void) { return new (std::nothrow) WxhVdb(); }
extern "C" void wxh_vdb_free(WxhVdb* v) { delete v; }
// void / pointer-returning entry points: no exception crosses the C ABI either -- a failure (out of host memory) leaves its
// text in wxh_last_error() and, where there is a pointer to return, returns NULL.
extern "C" void wxh_vdb_set_voxel(WxhVdb* v, int32_t x, int32_t y, int32_t z, uint32_t value) {
  (void)guarded([&]() { return v->v.set_voxel({x, y, z}, value), 0; });
}
extern "C" void wxh_vdb_set_voxels(WxhVdb* v, const int32_t* xyz, size_t n, uint32_t value) {
  (void)guarded([&]() {
    for (size_t i = 0; i < n; ++i) v->v.set_voxel({xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, value);
    return 0;
  });
}
extern "C" int wxh_vdb_get_voxel(const WxhVdb* v, int32_t x, int32_t y, int32_t z, uint32_t* value, int* level) {
  const vdb::VdbEndpoint e = v->v.get_voxel({x, y, z});
  if (value) *value = e.value;
  if (level) *level = e.level;
  return (int)e.kind;
}
extern "C" void wxh_vdb_count_nodes(const WxhVdb* v, uint64_t out[3]) {
  const auto c = v->v.count_nodes();
  out[0] = c[0], out[1] = c[1], out[2] = c[2];
}
extern "C" uint64_t wxh_vdb_count_leaf_values(const WxhVdb* v) { return v->v.count_leaf_values(); }
extern "C" void wxh_vdb_compute_sdf(WxhVdb* v) {
  (void)guarded([&]() { return v->v.compute_sdf(), 0; });
}

extern "C" int wxh_blosc_decompress(const uint8_t* frame, size_t n, uint8_t* out, size_t cap, size_t* out_len) {
  if (!frame || !out_len || (!out && cap)) return WXH_ERR_INVALID_ARGUMENT;
  return guarded([&]() {
    *out_len = n >= 16 ? ((size_t)frame[4] | ((size_t)frame[5] << 8) | ((size_t)frame[6] << 16) | ((size_t)frame[7] << 24)) : 0;  // room needed, per the header
    const std::vector<uint8_t> v = vdb::decompress_blosc_frame(frame, n, cap);  // refuses before allocating when the header claims more than cap
    *out_len = v.size();
    if (v.size() > cap) return (int)WXH_ERR_INVALID_ARGUMENT;  // *out_len says how much room is needed
    if (!v.empty()) memcpy(out, v.data(), v.size());
    return 0;
  });
}

extern "C" int wxh_vdb_read(const char* path, const char* grid_name, WxhVdb** out, WxhVdbInfo* info) {
  if (!path || !grid_name || !out) return WXH_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded([&]() {
    vdb::VdbReader reader{std::string(path)};
    WxhVdb* v = new WxhVdb();
    try {
      v->v = reader.read_vdb345_grid(grid_name);
    } catch (...) {
      delete v;
      throw;
    }
    if (info) {
      const vdb::GridDescriptor& gd = v->v.grid_descriptor;
      memset(info, 0, sizeof(*info));
      info->file_version = reader.header.file_version;
      info->library_major = reader.header.library_major, info->library_minor = reader.header.library_minor;
      info->grid_count = reader.header.grid_number;
      info->grid_compression = gd.compression;
      info->is_half_float = gd.meta_data.is_half_float();
      auto it = gd.meta_data.ints.find("file_voxel_count");
      info->file_voxel_count = it == gd.meta_data.ints.end() ? -1 : it->second;
      info->grid_pos = gd.grid_pos, info->block_pos = gd.block_pos, info->end_pos = gd.end_pos;
    }
    *out = v;
    return 0;
  });
}

extern "C" WxhFlat* wxh_vdb_to_flat(const WxhVdb* v, int narrow_leaves) {
  WxhFlat* f = new (std::nothrow) WxhFlat();
  if (f && guarded([&]() { return f->f = v->v.to_flat(narrow_leaves != 0), 0; }) != 0) delete f, f = nullptr;
  return f;
}
extern "C" void wxh_flat_free(WxhFlat* f) { delete f; }
extern "C" void wxh_flat_desc(const WxhFlat* f, WxTreeDesc* out) { *out = f->f.desc(); }

extern "C" WxhVdb* wxh_build_sphere(int32_t half, double radius, double band) {
  WxhVdb* v = new (std::nothrow) WxhVdb();
  if (v && guarded([&]() { return v->v = procedural::sphere_shell(half, radius, band), 0; }) != 0) delete v, v = nullptr;
  return v;
}
extern "C" WxhVdb* wxh_build_torus(int32_t half, double major, double minor, double band) {
  WxhVdb* v = new (std::nothrow) WxhVdb();
  if (v && guarded([&]() { return v->v = procedural::torus_shell(half, major, minor, band), 0; }) != 0) delete v, v = nullptr;
  return v;
}
extern "C" WxhVdb* wxh_build_fog(int32_t half, double tau, double* occupancy) {
  WxhVdb* v = new (std::nothrow) WxhVdb();
  if (v && guarded([&]() { return v->v = procedural::fbm_fog(half, tau, occupancy), 0; }) != 0) delete v, v = nullptr;
  return v;
}

static scene::Camera camera_of(const float eye[3], const float target[3], const float up[3], float aspect, float fovy) {
  scene::Camera c;
  c.eye = {eye[0], eye[1], eye[2]};
  c.target = {target[0], target[1], target[2]};
  c.up = {up[0], up[1], up[2]};
  c.aspect = aspect, c.fovy = fovy;
  return c;
}

extern "C" int wxh_compute_state_build(const float eye[3], const float target[3], const float up[3], float aspect, float fovy_deg,
                                       float resolution_width, uint32_t render_mode, const uint32_t show_grid[3],
                                       const float sun_dir3[3], const float sun_color3[3], float sun_intensity, WxState* out) {
  return guarded([&]() {
    const bool sg[3] = {show_grid[0] != 0, show_grid[1] != 0, show_grid[2] != 0};
    const render::ComputeState s = render::ComputeState::build(camera_of(eye, target, up, aspect, fovy_deg), resolution_width,
                                                               (render::RenderMode)render_mode, sg, sun_dir3, sun_color3, sun_intensity);
    *out = s;
    return 0;
  });
}
extern "C" void wxh_default_sun(float dir3[3], float color3[3], float* intensity) {
  const render::SunSettings s;
  memcpy(dir3, s.dir3, sizeof(s.dir3));
  memcpy(color3, s.color, sizeof(s.color));
  *intensity = s.intensity;
}

extern "C" int wxh_renderer_new(uint32_t width, uint32_t height, int n_devices, WxhRenderer** out) {
  if (!out) return WXH_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  return guarded([&]() {
    *out = new WxhRenderer(width, height, n_devices);
    return 0;
  });
}
extern "C" void wxh_renderer_free(WxhRenderer* r) { delete r; }
extern "C" int wxh_renderer_change_vdb_model(WxhRenderer* r, WxhVdb* v, int run_compute_sdf) {
  return guarded([&]() {
    r->r.change_vdb_model(v->v, run_compute_sdf != 0);
    return 0;
  });
}
extern "C" int wxh_renderer_change_vdb_model_file(WxhRenderer* r, const char* path, const char* grid_name) {
  return guarded([&]() {
    r->r.change_vdb_model(std::string(path), std::string(grid_name));
    return 0;
  });
}
extern "C" int wxh_renderer_set_options(WxhRenderer* r, uint32_t render_mode, const uint32_t show_grid[3], const float sun_dir3[3],
                                        const float sun_color3[3], float sun_intensity) {
  r->r.render_mode = (render::RenderMode)render_mode;
  for (int k = 0; k < 3; ++k) {
    if (show_grid) r->r.show_grid[k] = show_grid[k] != 0;
    if (sun_dir3) r->r.sun_settings.dir3[k] = sun_dir3[k];
    if (sun_color3) r->r.sun_settings.color[k] = sun_color3[k];
  }
  r->r.sun_settings.intensity = sun_intensity;
  return 0;
}
extern "C" int wxh_renderer_render(WxhRenderer* r, const float eye[3], const float target[3], const float up[3], float aspect,
                                   float fovy_deg, uint8_t* rgba_out) {
  return guarded([&]() {
    scene::Scene sc;
    sc.camera = camera_of(eye, target, up, aspect, fovy_deg);
    const render::Frame f = r->r.render(sc);
    memcpy(rgba_out, f.rgba.data(), f.rgba.size());
    return 0;
  });
}
extern "C" int wxh_renderer_set_sdf_on_gpu(WxhRenderer* r, int on) {
  r->r.sdf_on_gpu = on != 0;
  return 0;
}
extern "C" void wxh_renderer_last_sdf(const WxhRenderer* r, WxSdfInfo* out) { *out = r->r.last_sdf; }
extern "C" int wxh_flat_compute_sdf_gpu(WxhFlat* f, WxContext* ctx, WxSdfInfo* info) {
  return guarded([&]() { return render::Renderer::compute_sdf_gpu(ctx, f->f, info) ? 0 : (int)WX_ERR_UNSUPPORTED; });
}
extern "C" int wxh_write_ppm(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height) {
  return guarded([&]() { return render::write_ppm(path, rgb, width, height), 0; });
}
extern "C" int wxh_write_png(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height) {
  return guarded([&]() { return render::write_png(path, rgb, width, height), 0; });
}
extern "C" WxContext* wxh_renderer_context(WxhRenderer* r) { return r->r.context(); }
extern "C" WxTree* wxh_renderer_tree(WxhRenderer* r) { return r->r.tree(); }
