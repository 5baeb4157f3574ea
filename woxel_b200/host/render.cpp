// render.cpp -- ComputeState::build, camera matrices and the Renderer frame entry point.
// Float operation order follows the reference (compute_state.rs:87-131) and cgmath 0.18.0's
// look_at_rh / Matrix4::invert / InnerSpace::normalize and glam 0.24's Vec3::normalize, which the
// reference calls; all f32.
#include "render.hpp"

#include <cmath>
#include <cstring>

namespace woxel::scene {

namespace {
inline Vec3 sub(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline Vec3 normalize(Vec3 a) {  // v * (1 / |v|)
  const float r = 1.0f / std::sqrt(dot(a, a));
  return {a.x * r, a.y * r, a.z * r};
}
// determinant of the 3x3 matrix with columns c0, c1, c2
inline float det3(const float c0[3], const float c1[3], const float c2[3]) {
  return c0[0] * (c1[1] * c2[2] - c2[1] * c1[2]) - c1[0] * (c0[1] * c2[2] - c2[1] * c0[2]) + c2[0] * (c0[1] * c1[2] - c1[1] * c0[2]);
}
}  // namespace

Mat4 Camera::build_view_projection_matrix() const {
  const Vec3 f = normalize(sub(target, eye));
  const Vec3 s = normalize(cross(f, up));
  const Vec3 u = cross(s, f);
  Mat4 v;
  v.m[0][0] = s.x, v.m[0][1] = u.x, v.m[0][2] = -f.x, v.m[0][3] = 0.f;
  v.m[1][0] = s.y, v.m[1][1] = u.y, v.m[1][2] = -f.y, v.m[1][3] = 0.f;
  v.m[2][0] = s.z, v.m[2][1] = u.z, v.m[2][2] = -f.z, v.m[2][3] = 0.f;
  v.m[3][0] = -dot(eye, s), v.m[3][1] = -dot(eye, u), v.m[3][2] = dot(eye, f), v.m[3][3] = 1.f;
  return v;
}

bool Mat4::invert(Mat4& out) const {
  // minors of the first row give the determinant
  auto minor_cols = [this](int skip_col, int skip_row, float c[3][3]) {
    int k = 0;
    for (int col = 0; col < 4; ++col) {
      if (col == skip_col) continue;
      int r = 0;
      for (int row = 0; row < 4; ++row) {
        if (row == skip_row) continue;
        c[k][r++] = m[col][row];
      }
      ++k;
    }
  };
  float c[3][3];
  float d[4];
  for (int col = 0; col < 4; ++col) {
    minor_cols(col, 0, c);
    d[col] = det3(c[0], c[1], c[2]);
  }
  const float det = m[0][0] * d[0] - m[1][0] * d[1] + m[2][0] * d[2] - m[3][0] * d[3];
  if (det == 0.0f) return false;
  const float inv_det = 1.0f / det;
  // inverse(col i, row j) = cofactor of the transpose: minor of *this without row i and column j
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      // transpose t has t.m[col][row] = m[row][col]; drop t's column i and t's row j
      float t[3][3];
      int k = 0;
      for (int col = 0; col < 4; ++col) {
        if (col == i) continue;
        int r = 0;
        for (int row = 0; row < 4; ++row) {
          if (row == j) continue;
          t[k][r++] = m[row][col];
        }
        ++k;
      }
      const float sign = ((i + j) & 1) ? -1.0f : 1.0f;
      out.m[i][j] = det3(t[0], t[1], t[2]) * sign * inv_det;
    }
  return true;
}

}  // namespace woxel::scene

namespace woxel::render {

SunSettings::SunSettings() {
  const float x = 1.0f, y = -1.0f, z = 0.5f;
  const float r = 1.0f / std::sqrt(x * x + y * y + z * z);
  dir3[0] = x * r, dir3[1] = y * r, dir3[2] = z * r;
  color[0] = 255.f / 255.f, color[1] = 210.f / 255.f, color[2] = 160.f / 255.f;
  intensity = 1.0f;
}

ComputeState ComputeState::build(const scene::Camera& c, float resolution_width, RenderMode render_mode, const bool show_grid[3],
                                 const float sun_dir3[3], const float sun_color3[3], float sun_intensity) {
  ComputeState s;
  memset(static_cast<WxState*>(&s), 0, sizeof(WxState));
  const scene::Mat4 view = c.build_view_projection_matrix();
  scene::Mat4 c2w;
  if (!view.invert(c2w)) throw RenderError(WX_ERR_INVALID_ARGUMENT, "Could not invert camera matrix");
  memcpy(s.view_proj, view.m, sizeof(s.view_proj));
  memcpy(s.camera_to_world, c2w.m, sizeof(s.camera_to_world));
  s.eye[0] = c.eye.x, s.eye[1] = c.eye.y, s.eye[2] = c.eye.z, s.eye[3] = 0.f;
  const float height = resolution_width / c.aspect;
  const float tan_half = std::tan((c.fovy * (3.14159265358979323846f / 180.0f)) * 0.5f);
  for (int k = 0; k < 4; ++k) {
    const float u = c2w.m[0][k], v = c2w.m[1][k], w = c2w.m[2][k];
    s.u[k] = u;
    s.mv[k] = -v;
    // wp = (-W/2) u + (height/2) v - w (height/2) / tan(fovy/2)
    s.wp[k] = ((-resolution_width / 2.0f) * u + (height / 2.0f) * v) - (w * (height / 2.0f)) / tan_half;
  }
  s.render_mode[0] = (uint32_t)render_mode;
  for (int k = 0; k < 3; ++k) {
    s.show_345[k] = show_grid[k] ? 1u : 0u;
    s.sun_dir[k] = sun_dir3[k];
    s.sun_color[k] = sun_color3[k];
  }
  s.sun_color[3] = sun_intensity;
  return s;
}

static void check(WxContext* ctx, int rc, const char* what) {
  if (rc != WX_OK) throw RenderError(rc, std::string(what) + ": " + wx_strerror(rc) + " -- " + wx_last_error(ctx));
}

Renderer::Renderer(uint32_t width, uint32_t height, int n_devices) : width_(width), height_(height) {
  check(nullptr, wx_init(n_devices, nullptr, &ctx_), "wx_init");
}

Renderer::~Renderer() {
  if (tree_) wx_tree_free(ctx_, tree_);
  if (ctx_) wx_shutdown(ctx_);
}

bool Renderer::compute_sdf_gpu(WxContext* ctx, vdb::FlatTree& flat, WxSdfInfo* info) {
  // narrow (u8) leaf distances first -- what every level set produces -- then u32
  for (int attempt = 0; attempt < 2; ++attempt) {
    flat.narrow = attempt == 0;
    if (flat.narrow) flat.tab3_u8.assign((size_t)flat.n3 * 512, 0), flat.tab3.clear();
    else flat.tab3.assign((size_t)flat.n3 * 512, 0), flat.tab3_u8.clear();
    const WxTreeDesc d = flat.desc();
    void* t3 = flat.narrow ? (void*)flat.tab3_u8.data() : (void*)flat.tab3.data();
    const int rc = wx_compute_sdf(ctx, &d, flat.tab5.data(), flat.tab4.data(), t3, flat.narrow ? 1 : 4, info);
    if (rc == WX_OK) return true;
    if (rc != WX_ERR_UNSUPPORTED) check(ctx, rc, "wx_compute_sdf");
  }
  return false;
}

void Renderer::change_vdb_model(vdb::VDB345& vdb, bool run_compute_sdf) {
  vdb::FlatTree flat;
  bool swept = false;
  last_sdf = WxSdfInfo{};
  if (run_compute_sdf && sdf_on_gpu) {
    flat = vdb.to_flat();  // topology; the distances are computed on the device
    const WxTreeDesc topo = flat.desc();
    WxTree* built = nullptr;
    const int rc = wx_tree_build(ctx_, &topo, &built, &last_sdf);  // sweep + device tables in one go
    if (rc == WX_OK) {
      if (tree_) wx_tree_free(ctx_, tree_);
      tree_ = built;
      return;
    }
    if (rc != WX_ERR_UNSUPPORTED) check(ctx_, rc, "wx_tree_build");
    swept = compute_sdf_gpu(ctx_, flat, &last_sdf);  // leaf distances above 255: the two-call path has the u32 layout
    if (!swept) last_sdf = WxSdfInfo{};
  }
  if (!swept) {
    if (run_compute_sdf) vdb.compute_sdf();
    flat = vdb.to_flat();
  }
  const WxTreeDesc d = flat.desc();
  WxTree* t = nullptr;
  check(ctx_, wx_tree_upload(ctx_, &d, &t), "wx_tree_upload");
  if (tree_) wx_tree_free(ctx_, tree_);
  tree_ = t;
}

void Renderer::change_vdb_model(const std::string& path, const std::string& grid) {
  vdb::VdbReader reader(path);
  vdb::VDB345 vdb = reader.read_vdb345_grid(grid);
  change_vdb_model(vdb, true);
}

Frame Renderer::render(const scene::Scene& scene) {
  if (!tree_) throw RenderError(WX_ERR_INVALID_ARGUMENT, "render: no model loaded");
  const ComputeState s = ComputeState::build(scene.camera, (float)width_, render_mode, show_grid, sun_settings.dir3,
                                             sun_settings.color, sun_settings.intensity);
  Frame f;
  f.width = width_, f.height = height_;
  f.rgba.resize((size_t)width_ * height_ * 4);
  check(ctx_, wx_render(ctx_, tree_, &s, 1, width_, height_, f.rgba.data(), nullptr), "wx_render");
  return f;
}

}  // namespace woxel::render
