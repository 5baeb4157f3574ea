// render.hpp -- the reference's `src/render` surface for the raycast path:
//   RenderMode (src/render/egui_dev.rs:11-18), SunSettings (:348-368), show_grid (:41),
//   ComputeState::build (src/render/gpu_types/compute_state.rs:87-131),
//   the frame entry point WgpuContext::render / change_vdb_model (src/render/wgpu_context.rs:207-292, :506-573)
// with wgpu replaced by the C ABI of include/woxel_b200.h.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/woxel_b200.h"
#include "scene.hpp"
#include "vdb.hpp"

namespace woxel::render {

enum class RenderMode : uint32_t { Gray = 0, Rgb = 1, Ray = 2, Diffuse = 3, Glossy = 4 };

struct SunSettings {
  float dir3[3];
  float color[3];
  float intensity;
  SunSettings();  // egui_dev.rs:355-367: normalize(1,-1,0.5), (1, 210/255, 160/255), 1.0
};

// == WxState; the name and the builder follow compute_state.rs
struct ComputeState : WxState {
  static ComputeState build(const scene::Camera& c, float resolution_width, RenderMode render_mode, const bool show_grid[3],
                            const float sun_dir3[3], const float sun_color3[3], float sun_intensity);
};
static_assert(sizeof(ComputeState) == 256, "ComputeState must stay the 256-byte uniform of the reference");

class RenderError : public std::runtime_error {
 public:
  RenderError(int status, const std::string& what) : std::runtime_error(what), status(status) {}
  int status;
};

struct Frame {
  uint32_t width = 0, height = 0;
  std::vector<uint8_t> rgba;  // height x width x 4, rgba8unorm
};

// The WgpuContext of the new build: owns the device context, the uploaded model and the GUI-side
// render options (egui_dev.rs:37-51 keeps them next to the context in the reference too).
class Renderer {
 public:
  Renderer(uint32_t width, uint32_t height, int n_devices = 0);
  ~Renderer();
  Renderer(const Renderer&) = delete;
  Renderer& operator=(const Renderer&) = delete;

  // wgpu_context.rs:506-573: compute_sdf + serialise + upload
  void change_vdb_model(vdb::VDB345& vdb, bool run_compute_sdf = true);
  void change_vdb_model(const std::string& path, const std::string& grid);
  // wgpu_context.rs:207-292 (compute part) + the capture read-back (:374-405)
  Frame render(const scene::Scene& scene);
  void resize(uint32_t width, uint32_t height) { width_ = width, height_ = height; }

  RenderMode render_mode = RenderMode::Diffuse;  // egui_dev.rs:59
  bool show_grid[3] = {false, false, false};     // egui_dev.rs:60
  SunSettings sun_settings;                      // egui_dev.rs:61
  // compute_sdf (wgpu_context.rs:104, :513) runs on the GPU (wx_compute_sdf, identical values); the host sweep
  // VDB345::compute_sdf takes over when a distance does not fit that path, or when this is false
  bool sdf_on_gpu = true;
  WxSdfInfo last_sdf{};    // timing of the last GPU sweep (device_ms == 0: the host sweep ran)

  WxContext* context() const { return ctx_; }
  WxTree* tree() const { return tree_; }

  // Fills the distances of `flat` (tab5 / tab4 tiles, tab3) on the GPU.  False: not representable there.
  static bool compute_sdf_gpu(WxContext* ctx, vdb::FlatTree& flat, WxSdfInfo* info);

 private:
  uint32_t width_, height_;
  WxContext* ctx_ = nullptr;
  WxTree* tree_ = nullptr;
};

// Frame dumps of the capture step (frame_dump.cpp): RGB8 [height][width][3] as binary PPM or PNG.  Throw std::exception on I/O errors.
void write_ppm(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height);
void write_png(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height);

}  // namespace woxel::render
