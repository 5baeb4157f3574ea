// scene.hpp -- the reference's `src/scene` surface plus the camera it owns:
//   Camera (src/render/camera.rs:6-35), Scene (src/scene/scene.rs:7-24), State (src/scene/state.rs).
// Input handling (CameraController, cursor/keyboard events, frame timing) is interactive
// windowing and is out of scope; only the data the raycast consumes is kept.
#pragma once
#include <array>
#include <cstdint>

namespace woxel::scene {

struct Vec3 {
  float x = 0, y = 0, z = 0;
};

// cgmath::Matrix4<f32>, column-major: m[col][row]
struct Mat4 {
  float m[4][4] = {};
  // SquareMatrix::invert; returns false when the determinant is zero (the reference panics)
  bool invert(Mat4& out) const;
};

struct Camera {
  Vec3 eye, target, up{0.f, 1.f, 0.f};
  float aspect = 1.f;
  float fovy = 45.f;  // degrees

  // camera.rs:16-29
  static Camera quick_camera(float aspect) {
    Camera c;
    c.eye = {0.5f, 0.5f, -500.5f};
    c.target = {0.5f, 0.5f, -498.5f};
    c.up = {0.f, 1.f, 0.f};
    c.aspect = aspect;
    c.fovy = 45.f;
    return c;
  }
  // camera.rs:31-35: despite the name this is the view matrix only (look_at_rh)
  Mat4 build_view_projection_matrix() const;
};

// src/scene/state.rs: resolution + timing; only the resolution matters to the raycast
struct State {
  std::array<float, 2> resolution{0.f, 0.f};
};

// src/scene/scene.rs:7-24
struct Scene {
  State state;
  Camera camera;
  static Scene make(uint32_t width, uint32_t height) {
    Scene s;
    s.state.resolution = {(float)width, (float)height};
    s.camera = Camera::quick_camera((float)width / (float)height);
    return s;
  }
};

}  // namespace woxel::scene
