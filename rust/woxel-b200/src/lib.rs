//! Safe layer over `woxel-b200-sys`.
//!
//! UNCOMPILED in this repository's build environment (no rustc); the same call sequence is compiled and tested from C++
//! (`woxel_b200/host/render.cpp`), C99 (`tests/c/abi_smoke.c`) and Python (`woxel_b200/render.py`).
//!
//! Surface kept from woxel (file:line of the reference):
//!  * `ComputeState` + `ComputeState::build` -- `src/render/gpu_types/compute_state.rs:9-29, :87-131`
//!  * `Camera` (`quick_camera`, `build_view_projection_matrix`) -- `src/render/camera.rs:7-35`
//!  * `RenderMode` -- `src/render/egui_dev.rs:11-18`
//!  * `FlatTree` -- what `VDB345::to_flat()` fills; replaces `origins()/masks()/atlas()` of `src/vdb/vdb345.rs:108-264`
//!  * `Renderer::{new, change_vdb_model, render}` -- `WgpuContext::{new, change_vdb_model, render}`,
//!    `src/render/wgpu_context.rs:33, :506, :207`
use std::ffi::CStr;
use std::os::raw::c_void;
use std::ptr;

pub use woxel_b200_sys as sys;
use sys::*;

#[derive(Debug)]
pub struct RenderError {
    pub status: i32,
    pub detail: String,
}

impl std::fmt::Display for RenderError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        let name = unsafe { CStr::from_ptr(wx_strerror(self.status)) }.to_string_lossy();
        write!(f, "{} ({}): {}", name, self.status, self.detail)
    }
}
impl std::error::Error for RenderError {}

fn check(ctx: *const WxContext, rc: i32) -> Result<(), RenderError> {
    if rc == WX_OK {
        return Ok(());
    }
    let detail = unsafe { CStr::from_ptr(wx_last_error(ctx)) }.to_string_lossy().into_owned();
    Err(RenderError { status: rc, detail })
}

/// `RenderMode` of `src/render/egui_dev.rs:11-18` (the value the shader switches on, `raycast.comp.wgsl:168`).
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
#[repr(u32)]
pub enum RenderMode {
    Gray = 0,
    Rgb = 1,
    Ray = 2,
    Diffuse = 3,
    Glossy = 4,
}

/// `Camera` of `src/render/camera.rs:7-29`.
#[derive(Debug, Clone, Copy)]
pub struct Camera {
    pub eye: [f32; 3],
    pub target: [f32; 3],
    pub up: [f32; 3],
    pub aspect: f32,
    /// y-axis field of view in degrees
    pub fovy: f32,
}

fn sub(a: [f32; 3], b: [f32; 3]) -> [f32; 3] {
    [a[0] - b[0], a[1] - b[1], a[2] - b[2]]
}
fn dot(a: [f32; 3], b: [f32; 3]) -> f32 {
    a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
}
fn cross(a: [f32; 3], b: [f32; 3]) -> [f32; 3] {
    [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]
}
fn normalize(a: [f32; 3]) -> [f32; 3] {
    let r = 1.0 / dot(a, a).sqrt();
    [a[0] * r, a[1] * r, a[2] * r]
}
/// determinant of the 3x3 matrix with columns c0, c1, c2
fn det3(c0: [f32; 3], c1: [f32; 3], c2: [f32; 3]) -> f32 {
    c0[0] * (c1[1] * c2[2] - c2[1] * c1[2]) - c1[0] * (c0[1] * c2[2] - c2[1] * c0[2]) + c2[0] * (c0[1] * c1[2] - c1[1] * c0[2])
}

/// Column-major 4x4 (`m[col][row]`), like cgmath's `Matrix4`.
pub type Mat4 = [[f32; 4]; 4];

/// General inverse by cofactors, the operation order of `woxel_b200/host/render.cpp::Mat4::invert` (which follows cgmath 0.18).
pub fn invert(m: &Mat4) -> Option<Mat4> {
    let minor = |skip_col: usize, skip_row: usize, transposed: bool| -> f32 {
        let mut c = [[0f32; 3]; 3];
        let mut k = 0;
        for col in 0..4 {
            if col == skip_col {
                continue;
            }
            let mut r = 0;
            for row in 0..4 {
                if row == skip_row {
                    continue;
                }
                c[k][r] = if transposed { m[row][col] } else { m[col][row] };
                r += 1;
            }
            k += 1;
        }
        det3(c[0], c[1], c[2])
    };
    let d = [minor(0, 0, false), minor(1, 0, false), minor(2, 0, false), minor(3, 0, false)];
    let det = m[0][0] * d[0] - m[1][0] * d[1] + m[2][0] * d[2] - m[3][0] * d[3];
    if det == 0.0 {
        return None;
    }
    let inv_det = 1.0 / det;
    let mut out = [[0f32; 4]; 4];
    for i in 0..4 {
        for j in 0..4 {
            let sign = if (i + j) & 1 == 1 { -1.0 } else { 1.0 };
            out[i][j] = minor(i, j, true) * sign * inv_det;
        }
    }
    Some(out)
}

impl Camera {
    /// `Camera::quick_camera` (`src/render/camera.rs:16-29`).
    pub fn quick_camera(aspect: f32) -> Self {
        Camera { eye: [0.5, 0.5, -500.5], target: [0.5, 0.5, -498.5], up: [0.0, 1.0, 0.0], aspect, fovy: 45.0 }
    }

    /// `look_at_rh(eye, target, up)` (`src/render/camera.rs:31-35`): the view matrix, no projection.
    pub fn build_view_projection_matrix(&self) -> Mat4 {
        let f = normalize(sub(self.target, self.eye));
        let s = normalize(cross(f, self.up));
        let u = cross(s, f);
        [
            [s[0], u[0], -f[0], 0.0],
            [s[1], u[1], -f[1], 0.0],
            [s[2], u[2], -f[2], 0.0],
            [-dot(self.eye, s), -dot(self.eye, u), dot(self.eye, f), 1.0],
        ]
    }
}

/// `ComputeState` (`src/render/gpu_types/compute_state.rs:9-29`): layout-identical to `WxState`.
pub type ComputeState = WxState;

/// `ComputeState::build` (`src/render/gpu_types/compute_state.rs:87-131`).
pub fn compute_state_build(
    c: &Camera,
    resolution_width: f32,
    render_mode: RenderMode,
    show_grid: [bool; 3],
    sun_dir3: [f32; 3],
    sun_color3: [f32; 3],
    sun_intensity: f32,
) -> Result<ComputeState, RenderError> {
    let view = c.build_view_projection_matrix();
    let c2w = invert(&view).ok_or(RenderError { status: WX_ERR_INVALID_ARGUMENT, detail: "Could not invert camera matrix".into() })?;
    let height = resolution_width / c.aspect;
    let tan_half = (c.fovy.to_radians() * 0.5).tan();
    let mut s: ComputeState = unsafe { std::mem::zeroed() };
    for col in 0..4 {
        for row in 0..4 {
            s.view_proj[col * 4 + row] = view[col][row];
            s.camera_to_world[col * 4 + row] = c2w[col][row];
        }
    }
    s.eye = [c.eye[0], c.eye[1], c.eye[2], 0.0];
    for k in 0..4 {
        let (u, v, w) = (c2w[0][k], c2w[1][k], c2w[2][k]);
        s.u[k] = u;
        s.mv[k] = -v;
        // wp = (-W/2) u + (height/2) v - w (height/2) / tan(fovy/2)
        s.wp[k] = ((-resolution_width / 2.0) * u + (height / 2.0) * v) - (w * (height / 2.0)) / tan_half;
    }
    s.render_mode = [render_mode as u32, 0, 0, 0];
    s.show_345 = [show_grid[0] as u32, show_grid[1] as u32, show_grid[2] as u32, 0];
    s.sun_dir = [sun_dir3[0], sun_dir3[1], sun_dir3[2], 0.0];
    s.sun_color = [sun_color3[0], sun_color3[1], sun_color3[2], sun_intensity];
    Ok(s)
}

/// What `VDB345::to_flat()` fills: the reference's DFS node order, masks as u64 words, one u32 per
/// slot (child index or SDF distance).  Replaces `origins()/masks()/atlas()` (`src/vdb/vdb345.rs:108-264`).
#[derive(Debug, Default, Clone)]
pub struct FlatTree {
    pub origins: Vec<[i32; 3]>,
    pub kids5: Vec<u64>,
    pub vals5: Vec<u64>,
    pub tab5: Vec<u32>,
    pub kids4: Vec<u64>,
    pub vals4: Vec<u64>,
    pub tab4: Vec<u32>,
    pub vals3: Vec<u64>,
    pub tab3: Vec<u32>,
}

impl FlatTree {
    pub fn desc(&self) -> WxTreeDesc {
        WxTreeDesc {
            n5: self.origins.len() as u32,
            n4: (self.kids4.len() / 64) as u32,
            n3: (self.vals3.len() / 8) as u32,
            origins: self.origins.as_ptr() as *const i32,
            kids5: self.kids5.as_ptr(),
            vals5: self.vals5.as_ptr(),
            tab5: self.tab5.as_ptr(),
            kids4: self.kids4.as_ptr(),
            vals4: self.vals4.as_ptr(),
            tab4: self.tab4.as_ptr(),
            vals3: self.vals3.as_ptr(),
            tab3: self.tab3.as_ptr() as *const c_void,
            tab3_elem_bytes: 4,
            reserved: 0,
        }
    }
}

/// `wx_init` / `wx_shutdown`.
pub struct Context {
    raw: *mut WxContext,
}

impl Context {
    pub fn new(device_ids: &[i32]) -> Result<Self, RenderError> {
        let mut raw = ptr::null_mut();
        let ids = if device_ids.is_empty() { ptr::null() } else { device_ids.as_ptr() };
        check(ptr::null(), unsafe { wx_init(device_ids.len() as i32, ids, &mut raw) })?;
        Ok(Context { raw })
    }
    pub fn set_option(&mut self, option: i32, value: i64) -> Result<(), RenderError> {
        check(self.raw, unsafe { wx_set_option(self.raw, option, value) })
    }
    /// `wx_tree_build`: compute_sdf on the GPU + device tables in one call.
    pub fn build_tree(&mut self, flat: &FlatTree) -> Result<Tree, RenderError> {
        let mut t = ptr::null_mut();
        let mut info = WxSdfInfo::default();
        check(self.raw, unsafe { wx_tree_build(self.raw, &flat.desc(), &mut t, &mut info) })?;
        Ok(Tree { ctx: self.raw, raw: t, sdf: Some(info) })
    }
    /// `wx_tree_upload`: the distances of `flat` as they are (after `VDB345::compute_sdf()` on the host).
    pub fn upload_tree(&mut self, flat: &FlatTree) -> Result<Tree, RenderError> {
        let mut t = ptr::null_mut();
        check(self.raw, unsafe { wx_tree_upload(self.raw, &flat.desc(), &mut t) })?;
        Ok(Tree { ctx: self.raw, raw: t, sdf: None })
    }
    /// `wx_render`: one frame per state into host memory (`rgba.len() == states.len() * width * height * 4`).
    pub fn render(&mut self, tree: &Tree, states: &[ComputeState], width: u32, height: u32, rgba: &mut [u8]) -> Result<(), RenderError> {
        assert_eq!(rgba.len(), states.len() * width as usize * height as usize * 4);
        check(self.raw, unsafe { wx_render(self.raw, tree.raw, states.as_ptr(), states.len() as u32, width, height, rgba.as_mut_ptr(), ptr::null()) })
    }
    pub fn last_render_info(&self) -> Result<WxRenderInfo, RenderError> {
        let mut info = WxRenderInfo::default();
        check(self.raw, unsafe { wx_last_render_info(self.raw, &mut info) })?;
        Ok(info)
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { wx_shutdown(self.raw) };
    }
}

pub struct Tree {
    ctx: *mut WxContext,
    raw: *mut WxTree,
    pub sdf: Option<WxSdfInfo>,
}

impl Drop for Tree {
    fn drop(&mut self) {
        unsafe { wx_tree_free(self.ctx, self.raw) };
    }
}

/// The frame entry point woxel keeps (`WgpuContext`, `src/render/wgpu_context.rs:16-30`): owns the context, the current model
/// and a frame buffer.  Field order matters: the tree must drop before the context.
pub struct Renderer {
    tree: Option<Tree>,
    ctx: Context,
    pub width: u32,
    pub height: u32,
    frame: Vec<u8>,
}

impl Renderer {
    /// `WgpuContext::new` (`:33-99`) without the window.
    pub fn new(width: u32, height: u32) -> Result<Self, RenderError> {
        Ok(Renderer { tree: None, ctx: Context::new(&[])?, width, height, frame: vec![0; width as usize * height as usize * 4] })
    }
    /// `WgpuContext::change_vdb_model` (`:506-573`): `vdb.compute_sdf(); vdb.atlas(); MaskUniform::from(&vdb)` + uploads.
    pub fn change_vdb_model(&mut self, flat: &FlatTree) -> Result<(), RenderError> {
        self.tree = None;
        self.tree = Some(self.ctx.build_tree(flat)?);
        Ok(())
    }
    /// `WgpuContext::render` (`:207-292`): `ComputeState::build` + the compute pass.  Returns the RGBA8 frame.
    pub fn render(&mut self, camera: &Camera, mode: RenderMode, show_grid: [bool; 3], sun_dir: [f32; 3], sun_color: [f32; 3], sun_intensity: f32) -> Result<&[u8], RenderError> {
        let tree = self.tree.as_ref().ok_or(RenderError { status: WX_ERR_INVALID_ARGUMENT, detail: "no model loaded".into() })?;
        let state = compute_state_build(camera, self.width as f32, mode, show_grid, sun_dir, sun_color, sun_intensity)?;
        let (w, h) = (self.width, self.height);
        let frame = &mut self.frame;
        check(self.ctx.raw, unsafe { wx_render(self.ctx.raw, tree.raw, &state, 1, w, h, frame.as_mut_ptr(), ptr::null()) })?;
        Ok(&self.frame)
    }
}
