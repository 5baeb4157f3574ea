//! Raw bindings to `include/woxel_b200.h` (C ABI version 2).
//!
//! UNCOMPILED in this repository's build environment (no rustc).  `tests/test_rust_crate.py` checks every struct and
//! function below against the C header, field by field and argument by argument.
//!
//! What each call replaces in woxel: see the table at the top of `include/woxel_b200.h`
//! (`src/render/wgpu_context.rs:33-159, :207-292, :374-405` of the reference).
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const WX_ABI_VERSION: c_int = 2;

pub const WX_OK: c_int = 0;
pub const WX_ERR_INVALID_ARGUMENT: c_int = -1;
pub const WX_ERR_NO_DEVICE: c_int = -2;
pub const WX_ERR_CUDA: c_int = -3;
pub const WX_ERR_OUT_OF_MEMORY: c_int = -4;
pub const WX_ERR_BAD_TREE: c_int = -5;
pub const WX_ERR_UNSUPPORTED: c_int = -6;
pub const WX_ERR_PEER_ACCESS: c_int = -7;

pub const WX_OPT_MARCH: c_int = 1;
pub const WX_OPT_KERNEL: c_int = 2;
pub const WX_OPT_RENDER_CHUNKS: c_int = 3;
pub const WX_OPT_SMEM_PAD: c_int = 4;
pub const WX_OPT_NVTX: c_int = 5;
pub const WX_OPT_LONG_FIRST: c_int = 6;
pub const WX_OPT_LONG_THRESHOLD: c_int = 7;

#[repr(C)]
pub struct WxContext {
    _private: [u8; 0],
}

#[repr(C)]
pub struct WxTree {
    _private: [u8; 0],
}

/// Host-side description of a VDB345 tree in the reference's node order (`src/vdb/vdb345.rs:134-158, :186-261`).
#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct WxTreeDesc {
    pub n5: u32,
    pub n4: u32,
    pub n3: u32,
    pub origins: *const i32,
    pub kids5: *const u64,
    pub vals5: *const u64,
    pub tab5: *const u32,
    pub kids4: *const u64,
    pub vals4: *const u64,
    pub tab4: *const u32,
    pub vals3: *const u64,
    pub tab3: *const c_void,
    pub tab3_elem_bytes: u32,
    pub reserved: u32,
}

/// == `ComputeState` (`src/render/gpu_types/compute_state.rs:9-29`), 256 bytes.
#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct WxState {
    pub view_proj: [f32; 16],
    pub camera_to_world: [f32; 16],
    pub eye: [f32; 4],
    pub u: [f32; 4],
    pub mv: [f32; 4],
    pub wp: [f32; 4],
    pub render_mode: [u32; 4],
    pub show_345: [u32; 4],
    pub sun_dir: [f32; 4],
    pub sun_color: [f32; 4],
}

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct WxAov {
    pub state: *mut u8,
    pub voxel: *mut i32,
    pub leaf: *mut i32,
    pub level: *mut u8,
    pub iters: *mut u32,
    pub depth: *mut f32,
    pub mask: *mut u8,
    pub pos: *mut f32,
}

#[repr(C)]
#[derive(Debug, Clone, Copy, Default)]
pub struct WxShard {
    pub index: u32,
    pub count: u32,
    pub band_rows: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Debug, Clone, Copy, Default)]
pub struct WxTreeInfo {
    pub n5: u32,
    pub n4: u32,
    pub n3: u32,
    pub leaf_bits: u32,
    pub device_bytes: u64,
    pub max_dist: [u32; 3],
    pub n_devices: u32,
}

#[repr(C)]
#[derive(Debug, Clone, Copy, Default)]
pub struct WxRenderInfo {
    pub kernel_ms: f32,
    pub total_ms: f32,
    pub rays: u64,
    pub launches: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Debug, Clone, Copy, Default)]
pub struct WxSdfInfo {
    pub max_dist: [u32; 3],
    pub rounds: u32,
    pub device_ms: f32,
    pub total_ms: f32,
}

extern "C" {
    pub fn wx_abi_version() -> c_int;
    pub fn wx_strerror(status: c_int) -> *const c_char;
    pub fn wx_last_error(ctx: *const WxContext) -> *const c_char;

    pub fn wx_init(n_devices: c_int, device_ids: *const c_int, out: *mut *mut WxContext) -> c_int;
    pub fn wx_shutdown(ctx: *mut WxContext) -> c_int;
    pub fn wx_device_count(ctx: *const WxContext) -> c_int;

    pub fn wx_tree_upload(ctx: *mut WxContext, desc: *const WxTreeDesc, out: *mut *mut WxTree) -> c_int;
    pub fn wx_tree_free(ctx: *mut WxContext, tree: *mut WxTree) -> c_int;
    pub fn wx_tree_info(tree: *const WxTree, info: *mut WxTreeInfo) -> c_int;

    pub fn wx_compute_sdf(
        ctx: *mut WxContext,
        topo: *const WxTreeDesc,
        tab5_out: *mut u32,
        tab4_out: *mut u32,
        tab3_out: *mut c_void,
        tab3_elem_bytes: u32,
        info: *mut WxSdfInfo,
    ) -> c_int;
    pub fn wx_tree_build(ctx: *mut WxContext, topo: *const WxTreeDesc, out: *mut *mut WxTree, info: *mut WxSdfInfo) -> c_int;

    pub fn wx_render(
        ctx: *mut WxContext,
        tree: *const WxTree,
        states: *const WxState,
        n_states: u32,
        width: u32,
        height: u32,
        rgba_out: *mut u8,
        aov_out: *const WxAov,
    ) -> c_int;
    pub fn wx_render_device(
        ctx: *mut WxContext,
        device_index: c_int,
        tree: *const WxTree,
        states: *const WxState,
        n_states: u32,
        width: u32,
        height: u32,
        rgba_dev: *mut u8,
        aov_dev: *const WxAov,
        shard: *const WxShard,
        stream: *mut c_void,
    ) -> c_int;
    pub fn wx_render_shard(
        ctx: *mut WxContext,
        tree: *const WxTree,
        states: *const WxState,
        n_states: u32,
        width: u32,
        height: u32,
        shard: *const WxShard,
        rgba_out: *mut u8,
    ) -> c_int;
    pub fn wx_last_render_info(ctx: *const WxContext, info: *mut WxRenderInfo) -> c_int;

    pub fn wx_set_option(ctx: *mut WxContext, option: c_int, value: i64) -> c_int;
    pub fn wx_get_option(ctx: *const WxContext, option: c_int, value_out: *mut i64) -> c_int;

    pub fn wx_capture_srgb(ctx: *mut WxContext, n_states: u32, width: u32, height: u32, rgb_out: *mut u8) -> c_int;
    pub fn wx_srgb_table(table_out: *mut u8) -> c_int;

    pub fn wx_shard_rows(height: u32, shard: *const WxShard, row_mask_out: *mut u8) -> c_int;

    pub fn wx_device_alloc(ctx: *mut WxContext, device_index: c_int, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn wx_device_free(ctx: *mut WxContext, device_index: c_int, ptr: *mut c_void) -> c_int;
    pub fn wx_host_alloc_pinned(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn wx_host_free_pinned(ptr: *mut c_void) -> c_int;
    pub fn wx_memcpy_d2h(
        ctx: *mut WxContext,
        device_index: c_int,
        dst_host: *mut c_void,
        src_dev: *const c_void,
        bytes: usize,
        stream: *mut c_void,
    ) -> c_int;
    pub fn wx_stream_synchronize(ctx: *mut WxContext, device_index: c_int, stream: *mut c_void) -> c_int;
    pub fn wx_ipc_export(ctx: *mut WxContext, device_index: c_int, ptr: *mut c_void, handle_out: *mut u8) -> c_int;
    pub fn wx_ipc_open(ctx: *mut WxContext, device_index: c_int, handle: *const u8, out: *mut *mut c_void) -> c_int;
    pub fn wx_ipc_close(ctx: *mut WxContext, device_index: c_int, ptr: *mut c_void) -> c_int;
}

// layout guards: the same numbers tests/c/abi_smoke.c asserts on the C side
const _: () = {
    assert!(core::mem::size_of::<WxState>() == 256);
    assert!(core::mem::size_of::<WxShard>() == 16);
    assert!(core::mem::size_of::<WxRenderInfo>() == 24);
    assert!(core::mem::size_of::<WxSdfInfo>() == 24);
};
