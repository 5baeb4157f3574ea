"""ctypes declarations for libwoxel_b200.so (include/woxel_b200.h) and libwoxel_host.so (include/woxel_host.h).

The libraries are built in-tree by `__graft_entry__.build()` (or `make -C woxel_b200/csrc && make -C woxel_b200/host`).
Loading fails loudly when they are missing -- there is no Python or CPU fallback for the render path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# WOXEL_B200_LIB: load another build of the same library (kernel experiments, tools/variants.sh)
CUDA_LIB_PATH = os.environ.get("WOXEL_B200_LIB") or os.path.join(_HERE, "libwoxel_b200.so")
HOST_LIB_PATH = os.environ.get("WOXEL_HOST_LIB") or os.path.join(_HERE, "libwoxel_host.so")  # override: sanitizer builds (tools/asan_host.sh)


class WxTreeDesc(C.Structure):
    _fields_ = [
        ("n5", C.c_uint32), ("n4", C.c_uint32), ("n3", C.c_uint32),
        ("origins", C.c_void_p), ("kids5", C.c_void_p), ("vals5", C.c_void_p), ("tab5", C.c_void_p),
        ("kids4", C.c_void_p), ("vals4", C.c_void_p), ("tab4", C.c_void_p), ("vals3", C.c_void_p),
        ("tab3", C.c_void_p), ("tab3_elem_bytes", C.c_uint32), ("reserved", C.c_uint32),
    ]


class WxState(C.Structure):
    """== ComputeState (compute_state.rs:9-29), 256 bytes."""
    _fields_ = [
        ("view_proj", C.c_float * 16), ("camera_to_world", C.c_float * 16), ("eye", C.c_float * 4),
        ("u", C.c_float * 4), ("mv", C.c_float * 4), ("wp", C.c_float * 4),
        ("render_mode", C.c_uint32 * 4), ("show_345", C.c_uint32 * 4),
        ("sun_dir", C.c_float * 4), ("sun_color", C.c_float * 4),
    ]


assert C.sizeof(WxState) == 256


class WxAov(C.Structure):
    _fields_ = [
        ("state", C.c_void_p), ("voxel", C.c_void_p), ("leaf", C.c_void_p), ("level", C.c_void_p),
        ("iters", C.c_void_p), ("depth", C.c_void_p), ("mask", C.c_void_p), ("pos", C.c_void_p),
    ]


class WxShard(C.Structure):
    _fields_ = [("index", C.c_uint32), ("count", C.c_uint32), ("band_rows", C.c_uint32), ("reserved", C.c_uint32)]


class WxTreeInfo(C.Structure):
    _fields_ = [
        ("n5", C.c_uint32), ("n4", C.c_uint32), ("n3", C.c_uint32), ("leaf_bits", C.c_uint32),
        ("device_bytes", C.c_uint64), ("max_dist", C.c_uint32 * 3), ("n_devices", C.c_uint32),
    ]


class WxSdfInfo(C.Structure):
    _fields_ = [("max_dist", C.c_uint32 * 3), ("rounds", C.c_uint32), ("device_ms", C.c_float), ("total_ms", C.c_float)]


class WxRenderInfo(C.Structure):
    _fields_ = [("kernel_ms", C.c_float), ("total_ms", C.c_float), ("rays", C.c_uint64), ("launches", C.c_uint32),
                ("reserved", C.c_uint32)]


class WxhVdbInfo(C.Structure):
    _fields_ = [
        ("file_version", C.c_uint32), ("library_major", C.c_uint32), ("library_minor", C.c_uint32),
        ("grid_count", C.c_uint32), ("grid_compression", C.c_uint32), ("is_half_float", C.c_int32),
        ("file_voxel_count", C.c_int64), ("grid_pos", C.c_uint64), ("block_pos", C.c_uint64), ("end_pos", C.c_uint64),
    ]


# name -> (restype, argtypes); doubles as the list the symbol-export test checks against the headers
vp, u8p = C.c_void_p, C.POINTER(C.c_uint8)
CUDA_API = {
    "wx_abi_version": (C.c_int, []),
    "wx_strerror": (C.c_char_p, [C.c_int]),
    "wx_last_error": (C.c_char_p, [vp]),
    "wx_init": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(vp)]),
    "wx_shutdown": (C.c_int, [vp]),
    "wx_device_count": (C.c_int, [vp]),
    "wx_tree_upload": (C.c_int, [vp, C.POINTER(WxTreeDesc), C.POINTER(vp)]),
    "wx_tree_free": (C.c_int, [vp, vp]),
    "wx_tree_info": (C.c_int, [vp, C.POINTER(WxTreeInfo)]),
    "wx_render": (C.c_int, [vp, vp, C.POINTER(WxState), C.c_uint32, C.c_uint32, C.c_uint32, vp, C.POINTER(WxAov)]),
    "wx_render_device": (C.c_int, [vp, C.c_int, vp, C.POINTER(WxState), C.c_uint32, C.c_uint32, C.c_uint32, vp,
                                   C.POINTER(WxAov), C.POINTER(WxShard), vp]),
    "wx_render_shard": (C.c_int, [vp, vp, C.POINTER(WxState), C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(WxShard), vp]),
    "wx_last_render_info": (C.c_int, [vp, C.POINTER(WxRenderInfo)]),
    "wx_device_alloc": (C.c_int, [vp, C.c_int, C.c_size_t, C.POINTER(vp)]),
    "wx_device_free": (C.c_int, [vp, C.c_int, vp]),
    "wx_host_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "wx_host_free_pinned": (C.c_int, [vp]),
    "wx_memcpy_d2h": (C.c_int, [vp, C.c_int, vp, vp, C.c_size_t, vp]),
    "wx_stream_synchronize": (C.c_int, [vp, C.c_int, vp]),
    "wx_ipc_export": (C.c_int, [vp, C.c_int, vp, u8p]),
    "wx_ipc_open": (C.c_int, [vp, C.c_int, u8p, C.POINTER(vp)]),
    "wx_ipc_close": (C.c_int, [vp, C.c_int, vp]),
    "wx_shard_rows": (C.c_int, [C.c_uint32, C.POINTER(WxShard), vp]),
    "wx_tree_build": (C.c_int, [vp, C.POINTER(WxTreeDesc), C.POINTER(vp), C.POINTER(WxSdfInfo)]),
    "wx_compute_sdf": (C.c_int, [vp, C.POINTER(WxTreeDesc), vp, vp, vp, C.c_uint32, C.POINTER(WxSdfInfo)]),
    "wx_capture_srgb": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp]),
    "wx_srgb_table": (C.c_int, [vp]),
    "wx_set_option": (C.c_int, [vp, C.c_int, C.c_int64]),
    "wx_get_option": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int64)]),
}
# WxOption (include/woxel_b200.h)
WX_OPT_MARCH, WX_OPT_KERNEL, WX_OPT_RENDER_CHUNKS, WX_OPT_SMEM_PAD, WX_OPT_NVTX, WX_OPT_LONG_FIRST, WX_OPT_LONG_THRESHOLD = 1, 2, 3, 4, 5, 6, 7
f3, u3, i3 = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
HOST_API = {
    "wxh_last_error": (C.c_char_p, []),
    "wxh_write_ppm": (C.c_int, [C.c_char_p, vp, C.c_uint32, C.c_uint32]),
    "wxh_write_png": (C.c_int, [C.c_char_p, vp, C.c_uint32, C.c_uint32]),
    "wxh_global_to_node": (C.c_int, [C.c_int, i3, i3]),
    "wxh_global_to_offset": (C.c_int64, [C.c_int, i3]),
    "wxh_offset_to_child": (C.c_int, [C.c_int, C.c_uint32, u3]),
    "wxh_child_to_offset": (C.c_int64, [C.c_int, u3]),
    "wxh_vdb_new": (vp, []),
    "wxh_vdb_free": (None, [vp]),
    "wxh_vdb_set_voxel": (None, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_uint32]),
    "wxh_vdb_set_voxels": (None, [vp, vp, C.c_size_t, C.c_uint32]),
    "wxh_vdb_get_voxel": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, u3, C.POINTER(C.c_int)]),
    "wxh_vdb_count_nodes": (None, [vp, C.POINTER(C.c_uint64)]),
    "wxh_vdb_count_leaf_values": (C.c_uint64, [vp]),
    "wxh_vdb_compute_sdf": (None, [vp]),
    "wxh_blosc_decompress": (C.c_int, [C.c_char_p, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "wxh_vdb_read": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(vp), C.POINTER(WxhVdbInfo)]),
    "wxh_vdb_to_flat": (vp, [vp, C.c_int]),
    "wxh_flat_free": (None, [vp]),
    "wxh_flat_desc": (None, [vp, C.POINTER(WxTreeDesc)]),
    "wxh_build_sphere": (vp, [C.c_int32, C.c_double, C.c_double]),
    "wxh_build_torus": (vp, [C.c_int32, C.c_double, C.c_double, C.c_double]),
    "wxh_build_fog": (vp, [C.c_int32, C.c_double, C.POINTER(C.c_double)]),
    "wxh_compute_state_build": (C.c_int, [f3, f3, f3, C.c_float, C.c_float, C.c_float, C.c_uint32, u3, f3, f3, C.c_float,
                                          C.POINTER(WxState)]),
    "wxh_default_sun": (None, [f3, f3, C.POINTER(C.c_float)]),
    "wxh_renderer_new": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, C.POINTER(vp)]),
    "wxh_renderer_free": (None, [vp]),
    "wxh_renderer_change_vdb_model": (C.c_int, [vp, vp, C.c_int]),
    "wxh_renderer_change_vdb_model_file": (C.c_int, [vp, C.c_char_p, C.c_char_p]),
    "wxh_renderer_set_options": (C.c_int, [vp, C.c_uint32, u3, f3, f3, C.c_float]),
    "wxh_renderer_render": (C.c_int, [vp, f3, f3, f3, C.c_float, C.c_float, vp]),
    "wxh_renderer_set_sdf_on_gpu": (C.c_int, [vp, C.c_int]),
    "wxh_renderer_last_sdf": (None, [vp, C.POINTER(WxSdfInfo)]),
    "wxh_flat_compute_sdf_gpu": (C.c_int, [vp, vp, C.POINTER(WxSdfInfo)]),
    "wxh_renderer_context": (vp, [vp]),
    "wxh_renderer_tree": (vp, [vp]),
}


def _load(path: str, api: dict) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  woxel_b200 has no CPU fallback for the render path.")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in api.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    return lib


_cuda = None
_host = None


def cuda_lib_path() -> str:
    """Path of the CUDA library this process loads (WOXEL_B200_LIB selects a build-time variant for A/B runs)."""
    return CUDA_LIB_PATH


def cuda_lib() -> C.CDLL:
    global _cuda
    if _cuda is None:
        _cuda = _load(CUDA_LIB_PATH, CUDA_API)
    return _cuda


def host_lib() -> C.CDLL:
    global _host
    if _host is None:
        cuda_lib()  # libwoxel_host.so links against it
        _host = _load(HOST_LIB_PATH, HOST_API)
    return _host
