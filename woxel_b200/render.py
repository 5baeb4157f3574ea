"""`src/render` of the reference for the raycast path, over the C ABI.

  ComputeState.build      compute_state.rs:87-131
  RenderMode, SunSettings egui_dev.rs:11-18, :348-368
  Context / Tree          the device objects behind wx_init / wx_tree_upload
  Renderer                WgpuContext::{new, change_vdb_model, render} (wgpu_context.rs:33, :506, :207)
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _ffi
from .scene import Camera, Scene
from .vdb import VDB345, FlatTree, VdbReader


class WxError(RuntimeError):
    def __init__(self, status: int, detail: str = ""):
        text = _ffi.cuda_lib().wx_strerror(status).decode()
        super().__init__(f"[{status}] {text}" + (f": {detail}" if detail else ""))
        self.status = status


class RenderMode(enum.IntEnum):
    Gray = 0
    Rgb = 1
    Ray = 2
    Diffuse = 3
    Glossy = 4


class SunSettings:
    """egui_dev.rs:348-368."""

    def __init__(self):
        d = (C.c_float * 3)()
        c = (C.c_float * 3)()
        i = C.c_float()
        _ffi.host_lib().wxh_default_sun(d, c, C.byref(i))
        self.dir3 = list(d)
        self.color = list(c)
        self.intensity = float(i.value)


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class ComputeState(_ffi.WxState):
    """The 256-byte uniform; `build` follows compute_state.rs:87-131."""

    @classmethod
    def build(cls, camera: Camera, resolution_width: float, render_mode=RenderMode.Diffuse, show_grid=(False, False, False),
              sun_dir3=None, sun_color3=None, sun_intensity=None) -> "ComputeState":
        sun = SunSettings()
        s = cls()
        rc = _ffi.host_lib().wxh_compute_state_build(
            _f3(camera.eye), _f3(camera.target), _f3(camera.up), float(camera.aspect), float(camera.fovy),
            float(resolution_width), int(render_mode), (C.c_uint32 * 3)(*[1 if g else 0 for g in show_grid]),
            _f3(sun_dir3 if sun_dir3 is not None else sun.dir3), _f3(sun_color3 if sun_color3 is not None else sun.color),
            float(sun_intensity if sun_intensity is not None else sun.intensity), C.byref(s))
        if rc != 0:
            raise WxError(rc, _ffi.host_lib().wxh_last_error().decode())
        return s


AOV_SPEC = (("state", np.uint8, ()), ("voxel", np.int32, (3,)), ("leaf", np.int32, ()), ("level", np.uint8, ()),
            ("iters", np.uint32, ()), ("depth", np.float32, ()), ("mask", np.uint8, ()), ("pos", np.float32, (3,)))


class Context:
    """wx_init / wx_shutdown."""

    def __init__(self, n_devices: int = 0, device_ids=None):
        self._h = C.c_void_p()
        ids = (C.c_int * len(device_ids))(*device_ids) if device_ids else None
        rc = _ffi.cuda_lib().wx_init(int(n_devices), ids, C.byref(self._h))
        if rc != 0:
            raise WxError(rc, _ffi.cuda_lib().wx_last_error(None).decode())

    def close(self):
        if getattr(self, "_h", None):
            _ffi.cuda_lib().wx_shutdown(self._h)
            self._h = None

    __del__ = close

    def check(self, rc: int):
        if rc != 0:
            raise WxError(rc, _ffi.cuda_lib().wx_last_error(self._h).decode())

    @property
    def device_count(self) -> int:
        return int(_ffi.cuda_lib().wx_device_count(self._h))

    def set_option(self, option: int, value: int) -> None:
        """wx_set_option (WX_OPT_* of include/woxel_b200.h / woxel_b200._ffi)."""
        self.check(_ffi.cuda_lib().wx_set_option(self._h, int(option), int(value)))

    def get_option(self, option: int) -> int:
        v = C.c_int64(0)
        self.check(_ffi.cuda_lib().wx_get_option(self._h, int(option), C.byref(v)))
        return int(v.value)

    def upload(self, flat_or_desc) -> "Tree":
        return Tree(self, flat_or_desc)

    def build(self, flat_or_desc) -> "Tree":
        """wx_tree_build: compute_sdf on the GPU + device tables in one call (the flat tree's distances are ignored).
        The tree's `sdf` attribute holds the WxSdfInfo.  Raises WxError(-6) when a leaf distance exceeds 255."""
        return Tree(self, flat_or_desc, build=True)

    def last_render_info(self) -> _ffi.WxRenderInfo:
        info = _ffi.WxRenderInfo()
        self.check(_ffi.cuda_lib().wx_last_render_info(self._h, C.byref(info)))
        return info

    # ---- frames -------------------------------------------------------------------------------
    def render(self, tree: "Tree", states, width: int, height: int, aov: bool = False, out: np.ndarray | None = None):
        """wx_render: host buffers in, host buffers out.  Returns (rgba[n,H,W,4], aov dict | None)."""
        states = list(states) if isinstance(states, (list, tuple)) else [states]
        n = len(states)
        arr = (_ffi.WxState * n)(*states)
        rgba = out if out is not None else np.empty((n, height, width, 4), np.uint8)
        assert rgba.flags.c_contiguous and rgba.nbytes == n * height * width * 4
        aovs, a = None, None
        if aov:
            aovs = {k: np.zeros((n, height, width) + shp, dt) for k, dt, shp in AOV_SPEC}
            a = _ffi.WxAov(*[aovs[k].ctypes.data for k, _, _ in AOV_SPEC])
        self.check(_ffi.cuda_lib().wx_render(self._h, tree._h, arr, n, width, height, rgba.ctypes.data,
                                             C.byref(a) if a is not None else None))
        return rgba, aovs

    def render_to(self, tree: "Tree", states, width: int, height: int, dst_ptr: int) -> None:
        """wx_render with a raw destination address: host memory or device memory of any GPU (also an IPC-mapped buffer of
        another process's GPU).  Blocking; the frame is delivered chunk by chunk while later chunks render."""
        states = list(states) if isinstance(states, (list, tuple)) else [states]
        arr = (_ffi.WxState * len(states))(*states)
        self.check(_ffi.cuda_lib().wx_render(self._h, tree._h, arr, len(states), width, height, C.c_void_p(dst_ptr), None))

    def render_shard_to(self, tree: "Tree", states, width: int, height: int, shard: tuple, dst_ptr: int) -> None:
        """wx_render_shard: the rows of shard (index, count) of every frame, delivered to the frame stack at the raw address
        dst_ptr (host memory or device memory of any GPU).  Blocking."""
        states = list(states) if isinstance(states, (list, tuple)) else [states]
        arr = (_ffi.WxState * len(states))(*states)
        sh = _ffi.WxShard(shard[0], shard[1], 8, 0)
        self.check(_ffi.cuda_lib().wx_render_shard(self._h, tree._h, arr, len(states), width, height, C.byref(sh), C.c_void_p(dst_ptr)))

    def compute_sdf(self, flat_or_desc, narrow_leaves: bool = False):
        """wx_compute_sdf: VDB345::compute_sdf (vdb345.rs:290-628) on the GPU for a flat tree (its tile / voxel
        distances are ignored on input).  Returns (tab5, tab4, tab3, WxSdfInfo) in the layout of FlatTree."""
        d = flat_or_desc.desc if hasattr(flat_or_desc, "desc") else flat_or_desc
        tab5 = np.empty((d.n5, 32768), np.uint32)
        tab4 = np.empty((d.n4, 4096), np.uint32)
        tab3 = np.empty((d.n3, 512), np.uint8 if narrow_leaves else np.uint32)
        info = _ffi.WxSdfInfo()
        self.check(_ffi.cuda_lib().wx_compute_sdf(self._h, C.byref(d), tab5.ctypes.data, tab4.ctypes.data, tab3.ctypes.data,
                                                  tab3.dtype.itemsize, C.byref(info)))
        return tab5, tab4, tab3, info

    def capture_srgb(self, n_states: int, width: int, height: int) -> np.ndarray:
        """wx_capture_srgb: the frame(s) of the last render() as RGB8 after the reference's linear_to_srgb
        (recorder.rs:20-37, :132-140).  Returns uint8 [n, H, W, 3]."""
        rgb = np.empty((n_states, height, width, 3), np.uint8)
        self.check(_ffi.cuda_lib().wx_capture_srgb(self._h, n_states, width, height, rgb.ctypes.data))
        return rgb

    def render_device(self, tree: "Tree", states, width: int, height: int, rgba_ptr: int, aov_ptrs: dict | None = None,
                      shard: tuple | None = None, stream: int = 0, device_index: int = 0):
        """wx_render_device: asynchronous, device-resident output (pointers are raw device addresses)."""
        states = list(states) if isinstance(states, (list, tuple)) else [states]
        n = len(states)
        arr = (_ffi.WxState * n)(*states)
        a = None
        if aov_ptrs:
            a = _ffi.WxAov(*[aov_ptrs.get(k, None) for k, _, _ in AOV_SPEC])
        sh = _ffi.WxShard(shard[0], shard[1], shard[2], 0) if shard else None
        self.check(_ffi.cuda_lib().wx_render_device(self._h, device_index, tree._h, arr, n, width, height, rgba_ptr,
                                                    C.byref(a) if a is not None else None,
                                                    C.byref(sh) if sh is not None else None, stream))


class Tree:
    """wx_tree_upload (or wx_tree_build) / wx_tree_free."""

    def __init__(self, ctx: Context, flat_or_desc, build: bool = False):
        self._ctx = ctx
        self._keep = flat_or_desc  # keeps the host arrays alive during the call
        desc = flat_or_desc.desc if isinstance(flat_or_desc, FlatTree) else flat_or_desc
        self._h = C.c_void_p()
        self.sdf = None
        if build:
            self.sdf = _ffi.WxSdfInfo()
            ctx.check(_ffi.cuda_lib().wx_tree_build(ctx._h, C.byref(desc), C.byref(self._h), C.byref(self.sdf)))
        else:
            ctx.check(_ffi.cuda_lib().wx_tree_upload(ctx._h, C.byref(desc), C.byref(self._h)))
        self._keep = None
        self.info = _ffi.WxTreeInfo()
        ctx.check(_ffi.cuda_lib().wx_tree_info(self._h, C.byref(self.info)))

    def free(self):
        if getattr(self, "_h", None) and getattr(self._ctx, "_h", None):
            _ffi.cuda_lib().wx_tree_free(self._ctx._h, self._h)
        self._h = None

    __del__ = free


def make_desc(origins, kids5, vals5, tab5, kids4, vals4, tab4, vals3, tab3):
    """Build a WxTreeDesc from numpy arrays (kept alive on the returned object)."""
    arrs = {
        "origins": np.ascontiguousarray(origins, np.int32), "kids5": np.ascontiguousarray(kids5, np.uint64),
        "vals5": np.ascontiguousarray(vals5, np.uint64), "tab5": np.ascontiguousarray(tab5, np.uint32),
        "kids4": np.ascontiguousarray(kids4, np.uint64), "vals4": np.ascontiguousarray(vals4, np.uint64),
        "tab4": np.ascontiguousarray(tab4, np.uint32), "vals3": np.ascontiguousarray(vals3, np.uint64),
    }
    t3 = np.ascontiguousarray(tab3)
    if t3.dtype not in (np.uint8, np.uint32):
        t3 = t3.astype(np.uint32)
    arrs["tab3"] = t3
    d = _ffi.WxTreeDesc()
    d.n5 = arrs["origins"].size // 3
    d.n4 = arrs["kids4"].size // 64
    d.n3 = arrs["vals3"].size // 8
    for k, a in arrs.items():
        setattr(d, k, a.ctypes.data if a.size else None)
    d.tab3_elem_bytes = t3.dtype.itemsize
    d._keepalive = arrs
    return d


class Renderer:
    """WgpuContext of the new build: owns the device context, the uploaded model and the render options."""

    def __init__(self, width: int, height: int, n_devices: int = 0):
        self.width, self.height = width, height
        self._h = C.c_void_p()
        rc = _ffi.host_lib().wxh_renderer_new(width, height, n_devices, C.byref(self._h))
        if rc != 0:
            raise WxError(rc, _ffi.host_lib().wxh_last_error().decode())
        self.render_mode = RenderMode.Diffuse  # egui_dev.rs:59
        self.show_grid = [False, False, False]
        self.sun_settings = SunSettings()

    def __del__(self):
        if getattr(self, "_h", None):
            _ffi.host_lib().wxh_renderer_free(self._h)
            self._h = None

    def _check(self, rc):
        if rc != 0:
            raise WxError(rc, _ffi.host_lib().wxh_last_error().decode())

    def change_vdb_model(self, vdb_or_path, grid: str | None = None, compute_sdf: bool = True, sdf_on_gpu: bool = True):
        """wgpu_context.rs:506-573.  compute_sdf runs on the GPU (identical values) unless sdf_on_gpu is False."""
        _ffi.host_lib().wxh_renderer_set_sdf_on_gpu(self._h, 1 if sdf_on_gpu else 0)
        if isinstance(vdb_or_path, VDB345):
            self._check(_ffi.host_lib().wxh_renderer_change_vdb_model(self._h, vdb_or_path._h, 1 if compute_sdf else 0))
        else:
            self._check(_ffi.host_lib().wxh_renderer_change_vdb_model_file(self._h, str(vdb_or_path).encode(), grid.encode()))

    @property
    def last_sdf(self) -> _ffi.WxSdfInfo:
        info = _ffi.WxSdfInfo()
        _ffi.host_lib().wxh_renderer_last_sdf(self._h, C.byref(info))
        return info

    def render(self, scene: Scene) -> np.ndarray:
        """wgpu_context.rs:207-292: one frame, rgba8 [H, W, 4]."""
        s = self.sun_settings
        self._check(_ffi.host_lib().wxh_renderer_set_options(
            self._h, int(self.render_mode), (C.c_uint32 * 3)(*[1 if g else 0 for g in self.show_grid]), _f3(s.dir3),
            _f3(s.color), float(s.intensity)))
        out = np.empty((self.height, self.width, 4), np.uint8)
        c = scene.camera
        self._check(_ffi.host_lib().wxh_renderer_render(self._h, _f3(c.eye), _f3(c.target), _f3(c.up), float(c.aspect),
                                                        float(c.fovy), out.ctypes.data))
        return out


def _dump(fn_name: str, path: str, rgb: np.ndarray) -> None:
    rgb = np.ascontiguousarray(rgb, np.uint8)
    if rgb.ndim != 3 or rgb.shape[2] != 3:
        raise ValueError("expected an RGB8 frame [H, W, 3]")
    rc = getattr(_ffi.host_lib(), fn_name)(path.encode(), rgb.ctypes.data, rgb.shape[1], rgb.shape[0])
    if rc != 0:
        raise WxError(rc, _ffi.host_lib().wxh_last_error().decode())


def write_ppm(path: str, rgb: np.ndarray) -> None:
    """Binary PPM (P6) of an RGB8 frame [H, W, 3]: the recorder's frame dump without ffmpeg (recorder.rs:67-105)."""
    _dump("wxh_write_ppm", path, rgb)


def write_png(path: str, rgb: np.ndarray) -> None:
    """PNG (8-bit RGB, sRGB chunk) of an RGB8 frame [H, W, 3] (woxel_b200/host/frame_dump.cpp)."""
    _dump("wxh_write_png", path, rgb)


__all__ = ["WxError", "RenderMode", "SunSettings", "ComputeState", "Context", "Tree", "Renderer", "make_desc",
           "VdbReader", "AOV_SPEC", "write_ppm", "write_png"]
