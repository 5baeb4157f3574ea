"""ctypes binding of the CPU oracle (oracle/libwxo.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (woxel_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
# WXO_VARIANT=fma loads the FMA-contracted sensitivity build (oracle/Makefile); the default is the strict-IEEE oracle
# WXO_VARIANT=asan loads the AddressSanitizer/UBSan build (tools/asan_host.sh)
_VARIANT = {"fma": "libwxo_fma.so", "asan": "libwxo_asan.so"}.get(os.environ.get("WXO_VARIANT", ""), "libwxo.so")
LIB_PATH = os.path.join(ORACLE_DIR, _VARIANT)

EP_OFFS, EP_LEAF, EP_INNR5, EP_INNR4, EP_ROOT, EP_BKGR = range(6)


def build_oracle(force: bool = False) -> str:
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h")) or f == "Makefile"]
    stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", ORACLE_DIR, _VARIANT], check=True, capture_output=True)
    return LIB_PATH


class VdbInfo(C.Structure):
    _fields_ = [
        ("file_version", C.c_uint32), ("library_major", C.c_uint32), ("library_minor", C.c_uint32),
        ("grid_count", C.c_uint32), ("grid_compression", C.c_uint32), ("is_half_float", C.c_int32),
        ("file_voxel_count", C.c_int64), ("grid_pos", C.c_uint64), ("block_pos", C.c_uint64),
        ("end_pos", C.c_uint64), ("topology_end_pos", C.c_uint64), ("root_tiles", C.c_uint32),
        ("root_nodes", C.c_uint32),
    ]


class State(C.Structure):
    """The 256-byte uniform (compute_state.rs:9-29)."""
    _fields_ = [
        ("view_proj", C.c_float * 16), ("camera_to_world", C.c_float * 16), ("eye", C.c_float * 4),
        ("u", C.c_float * 4), ("mv", C.c_float * 4), ("wp", C.c_float * 4),
        ("render_mode", C.c_uint32 * 4), ("show_345", C.c_uint32 * 4),
        ("sun_dir", C.c_float * 4), ("sun_color", C.c_float * 4),
    ]

    def to_bytes(self) -> bytes:
        return bytes(self)


class Aov(C.Structure):
    _fields_ = [
        ("state", C.c_void_p), ("voxel", C.c_void_p), ("leaf", C.c_void_p), ("level", C.c_void_p),
        ("iters", C.c_void_p), ("depth", C.c_void_p), ("mask", C.c_void_p), ("pos", C.c_void_p),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64), ("primary_rays", C.c_uint64), ("lookups", C.c_uint64 * 4),
        ("primary_lookups", C.c_uint64 * 4), ("alg_bytes", C.c_uint64), ("primary_alg_bytes", C.c_uint64),
        ("max_iters", C.c_uint32), ("hit", C.c_uint64), ("oob", C.c_uint64), ("maxed", C.c_uint64),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build_oracle()
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
    u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
    u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
    L.wxo_global_to_node.argtypes = [C.c_int, i32p, i32p]
    L.wxo_global_to_offset.argtypes = [C.c_int, i32p]
    L.wxo_global_to_offset.restype = C.c_uint32
    L.wxo_offset_to_child.argtypes = [C.c_int, C.c_uint32, u32p]
    L.wxo_child_to_offset.argtypes = [C.c_int, u32p]
    L.wxo_child_to_offset.restype = C.c_uint32
    L.wxo_tree_new.restype = vp
    L.wxo_tree_free.argtypes = [vp]
    L.wxo_set_voxel.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_uint32]
    L.wxo_set_voxels.argtypes = [vp, i32p, C.c_size_t, C.c_uint32]
    L.wxo_get_voxel.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]
    L.wxo_get_voxel.restype = C.c_int
    L.wxo_count_nodes.argtypes = [vp, u64p]
    L.wxo_count_leaf_values.argtypes = [vp]
    L.wxo_count_leaf_values.restype = C.c_uint64
    L.wxo_compute_sdf.argtypes = [vp]
    L.wxo_tree_from_topology.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, i32p, u64p, u64p, u64p, u64p, u64p]
    L.wxo_tree_from_topology.restype = vp
    L.wxo_vdb_read.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(vp), C.POINTER(VdbInfo)]
    L.wxo_vdb_read.restype = C.c_int
    L.wxo_serialise.argtypes = [vp]
    L.wxo_serialise.restype = vp
    L.wxo_gpudata_from_tables.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, i32p, u64p, u64p, u32p, u64p, u64p, u32p, u64p, u32p]
    L.wxo_gpudata_from_tables.restype = vp
    L.wxo_gpudata_free.argtypes = [vp]
    L.wxo_gpudata_counts.argtypes = [vp, u32p, u32p]
    L.wxo_gpudata_origins.argtypes = [vp]
    L.wxo_gpudata_origins.restype = C.POINTER(C.c_int32)
    L.wxo_gpudata_mask.argtypes = [vp, C.c_int]
    L.wxo_gpudata_mask.restype = C.POINTER(C.c_uint32)
    L.wxo_gpudata_atlas.argtypes = [vp, C.c_int]
    L.wxo_gpudata_atlas.restype = C.POINTER(C.c_uint32)
    L.wxo_gpudata_tables.argtypes = [vp, u32p, u32p, u32p]
    L.wxo_compute_state_build.argtypes = [f32p, f32p, f32p, C.c_float, C.c_float, C.c_float, C.c_uint32, u32p,
                                          f32p, f32p, C.c_float, C.POINTER(State)]
    L.wxo_default_sun.argtypes = [f32p, f32p, C.POINTER(C.c_float)]
    L.wxo_render.argtypes = [vp, C.POINTER(State), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp,
                             C.POINTER(Aov), C.c_int, C.POINTER(Stats)]
    _lib = L
    return L


def _i3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.int32).reshape(3))


def global_to_node(level, g):
    out = np.zeros(3, np.int32)
    lib().wxo_global_to_node(level, _i3(g), out)
    return out.tolist()


def global_to_offset(level, g):
    return int(lib().wxo_global_to_offset(level, _i3(g)))


def offset_to_child(level, off):
    out = np.zeros(3, np.uint32)
    lib().wxo_offset_to_child(level, off, out)
    return out.tolist()


def child_to_offset(level, c):
    return int(lib().wxo_child_to_offset(level, np.ascontiguousarray(np.asarray(c, np.uint32))))


class GpuData:
    """What vdb.origins()/masks()/atlas() hand to the reference shader."""

    def __init__(self, handle):
        self._h = handle
        n = np.zeros(3, np.uint32)
        d = np.zeros(3, np.uint32)
        lib().wxo_gpudata_counts(self._h, n, d)
        self.n5, self.n4, self.n3 = (int(x) for x in n)
        self.atlas_dim = [int(x) for x in d]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().wxo_gpudata_free(self._h)
            self._h = None

    @property
    def origins(self) -> np.ndarray:
        p = lib().wxo_gpudata_origins(self._h)
        return np.ctypeslib.as_array(p, shape=(self.n5, 4)).copy() if self.n5 else np.zeros((0, 4), np.int32)

    def mask(self, which: int) -> np.ndarray:
        """0 kids5, 1 vals5, 2 kids4, 3 vals4, 4 vals3 -- as the u32 words the shader binds."""
        n, w = [(self.n5, 1024), (self.n5, 1024), (self.n4, 128), (self.n4, 128), (self.n3, 16)][which]
        if n == 0:
            return np.zeros((0, w), np.uint32)
        return np.ctypeslib.as_array(lib().wxo_gpudata_mask(self._h, which), shape=(n, w)).copy()

    def mask64(self, which: int) -> np.ndarray:
        m = self.mask(which)
        return np.ascontiguousarray(m).view(np.uint64)

    def atlas(self, level: int) -> np.ndarray:
        side = [32, 16, 8][level] * self.atlas_dim[level]
        if side == 0:
            return np.zeros((0, 0, 0), np.uint32)
        return np.ctypeslib.as_array(lib().wxo_gpudata_atlas(self._h, level), shape=(side, side, side)).copy()

    def tables(self):
        t5 = np.zeros((self.n5, 32768), np.uint32)
        t4 = np.zeros((self.n4, 4096), np.uint32)
        t3 = np.zeros((self.n3, 512), np.uint32)
        lib().wxo_gpudata_tables(self._h, t5, t4, t3)
        return t5, t4, t3

    def render(self, state: State, width: int, height: int, aov: bool = True, threads: int | None = None,
               rows: tuple[int, int] | None = None):
        """Returns (rgba[H,W,4] u8, aov dict or None, Stats)."""
        threads = threads or (os.cpu_count() or 1)
        rgba = np.zeros((height, width, 4), np.uint8)
        out = None
        a = None
        if aov:
            out = {
                "state": np.full((height, width), 255, np.uint8),
                "voxel": np.zeros((height, width, 3), np.int32),
                "leaf": np.full((height, width), -1, np.int32),
                "level": np.zeros((height, width), np.uint8),
                "iters": np.zeros((height, width), np.uint32),
                "depth": np.zeros((height, width), np.float32),
                "mask": np.zeros((height, width), np.uint8),
                "pos": np.zeros((height, width, 3), np.float32),
            }
            a = Aov(*[out[k].ctypes.data for k in ("state", "voxel", "leaf", "level", "iters", "depth", "mask", "pos")])
        st = Stats()
        y0, y1 = rows if rows else (0, height)
        lib().wxo_render(self._h, C.byref(state), width, height, y0, y1, rgba.ctypes.data,
                         C.byref(a) if a is not None else None, threads, C.byref(st))
        return rgba, out, st


class Tree:
    """The reference's VDB345<u32>."""

    def __init__(self, handle=None):
        self._h = handle if handle is not None else lib().wxo_tree_new()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().wxo_tree_free(self._h)
            self._h = None

    @classmethod
    def read(cls, path: str, grid: str):
        h = C.c_void_p()
        info = VdbInfo()
        rc = lib().wxo_vdb_read(path.encode(), grid.encode(), C.byref(h), C.byref(info))
        if rc != 0:
            raise IOError(f"oracle reader failed with {rc} on {path}:{grid}")
        return cls(h), info

    @classmethod
    def from_topology(cls, origins, kids5, vals5, kids4, vals4, vals3):
        origins = np.ascontiguousarray(np.asarray(origins, np.int32).reshape(-1, 3))
        arrs = [np.ascontiguousarray(np.asarray(a, np.uint64)).reshape(-1) for a in (kids5, vals5, kids4, vals4, vals3)]
        n5, n4, n3 = len(origins), arrs[2].size // 64, arrs[4].size // 8
        h = lib().wxo_tree_from_topology(n5, n4, n3, origins.reshape(-1), *arrs)
        if not h:
            raise ValueError("inconsistent topology")
        return cls(h)

    def set_voxel(self, p, v=1):
        lib().wxo_set_voxel(self._h, int(p[0]), int(p[1]), int(p[2]), int(v))

    def set_voxels(self, xyz, v=1):
        xyz = np.ascontiguousarray(np.asarray(xyz, np.int32).reshape(-1, 3))
        lib().wxo_set_voxels(self._h, xyz.reshape(-1), len(xyz), int(v))

    def get_voxel(self, p):
        val = C.c_uint64()
        kind = lib().wxo_get_voxel(self._h, int(p[0]), int(p[1]), int(p[2]), C.byref(val))
        return kind, int(val.value)

    def count_nodes(self):
        out = np.zeros(3, np.uint64)
        lib().wxo_count_nodes(self._h, out)
        return [int(x) for x in out]

    def count_leaf_values(self) -> int:
        return int(lib().wxo_count_leaf_values(self._h))

    def compute_sdf(self):
        lib().wxo_compute_sdf(self._h)

    def serialise(self) -> GpuData:
        return GpuData(lib().wxo_serialise(self._h))


def gpudata_from_tables(origins, kids5, vals5, tab5, kids4, vals4, tab4, vals3, tab3) -> GpuData:
    origins = np.ascontiguousarray(np.asarray(origins, np.int32).reshape(-1, 3))
    u64 = lambda a: np.ascontiguousarray(np.asarray(a, np.uint64)).reshape(-1)
    u32 = lambda a: np.ascontiguousarray(np.asarray(a, np.uint32)).reshape(-1)
    n5, n4, n3 = len(origins), u64(kids4).size // 64, u64(vals3).size // 8
    h = lib().wxo_gpudata_from_tables(n5, n4, n3, origins.reshape(-1), u64(kids5), u64(vals5), u32(tab5),
                                      u64(kids4), u64(vals4), u32(tab4), u64(vals3), u32(tab3))
    return GpuData(h)


def default_sun():
    d = np.zeros(3, np.float32)
    c = np.zeros(3, np.float32)
    i = C.c_float()
    lib().wxo_default_sun(d, c, C.byref(i))
    return d, c, float(i.value)


def compute_state(eye, target, up=(0, 1, 0), aspect=None, fovy=45.0, width=640, height=480, render_mode=0,
                  show_grid=(0, 0, 0), sun_dir=None, sun_color=None, sun_intensity=None) -> State:
    """ComputeState::build (compute_state.rs:87-131) with the defaults of camera.rs:16-29 / egui_dev.rs:355-367."""
    d, c, i = default_sun()
    f32 = lambda v: np.ascontiguousarray(np.asarray(v, np.float32).reshape(3))
    st = State()
    lib().wxo_compute_state_build(
        f32(eye), f32(target), f32(up), np.float32(aspect if aspect is not None else width / height),
        np.float32(fovy), np.float32(width), int(render_mode),
        np.ascontiguousarray(np.asarray(show_grid, np.uint32)), f32(sun_dir if sun_dir is not None else d),
        f32(sun_color if sun_color is not None else c), np.float32(sun_intensity if sun_intensity is not None else i),
        C.byref(st))
    return st
