"""`src/vdb` of the reference as seen from Python: VDB345, VdbReader, index maths, FlatTree.

Each class forwards to the C++ host (woxel_b200/host/vdb.{hpp,cpp}, vdb_read.cpp) through
include/woxel_host.h.  Names and argument meaning follow the Rust items they replace.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _ffi


class VdbError(IOError):
    """ErrorKind of read.rs:31-53."""

    def __init__(self, status: int, text: str):
        super().__init__(f"[{status}] {text}")
        self.status = status


class _NodeMath:
    """The reference's `Node` trait for one level (data_structure.rs:17-92)."""

    def __init__(self, level: int, log2_d: int, total_log2_d: int):
        self.level = level
        self.LOG2_D = log2_d
        self.TOTAL_LOG2_D = total_log2_d
        self.DIM = 1 << log2_d
        self.TOTAL_DIM = 1 << total_log2_d
        self.SIZE = 1 << (3 * log2_d)

    def global_to_node(self, g):
        a = (C.c_int32 * 3)(*[int(v) for v in g])
        out = (C.c_int32 * 3)()
        _ffi.host_lib().wxh_global_to_node(self.level, a, out)
        return list(out)

    def global_to_offset(self, g) -> int:
        a = (C.c_int32 * 3)(*[int(v) for v in g])
        return int(_ffi.host_lib().wxh_global_to_offset(self.level, a))

    def offset_to_child(self, offset: int):
        out = (C.c_uint32 * 3)()
        _ffi.host_lib().wxh_offset_to_child(self.level, int(offset), out)
        return list(out)

    def child_to_offset(self, c) -> int:
        a = (C.c_uint32 * 3)(*[int(v) for v in c])
        return int(_ffi.host_lib().wxh_child_to_offset(self.level, a))


N3 = _NodeMath(3, 3, 3)
N4 = _NodeMath(4, 4, 7)
N5 = _NodeMath(5, 5, 12)


@dataclass
class VdbEndpoint:
    """data_structure.rs:337-344."""
    kind: str  # "Offs" | "Leaf" | "Innr" | "Root" | "Bkgr"
    value: int
    level: int = 0

    KINDS = ("Offs", "Leaf", "Innr", "Root", "Bkgr")


class FlatTree:
    """origins() + masks() + atlas() of the reference, flattened in its DFS node order (vdb345.rs:108-264)."""

    def __init__(self, handle):
        self._h = handle
        self.desc = _ffi.WxTreeDesc()
        _ffi.host_lib().wxh_flat_desc(self._h, C.byref(self.desc))
        self.n5, self.n4, self.n3 = self.desc.n5, self.desc.n4, self.desc.n3

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                _ffi.host_lib().wxh_flat_free(self._h)
            except TypeError:  # interpreter teardown: the module globals are already gone
                pass
            self._h = None

    def _arr(self, ptr, dtype, shape):
        n = int(np.prod(shape))
        if n == 0 or not ptr:
            return np.zeros(shape, dtype)
        buf = (C.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)  # view: valid while self lives

    def compute_sdf_gpu(self, ctx):
        """Fill this flat tree's distances with wx_compute_sdf (take the flat tree BEFORE VDB345.compute_sdf).
        Returns the WxSdfInfo; raises WxError(-6) when a distance does not fit the GPU path."""
        from .render import WxError
        info = _ffi.WxSdfInfo()
        rc = _ffi.host_lib().wxh_flat_compute_sdf_gpu(self._h, ctx._h, C.byref(info))
        if rc != 0:
            raise WxError(rc, _ffi.host_lib().wxh_last_error().decode())
        _ffi.host_lib().wxh_flat_desc(self._h, C.byref(self.desc))  # tab3 may have moved (u8 <-> u32)
        return info

    @property
    def origins(self):
        return self._arr(self.desc.origins, np.int32, (self.n5, 3))

    @property
    def kids5(self):
        return self._arr(self.desc.kids5, np.uint64, (self.n5, 512))

    @property
    def vals5(self):
        return self._arr(self.desc.vals5, np.uint64, (self.n5, 512))

    @property
    def tab5(self):
        return self._arr(self.desc.tab5, np.uint32, (self.n5, 32768))

    @property
    def kids4(self):
        return self._arr(self.desc.kids4, np.uint64, (self.n4, 64))

    @property
    def vals4(self):
        return self._arr(self.desc.vals4, np.uint64, (self.n4, 64))

    @property
    def tab4(self):
        return self._arr(self.desc.tab4, np.uint32, (self.n4, 4096))

    @property
    def vals3(self):
        return self._arr(self.desc.vals3, np.uint64, (self.n3, 8))

    @property
    def tab3(self):
        dt = np.uint8 if self.desc.tab3_elem_bytes == 1 else np.uint32
        return self._arr(self.desc.tab3, dt, (self.n3, 512))


class VDB345:
    """VDB345<u32> (vdb345.rs:12)."""

    def __init__(self, handle=None):
        self._h = handle if handle is not None else _ffi.host_lib().wxh_vdb_new()
        if not self._h:
            raise MemoryError("wxh_vdb_new failed")

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                _ffi.host_lib().wxh_vdb_free(self._h)
            except TypeError:  # interpreter teardown
                pass
            self._h = None

    def set_voxel(self, p, v: int = 1):
        _ffi.host_lib().wxh_vdb_set_voxel(self._h, int(p[0]), int(p[1]), int(p[2]), int(v))

    def set_voxels(self, xyz, v: int = 1):
        a = np.ascontiguousarray(np.asarray(xyz, np.int32).reshape(-1, 3))
        _ffi.host_lib().wxh_vdb_set_voxels(self._h, a.ctypes.data, len(a), int(v))

    def get_voxel(self, p) -> VdbEndpoint:
        val = C.c_uint32()
        lvl = C.c_int()
        k = _ffi.host_lib().wxh_vdb_get_voxel(self._h, int(p[0]), int(p[1]), int(p[2]), C.byref(val), C.byref(lvl))
        return VdbEndpoint(VdbEndpoint.KINDS[k], int(val.value), int(lvl.value))

    def count_nodes(self):
        out = (C.c_uint64 * 3)()
        _ffi.host_lib().wxh_vdb_count_nodes(self._h, out)
        return [int(x) for x in out]

    def count_leaf_values(self) -> int:
        return int(_ffi.host_lib().wxh_vdb_count_leaf_values(self._h))

    def compute_sdf(self):
        _ffi.host_lib().wxh_vdb_compute_sdf(self._h)

    def to_flat(self, narrow_leaves: bool = True) -> FlatTree:
        h = _ffi.host_lib().wxh_vdb_to_flat(self._h, 1 if narrow_leaves else 0)
        if not h:
            raise MemoryError("wxh_vdb_to_flat failed")
        return FlatTree(h)

    # procedural scenes of the benchmark configs (not in the reference)
    @classmethod
    def sphere(cls, half: int = 1024, radius: float = 1000.0, band: float = 3.0):
        return cls(_ffi.host_lib().wxh_build_sphere(half, radius, band))

    @classmethod
    def torus(cls, half: int = 1024, major: float = 700.0, minor: float = 250.0, band: float = 3.0):
        return cls(_ffi.host_lib().wxh_build_torus(half, major, minor, band))

    @classmethod
    def fog(cls, half: int, tau: float):
        occ = C.c_double()
        v = cls(_ffi.host_lib().wxh_build_fog(half, tau, C.byref(occ)))
        v.occupancy = float(occ.value)
        return v


def blosc_decompress(frame: bytes) -> bytes:
    """One c-blosc 1.x frame as the reader meets them inside a .vdb (read.rs:514-533): BloscLZ / LZ4 / Snappy / zlib / Zstd, byte or bit shuffle."""
    need = C.c_size_t(0)
    if len(frame) >= 16:
        need.value = int.from_bytes(frame[4:8], "little")
    buf = C.create_string_buffer(max(need.value, 1))
    n = C.c_size_t(0)
    rc = _ffi.host_lib().wxh_blosc_decompress(frame, len(frame), buf, need.value, C.byref(n))
    if rc != 0:
        raise VdbError(rc, _ffi.host_lib().wxh_last_error().decode())
    return buf.raw[:n.value]


class VdbReader:
    """VdbReader (read.rs:55-141).  `read_vdb345_grid(name)` returns a VDB345."""

    def __init__(self, path: str):
        self.path = path
        self.info = None

    def read_vdb345_grid(self, name: str) -> VDB345:
        h = C.c_void_p()
        info = _ffi.WxhVdbInfo()
        rc = _ffi.host_lib().wxh_vdb_read(self.path.encode(), name.encode(), C.byref(h), C.byref(info))
        if rc != 0:
            raise VdbError(rc, _ffi.host_lib().wxh_last_error().decode())
        self.info = info
        return VDB345(h)
