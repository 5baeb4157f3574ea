"""woxel_b200 -- B200-native drop-in for woxel's per-pixel VDB345 HDDA+SDF raycast.

Layout:
  csrc/   hand-written sm_100a kernels + the C ABI (libwoxel_b200.so, include/woxel_b200.h)
  host/   C++ host mirroring the reference's src/vdb, src/scene, src/render (libwoxel_host.so)
  vdb.py, scene.py, render.py   thin ctypes views with the reference's names

There is no CPU fallback: rendering without the CUDA library or without a GPU raises.
"""
from . import _ffi  # noqa: F401
from .vdb import VDB345, VdbReader, FlatTree, VdbEndpoint, N3, N4, N5  # noqa: F401
from .scene import Camera, Scene  # noqa: F401
from .render import ComputeState, RenderMode, SunSettings, Renderer, Context, Tree, WxError, write_ppm, write_png  # noqa: F401

__all__ = ["VDB345", "VdbReader", "FlatTree", "VdbEndpoint", "N3", "N4", "N5", "Camera", "Scene", "ComputeState",
           "RenderMode", "SunSettings", "Renderer", "Context", "Tree", "WxError", "write_ppm", "write_png"]
