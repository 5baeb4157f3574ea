"""ctypes view of tests/emu/libwx_emu.so -- TEST INFRASTRUCTURE: the device code of woxel_b200/csrc/wx_device.cuh compiled
for the host (tests/emu/cuda_shim.h) and run lane by lane on the CPU over the tables wx_tree_upload would send to the GPU."""
from __future__ import annotations

import ctypes as C
import functools
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")

STAT_NAMES = ["rays", "warps", "lane_steps", "warp_steps", "root_blocks", "n5_blocks", "n4_blocks", "leaf_blocks", "generic_iters",
              *[f"combo{i}" for i in range(8)], "lane_table_reads", "truncated"]


@functools.lru_cache(maxsize=None)
def lib() -> C.CDLL:
    # WX_EMU_EXTRA="-DWX_..." builds (and loads) the emulation of a build-time variant of the kernel (measurement knobs)
    extra = os.environ.get("WX_EMU_EXTRA", "")
    out = "libwx_emu.so" if not extra else "libwx_emu_" + "".join(ch if ch.isalnum() else "_" for ch in extra) + ".so"
    r = subprocess.run(["make", "-C", EMU_DIR, f"EXTRA={extra}", f"OUT={out}"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"make -C tests/emu failed:\n{r.stdout}\n{r.stderr}")
    L = C.CDLL(os.path.join(EMU_DIR, out))
    L.wxe_render.restype = C.c_int
    L.wxe_render.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32,
                             C.c_void_p]
    L.wxe_queue_sim.restype = C.c_int
    L.wxe_queue_sim.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
    L.wxe_render_cta_queue.restype = C.c_int
    L.wxe_render_cta_queue.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.wxe_chunk_coverage.restype = C.c_int
    L.wxe_chunk_coverage.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
    L.wxe_n_stats.restype = C.c_int
    L.wxe_set_march.restype = None
    L.wxe_set_march.argtypes = [C.c_int]
    assert L.wxe_n_stats() == len(STAT_NAMES)
    return L


class Aov(C.Structure):  # == WxAov (include/woxel_b200.h)
    _fields_ = [(k, C.c_void_p) for k in ("state", "voxel", "leaf", "level", "iters", "depth", "mask", "pos")]


def render(desc, states, width: int, height: int, aov: bool = True, rcp_bump: int = 0, warp_w: int = 4, stats: bool = False,
           march: int = 0):
    """desc: WxTreeDesc (woxel_b200.render.make_desc); states: one or a list of 256-byte ComputeState objects.
    Returns (rgba[n,H,W,4], aov dict of [n,...] arrays or None, stats dict or None)."""
    if not isinstance(states, (list, tuple)):
        states = [states]
    n = len(states)
    buf = (C.c_char * (256 * n))()
    for i, s in enumerate(states):
        b = bytes(s)
        assert len(b) == 256
        buf[256 * i:256 * (i + 1)] = b
    rgba = np.zeros((n, height, width, 4), np.uint8)
    out, a = None, None
    if aov:
        out = {
            "state": np.zeros((n, height, width), np.uint8), "voxel": np.zeros((n, height, width, 3), np.int32),
            "leaf": np.zeros((n, height, width), np.int32), "level": np.zeros((n, height, width), np.uint8),
            "iters": np.zeros((n, height, width), np.uint32), "depth": np.zeros((n, height, width), np.float32),
            "mask": np.zeros((n, height, width), np.uint8), "pos": np.zeros((n, height, width, 3), np.float32),
        }
        a = Aov(*[out[k].ctypes.data for k in ("state", "voxel", "leaf", "level", "iters", "depth", "mask", "pos")])
    st = np.zeros(len(STAT_NAMES), np.uint64) if stats else None
    lib().wxe_set_march(march)  # 1: the tolerance-mode march (WX_OPT_MARCH); a process-wide switch of the emulation library
    try:
        rc = lib().wxe_render(C.addressof(desc), C.addressof(buf), n, width, height, rgba.ctypes.data, C.addressof(a) if a is not None else None,
                              rcp_bump, warp_w, st.ctypes.data if st is not None else None)
    finally:
        lib().wxe_set_march(0)
    if rc != 0:
        raise RuntimeError(f"wxe_render failed: {rc}")
    return rgba, out, (dict(zip(STAT_NAMES, (int(v) for v in st))) if st is not None else None)


def queue_sim(n_chunks: int, n_ctas: int, warps_per_cta: int = 4, seed: int = 1):
    """The ticket protocol of the CTA-level work queue run by real threads.  Returns (counts[n_chunks, 16], private chunks)."""
    counts = np.zeros((n_chunks, 16), np.uint32)
    r = lib().wxe_queue_sim(n_chunks, n_ctas, warps_per_cta, counts.ctypes.data, seed)
    if r < 0:
        raise RuntimeError("wxe_queue_sim: a ticket outside the frame was handed out")
    return counts, r


def chunk_coverage(width: int, height: int, shard_index: int = 0, shard_count: int = 1, band_rows: int = 0, n_cams: int = 1):
    """How often each pixel is reached by the (chunk, tile, lane) mapping of the CTA-level work queue for one shard."""
    hits = np.zeros((n_cams, height, width), np.uint32)
    lib().wxe_chunk_coverage(width, height, shard_index, shard_count, band_rows, n_cams, hits.ctypes.data)
    return hits


def render_cta_queue(desc, states, width: int, height: int, n_ctas: int = 6, warps_per_cta: int = 4):
    """A frame (or a batch of frames in one mode) through the work distribution of raycast_persistent_cta, played by CPU threads.
    Returns (rgba[n,H,W,4], aov dict)."""
    if not isinstance(states, (list, tuple)):
        states = [states]
    n = len(states)
    buf = (C.c_char * (256 * n))()
    for i, s in enumerate(states):
        buf[256 * i:256 * (i + 1)] = bytes(s)
    rgba = np.zeros((n, height, width, 4), np.uint8)
    out = {
        "state": np.zeros((n, height, width), np.uint8), "voxel": np.zeros((n, height, width, 3), np.int32),
        "leaf": np.zeros((n, height, width), np.int32), "level": np.zeros((n, height, width), np.uint8),
        "iters": np.zeros((n, height, width), np.uint32), "depth": np.zeros((n, height, width), np.float32),
        "mask": np.zeros((n, height, width), np.uint8), "pos": np.zeros((n, height, width, 3), np.float32),
    }
    a = Aov(*[out[k].ctypes.data for k in ("state", "voxel", "leaf", "level", "iters", "depth", "mask", "pos")])
    rc = lib().wxe_render_cta_queue(C.addressof(desc), C.addressof(buf), n, width, height, rgba.ctypes.data, C.addressof(a), n_ctas, warps_per_cta)
    if rc != 0:
        raise RuntimeError(f"wxe_render_cta_queue failed: {rc}")
    return rgba, out
