"""Large procedural scenes built on the GPU (BENCH / TEST INFRASTRUCTURE; see tools/scenegen/fog_gen.cu).

`fog_topology(half, tau)` returns the flat topology (the arrays of a WxTreeDesc without distances) of the value-noise
fog of BASELINE config 4 over [-half, half)^3 -- the same tree `W.VDB345.fog(half, tau).to_flat()` gives, without
the host's per-voxel loop and without the 2 KB-per-leaf host tree (16.7 M leaves at half = 1024)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "tools", "scenegen", "libwx_scenegen.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run __graft_entry__.build()")
        _LIB = C.CDLL(path)
        _LIB.wxs_fog_masks.argtypes = [C.c_int, C.c_int32, C.c_double, C.c_void_p, C.POINTER(C.c_uint64)]
        _LIB.wxs_fog_occupancy.argtypes = [C.c_int, C.c_int32, C.c_double, C.POINTER(C.c_double)]
    return _LIB


def fog_occupancy(half: int, tau: float, device: int = 0) -> float:
    occ = C.c_double()
    rc = lib().wxs_fog_occupancy(device, half, tau, C.byref(occ))
    if rc != 0:
        raise RuntimeError(f"wxs_fog_occupancy failed: {rc}")
    return float(occ.value)


def topology_from_dense_masks(masks: np.ndarray, half: int) -> dict:
    """masks: [(2*half/8)^3, 8] u64 in DFS order over the dense node set (8 N5s x hn^3 N4s x 4096 leaves, hn = half/128).
    Drops the empty leaves / N4s / N5s and numbers the children in DFS order (vdb345.rs:134-158)."""
    hn = half // 128
    dense4 = 8 * hn ** 3
    leaf_any = masks.any(axis=1)
    vals3 = np.ascontiguousarray(masks[leaf_any])
    la = leaf_any.reshape(dense4, 4096)
    n4_any = la.any(axis=1)
    la4 = np.ascontiguousarray(la[n4_any])                       # [n4, 4096] child bits
    kids4 = np.packbits(la4, axis=1, bitorder="little").view(np.uint64).reshape(-1, 64)
    tab4 = np.zeros(la4.shape, np.uint32)
    tab4[la4] = np.arange(int(la4.sum()), dtype=np.uint32)       # running leaf index in DFS order
    a = np.arange(hn)
    origins, kids5, tab5 = [], [], []
    n4_run = 0
    per5 = n4_any.reshape(8, hn, hn, hn)
    for i5 in range(8):
        if not per5[i5].any():
            continue
        side = [(i5 >> 2) & 1, (i5 >> 1) & 1, i5 & 1]            # 0: the N5 at -4096 (its last hn cells), 1: the N5 at 0
        origins.append([0 if s else -4096 for s in side])
        c = [a + (0 if s else 32 - hn) for s in side]
        o5 = (c[0][:, None, None] << 10) | (c[1][None, :, None] << 5) | c[2][None, None, :]
        bits = np.zeros(32768, bool)
        tab = np.zeros(32768, np.uint32)
        sel = per5[i5]
        bits[o5[sel]] = True
        tab[o5[sel]] = n4_run + np.arange(int(sel.sum()), dtype=np.uint32)   # ax-major order == ascending o5
        n4_run += int(sel.sum())
        kids5.append(np.packbits(bits, bitorder="little").view(np.uint64))
        tab5.append(tab)
    n5 = len(origins)
    return {
        "origins": np.array(origins, np.int32).reshape(n5, 3), "kids5": np.array(kids5, np.uint64).reshape(n5, 512),
        "vals5": np.zeros((n5, 512), np.uint64), "tab5": np.array(tab5, np.uint32).reshape(n5, 32768),
        "kids4": kids4, "vals4": np.zeros_like(kids4), "tab4": tab4, "vals3": vals3,
    }


def fog_topology(half: int, tau: float, device: int = 0) -> dict:
    n_dense = (2 * half // 8) ** 3
    masks = np.empty((n_dense, 8), np.uint64)
    n_act = C.c_uint64()
    rc = lib().wxs_fog_masks(device, half, tau, masks.ctypes.data, C.byref(n_act))
    if rc != 0:
        raise RuntimeError(f"wxs_fog_masks failed: {rc}")
    t = topology_from_dense_masks(masks, half)
    t["occupancy"] = n_act.value / float((2 * half) ** 3)
    return t


def desc_of(topo: dict):
    """WxTreeDesc of a topology (no distances: for wx_tree_build / wx_compute_sdf)."""
    from woxel_b200.render import make_desc
    return make_desc(topo["origins"], topo["kids5"], topo["vals5"], topo["tab5"], topo["kids4"], topo["vals4"], topo["tab4"],
                     topo["vals3"], np.zeros(0, np.uint8))
