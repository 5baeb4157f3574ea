"""Scene fixtures shared by the tests: golden asset topologies + synthetic edge cases, prepared by the ORACLE."""
from __future__ import annotations

import functools
import os

import numpy as np

import oracle_ffi as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CAMERAS = {
    "default": ((0.5, 0.5, -500.5), (0.5, 0.5, -498.5)),
    "oblique_a": ((300.0, 200.0, -350.0), (0.0, 0.0, 0.0)),
    "oblique_b": ((250.0, 180.0, -300.0), (0.0, 0.0, 0.0)),
}


def load_topo(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN, f"{name}.topo.npz"))
    return {k: z[k] for k in z.files}


class OracleScene:
    """A tree whose SDF and GPU serialisation were produced by the oracle."""

    def __init__(self, tree: O.Tree, sdf: bool = True):
        self.tree = tree
        if sdf:
            tree.compute_sdf()
        self.gpu = tree.serialise()
        g = self.gpu
        self.origins = g.origins[:, :3].copy()
        self.kids5, self.vals5, self.kids4, self.vals4, self.vals3 = (g.mask64(i) for i in range(5))
        self.tab5, self.tab4, self.tab3 = g.tables()

    def desc(self):
        from woxel_b200.render import make_desc
        return make_desc(self.origins, self.kids5, self.vals5, self.tab5, self.kids4, self.vals4, self.tab4, self.vals3, self.tab3)


@functools.lru_cache(maxsize=None)
def asset(name: str) -> OracleScene:
    t = load_topo(name)
    return OracleScene(O.Tree.from_topology(t["origins"], t["kids5"], t["vals5"], t["kids4"], t["vals4"], t["vals3"]))


def _shell_points(radius: float, band: float, centre=(0, 0, 0)):
    r = int(radius + band + 2)
    ax = np.arange(-r, r + 1)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    d = np.sqrt((x + 0.5) ** 2 + (y + 0.5) ** 2 + (z + 0.5) ** 2)
    m = np.abs(d - radius) <= band
    pts = np.stack([x[m], y[m], z[m]], 1).astype(np.int32)
    return pts + np.asarray(centre, np.int32)


@functools.lru_cache(maxsize=None)
def synthetic(name: str) -> OracleScene:
    t = O.Tree()
    if name == "small_sphere":  # straddles all 8 N5s around the origin
        t.set_voxels(_shell_points(60.0, 1.5))
    elif name == "offcentre_sphere":  # lives in one N5, crosses N4 borders
        t.set_voxels(_shell_points(40.0, 1.0, centre=(300, 140, -260)))
    elif name == "scattered":  # isolated voxels, negative coordinates, far-apart N5s
        rng = np.random.default_rng(7)
        pts = rng.integers(-700, 700, size=(400, 3)).astype(np.int32)
        t.set_voxels(pts)
        for p in ([0, 0, 0], [-1, -1, 0], [5, 6, 7], [123, 78, 3], [-4096, -4096, -4096], [4095, 4095, 4095]):
            t.set_voxel(p)
    elif name == "single_voxel":  # the reference's compute_sdf_test scenario (vdb345.rs:726-741)
        t.set_voxel([5, 6, 7])
    elif name == "beyond_bounds":  # N5s outside the +-4096 world and more than 8 root nodes
        t.set_voxels(_shell_points(30.0, 1.0))
        for k, p in enumerate(([4100, 3, 3], [-8000, 10, 10], [10, 5000, 10], [10, 10, -9000], [4100, 4100, 4100],
                               [9000, 9000, 9000])):
            t.set_voxel(p, k + 1)
    elif name == "slab":  # a thick axis-aligned slab: grazing rays, exact ties
        ax = np.arange(-96, 96)
        x, y, z = np.meshgrid(ax, np.arange(-4, 4), ax, indexing="ij")
        t.set_voxels(np.stack([x.ravel(), y.ravel(), z.ravel()], 1).astype(np.int32))
    elif name == "long_slab":  # long enough for a grazing ray to exhaust the 1000-step budget
        x, y, z = np.meshgrid(np.arange(-600, 600), np.arange(-4, 4), np.arange(-8, 8), indexing="ij")
        t.set_voxels(np.stack([x.ravel(), y.ravel(), z.ravel()], 1).astype(np.int32))
    else:
        raise KeyError(name)
    return OracleScene(t)


@functools.lru_cache(maxsize=None)
def handmade(name: str) -> OracleScene:
    """Topologies set_voxel cannot produce: active tiles, empty leaves (built from raw masks)."""
    base = synthetic("small_sphere")
    k5, v5, k4, v4, v3 = (a.copy() for a in (base.kids5, base.vals5, base.kids4, base.vals4, base.vals3))
    if name == "active_tiles":  # Q4: value bit on internal slots without a child = hit, but no SDF seed
        # an N5 slot and an N4 slot that have no child
        o5 = int(np.flatnonzero(~_bits(k5[7]))[1000])
        v5[7, o5 >> 6] |= np.uint64(1) << np.uint64(o5 & 63)
        o4 = int(np.flatnonzero(~_bits(k4[0]))[5])
        v4[0, o4 >> 6] |= np.uint64(1) << np.uint64(o4 & 63)
        # and one slot that has BOTH bits: the value bit wins (raycast.comp.wgsl:431-437)
        o4b = int(np.flatnonzero(_bits(k4[1]))[0])
        v4[1, o4b >> 6] |= np.uint64(1) << np.uint64(o4b & 63)
    elif name == "empty_leaf":  # leaves with no active voxel at all
        v3[::3] = 0
    else:
        raise KeyError(name)
    return OracleScene(O.Tree.from_topology(base.origins, k5, v5, k4, v4, v3))


def _bits(row64: np.ndarray) -> np.ndarray:
    return np.unpackbits(np.ascontiguousarray(row64).view(np.uint8), bitorder="little").astype(bool)


def bits2d(m64: np.ndarray) -> np.ndarray:
    return np.unpackbits(np.ascontiguousarray(m64).view(np.uint8), bitorder="little").reshape(m64.shape[0], -1).astype(bool)


def get_scene(name: str) -> OracleScene:
    if name in ("cube", "icosahedron"):
        return asset(name)
    if name in ("active_tiles", "empty_leaf"):
        return handmade(name)
    return synthetic(name)


def state_for(eye, target, width, height, mode=0, show_grid=(0, 0, 0), **kw) -> O.State:
    return O.compute_state(eye, target, width=width, height=height, render_mode=mode, show_grid=show_grid, **kw)


def host_tree_from_scene(s):
    """Rebuild a scene in the product's VDB345 through set_voxel (topology only)."""
    import woxel_b200 as W
    v = W.VDB345()
    b3 = bits2d(s.vals3)
    # walk the oracle's DFS order to recover global coordinates of every active voxel
    pts = []
    i4 = i3 = 0
    b5, b4 = bits2d(s.kids5), bits2d(s.kids4)
    for i5, org in enumerate(s.origins):
        for o5 in np.flatnonzero(b5[i5]):
            c5 = np.array([o5 >> 10, (o5 >> 5) & 31, o5 & 31]) * 128
            for o4 in np.flatnonzero(b4[i4]):
                c4 = np.array([o4 >> 8, (o4 >> 4) & 15, o4 & 15]) * 8
                o3 = np.flatnonzero(b3[i3])
                if len(o3):
                    c3 = np.stack([o3 >> 6, (o3 >> 3) & 7, o3 & 7], 1)
                    pts.append(org + c5 + c4 + c3)
                i3 += 1
            i4 += 1
    v.set_voxels(np.concatenate(pts).astype(np.int32))
    return v
