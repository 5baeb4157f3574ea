"""oracle/sdf_python.py -- TEST INFRASTRUCTURE (imported by tests/ only; never by the product).

A second, independent restatement of VDB345::compute_sdf (src/vdb/vdb345.rs:290-628), written from the Rust text in plain
Python on a pointer-style tree (dict of N5 -> list of Tile / N4 -> list of Tile / leaf), the way the reference holds it --
not on the flat arrays oracle/wxo_tree.c, the product host (woxel_b200/host/vdb.cpp) and the GPU sweep (wx_sdf.cu) share.
Small scenes only (pure-Python loops).  tests/test_oracle_cross_check.py requires its distances to equal the C oracle's.

The reference's algorithm, as written there:
  * :292-326  every tile slot (N5, N4: u32; leaf voxels that are LeafData::Tile: usize) is set to MAX - 1;
  * :328-345  13 "forward" neighbours (dx = -1 with any dy, dz; dx = 0, dy = -1 with any dz; (0, 0, -1)) in that order,
              and their negations as the "backward" neighbours;
  * :355-497  forward pass: roots in ascending key order, N5 slots in ascending offset; a child slot is descended into at
              once (its N4 slots in ascending offset, each leaf's voxels in ascending offset) before the next N5 slot;
              a tile slot looks at its forward neighbours IN PLACE: inside the same node a child neighbour gives 1 and ends
              the neighbour loop (`break`), a tile neighbour gives min(own, v + 1); outside the node the neighbour position is
              resolved with get_voxel (:67-106): an endpoint of the SAME kind (N5 tile for an N5 tile, N4 tile for an N4 tile,
              a leaf's inactive voxel for a voxel) gives min(own, v + 1), ANYTHING else (background, another level, an active
              voxel) gives 1 -- without ending the loop;
  * :499-628  backward pass: the same in descending order with the backward neighbours.
"""
from __future__ import annotations

U32_INIT = 0xFFFFFFFF - 1
USIZE_INIT = 0xFFFFFFFFFFFFFFFF - 1

F_NEIGHBOURS = [(-1, dy, dz) for dy in (-1, 0, 1) for dz in (-1, 0, 1)] + [(0, -1, dz) for dz in (-1, 0, 1)] + [(0, 0, -1)]
B_NEIGHBOURS = [(1, dy, dz) for dy in (-1, 0, 1) for dz in (-1, 0, 1)] + [(0, 1, dz) for dz in (-1, 0, 1)] + [(0, 0, 1)]


class Leaf:  # LeafNode<_, 3>: data[512] of LeafData::Tile(usize) | LeafData::Value
    def __init__(self, active):
        self.active = active          # list[bool], offset (x << 6 | y << 3 | z)
        self.tile = [0] * 512         # LeafData::Tile payload where not active


class Internal:  # InternalNode: data[N] of InternalData::Tile(u32) | InternalData::Node
    def __init__(self, n):
        self.child = [None] * n
        self.tile = [0] * n


def build(origins, kids5, kids4, vals3):
    """Pointer-style tree from the DFS-ordered topology (bool arrays per node)."""
    root, i4, i3 = {}, 0, 0
    for k, org in enumerate(origins):
        n5 = Internal(32768)
        for o5 in range(32768):
            if not kids5[k][o5]:
                continue
            n4 = Internal(4096)
            for o4 in range(4096):
                if kids4[i4][o4]:
                    n4.child[o4] = Leaf([bool(b) for b in vals3[i3]])
                    i3 += 1
            n5.child[o5] = n4
            i4 += 1
        root[tuple(int(v) for v in org)] = n5
    return root


def get_voxel(root, p):
    """get_voxel (:67-106): ('bkgr',) | ('innr', v, 5) | ('innr', v, 4) | ('offs', v) | ('leaf',)"""
    key = ((p[0] >> 12) << 12, (p[1] >> 12) << 12, (p[2] >> 12) << 12)
    n5 = root.get(key)
    if n5 is None:
        return ("bkgr",)
    o5 = (((p[0] & 4095) >> 7) << 10) | (((p[1] & 4095) >> 7) << 5) | ((p[2] & 4095) >> 7)
    n4 = n5.child[o5]
    if n4 is None:
        return ("innr", n5.tile[o5], 5)
    o4 = (((p[0] & 127) >> 3) << 8) | (((p[1] & 127) >> 3) << 4) | ((p[2] & 127) >> 3)
    leaf = n4.child[o4]
    if leaf is None:
        return ("innr", n4.tile[o4], 4)
    o3 = ((p[0] & 7) << 6) | ((p[1] & 7) << 3) | (p[2] & 7)
    return ("leaf",) if leaf.active[o3] else ("offs", leaf.tile[o3])


def _relax_internal(root, node, off, child, glob, log_d, cell, level, neighbours):
    """One tile slot of an N5 (log_d 5, cell 128, level 5) or N4 (log_d 4, cell 8, level 4)."""
    dim = 1 << log_d
    v = node.tile[off]
    for dn in neighbours:
        nc = (child[0] + dn[0], child[1] + dn[1], child[2] + dn[2])
        if 0 <= nc[0] < dim and 0 <= nc[1] < dim and 0 <= nc[2] < dim:  # global_to_node(nglobal) == global_to_node(global)
            nid = (nc[0] << (2 * log_d)) | (nc[1] << log_d) | nc[2]
            if node.child[nid] is not None:
                v = 1
                break
            v = min(v, node.tile[nid] + 1)
            continue
        e = get_voxel(root, (glob[0] + dn[0] * cell, glob[1] + dn[1] * cell, glob[2] + dn[2] * cell))
        v = min(v, e[1] + 1) if (e[0] == "innr" and e[2] == level) else 1
    node.tile[off] = v


def _relax_voxel(root, leaf, off, child, glob, neighbours):
    v = leaf.tile[off]
    for dn in neighbours:
        nc = (child[0] + dn[0], child[1] + dn[1], child[2] + dn[2])
        if 0 <= nc[0] < 8 and 0 <= nc[1] < 8 and 0 <= nc[2] < 8:
            nid = (nc[0] << 6) | (nc[1] << 3) | nc[2]
            if leaf.active[nid]:
                v = 1
                break
            v = min(v, leaf.tile[nid] + 1)
            continue
        e = get_voxel(root, (glob[0] + dn[0], glob[1] + dn[1], glob[2] + dn[2]))
        v = min(v, e[1] + 1) if e[0] == "offs" else 1
    leaf.tile[off] = v


def compute_sdf(root):
    for n5 in root.values():  # :292-326
        for o5 in range(32768):
            n4 = n5.child[o5]
            if n4 is None:
                n5.tile[o5] = U32_INIT
                continue
            for o4 in range(4096):
                leaf = n4.child[o4]
                if leaf is None:
                    n4.tile[o4] = U32_INIT
                    continue
                for o3 in range(512):
                    if not leaf.active[o3]:
                        leaf.tile[o3] = USIZE_INIT
    for neighbours, backward in ((F_NEIGHBOURS, False), (B_NEIGHBOURS, True)):
        order = (lambda n: range(n - 1, -1, -1)) if backward else (lambda n: range(n))
        for key in sorted(root.keys(), reverse=backward):
            n5 = root[key]
            for o5 in order(32768):
                c5 = (o5 >> 10, (o5 >> 5) & 31, o5 & 31)
                g5 = (key[0] + c5[0] * 128, key[1] + c5[1] * 128, key[2] + c5[2] * 128)
                n4 = n5.child[o5]
                if n4 is None:
                    _relax_internal(root, n5, o5, c5, g5, 5, 128, 5, neighbours)
                    continue
                for o4 in order(4096):
                    c4 = (o4 >> 8, (o4 >> 4) & 15, o4 & 15)
                    g4 = (g5[0] + c4[0] * 8, g5[1] + c4[1] * 8, g5[2] + c4[2] * 8)
                    leaf = n4.child[o4]
                    if leaf is None:
                        _relax_internal(root, n4, o4, c4, g4, 4, 8, 4, neighbours)
                        continue
                    for o3 in order(512):
                        if leaf.active[o3]:
                            continue
                        c3 = (o3 >> 6, (o3 >> 3) & 7, o3 & 7)
                        _relax_voxel(root, leaf, o3, c3, (g4[0] + c3[0], g4[1] + c3[1], g4[2] + c3[2]), neighbours)
    return root


def tables(root):
    """Distances in the DFS order of the flat serialisation: lists per N5 / N4 / leaf (None where the slot is a child / active)."""
    t5, t4, t3 = [], [], []
    for key in sorted(root.keys()):
        n5 = root[key]
        t5.append([None if n5.child[o] is not None else n5.tile[o] for o in range(32768)])
        for o5 in range(32768):
            n4 = n5.child[o5]
            if n4 is None:
                continue
            t4.append([None if n4.child[o] is not None else n4.tile[o] for o in range(4096)])
            for o4 in range(4096):
                leaf = n4.child[o4]
                if leaf is not None:
                    t3.append([None if leaf.active[o] else leaf.tile[o] & 0xFFFFFFFF for o in range(512)])  # `as u32` at :245
    return t5, t4, t3
