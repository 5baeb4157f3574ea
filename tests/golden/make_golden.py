"""Generates the committed fixtures under tests/golden/ (run HERE, where /root/reference exists):

  <asset>.topo.npz   topology of the reference's shipped .vdb assets (origins + child/value masks in
                     the reference's DFS order) as parsed by the oracle's restatement of src/vdb/read.rs.
                     The GPU box has no /root/reference, so tests and bench load these instead of the
                     .vdb files.  Only topology is kept: leaf values never influence a pixel
                     (raycast.comp.wgsl:485-493) and the SDF is recomputed.
  golden.json        known answers: the reference's own unit-test vectors (cited), file metadata
                     voxel counts, and oracle-derived regression numbers (SDF sums, frame statistics).
  frames_<asset>.npz small oracle frames (rgba + AOVs) per render mode, for regression + GPU parity.

Usage: python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_ffi as O  # noqa: E402

ASSETS = "/root/reference/assets"
CAMERAS = {
    "default": ((0.5, 0.5, -500.5), (0.5, 0.5, -498.5)),
    "oblique_a": ((300.0, 200.0, -350.0), (0.0, 0.0, 0.0)),
    "oblique_b": ((250.0, 180.0, -300.0), (0.0, 0.0, 0.0)),
}


def bits(m):
    return np.unpackbits(m.view(np.uint8), bitorder="little").reshape(m.shape[0], -1).astype(bool)


def main():
    golden = {
        "reference_unit_tests": {
            "global_to_node": [  # src/vdb/data_structure.rs:424-437
                {"level": 3, "in": [-1, 0, 0], "out": [-8, 0, 0]},
                {"level": 4, "in": [-142, 2431, 102], "out": [-256, 2304, 0]},
                {"level": 5, "in": [-1, 0, -42141], "out": [-4096, 0, -45056]},
            ],
            "mask_words": {"3": 8, "4": 64, "5": 512},  # :440-459
            "total_dim": {"3": 8, "4": 128, "5": 4096},  # :462-466
            "global_to_offset": [  # :469-474
                {"level": 3, "in": [0, 0, 0], "out": 0},
                {"level": 3, "in": [1, 2, 3], "out": 83},
                {"level": 4, "in": [121321, 212123, 3121], "out": 3382},
                {"level": 5, "in": [1, 2, 3], "out": 0},
            ],
            "local_to_offset_roundtrip_n4": [[1, 2, 3], [15, 15, 0], [8, 9, 10]],  # :477-485
            "set_get_voxel_points": [[0, 0, 0], [123, 78, 3], [34, 123, 46], [102, 79, 28]],  # vdb345.rs:703-723
            "compute_sdf_test_point": [5, 6, 7],  # vdb345.rs:726-741 (smoke only in the reference)
        },
        "assets": {},
    }
    for name in ("cube", "icosahedron"):
        path = f"{ASSETS}/{name}.vdb"
        tree, info = O.Tree.read(path, "ls_" + name)
        assert info.topology_end_pos == info.block_pos
        assert tree.count_leaf_values() == info.file_voxel_count  # what read.rs:796-806 asserts
        g0 = tree.serialise()
        np.savez_compressed(
            f"{HERE}/{name}.topo.npz", origins=g0.origins[:, :3].copy(), kids5=g0.mask64(0), vals5=g0.mask64(1),
            kids4=g0.mask64(2), vals4=g0.mask64(3), vals3=g0.mask64(4))
        tree.compute_sdf()
        g = tree.serialise()
        t5, t4, t3 = g.tables()
        b5, b4, b3 = bits(g.mask64(0)), bits(g.mask64(2)), bits(g.mask64(4))
        a = {
            "sha256_vdb": hashlib.sha256(open(path, "rb").read()).hexdigest(),
            "file_version": info.file_version, "grid_compression": info.grid_compression,
            "is_half_float": info.is_half_float,
            "file_voxel_count": info.file_voxel_count,  # read.rs:796-806 (metadata of the file)
            "nodes": tree.count_nodes(), "atlas_dim": g.atlas_dim,
            "sdf_sum": [int(t5[~b5].sum()), int(t4[~b4].sum()), int(t3[~b3].sum())],
            "sdf_max": [int(t5[~b5].max()), int(t4[~b4].max()), int(t3[~b3].max())],
            "leaf_dist_hist": np.bincount(t3[~b3], minlength=8)[:8].tolist(),
            "frames": {},
        }
        frames = {}
        for cam, (eye, target) in CAMERAS.items():
            for (w, h, mode) in ((640, 480, 3), (1920, 1080, 0)):
                if cam != "default" and w != 640:
                    continue
                st = O.compute_state(eye, target, width=w, height=h, render_mode=mode)
                rgba, aov, s = g.render(st, w, h)
                n = s.primary_rays
                a["frames"][f"{cam}_{w}x{h}_m{mode}"] = {
                    "hit": s.hit, "oob": s.oob, "maxed": s.maxed, "rays": s.rays,
                    "primary_lookups": list(s.primary_lookups), "primary_alg_bytes": s.primary_alg_bytes,
                    "max_iters": s.max_iters, "rgba_sha256": hashlib.sha256(rgba.tobytes()).hexdigest(),
                    "iters_sum": int(aov["iters"].sum()),
                }
            for mode in range(5):  # small frames kept in full
                w, h = 96, 64
                st = O.compute_state(eye, target, width=w, height=h, render_mode=mode, show_grid=(1, 1, 1) if mode < 3 else (0, 0, 0))
                rgba, aov, s = g.render(st, w, h)
                frames[f"{cam}_m{mode}_state"] = np.frombuffer(bytes(st), np.uint8).copy()
                frames[f"{cam}_m{mode}_rgba"] = rgba
                if mode == 0:
                    for k in ("state", "voxel", "leaf", "iters", "depth", "mask"):
                        frames[f"{cam}_{k}"] = aov[k]
        np.savez_compressed(f"{HERE}/frames_{name}.npz", **frames)
        golden["assets"][name] = a
        print(name, "done", {k: v for k, v in a.items() if k != "frames"})
    json.dump(golden, open(f"{HERE}/golden.json", "w"), indent=1)


if __name__ == "__main__":
    main()
