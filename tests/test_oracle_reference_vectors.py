"""The oracle against every known-answer vector the reference's own tests hold for this path
 and against the metadata of its shipped assets."""
import json
import os

import numpy as np
import pytest

import oracle_ffi as O
import scenes

G = json.load(open(os.path.join(scenes.GOLDEN, "golden.json")))
RT = G["reference_unit_tests"]
REF_ASSETS = "/root/reference/assets"


def test_global_to_node():  # data_structure.rs:424-437
    for v in RT["global_to_node"]:
        assert O.global_to_node(v["level"], v["in"]) == v["out"]


def test_bit_index():  # data_structure.rs:469-474
    for v in RT["global_to_offset"]:
        assert O.global_to_offset(v["level"], v["in"]) == v["out"]


def test_local_to_offset_roundtrip():  # data_structure.rs:477-485
    for c in RT["local_to_offset_roundtrip_n4"]:
        assert O.offset_to_child(4, O.child_to_offset(4, c)) == c


def test_mask_sizes_and_dims():  # data_structure.rs:440-466
    s = scenes.get_scene("single_voxel")
    assert s.gpu.mask(4).shape[1] * 32 == 64 * RT["mask_words"]["3"]
    assert s.gpu.mask(2).shape[1] * 32 == 64 * RT["mask_words"]["4"]
    assert s.gpu.mask(0).shape[1] * 32 == 64 * RT["mask_words"]["5"]
    for lvl, dim in RT["total_dim"].items():
        assert O.global_to_node(int(lvl), [dim, dim, dim]) == [dim, dim, dim]
        assert O.global_to_node(int(lvl), [dim - 1] * 3) == [0, 0, 0]


def test_set_get_voxel():  # vdb345.rs:703-723
    t = O.Tree()
    pts = RT["set_get_voxel_points"]
    for i, p in enumerate(pts):
        t.set_voxel(p, i)
    for i, p in enumerate(pts):
        assert t.get_voxel(p) == (O.EP_LEAF, i)
    assert t.get_voxel([1, 0, 0])[0] == O.EP_OFFS  # same leaf, inactive
    assert t.get_voxel([60, 60, 60])[0] == O.EP_INNR4  # same N4, no leaf
    assert t.get_voxel([1000, 0, 0])[0] == O.EP_INNR5  # same N5, no N4
    assert t.get_voxel([5000, 0, 0])[0] == O.EP_BKGR  # no N5


def test_compute_sdf_scenario():  # vdb345.rs:726-741 (the reference asserts nothing; we check the obvious)
    s = scenes.get_scene("single_voxel")
    p = np.array(RT["compute_sdf_test_point"])
    assert s.tree.count_nodes() == [1, 1, 1]
    t3 = s.tab3[0].reshape(8, 8, 8)
    x, y, z = np.meshgrid(*[np.arange(8)] * 3, indexing="ij")
    cheb = np.maximum(np.maximum(abs(x - p[0]), abs(y - p[1])), abs(z - p[2]))
    # inside one leaf the two-pass chamfer with unit weights is the Chebyshev distance, capped by the
    # "unknown neighbour => 1" rule at the leaf faces
    face = np.minimum.reduce([x, y, z, 7 - x, 7 - y, 7 - z]) + 1
    inactive = cheb > 0
    assert np.array_equal(t3[inactive], np.minimum(cheb, face)[inactive])
    # the same rule one and two levels up: the only child sits in slot (0,0,0) of its N4 and of its N5
    for tab, dim in ((s.tab4[0], 16), (s.tab5[0], 32)):
        t = tab.reshape(dim, dim, dim)
        x, y, z = np.meshgrid(*[np.arange(dim)] * 3, indexing="ij")
        cheb = np.maximum(np.maximum(x, y), z)
        face = np.minimum.reduce([x, y, z, dim - 1 - x, dim - 1 - y, dim - 1 - z]) + 1
        assert np.array_equal(t[cheb > 0], np.minimum(cheb, face)[cheb > 0])


@pytest.mark.skipif(not os.path.isdir(REF_ASSETS), reason="reference assets only exist in the authoring container")
@pytest.mark.parametrize("name", ["cube", "icosahedron"])
def test_reader_voxel_count_and_golden_topology(name):  # read.rs:734-807
    tree, info = O.Tree.read(f"{REF_ASSETS}/{name}.vdb", "ls_" + name)
    a = G["assets"][name]
    assert info.file_version == a["file_version"] and info.grid_compression == a["grid_compression"]
    assert tree.count_leaf_values() == info.file_voxel_count == a["file_voxel_count"]
    assert info.topology_end_pos == info.block_pos  # the topology walk lands exactly on the leaf buffers
    g = tree.serialise()
    t = scenes.load_topo(name)
    assert np.array_equal(g.origins[:, :3], t["origins"])
    for i, k in enumerate(("kids5", "vals5", "kids4", "vals4", "vals3")):
        assert np.array_equal(g.mask64(i), t[k]), k


def test_reader_errors(tmp_path):
    p = tmp_path / "bad.vdb"
    p.write_bytes(b"\x00" * 64)
    with pytest.raises(IOError):
        O.Tree.read(str(p), "x")
    if os.path.isdir(REF_ASSETS):
        with pytest.raises(IOError):
            O.Tree.read(f"{REF_ASSETS}/cube.vdb", "no_such_grid")


@pytest.mark.parametrize("name", ["cube", "icosahedron"])
def test_golden_sdf_and_determinism(name):
    s = scenes.get_scene(name)
    a = G["assets"][name]
    assert s.tree.count_nodes() == a["nodes"] and s.gpu.atlas_dim == a["atlas_dim"]
    b5, b4, b3 = scenes.bits2d(s.kids5), scenes.bits2d(s.kids4), scenes.bits2d(s.vals3)
    assert [int(s.tab5[~b5].sum()), int(s.tab4[~b4].sum()), int(s.tab3[~b3].sum())] == a["sdf_sum"]
    assert [int(s.tab5[~b5].max()), int(s.tab4[~b4].max()), int(s.tab3[~b3].max())] == a["sdf_max"]
    assert np.bincount(s.tab3[~b3], minlength=8)[:8].tolist() == a["leaf_dist_hist"]
    # examples/load_vdb.rs:6-24: serialisation is deterministic across rebuilds
    t = scenes.load_topo(name)
    again = scenes.OracleScene(O.Tree.from_topology(t["origins"], t["kids5"], t["vals5"], t["kids4"], t["vals4"], t["vals3"]))
    for k in range(3):
        assert np.array_equal(again.gpu.atlas(k), s.gpu.atlas(k))


@pytest.mark.parametrize("name", ["cube", "icosahedron"])
def test_golden_frames(name):
    """Oracle regression: committed small frames (made by tests/golden/make_golden.py) reproduce bit-exactly."""
    s = scenes.get_scene(name)
    z = np.load(os.path.join(scenes.GOLDEN, f"frames_{name}.npz"))
    for cam in scenes.CAMERAS:
        for mode in range(5):
            st = O.State.from_buffer_copy(z[f"{cam}_m{mode}_state"].tobytes())
            rgba, aov, _ = s.gpu.render(st, 96, 64)
            assert np.array_equal(rgba, z[f"{cam}_m{mode}_rgba"]), (cam, mode)
            if mode == 0:
                for k in ("state", "voxel", "leaf", "iters", "mask"):
                    assert np.array_equal(aov[k], z[f"{cam}_{k}"]), (cam, k)
                assert np.array_equal(aov["depth"].view(np.uint32), z[f"{cam}_depth"].view(np.uint32))


@pytest.mark.parametrize("name,key", [("cube", "default_640x480_m3"), ("icosahedron", "default_640x480_m3")])
def test_golden_frame_statistics(name, key):
    s = scenes.get_scene(name)
    f = G["assets"][name]["frames"][key]
    eye, target = scenes.CAMERAS["default"]
    st = scenes.state_for(eye, target, 640, 480, mode=3)
    _, _, stats = s.gpu.render(st, 640, 480)
    assert (stats.hit, stats.oob, stats.maxed, stats.rays) == (f["hit"], f["oob"], f["maxed"], f["rays"])
    assert list(stats.primary_lookups) == f["primary_lookups"] and stats.primary_alg_bytes == f["primary_alg_bytes"]
