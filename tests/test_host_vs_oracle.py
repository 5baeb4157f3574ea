"""The product's C++ host (woxel_b200/host: src/vdb, src/scene, src/render mirror) against the oracle.
Runs without a GPU: nothing here launches a kernel."""
import os

import numpy as np
import pytest

import oracle_ffi as O
import scenes
import woxel_b200 as W
from woxel_b200 import vdb as V

REF_ASSETS = "/root/reference/assets"


host_tree_from_scene = scenes.host_tree_from_scene


def assert_flat_equals_oracle(f: W.FlatTree, s: scenes.OracleScene, values: bool = False):
    assert (f.n5, f.n4, f.n3) == (len(s.origins), s.kids4.shape[0], s.vals3.shape[0])
    assert np.array_equal(f.origins, s.origins)
    for a, b, k in ((f.kids5, s.kids5, "kids5"), (f.vals5, s.vals5, "vals5"), (f.kids4, s.kids4, "kids4"),
                    (f.vals4, s.vals4, "vals4"), (f.vals3, s.vals3, "vals3"), (f.tab5, s.tab5, "tab5"), (f.tab4, s.tab4, "tab4")):
        assert np.array_equal(a, b), k
    inactive = ~scenes.bits2d(s.vals3)
    assert np.array_equal(np.asarray(f.tab3, np.uint32)[inactive], s.tab3[inactive]), "leaf distances"
    if values:
        assert np.array_equal(np.asarray(f.tab3, np.uint32), s.tab3), "leaf slots incl. values"


def test_index_maths_match_reference_vectors():
    assert V.N3.global_to_node([-1, 0, 0]) == [-8, 0, 0]
    assert V.N4.global_to_node([-142, 2431, 102]) == [-256, 2304, 0]
    assert V.N5.global_to_node([-1, 0, -42141]) == [-4096, 0, -45056]
    assert V.N3.global_to_offset([1, 2, 3]) == 83 and V.N4.global_to_offset([121321, 212123, 3121]) == 3382
    assert V.N5.global_to_offset([1, 2, 3]) == 0
    rng = np.random.default_rng(1)
    for nm, lvl in ((V.N3, 3), (V.N4, 4), (V.N5, 5)):
        for g in rng.integers(-2**20, 2**20, size=(50, 3)):
            assert nm.global_to_node(g) == O.global_to_node(lvl, g)
            assert nm.global_to_offset(g) == O.global_to_offset(lvl, g)
        for off in rng.integers(0, nm.SIZE, size=20):
            assert nm.offset_to_child(off) == O.offset_to_child(lvl, int(off))
            assert nm.child_to_offset(nm.offset_to_child(off)) == off


def test_set_get_voxel():  # vdb345.rs:703-723
    v = W.VDB345()
    pts = [[0, 0, 0], [123, 78, 3], [34, 123, 46], [102, 79, 28]]
    for i, p in enumerate(pts):
        v.set_voxel(p, i)
    for i, p in enumerate(pts):
        e = v.get_voxel(p)
        assert (e.kind, e.value) == ("Leaf", i)
    assert v.get_voxel([1, 0, 0]).kind == "Offs"
    e = v.get_voxel([60, 60, 60])
    assert (e.kind, e.level) == ("Innr", 4)
    e = v.get_voxel([1000, 0, 0])
    assert (e.kind, e.level) == ("Innr", 5)
    assert v.get_voxel([5000, 0, 0]).kind == "Bkgr"
    assert v.count_nodes() == [1, 1, 4]


@pytest.mark.parametrize("name", ["single_voxel", "scattered", "small_sphere", "offcentre_sphere", "beyond_bounds", "slab"])
def test_compute_sdf_and_to_flat_match_oracle(name):
    s = scenes.get_scene(name)
    v = host_tree_from_scene(s)
    assert v.count_nodes() == s.tree.count_nodes()
    v.compute_sdf()
    assert_flat_equals_oracle(v.to_flat(narrow_leaves=False), s)
    f8 = v.to_flat(narrow_leaves=True)
    assert f8.desc.tab3_elem_bytes == 1
    assert_flat_equals_oracle(f8, s)


@pytest.mark.parametrize("name", ["cube", "icosahedron"])
def test_assets_sdf_match_oracle(name):
    s = scenes.get_scene(name)
    v = host_tree_from_scene(s)
    v.compute_sdf()
    assert_flat_equals_oracle(v.to_flat(narrow_leaves=False), s)


@pytest.mark.skipif(not os.path.isdir(REF_ASSETS), reason="reference assets only exist in the authoring container")
@pytest.mark.parametrize("name", ["cube", "icosahedron"])
def test_reader_matches_oracle_reader(name):  # read.rs:734-807
    r = W.VdbReader(f"{REF_ASSETS}/{name}.vdb")
    v = r.read_vdb345_grid("ls_" + name)
    assert v.count_leaf_values() == r.info.file_voxel_count
    ot, info = O.Tree.read(f"{REF_ASSETS}/{name}.vdb", "ls_" + name)
    assert (r.info.file_version, r.info.grid_compression, r.info.block_pos) == (info.file_version, info.grid_compression, info.block_pos)
    v.compute_sdf()
    assert_flat_equals_oracle(v.to_flat(narrow_leaves=False), scenes.OracleScene(ot), values=True)


def test_reader_errors(tmp_path):
    p = tmp_path / "bad.vdb"
    p.write_bytes(b"\x00" * 64)
    with pytest.raises(V.VdbError) as e:
        W.VdbReader(str(p)).read_vdb345_grid("x")
    assert e.value.status == -102  # MagicMismatch
    with pytest.raises(V.VdbError):
        W.VdbReader(str(tmp_path / "missing.vdb")).read_vdb345_grid("x")
    if os.path.isdir(REF_ASSETS):
        with pytest.raises(V.VdbError) as e:
            W.VdbReader(f"{REF_ASSETS}/cube.vdb").read_vdb345_grid("nope")
        assert e.value.status == -105  # InvalidGridName


def test_compute_state_matches_oracle_bitwise():  # compute_state.rs:87-131
    rng = np.random.default_rng(3)
    cams = [scenes.CAMERAS[k] for k in scenes.CAMERAS]
    for _ in range(40):
        eye = rng.uniform(-3000, 3000, 3)
        cams.append((tuple(eye), tuple(eye + rng.normal(size=3))))
    for eye, target in cams:
        for (w, h) in ((640, 480), (1920, 1080), (3840, 2160)):
            for mode in (0, 3):
                so = O.compute_state(eye, target, width=w, height=h, render_mode=mode, show_grid=(1, 0, 1))
                sp = W.ComputeState.build(W.Camera(eye=eye, target=target, aspect=w / h), w, mode, (True, False, True))
                assert bytes(so) == bytes(sp)
    s = W.ComputeState.build(W.Camera.quick_camera(640 / 480), 640)
    assert s.render_mode[0] == 3 and list(s.eye) == [0.5, 0.5, -500.5, 0.0]
    assert np.allclose(list(s.sun_dir)[:3], np.array([1, -1, 0.5]) / 1.5, atol=1e-6)
    assert list(s.sun_color) == [1.0, np.float32(210 / 255), np.float32(160 / 255), 1.0]


def test_procedural_sphere_matches_bruteforce():
    v = W.VDB345.sphere(half=64, radius=50.0, band=1.5)
    ax = np.arange(-64, 64)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    d = np.sqrt((x + 0.5) ** 2 + (y + 0.5) ** 2 + (z + 0.5) ** 2)
    m = np.abs(d - 50.0) <= 1.5
    assert v.count_leaf_values() == int(m.sum())
    t = O.Tree()
    t.set_voxels(np.stack([x[m], y[m], z[m]], 1).astype(np.int32))
    v.compute_sdf()
    assert_flat_equals_oracle(v.to_flat(False), scenes.OracleScene(t))


def test_procedural_torus_matches_bruteforce():
    v = W.VDB345.torus(half=64, major=40.0, minor=12.0, band=1.0)
    ax = np.arange(-64, 64)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    q = np.sqrt((x + 0.5) ** 2 + (z + 0.5) ** 2) - 40.0
    m = np.abs(np.sqrt(q * q + (y + 0.5) ** 2) - 12.0) <= 1.0
    assert v.count_leaf_values() == int(m.sum())
