"""Two independent restatements of raycast.comp.wgsl must agree bit for bit: oracle/wxo_render.c (C, per pixel, with the
shader's parent cache, on flat tables) and oracle/wgsl_numpy.py (numpy float32, all rays at once, every lookup from the
root, on the reference's own atlas textures + u32 mask words + origins).  A transcription error in either one -- or in the
atlas / mask serialisation they do NOT share -- shows up here.  (Neither is pinned to an execution of the reference:
"parity unpinned")"""
import os
import sys

import numpy as np
import pytest

import scenes

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import wgsl_numpy as WN  # noqa: E402


def numpy_scene(s):
    g = s.gpu
    return WN.Scene(g.atlas(0), g.atlas(1), g.atlas(2), *(g.mask(i) for i in range(5)), g.origins)


def compare(name, st, w, h):
    s = scenes.get_scene(name)
    rgba, hit = WN.cp_main(numpy_scene(s), bytes(st), w, h)
    ref_rgba, ref, _ = s.gpu.render(st, w, h)
    dw, dh = (w // 8) * 8, (h // 4) * 4
    assert np.array_equal(hit["state"], ref["state"][:dh, :dw]), name
    assert np.array_equal(hit["i"], ref["iters"][:dh, :dw]), name
    assert np.array_equal(hit["num_parents"], ref["level"][:dh, :dw]), name
    bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
    nan = np.isnan(hit["p"])
    assert np.array_equal(nan, np.isnan(ref["pos"][:dh, :dw]))
    assert np.array_equal(bits(hit["p"])[~nan], bits(ref["pos"][:dh, :dw])[~nan]), name
    m = hit["mask"]
    assert np.array_equal(m[..., 0] | (m[..., 1].astype(np.uint8) << 1) | (m[..., 2].astype(np.uint8) << 2), ref["mask"][:dh, :dw]), name
    leaf = np.where(hit["num_parents"] == 3, hit["leaf"].astype(np.int64), -1)
    assert np.array_equal(leaf, ref["leaf"][:dh, :dw]), name
    assert np.array_equal(np.floor(hit["p"]).astype(np.int64)[~nan], ref["voxel"][:dh, :dw].astype(np.int64)[~nan]), name
    assert np.array_equal(rgba, ref_rgba), (name, int((rgba != ref_rgba).any(-1).sum()))


@pytest.mark.parametrize("name", ["cube", "icosahedron"])
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_assets_all_modes(name, mode):
    eye, target = scenes.CAMERAS["oblique_a" if mode in (1, 3) else ("oblique_b" if mode == 4 else "default")]
    st = scenes.state_for(eye, target, 96, 64, mode=mode, show_grid=(1, 1, 1) if mode < 3 else (0, 0, 0))
    compare(name, st, 96, 64)


@pytest.mark.parametrize("name,eye,target", [
    ("small_sphere", (90.5, 70.5, -120.5), (0.0, 0.0, 0.0)),
    ("scattered", (10.5, 20.5, -900.5), (0.0, 0.0, 0.0)),
    ("beyond_bounds", (4000.5, 30.5, -150.5), (4100.0, 3.0, 3.0)),
    ("slab", (0.5, 4.0, -300.5), (0.5, 4.0, 0.5)),
    ("long_slab", (-900.5, 4.0005, 0.5), (0.5, 4.0005, 0.5)),
    ("active_tiles", (90.5, 70.5, -120.5), (0.0, 0.0, 0.0)),
    ("empty_leaf", (90.5, 70.5, -120.5), (0.0, 0.0, 0.0)),
])
@pytest.mark.parametrize("mode", [0, 4])
def test_edge_scenes(name, eye, target, mode):
    st = scenes.state_for(eye, target, 64, 40, mode=mode)
    compare(name, st, 64, 40)


def test_ragged_frame_and_outside_eye():
    compare("cube", scenes.state_for(*scenes.CAMERAS["default"], 45, 30, mode=3), 45, 30)
    compare("small_sphere", scenes.state_for((0.5, 0.5, -5000.5), (0.5, 0.5, 0.5), 32, 16, mode=2), 32, 16)


@pytest.mark.parametrize("name", ["single_voxel", "offcentre_sphere", "small_sphere", "scattered"])
def test_compute_sdf_two_restatements_agree(name):
    """compute_sdf (vdb345.rs:290-628): the C oracle's sweep over flat arrays against oracle/sdf_python.py, a plain-Python
    restatement on a pointer-style tree (written from the Rust text), value for value: N5 tiles, N4 tiles and every inactive
    voxel.  One N5 (single_voxel, offcentre_sphere), eight N5s with the shell crossing all of them (small_sphere) and isolated
    voxels in far-apart nodes (scattered) exercise the in-node, cross-node and background cases."""
    import sdf_python as SP
    s = scenes.get_scene(name)
    k5, k4, v3 = scenes.bits2d(s.kids5), scenes.bits2d(s.kids4), scenes.bits2d(s.vals3)
    root = SP.compute_sdf(SP.build(s.origins, k5, k4, v3))
    t5, t4, t3 = SP.tables(root)
    mine5 = np.array([[0 if v is None else v for v in row] for row in t5], np.uint64)
    mine4 = np.array([[0 if v is None else v for v in row] for row in t4], np.uint64).reshape(-1, 4096)
    mine3 = np.array([[0 if v is None else v for v in row] for row in t3], np.uint64).reshape(-1, 512)
    assert mine5.shape == s.tab5.shape and mine4.shape == s.tab4.shape and mine3.shape == s.tab3.shape
    assert np.array_equal(mine5[~k5], s.tab5[~k5].astype(np.uint64))
    assert np.array_equal(mine4[~k4], s.tab4[~k4].astype(np.uint64))
    assert np.array_equal(mine3[~v3], s.tab3[~v3].astype(np.uint64))
    assert int(mine5[~k5].min()) >= 1 and int(mine3[~v3].min()) >= 1


def test_randomized_scenes_and_cameras_two_restatements_agree():
    """The same seeded fuzz the device code is held to (tests/test_device_emu.py, tests/test_parity_gpu.py), between the two
    restatements of the shader: random sparse trees incl. out-of-world N5s and solid blocks, random cameras (inside the volume,
    on lattice planes, axis-aligned, near the bounds), random render modes."""
    import oracle_ffi as O
    rng = np.random.default_rng(424242)
    for case in range(5):
        t = O.Tree()
        for _ in range(int(rng.integers(1, 5))):
            centre = rng.integers(-900, 900, 3) if rng.random() < 0.8 else rng.integers(-6000, 6000, 3)
            ext = int(rng.integers(1, 40))
            t.set_voxels((centre + rng.integers(-ext, ext + 1, size=(int(rng.integers(1, 300)), 3))).astype(np.int32))
        if case % 2 == 0:
            c0 = rng.integers(-100, 100, 3)
            ax = [np.arange(c, c + int(rng.integers(2, 20))) for c in c0]
            t.set_voxels(np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3).astype(np.int32))
        s = scenes.OracleScene(t)
        sc = numpy_scene(s)
        for kind in range(6):
            eye = [tuple(rng.uniform(-1500, 1500, 3)), tuple(float(v) for v in rng.integers(-300, 300, 3)),
                   (float(rng.integers(-50, 50)) + 0.5, float(rng.integers(-50, 50)) + 0.5, -700.5), tuple(rng.uniform(-4090, 4090, 3)),
                   tuple(rng.uniform(4000, 4300, 3)), tuple(rng.uniform(-60, 60, 3))][kind]
            target = (eye[0], eye[1], eye[2] + 100.0) if kind == 2 else tuple(rng.uniform(-200, 200, 3))
            st = scenes.state_for(eye, target, 48, 24, mode=int(rng.integers(0, 5)), show_grid=(1, 1, 1))
            rgba, hit = WN.cp_main(sc, bytes(st), 48, 24)
            ref_rgba, ref, _ = s.gpu.render(st, 48, 24)
            assert np.array_equal(hit["state"], ref["state"]) and np.array_equal(hit["i"], ref["iters"]), (case, kind, eye, target)
            nan = np.isnan(hit["p"])
            assert np.array_equal(hit["p"].view(np.uint32)[~nan], ref["pos"].view(np.uint32)[~nan]), (case, kind)
            assert np.array_equal(rgba, ref_rgba), (case, kind, eye, target)
