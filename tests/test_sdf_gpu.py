"""compute_sdf on the GPU (wx_compute_sdf) against the oracle's restatement of vdb345.rs:290-628, value for value."""
import numpy as np
import pytest

import oracle_ffi as O
import scenes
import woxel_b200 as W
from woxel_b200.render import make_desc

pytestmark = pytest.mark.gpu


def topo_desc(s):
    """The scene's topology with every distance wiped: tiles hold garbage, child slots their index."""
    k5, k4 = scenes.bits2d(s.kids5), scenes.bits2d(s.kids4)
    t5 = np.where(k5, s.tab5, np.uint32(0xDEADBEEF)).astype(np.uint32)
    t4 = np.where(k4, s.tab4, np.uint32(0xDEADBEEF)).astype(np.uint32)
    return make_desc(s.origins, s.kids5, s.vals5, t5, s.kids4, s.vals4, t4, s.vals3, np.zeros_like(s.tab3))


def check_scene(ctx, s, narrow=False):
    tab5, tab4, tab3, info = ctx.compute_sdf(topo_desc(s), narrow_leaves=narrow)
    assert np.array_equal(tab5, s.tab5), "N5 table (child indices + tile distances)"
    assert np.array_equal(tab4, s.tab4), "N4 table"
    inactive = ~scenes.bits2d(s.vals3)
    assert np.array_equal(tab3[inactive], s.tab3[inactive].astype(tab3.dtype)), "leaf distances"
    assert not tab3[~inactive].any()
    k5, k4 = scenes.bits2d(s.kids5), scenes.bits2d(s.kids4)
    if (~k5).any():
        assert info.max_dist[0] == s.tab5[~k5].max()
    if (~k4).any():
        assert info.max_dist[1] == s.tab4[~k4].max()
    if inactive.any():
        assert info.max_dist[2] == s.tab3[inactive].max()
    return info


@pytest.mark.parametrize("name", ["cube", "icosahedron", "single_voxel", "scattered", "small_sphere", "offcentre_sphere", "beyond_bounds",
                                  "slab", "long_slab", "active_tiles", "empty_leaf"])
def test_gpu_sdf_equals_oracle(gpu_ctx, name):
    s = scenes.get_scene(name)
    check_scene(gpu_ctx, s)
    check_scene(gpu_ctx, s, narrow=True)


def test_gpu_sdf_feeds_the_renderer(gpu_ctx):
    """model load with the GPU sweep: topology -> wx_compute_sdf -> wx_tree_upload -> frame == oracle frame."""
    s = scenes.get_scene("icosahedron")
    tab5, tab4, tab3, _ = gpu_ctx.compute_sdf(topo_desc(s), narrow_leaves=True)
    tree = gpu_ctx.upload(make_desc(s.origins, s.kids5, s.vals5, tab5, s.kids4, s.vals4, tab4, s.vals3, tab3))
    try:
        eye, target = scenes.CAMERAS["oblique_b"]
        st = scenes.state_for(eye, target, 320, 200, mode=3)
        rgba, _ = gpu_ctx.render(tree, W.ComputeState.from_buffer_copy(bytes(st)), 320, 200)
        ref, _, _ = s.gpu.render(st, 320, 200, aov=False)
        assert np.array_equal(rgba[0], ref)
    finally:
        tree.free()


def test_gpu_sdf_matches_host_on_a_procedural_sphere(gpu_ctx):
    """A scene with hundreds of N4 nodes (cross-node order dependence at every level): GPU == product host sweep."""
    v = W.VDB345.sphere(half=512, radius=480.0, band=2.0)
    f0 = v.to_flat(narrow_leaves=False)  # before compute_sdf: topology only
    tab5, tab4, tab3, info = gpu_ctx.compute_sdf(f0)
    v.compute_sdf()
    f = v.to_flat(narrow_leaves=False)
    assert np.array_equal(tab5, f.tab5) and np.array_equal(tab4, f.tab4)
    inactive = ~scenes.bits2d(f.vals3)
    assert np.array_equal(tab3[inactive], f.tab3[inactive])
    assert info.device_ms > 0


def test_gpu_sdf_rejects_bad_input(gpu_ctx):
    s = scenes.get_scene("single_voxel")
    d = topo_desc(s)
    d._keepalive["tab5"][0, int(np.flatnonzero(scenes.bits2d(s.kids5)[0])[0])] = 99  # child index beyond n4
    with pytest.raises(W.WxError) as e:
        gpu_ctx.compute_sdf(d)
    assert e.value.status == -5


@pytest.mark.parametrize("name", ["cube", "scattered", "beyond_bounds", "active_tiles", "empty_leaf"])
def test_tree_build_equals_sweep_plus_upload(gpu_ctx, name):
    """wx_tree_build (sweep + device-side packing) renders exactly like the oracle's tables through wx_tree_upload."""
    s = scenes.get_scene(name)
    built = gpu_ctx.build(topo_desc(s))
    plain = gpu_ctx.upload(s.desc())
    try:
        for k in ("n5", "n4", "n3", "leaf_bits", "device_bytes"):
            assert getattr(built.info, k) == getattr(plain.info, k), k
        assert list(built.info.max_dist) == list(plain.info.max_dist) and built.sdf.rounds > 0
        for (eye, target), mode in zip([scenes.CAMERAS["oblique_a"], ((3.3, 2.2, 1.1), (40.0, 30.0, -20.0)), ((0.5, 0.5, -300.5), (0.5, 0.5, 0.5))], (3, 0, 4)):
            st = W.ComputeState.from_buffer_copy(bytes(scenes.state_for(eye, target, 256, 144, mode=mode)))
            a, aov_a = gpu_ctx.render(built, st, 256, 144, aov=True)
            b, aov_b = gpu_ctx.render(plain, st, 256, 144, aov=True)
            assert np.array_equal(a, b)
            for key in ("state", "voxel", "leaf", "iters"):
                assert np.array_equal(aov_a[key], aov_b[key]), key
    finally:
        built.free(), plain.free()


def test_tree_build_replicates_to_every_device(gpu_ctx):
    """wx_tree_build on a multi-device context: the tables packed on device 0 are copied to the other devices (on a
    single-GPU box: the same GPU listed three times), and the band-sharded frame equals the single-device one."""
    import torch
    n = torch.cuda.device_count()
    ids = list(range(n)) if n >= 2 else [0, 0, 0]
    s = scenes.get_scene("icosahedron")
    many = W.Context(n_devices=len(ids), device_ids=ids)
    built1 = gpu_ctx.build(topo_desc(s))
    try:
        built = many.build(topo_desc(s))
        eye, target = scenes.CAMERAS["oblique_b"]
        st = W.ComputeState.from_buffer_copy(bytes(scenes.state_for(eye, target, 512, 264, mode=3)))
        a, aov_a = gpu_ctx.render(built1, st, 512, 264, aov=True)
        b, aov_b = many.render(built, st, 512, 264, aov=True)
        c, _ = many.render(built, st, 512, 264)
        assert np.array_equal(a, b) and np.array_equal(a, c)
        for key in ("state", "voxel", "leaf", "iters"):
            assert np.array_equal(aov_a[key], aov_b[key]), key
        built.free()
    finally:
        built1.free(), many.close()


def test_tree_build_reports_wide_leaves(gpu_ctx):
    """A chain of empty leaves pushes a leaf distance above 255: wx_tree_build says so, the two-call path handles it."""
    t = O.Tree()
    t.set_voxel([0, 0, 0])
    s0 = scenes.OracleScene(t, sdf=False)
    # one N4 whose 16 leaves along z exist but only the first holds a voxel: distances grow to 8 * 15 + 7
    k4 = s0.kids4.copy()
    k4[0, 0] |= np.uint64(0xFFFF)
    vals3 = np.zeros((16, 8), np.uint64)
    vals3[0] = s0.vals3[0]
    long_row = scenes.OracleScene(O.Tree.from_topology(s0.origins, s0.kids5, s0.vals5, k4, s0.vals4, vals3))
    assert long_row.tab3[~scenes.bits2d(long_row.vals3)].max() <= 255  # not wide yet: the path must still agree
    built = gpu_ctx.build(topo_desc(long_row))
    assert list(built.info.max_dist)[2] == long_row.tab3[~scenes.bits2d(long_row.vals3)].max()
    built.free()
