"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same inputs.

Bar (BASELINE.json north_star): hit voxel + leaf index agree on >= 99.9 % of pixels, depth within 1e-4
relative, RGB within 1/255.  The kernel keeps the oracle's operation order with IEEE arithmetic, so
these tests assert the stronger property -- every output bit-identical -- and print the north-star
figures when that ever fails."""
import ctypes as C

import numpy as np
import pytest

import oracle_ffi as O
import scenes
import woxel_b200 as W

pytestmark = pytest.mark.gpu

_trees = {}


def gpu_tree(ctx, name):
    if name not in _trees:
        _trees[name] = ctx.upload(scenes.get_scene(name).desc())
    return _trees[name]


def to_wx(st: O.State) -> W.ComputeState:
    return W.ComputeState.from_buffer_copy(bytes(st))


def report(rgba, aov, ref_rgba, ref_aov):
    same_hit = (aov["voxel"] == ref_aov["voxel"]).all(-1) & (aov["leaf"] == ref_aov["leaf"]) & (aov["state"] == ref_aov["state"])
    d, rd = aov["depth"].astype(np.float64), ref_aov["depth"].astype(np.float64)
    fin = np.isfinite(d) & np.isfinite(rd) & (rd != 0)
    rel = np.abs(d[fin] - rd[fin]) / np.abs(rd[fin])
    rgb = np.abs(rgba.astype(np.int32) - ref_rgba.astype(np.int32))
    bad = np.argwhere(~same_hit)
    return {
        "hit_agreement": float(same_hit.mean()), "depth_rel_max": float(rel.max()) if rel.size else 0.0,
        "rgb_max_diff": int(rgb.max()), "mismatching_pixels": bad[:20].tolist(), "n_mismatch": int(len(bad)),
    }


def _bits_nan_canonical(a):
    b = np.ascontiguousarray(a, np.float32).view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


def check_frame(ctx, name, st, w, h):
    s = scenes.get_scene(name)
    rgba, aov = ctx.render(gpu_tree(ctx, name), to_wx(st), w, h, aov=True)
    rgba, aov = rgba[0], {k: v[0] for k, v in aov.items()}
    ref_rgba, ref_aov, stats = s.gpu.render(st, w, h)
    # pixels outside the reference's dispatch (W % 8, H % 4 fringe) carry no AOV on either side
    dw, dh = (w // 8) * 8, (h // 4) * 4
    assert not rgba[dh:].any() and not rgba[:, dw:].any() and not ref_rgba[dh:].any() and not ref_rgba[:, dw:].any()
    aov = {k: v[:dh, :dw] for k, v in aov.items()}
    ref_aov = {k: v[:dh, :dw] for k, v in ref_aov.items()}
    r = report(rgba[:dh, :dw], aov, ref_rgba[:dh, :dw], ref_aov)
    # the north-star bar
    assert r["hit_agreement"] >= 0.999, r
    assert r["depth_rel_max"] <= 1e-4, r
    assert r["rgb_max_diff"] <= 1, r
    # the bar this implementation actually meets: bit-identical
    assert np.array_equal(rgba, ref_rgba), r
    for k in ("state", "voxel", "leaf", "level", "iters", "mask"):
        assert np.array_equal(aov[k], ref_aov[k]), (k, r)
    for k in ("depth", "pos"):  # bit patterns; a NaN matches a NaN (IEEE 754 leaves payload and sign of a generated NaN open)
        assert np.array_equal(_bits_nan_canonical(aov[k]), _bits_nan_canonical(ref_aov[k])), (k, r)
    return stats


@pytest.mark.parametrize("name", ["cube", "icosahedron"])
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("cam", ["default", "oblique_a", "oblique_b"])
def test_assets_all_modes(gpu_ctx, name, mode, cam):
    """BASELINE config 2 at reduced resolution (the 1080p frames are in test_assets_1080p)."""
    eye, target = scenes.CAMERAS[cam]
    st = scenes.state_for(eye, target, 480, 272, mode=mode, show_grid=(1, 1, 1) if mode < 3 else (0, 0, 0))
    check_frame(gpu_ctx, name, st, 480, 272)


@pytest.mark.parametrize("name", ["cube", "icosahedron"])
@pytest.mark.parametrize("mode", [0, 3])
def test_assets_1080p(gpu_ctx, name, mode):
    """BASELINE config 2: 1920x1080, default camera."""
    eye, target = scenes.CAMERAS["default"]
    stats = check_frame(gpu_ctx, name, scenes.state_for(eye, target, 1920, 1080, mode=mode), 1920, 1080)
    assert stats.maxed == 0 and stats.hit > 0 and stats.oob > 0


@pytest.mark.parametrize("name", ["cube", "icosahedron"])
def test_config1_640x480(gpu_ctx, name):
    """BASELINE config 1 stand-in (utahteapot.vdb is not shipped): 640x480, Diffuse, default camera."""
    eye, target = scenes.CAMERAS["default"]
    check_frame(gpu_ctx, name, scenes.state_for(eye, target, 640, 480, mode=3), 640, 480)


@pytest.mark.parametrize("name", ["single_voxel", "scattered", "small_sphere", "offcentre_sphere", "beyond_bounds", "slab",
                                  "active_tiles", "empty_leaf"])
@pytest.mark.parametrize("mode", [0, 2, 3, 4])
def test_synthetic_scenes(gpu_ctx, name, mode):
    cams = [((0.5, 0.5, -200.5), (0.5, 0.5, 0.5)), ((150.0, 90.0, -170.0), (0.0, 0.0, 0.0)), ((3.3, 2.2, 1.1), (40.0, 30.0, -20.0))]
    if name == "offcentre_sphere":
        cams.append(((300.0, 140.0, -400.0), (300.0, 140.0, -260.0)))
    for eye, target in cams:
        check_frame(gpu_ctx, name, scenes.state_for(eye, target, 320, 200, mode=mode, show_grid=(1, 1, 1)), 320, 200)


def test_edge_cases(gpu_ctx):
    # partial tiles: W % 8 != 0, H % 4 != 0 -> the undispatched fringe stays zero (wgpu_context.rs:281)
    eye, target = scenes.CAMERAS["oblique_a"]
    check_frame(gpu_ctx, "cube", scenes.state_for(eye, target, 70, 30, mode=3), 70, 30)
    check_frame(gpu_ctx, "cube", scenes.state_for(eye, target, 8, 4, mode=0), 8, 4)
    # eye outside the +-4096 world: every ray is out of bounds at once (raycast.comp.wgsl:100-103)
    check_frame(gpu_ctx, "cube", scenes.state_for((0.5, 0.5, -5000.5), (0.5, 0.5, 0.5), 128, 64, mode=4), 128, 64)
    # eye inside an active voxel: hit at i == 0, mask all false, normal = normalize(0)
    for mode in (0, 3, 4):
        check_frame(gpu_ctx, "slab", scenes.state_for((0.5, 0.5, 0.5), (10.0, 3.0, 5.0), 64, 32, mode=mode), 64, 32)
    # a 9th+ root node and nodes beyond the world bounds: the lookup precedes the bounds test
    check_frame(gpu_ctx, "beyond_bounds", scenes.state_for((4100.5, 3.5, 3.5), (0.0, 0.0, 0.0), 64, 32, mode=0), 64, 32)
    check_frame(gpu_ctx, "beyond_bounds", scenes.state_for((9000.5, 9000.5, 9000.5), (0.0, 0.0, 0.0), 64, 32, mode=3), 64, 32)


def test_exact_zero_and_negative_zero_directions(gpu_ctx):
    """dir components that are exactly +0 / -0: idir = +-inf, NaN positions -- both sides must follow IEEE alike."""
    st = scenes.state_for((0.5, 0.5, -200.5), (0.5, 0.5, 0.5), 256, 128, mode=0)
    px0 = np.float32(100) + np.float32(0.001)
    py0 = np.float32(50) + np.float32(0.001)
    for k, v in enumerate((1.0, 0.0, 0.0, 0.0)):
        st.u[k] = v
    for k, v in enumerate((0.0, -1.0, 0.0, 0.0)):
        st.mv[k] = v
    for k, v in enumerate((-float(px0), float(py0), 200.0, 0.0)):
        st.wp[k] = v
    for mode in (0, 2, 3):
        st.render_mode[0] = mode
        check_frame(gpu_ctx, "small_sphere", st, 256, 128)
    st.u[0], st.mv[0], st.wp[0] = -0.0, -0.0, -0.0  # dir.x == -0.0 on every pixel
    st.u[1] = 0.5
    for mode in (0, 4):
        st.render_mode[0] = mode
        check_frame(gpu_ctx, "small_sphere", st, 256, 128)


def test_max_steps_state(gpu_ctx):
    """Rays grazing a 1200-voxel slab one cell above its surface run out of the 1000-step budget (state 2)."""
    st = scenes.state_for((-599.5, 4.5, 0.5), (600.0, 4.5, 0.5), 128, 64, mode=0)
    stats = check_frame(gpu_ctx, "long_slab", st, 128, 64)
    assert stats.maxed > 0, "scene/camera no longer exercises HDDA_MAX_RAY_STEPS"
    st.render_mode[0] = 4
    check_frame(gpu_ctx, "long_slab", st, 128, 64)


def test_product_host_path_matches_oracle(gpu_ctx):
    """Reference-facing surface end to end: VDB345 (procedural) -> compute_sdf -> Renderer.render vs oracle."""
    v = W.VDB345.sphere(half=128, radius=100.0, band=2.0)
    r = W.Renderer(320, 200)
    r.change_vdb_model(v)  # the product's model load: to_flat + wx_compute_sdf (GPU sweep) + wx_tree_upload
    assert r.last_sdf.device_ms > 0 and list(r.last_sdf.max_dist)[2] > 0
    v.compute_sdf()  # host sweep: the tables the oracle renders from
    f = v.to_flat(narrow_leaves=False)
    g = O.gpudata_from_tables(f.origins, f.kids5, f.vals5, f.tab5, f.kids4, f.vals4, f.tab4, f.vals3, f.tab3)
    for mode in (W.RenderMode.Diffuse, W.RenderMode.Gray, W.RenderMode.Glossy):
        r.render_mode = mode
        sc = W.Scene(320, 200, camera=W.Camera(eye=(0.5, 0.5, -300.5), target=(0.5, 0.5, 0.5), aspect=320 / 200))
        img = r.render(sc)
        st = scenes.state_for(sc.camera.eye, sc.camera.target, 320, 200, mode=int(mode))
        ref, _, _ = g.render(st, 320, 200, aov=False)
        assert np.array_equal(img, ref), int(mode)
    r.change_vdb_model(v, compute_sdf=False)  # v now carries the host sweep's distances: same frames
    assert r.last_sdf.device_ms == 0
    assert np.array_equal(r.render(sc), img)


def test_camera_batch_and_shards_equal_single_frames(gpu_ctx):
    """n_states > 1 and row-band shards write exactly what single full-frame calls write."""
    from woxel_b200 import _ffi
    name, w, h = "icosahedron", 256, 136
    tree = gpu_tree(gpu_ctx, name)
    states = []
    for k in range(5):
        th = 2 * np.pi * k / 5
        eye = (300 * np.sin(th) + 0.5, 60.5, -300 * np.cos(th) + 0.5)
        states.append(to_wx(scenes.state_for(eye, (0.5, 0.5, 0.5), w, h, mode=[3, 3, 0, 0, 4][k])))
    singles = np.stack([gpu_ctx.render(tree, s, w, h)[0][0] for s in states])
    batch, _ = gpu_ctx.render(tree, states, w, h)
    assert np.array_equal(batch, singles)
    # shards: 3 "ranks" with 8-row bands into one device buffer
    lib = _ffi.cuda_lib()
    buf = C.c_void_p()
    nbytes = len(states) * w * h * 4
    gpu_ctx.check(lib.wx_device_alloc(gpu_ctx._h, 0, nbytes, C.byref(buf)))
    try:
        for idx in range(3):
            gpu_ctx.render_device(tree, states, w, h, buf.value, shard=(idx, 3, 8))
        out = np.zeros((len(states), h, w, 4), np.uint8)
        gpu_ctx.check(lib.wx_stream_synchronize(gpu_ctx._h, 0, None))
        gpu_ctx.check(lib.wx_memcpy_d2h(gpu_ctx._h, 0, out.ctypes.data, buf, nbytes, None))
        gpu_ctx.check(lib.wx_stream_synchronize(gpu_ctx._h, 0, None))
        assert np.array_equal(out, singles)
    finally:
        lib.wx_device_free(gpu_ctx._h, 0, buf)


def test_upload_validation(gpu_ctx):
    s = scenes.get_scene("single_voxel")
    from woxel_b200.render import make_desc
    bad = s.tab5.copy()
    bad[0, 0] = 77  # child index beyond n4
    with pytest.raises(W.WxError) as e:
        gpu_ctx.upload(make_desc(s.origins, s.kids5, s.vals5, bad, s.kids4, s.vals4, s.tab4, s.vals3, s.tab3))
    assert e.value.status == -5  # WX_ERR_BAD_TREE
    big = s.tab4.copy()
    big[0, 1] = 0x90000000  # distance >= 2^31
    with pytest.raises(W.WxError) as e:
        gpu_ctx.upload(make_desc(s.origins, s.kids5, s.vals5, s.tab5, s.kids4, s.vals4, big, s.vals3, s.tab3))
    assert e.value.status == -6  # WX_ERR_UNSUPPORTED
    # wide leaf distances select the 8-bit / 32-bit brick layouts and still render identically
    for extra, bits in ((200, 8), (70000, 32)):
        t3 = s.tab3.copy()
        t3[0, 0] = extra  # voxel (0,0,0) of the only leaf is inactive
        t = gpu_ctx.upload(make_desc(s.origins, s.kids5, s.vals5, s.tab5, s.kids4, s.vals4, s.tab4, s.vals3, t3))
        assert t.info.leaf_bits == bits
        g = O.gpudata_from_tables(s.origins, s.kids5, s.vals5, s.tab5, s.kids4, s.vals4, s.tab4, s.vals3, t3)
        st = scenes.state_for((20.5, 20.5, -30.5), (4.0, 4.0, 4.0), 128, 64, mode=0)
        rgba, aov = gpu_ctx.render(t, to_wx(st), 128, 64, aov=True)
        ref, ref_aov, _ = g.render(st, 128, 64)
        assert np.array_equal(rgba[0], ref) and np.array_equal(aov["iters"][0], ref_aov["iters"])
        t.free()


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def test_pipelined_readback_equals_single_launch(gpu_ctx):
    """wx_render without AOVs reads the frame back in chunks while later chunks render (>= 1 Mpx, or a camera
    batch); the pixels must be those of the single-launch path (the AOV call) and of the oracle."""
    name, w, h = "icosahedron", 1280, 824  # > 2^20 pixels, height not a multiple of the chunk height
    tree = gpu_tree(gpu_ctx, name)
    eye, target = scenes.CAMERAS["oblique_a"]
    for mode in (0, 3):
        st = scenes.state_for(eye, target, w, h, mode=mode)
        plain, _ = gpu_ctx.render(tree, to_wx(st), w, h)           # chunked + pipelined
        single, _ = gpu_ctx.render(tree, to_wx(st), w, h, aov=True)  # one launch (+ the long-tile kernel once this geometry has run before)
        assert gpu_ctx.last_render_info().launches in (1, 2)
        assert np.array_equal(plain, single)
        ref, _, _ = scenes.get_scene(name).gpu.render(st, w, h, aov=False)
        assert np.array_equal(plain[0], ref)


@pytest.mark.parametrize("shape", [(1280, 824, 1), (200, 120, 3)])
def test_wx_render_delivers_to_device_memory(gpu_ctx, shape):
    """wx_render's destination may be device memory (how bench.py gathers the ranks' frames on GPU 0): the chunked,
    pipelined delivery must put the same bytes there as into a host buffer."""
    import ctypes as C
    from woxel_b200 import _ffi
    lib = _ffi.cuda_lib()
    w, h, n = shape
    tree = gpu_tree(gpu_ctx, "icosahedron")
    cams = [scenes.CAMERAS["oblique_a"], scenes.CAMERAS["default"], scenes.CAMERAS["oblique_b"]]
    states = [to_wx(scenes.state_for(*cams[k], w, h, mode=(3, 0, 4)[k])) for k in range(n)]
    host, _ = gpu_ctx.render(tree, states, w, h)
    nbytes = n * w * h * 4
    dev = C.c_void_p()
    gpu_ctx.check(lib.wx_device_alloc(gpu_ctx._h, 0, nbytes, C.byref(dev)))
    try:
        gpu_ctx.render_to(tree, states, w, h, dev.value)
        back = np.empty((n, h, w, 4), np.uint8)
        gpu_ctx.check(lib.wx_memcpy_d2h(gpu_ctx._h, 0, back.ctypes.data, dev, nbytes, None))
        gpu_ctx.check(lib.wx_stream_synchronize(gpu_ctx._h, 0, None))
        assert np.array_equal(back, host)
    finally:
        lib.wx_device_free(gpu_ctx._h, 0, dev)


def _multi_contexts():
    """Device lists for the multi-device tests: 2 and all GPUs of the box; on a single-GPU box the same GPU listed 2 and
    3 times (wx_init gives every entry its own streams, frame and tree replica, so the sharding, the per-device
    read-back and the gather run exactly as on several GPUs)."""
    n = _device_count()
    if n >= 2:
        return [list(range(k)) for k in sorted({2, n})]
    return [[0, 0], [0, 0, 0]]


@pytest.mark.parametrize("which", [0, 1])
def test_multi_device_context_equals_single_device(which):
    """One frame / one camera batch split by row bands over every GPU of the box, stored into device 0's frame over
    NVLink by the kernels themselves: identical to the single-device frame."""
    lists = _multi_contexts()
    ids = lists[min(which, len(lists) - 1)]
    n = len(ids)
    name, w, h = "icosahedron", 640, 360
    s = scenes.get_scene(name)
    one = W.Context()
    many = W.Context(n_devices=n, device_ids=ids)
    try:
        assert many.device_count == n
        t1, tn = one.upload(s.desc()), many.upload(s.desc())
        states = [to_wx(scenes.state_for(e, t, w, h, mode=m)) for (e, t), m in
                  zip([scenes.CAMERAS["default"], scenes.CAMERAS["oblique_a"], scenes.CAMERAS["oblique_b"]], (0, 3, 4))]
        a, aov_a = one.render(t1, states, w, h, aov=True)
        b, aov_b = many.render(tn, states, w, h, aov=True)
        assert np.array_equal(a, b)
        for k in ("state", "voxel", "leaf", "iters"):
            assert np.array_equal(aov_a[k], aov_b[k]), k
        c, _ = many.render(tn, states[1], w, h)
        assert np.array_equal(c[0], a[1])
        t1.free(), tn.free()
    finally:
        one.close(), many.close()


@pytest.mark.parametrize("shape", [(640, 360, 1), (328, 203, 5), (1280, 824, 1), (96, 44, 19), (64, 4, 2)])
def test_multi_device_distributed_readback(shape):
    """Without AOVs every GPU reads its own row bands back over its own PCIe link (frames, heights that are not a
    multiple of the band height, fewer bands than GPUs, more cameras than pipeline chunks); wx_capture_srgb afterwards
    completes device 0's frame over NVLink.  Everything must equal the single-device results."""
    w, h, n_cam = shape
    s = scenes.get_scene("icosahedron")
    cams = [scenes.CAMERAS["default"], scenes.CAMERAS["oblique_a"], scenes.CAMERAS["oblique_b"]]
    states = [to_wx(scenes.state_for(*cams[k % 3], w, h, mode=(0, 3, 4, 1, 2)[k % 5])) for k in range(n_cam)]
    one = W.Context()
    try:
        t1 = one.upload(s.desc())
        a, _ = one.render(t1, states, w, h)
        a_rgb = one.capture_srgb(n_cam, w, h)
        t1.free()
    finally:
        one.close()
    for ids in _multi_contexts():
        n = len(ids)
        many = W.Context(n_devices=n, device_ids=ids)
        try:
            tn = many.upload(s.desc())
            out = np.full((n_cam, h, w, 4), 0xAB, np.uint8)
            b, _ = many.render(tn, states, w, h, out=out)
            assert np.array_equal(a, b), f"{n} devices"
            assert np.array_equal(many.capture_srgb(n_cam, w, h), a_rgb), f"{n} devices: capture after the lazy gather"
            b2, _ = many.render(tn, states, w, h)  # a second call reuses the per-device frames
            assert np.array_equal(a, b2)
            tn.free()
        finally:
            many.close()


def _product_scene(kind):
    """A procedural scene of the benchmark configs built by the PRODUCT host (set_voxel tree -> compute_sdf -> to_flat);
    the oracle renders from the same tables (its own compute_sdf is compared with the product's in test_host_vs_oracle)."""
    if kind == "torus":  # BASELINE config 3 at 1/8 scale
        v = W.VDB345.torus(half=128, major=88.0, minor=31.0, band=2.0)
    elif kind == "fog":  # BASELINE config 4 at 1/16 scale: dense value-noise fog, high leaf occupancy
        v = W.VDB345.fog(half=64, tau=0.32)
        assert 0.05 < v.occupancy < 0.95
    else:
        raise KeyError(kind)
    v.compute_sdf()
    f = v.to_flat(narrow_leaves=False)
    g = O.gpudata_from_tables(f.origins, f.kids5, f.vals5, f.tab5, f.kids4, f.vals4, f.tab4, f.vals3, f.tab3)
    return f, g


@pytest.mark.parametrize("kind", ["torus", "fog"])
def test_procedural_config_scenes(gpu_ctx, kind):
    f, g = _product_scene(kind)
    tree = gpu_ctx.upload(f)
    w, h = 384, 216
    cams = [((0.5, 0.5, -320.5), (0.5, 0.5, 0.5)), ((210.0, 140.0, -230.0), (0.0, 0.0, 0.0))]
    if kind == "fog":
        cams.append(((3.5, 2.5, 1.5), (60.0, 40.0, 50.0)))  # inside the volume: long divergent rays
    try:
        for eye, target in cams:
            for mode in (0, 2, 3, 4):
                st = scenes.state_for(eye, target, w, h, mode=mode)
                rgba, aov = gpu_ctx.render(tree, to_wx(st), w, h, aov=True)
                ref, ref_aov, _ = g.render(st, w, h)
                assert np.array_equal(rgba[0], ref), (kind, eye, mode)
                for k in ("state", "voxel", "leaf", "level", "iters", "mask"):
                    assert np.array_equal(aov[k][0], ref_aov[k]), (kind, eye, mode, k)
    finally:
        tree.free()


def test_orbit_camera_batch_config5(gpu_ctx):
    """BASELINE config 5 in small: a 16-camera orbit rendered as one batch == the per-camera frames == the oracle."""
    import bench
    name, w, h = "small_sphere", 192, 108
    tree = gpu_tree(gpu_ctx, name)
    s = scenes.get_scene(name)
    sts = []
    for k in range(16):
        th = 2 * np.pi * k / 16
        eye = (0.5 + 150 * np.cos(np.radians(20)) * np.sin(th), 0.5 + 150 * np.sin(np.radians(20)), 0.5 - 150 * np.cos(np.radians(20)) * np.cos(th))
        sts.append(scenes.state_for(eye, (0.5, 0.5, 0.5), w, h, mode=0))
    batch, _ = gpu_ctx.render(tree, [to_wx(st) for st in sts], w, h)
    for k in (0, 5, 11, 15):
        ref, _, _ = s.gpu.render(sts[k], w, h, aov=False)
        assert np.array_equal(batch[k], ref), k
    assert bench.orbit_eye(0) == (0.5, 0.5, -2500.5)


def test_persistent_work_queue_kernel_is_bit_identical():
    """WX_KERNEL=persistent selects the persistent-thread work-queue kernel (read once per process): the same
    parity tests must pass through it."""
    import os
    import subprocess
    import sys
    if os.environ.get("WX_KERNEL") == "persistent":
        pytest.skip("already running under the persistent kernel")
    env = dict(os.environ, WX_KERNEL="persistent")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_parity_gpu.py"), "-m", "gpu", "-x", "-q", "-k",
                        "assets_all_modes or edge_cases or camera_batch_and_shards or synthetic_scenes or zero_directions"],
                       env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_cta_queue_kernel_is_bit_identical():
    """WX_OPT_KERNEL = 2: the persistent kernel with a CTA-level chunk queue (wx_raycast.cu).  Written at the end of round 1, first
    run on a GPU in round 2 (profiles/r2_cta_queue.txt: bit-identical, 2.1x slower than the tiled grid -- kept selectable, not the
    default).  The parity subset below runs under it in a child process (tests/knobs.py turns WX_KERNEL into the option)."""
    import os
    import subprocess
    import sys
    if os.environ.get("WX_KERNEL") == "persistent_cta":
        pytest.skip("already running under the CTA-queue kernel")
    env = dict(os.environ, WX_KERNEL="persistent_cta")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_parity_gpu.py"), "-m", "gpu", "-x", "-q", "-k",
                        "assets_all_modes or edge_cases or camera_batch_and_shards or synthetic_scenes or zero_directions"],
                       env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_randomized_scenes_and_cameras(gpu_ctx):
    """Differential fuzz (seeded): random sparse trees incl. far-apart and out-of-world N5s, random cameras (inside
    the volume, on integer coordinates, axis-aligned, outside the +-4096 world), every render mode -- bit-identical
    frames and AOVs.  Exercises the exactness arguments of march_fast (floor trick at lattice planes, ties, nudges)."""
    rng = np.random.default_rng(20240229)
    from woxel_b200.render import make_desc
    for case in range(12):
        t = O.Tree()
        n_blobs = int(rng.integers(1, 6))
        for _ in range(n_blobs):
            centre = rng.integers(-900, 900, 3) if rng.random() < 0.8 else rng.integers(-6000, 6000, 3)
            ext = int(rng.integers(1, 40))
            pts = centre + rng.integers(-ext, ext + 1, size=(int(rng.integers(1, 400)), 3))
            t.set_voxels(pts.astype(np.int32))
        if case % 3 == 0:  # a solid block: exact ties on flat faces
            c0 = rng.integers(-100, 100, 3)
            ax = [np.arange(c, c + int(rng.integers(2, 20))) for c in c0]
            g = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
            t.set_voxels(g.astype(np.int32))
        s = scenes.OracleScene(t)
        tree = gpu_ctx.upload(make_desc(s.origins, s.kids5, s.vals5, s.tab5, s.kids4, s.vals4, s.tab4, s.vals3, s.tab3))
        try:
            for k in range(6):
                kind = k % 6
                if kind == 0:
                    eye = tuple(rng.uniform(-1500, 1500, 3))
                elif kind == 1:
                    eye = tuple(float(v) for v in rng.integers(-300, 300, 3))  # exactly on lattice planes
                elif kind == 2:
                    eye = (float(rng.integers(-50, 50)) + 0.5, float(rng.integers(-50, 50)) + 0.5, -700.5)  # axis-aligned view
                elif kind == 3:
                    eye = tuple(rng.uniform(-4090, 4090, 3))
                elif kind == 4:
                    eye = tuple(rng.uniform(4000, 4300, 3))  # partly outside the world
                else:
                    eye = tuple(rng.uniform(-60, 60, 3))     # inside / next to the voxels
                target = (eye[0], eye[1], eye[2] + 100.0) if kind == 2 else tuple(rng.uniform(-200, 200, 3))
                mode = int(rng.integers(0, 5))
                w, h = 64, 32
                st = scenes.state_for(eye, target, w, h, mode=mode, show_grid=(1, 1, 1))
                rgba, aov = gpu_ctx.render(tree, to_wx(st), w, h, aov=True)
                ref, ref_aov, _ = s.gpu.render(st, w, h)
                assert np.array_equal(rgba[0], ref), (case, k, eye, target, mode)
                for name in ("state", "voxel", "leaf", "level", "iters", "mask"):
                    assert np.array_equal(aov[name][0], ref_aov[name]), (case, k, name, eye, target, mode)
                assert np.array_equal(_bits_nan_canonical(aov["pos"][0]), _bits_nan_canonical(ref_aov["pos"])), (case, k)
        finally:
            tree.free()


def test_large_camera_batch_mixed_modes(gpu_ctx):
    """130 states: the pipelined read-back groups cameras 3 per chunk, with render-mode changes inside a chunk and a
    partial last chunk; every frame equals its single-frame render."""
    name, w, h = "small_sphere", 64, 32
    tree = gpu_tree(gpu_ctx, name)
    rng = np.random.default_rng(9)
    states = []
    for k in range(130):
        th = 2 * np.pi * k / 130
        eye = (0.5 + 170 * np.sin(th), 20.5 + 0.1 * k, 0.5 - 170 * np.cos(th))
        states.append(to_wx(scenes.state_for(eye, (0.5, 0.5, 0.5), w, h, mode=int(rng.integers(0, 5)))))
    batch, _ = gpu_ctx.render(tree, states, w, h)
    assert gpu_ctx.last_render_info().launches >= 44  # 44 chunks, more where the mode changes inside one
    for k in (0, 1, 2, 3, 64, 65, 127, 128, 129):
        single, _ = gpu_ctx.render(tree, states[k], w, h)
        assert np.array_equal(batch[k], single[0]), k
    ref, _, _ = scenes.get_scene(name).gpu.render(scenes.state_for((0.5, 20.5, -169.5), (0.5, 0.5, 0.5), w, h, mode=int(states[0].render_mode[0])), w, h, aov=False)
    assert np.array_equal(batch[0], ref)


@pytest.mark.parametrize("mode", [1, 2, 4])
def test_assets_1080p_remaining_modes(gpu_ctx, mode):
    """BASELINE config 2 at full size for the modes test_assets_1080p leaves out."""
    eye, target = scenes.CAMERAS["oblique_a"]
    check_frame(gpu_ctx, "icosahedron", scenes.state_for(eye, target, 1920, 1080, mode=mode, show_grid=(1, 1, 1)), 1920, 1080)


def test_config3_full_size_torus(gpu_ctx):
    """BASELINE config 3 at its full size: the procedural 2048^3 torus at 3840x2160.  Model load through wx_tree_build
    (GPU sweep), one frame against the oracle (rendering from the host sweep's tables), and the size-independent
    properties: rendering twice gives the same bytes; three row-band shards reassemble the frame; a 2-camera batch
    equals two single frames."""
    import ctypes as C
    from woxel_b200 import _ffi
    w, h = 3840, 2160
    v = W.VDB345.torus()
    tree = gpu_ctx.build(v.to_flat(narrow_leaves=False))
    v.compute_sdf()
    f = v.to_flat(narrow_leaves=False)
    g = O.gpudata_from_tables(f.origins, f.kids5, f.vals5, f.tab5, f.kids4, f.vals4, f.tab4, f.vals3, f.tab3)
    try:
        assert list(tree.info.max_dist) == [int(f.tab5[~scenes.bits2d(f.kids5)].max()), int(f.tab4[~scenes.bits2d(f.kids4)].max()),
                                             int(f.tab3[~scenes.bits2d(f.vals3)].max())]
        st = scenes.state_for((0.5, 0.5, -2500.5), (0.5, 0.5, 0.5), w, h, mode=0)
        a, _ = gpu_ctx.render(tree, to_wx(st), w, h)
        ref, _, stats = g.render(st, w, h, aov=False)
        assert stats.hit > 500000 and stats.maxed == 0
        assert np.array_equal(a[0], ref)
        b, _ = gpu_ctx.render(tree, to_wx(st), w, h)
        assert np.array_equal(a, b)
        st2 = scenes.state_for((1800.5, 900.5, -1700.5), (0.5, 0.5, 0.5), w, h, mode=3)
        pair, _ = gpu_ctx.render(tree, [to_wx(st), to_wx(st2)], w, h)
        assert np.array_equal(pair[0], a[0])
        single2, _ = gpu_ctx.render(tree, to_wx(st2), w, h)
        assert np.array_equal(pair[1], single2[0])
        lib = _ffi.cuda_lib()
        buf = C.c_void_p()
        gpu_ctx.check(lib.wx_device_alloc(gpu_ctx._h, 0, w * h * 4, C.byref(buf)))
        try:
            for idx in range(3):
                gpu_ctx.render_device(tree, to_wx(st), w, h, buf.value, shard=(idx, 3, 16))
            out = np.zeros((h, w, 4), np.uint8)
            gpu_ctx.check(lib.wx_stream_synchronize(gpu_ctx._h, 0, None))
            gpu_ctx.check(lib.wx_memcpy_d2h(gpu_ctx._h, 0, out.ctypes.data, buf, w * h * 4, None))
            gpu_ctx.check(lib.wx_stream_synchronize(gpu_ctx._h, 0, None))
            assert np.array_equal(out, a[0])
        finally:
            lib.wx_device_free(gpu_ctx._h, 0, buf)
    finally:
        tree.free()
