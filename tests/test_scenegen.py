"""The GPU scene generator of the large fog (tools/scenegen, tests/scenegen.py -- bench infrastructure) builds the tree
the host builder builds: the dense-mask -> flat-topology assembly on the CPU, the f64 noise kernel on the GPU."""
import numpy as np
import pytest

import scenegen
import woxel_b200 as W

TAU = 0.47


def host_flat(half, tau):
    v = W.VDB345.fog(half=half, tau=tau)
    f = v.to_flat(narrow_leaves=False)
    return v, f


def dense_masks_of(f, half):
    """Scatter a flat tree's leaf masks back to the dense DFS slots the generator kernel writes."""
    hn = half // 128
    masks = np.zeros((8 * hn ** 3 * 4096, 8), np.uint64)
    k5 = np.unpackbits(np.ascontiguousarray(f.kids5).view(np.uint8), bitorder="little").reshape(f.n5, 32768).astype(bool)
    k4 = np.unpackbits(np.ascontiguousarray(f.kids4).view(np.uint8), bitorder="little").reshape(f.n4, 4096).astype(bool)
    for i, org in enumerate(f.origins):
        side = [0 if o < 0 else 1 for o in org]
        i5 = side[0] * 4 + side[1] * 2 + side[2]
        for o5 in np.flatnonzero(k5[i]):
            c = [(o5 >> 10) & 31, (o5 >> 5) & 31, o5 & 31]
            a = [ci - (0 if s else 32 - hn) for ci, s in zip(c, side)]
            assert all(0 <= ai < hn for ai in a)
            d4 = (i5 * hn ** 3 + (a[0] * hn + a[1]) * hn + a[2]) * 4096
            i4 = int(f.tab5[i, o5])
            o4 = np.flatnonzero(k4[i4])
            masks[d4 + o4] = f.vals3[f.tab4[i4, o4]]
    return masks


def assert_same_topology(t, f):
    assert np.array_equal(t["origins"], f.origins)
    assert np.array_equal(t["kids5"], f.kids5) and np.array_equal(t["kids4"], f.kids4) and np.array_equal(t["vals3"], f.vals3)
    k5 = np.unpackbits(np.ascontiguousarray(f.kids5).view(np.uint8), bitorder="little").reshape(f.n5, 32768).astype(bool)
    k4 = np.unpackbits(np.ascontiguousarray(f.kids4).view(np.uint8), bitorder="little").reshape(f.n4, 4096).astype(bool)
    assert np.array_equal(t["tab5"][k5], f.tab5[k5]) and np.array_equal(t["tab4"][k4], f.tab4[k4])
    assert not t["vals5"].any() and not t["vals4"].any()


@pytest.mark.parametrize("half", [128, 256])
def test_assembly_rebuilds_the_host_tree(half):
    v, f = host_flat(half, TAU)
    assert 0.05 < v.occupancy < 0.95
    t = scenegen.topology_from_dense_masks(dense_masks_of(f, half), half)
    assert_same_topology(t, f)


@pytest.mark.gpu
@pytest.mark.parametrize("half,tau", [(128, TAU), (256, 0.52), (256, 0.30)])
def test_gpu_generator_equals_host_builder(half, tau):
    v, f = host_flat(half, tau)
    t = scenegen.fog_topology(half, tau)
    assert t["occupancy"] == pytest.approx(v.occupancy, abs=1e-12)
    assert abs(scenegen.fog_occupancy(half, tau) - v.occupancy) < 1e-12
    assert_same_topology(t, f)


@pytest.mark.gpu
def test_generated_fog_renders_like_the_host_built_one():
    half, tau = 256, 0.5
    v, f = host_flat(half, tau)
    ctx = W.Context()
    f.compute_sdf_gpu(ctx)
    ref_tree = ctx.upload(f)
    tree = ctx.build(scenegen.desc_of(scenegen.fog_topology(half, tau)))
    st = W.ComputeState.build(W.Camera(eye=(0.5, 0.5, -700.5), target=(0.5, 0.5, 0.5), aspect=640 / 360), 640, W.RenderMode(3))
    a, aa = ctx.render(ref_tree, st, 640, 360, aov=True)
    b, ba = ctx.render(tree, st, 640, 360, aov=True)
    assert np.array_equal(a, b)
    for k in ("state", "voxel", "leaf", "iters", "mask"):
        assert np.array_equal(aa[k], ba[k]), k
    assert (aa["state"] == 0).mean() > 0.05
    tree.free(), ref_tree.free(), ctx.close()
