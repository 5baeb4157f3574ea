"""Kernel times of the BASELINE.json configurations that are not the bench line (one B200): median of 9 launches
after 2 warm-ups, CUDA events around wx_render_device (device-resident output).  Prints one JSON object per line."""
import ctypes as C, json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
import scenes
import woxel_b200 as W
from woxel_b200 import _ffi

lib = _ffi.cuda_lib()
import knobs
ctx = knobs.apply_env(W.Context())


def upload_product(v):
    f = v.to_flat(narrow_leaves=False)
    info = f.compute_sdf_gpu(ctx)
    return ctx.upload(f), f, info


def timed(tree, states, w, h, reps=9):
    n = len(states)
    buf = C.c_void_p()
    ctx.check(lib.wx_device_alloc(ctx._h, 0, n * w * h * 4, C.byref(buf)))
    ms = []
    for k in range(reps + 2):
        ctx.render_device(tree, states, w, h, buf.value)
        ctx.check(lib.wx_stream_synchronize(ctx._h, 0, None))
        if k >= 2:
            ms.append(ctx.last_render_info().kernel_ms)
    lib.wx_device_free(ctx._h, 0, buf)
    return float(np.median(ms))


def state(eye, target, w, h, mode):
    return W.ComputeState.build(W.Camera(eye=eye, target=target, aspect=w / h), w, W.RenderMode(mode))


def report(config, scene, w, h, mode, n_frames, ms, extra=None):
    rays = n_frames * (w // 8 * 8) * (h // 4 * 4)
    out = {"config": config, "scene": scene, "frames": n_frames, "width": w, "height": h, "mode": mode, "kernel_ms": round(ms, 4),
           "primary_Mrays_per_s": round(rays / ms / 1e3, 1)}
    out.update(extra or {})
    print(json.dumps(out), flush=True)


# config 2: shipped assets at 1080p, all modes (trees from the golden topologies, SDF on the GPU)
for name in ("cube", "icosahedron"):
    v = scenes.host_tree_from_scene(bench._TopoView(scenes.load_topo(name)))
    tree, f, info = upload_product(v)
    for mode in (0, 1, 2, 3, 4):
        ms = timed(tree, [state((0.5, 0.5, -500.5), (0.5, 0.5, -498.5), 1920, 1080, mode)], 1920, 1080)
        report(2, f"assets/{name}.vdb", 1920, 1080, mode, 1, ms, {"sdf_gpu_ms": round(info.device_ms, 2)})
    if name == "cube":  # config 1 stand-in
        report(1, f"assets/{name}.vdb (teapot stand-in)", 640, 480, 3, 1,
               timed(tree, [state((0.5, 0.5, -500.5), (0.5, 0.5, -498.5), 640, 480, 3)], 640, 480))
    tree.free()

# config 3: 2048^3 sphere and torus level sets at 4K, modes 0 / 3 / 4
for name, v in (("sphere", W.VDB345.sphere()), ("torus", W.VDB345.torus())):
    tree, f, info = upload_product(v)
    for mode in (0, 3, 4):
        ms = timed(tree, [state((0.5, 0.5, -2500.5), (0.5, 0.5, 0.5), 3840, 2160, mode)], 3840, 2160)
        report(3, f"procedural 2048^3 {name} (n4={f.n4}, n3={f.n3})", 3840, 2160, mode, 1, ms, {"sdf_gpu_ms": round(info.device_ms, 2)})
    if name == "sphere":  # config 5 on one GPU: 64-camera 1080p orbit as one batch
        sts = []
        for k in range(64):
            th = 2 * math.pi * k / 64
            el = math.radians(20)
            eye = (0.5 + 2500 * math.cos(el) * math.sin(th), 0.5 + 2500 * math.sin(el), 0.5 - 2500 * math.cos(el) * math.cos(th))
            sts.append(state(eye, (0.5, 0.5, 0.5), 1920, 1080, 0))
        ms = timed(tree, sts, 1920, 1080, reps=5)
        report(5, "64-camera 1080p orbit over the sphere, one launch", 1920, 1080, 0, 64, ms)
    tree.free()

# config 4 at 1/4 scale: value-noise fog built by the HOST builder over 512^3 (the full 2048^3 volume is generated on
# the GPU and measured by tools/fog_bench.py), 4K
t0 = time.time()
v = W.VDB345.fog(half=256, tau=0.32)
tree, f, info = upload_product(v)
for eye, target, tag in (((0.5, 0.5, -700.5), (0.5, 0.5, 0.5), "outside"), ((3.5, 2.5, 1.5), (200.0, 120.0, 160.0), "inside the volume")):
    for mode in (0, 3):
        ms = timed(tree, [state(eye, target, 3840, 2160, mode)], 3840, 2160)
        report(4, f"value-noise fog 512^3, occupancy {v.occupancy:.3f} (n4={f.n4}, n3={f.n3}), camera {tag}", 3840, 2160, mode, 1, ms,
               {"sdf_gpu_ms": round(info.device_ms, 2)})
tree.free()
