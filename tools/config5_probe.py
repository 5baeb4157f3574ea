"""BASELINE config 5's launch on one GPU (64 cameras, 1920x1080, orbit over the 2048^3 sphere, ONE wx_render_device launch): kernel ms.
Knobs through tests/knobs.py (WX_LONG_FIRST=0|1, ...)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, knobs
import woxel_b200 as W
from woxel_b200 import _ffi
lib = _ffi.cuda_lib()
ctx = knobs.apply_env(W.Context())
v, flat, what, prep = bench.build_scene("sphere2048", ctx)
tree = ctx.upload(flat)
w, h, n = 1920, 1080, 64
sts = [W.ComputeState.build(W.Camera(eye=bench.orbit_eye(k, 64, 2500.0, 20.0), target=(0.5, 0.5, 0.5), aspect=w / h), w, W.RenderMode.Gray) for k in range(n)]
buf = C.c_void_p()
ctx.check(lib.wx_device_alloc(ctx._h, 0, n * w * h * 4, C.byref(buf)))
ms = []
for k in range(9):
    ctx.render_device(tree, sts, w, h, buf.value)
    ctx.check(lib.wx_stream_synchronize(ctx._h, 0, None))
    ms.append(ctx.last_render_info().kernel_ms)
print("config 5, one launch: kernel_ms", [round(m, 3) for m in ms], "launches", ctx.last_render_info().launches,
      "Mrays/s", round(n * w * h / float(np.median(ms[2:])) / 1e3, 1))
