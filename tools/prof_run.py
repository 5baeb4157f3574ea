"""Small driver for ncu: builds a scene, renders a few frames of the hot path (no torch, no oracle).

  ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 2 -c 1 -o gpurun_out/prof python tools/prof_run.py
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import woxel_b200 as W  # noqa: E402
from woxel_b200 import _ffi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="sphere2048")
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--mode", type=int, default=0)
a = ap.parse_args()
v, flat, what, prep = bench.build_scene(a.scene)
import knobs  # noqa: E402
ctx = knobs.apply_env(W.Context())
tree = ctx.upload(flat)
st = bench.make_state(a.scene, 0)
st.render_mode[0] = a.mode
lib = _ffi.cuda_lib()
buf = C.c_void_p()
ctx.check(lib.wx_device_alloc(ctx._h, 0, bench.WIDTH * bench.HEIGHT * 4, C.byref(buf)))
ms = []
for _ in range(a.frames):
    ctx.render_device(tree, st, bench.WIDTH, bench.HEIGHT, buf.value)
    ctx.check(lib.wx_stream_synchronize(ctx._h, 0, None))
    ms.append(ctx.last_render_info().kernel_ms)
import numpy as np  # noqa: E402
frame = np.zeros((bench.HEIGHT, bench.WIDTH), np.uint32)
ctx.check(lib.wx_memcpy_d2h(ctx._h, 0, frame.ctypes.data, buf, frame.nbytes, None))
ctx.check(lib.wx_stream_synchronize(ctx._h, 0, None))
checksum = int(frame.sum(dtype=np.uint64))  # equal checksums of two builds on the same scene/mode = (almost surely) equal frames
rays = (bench.WIDTH // 8 * 8) * (bench.HEIGHT // 4 * 4)
best = sorted(ms[1:])[len(ms[1:]) // 2] if len(ms) > 1 else ms[0]
print("lib", os.environ.get("WOXEL_B200_LIB", "default"), "scene", a.scene, "mode", a.mode, "median_ms", round(best, 4),
      "Mrays/s", round(rays / best / 1e3, 1), "checksum", checksum, "all_ms", [round(m, 3) for m in ms])
