"""Kernel time of ONE rank's share of a frame that is tile-partitioned over n GPUs (shard 0 of n, 8-row bands), on one GPU:
what rank 0 of `bench.py --gpus n`'s strong_single_frame record spends in its kernel(s).  For tuning without an n-GPU box.

  python tools/shard_probe.py [--scene sphere2048] [--shards 1 2 4 8]        (knobs: WX_LONG_FIRST, WX_LONG_THRESHOLD, ...)"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import knobs  # noqa: E402
import woxel_b200 as W  # noqa: E402
from woxel_b200 import _ffi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="sphere2048")
ap.add_argument("--shards", type=int, nargs="*", default=[1, 2, 4, 8])
ap.add_argument("--frames", type=int, default=24)
a = ap.parse_args()
v, flat, what, prep = bench.build_scene(a.scene)
ctx = knobs.apply_env(W.Context())
tree = ctx.upload(flat)
st = bench.make_state(a.scene, 0)
lib = _ffi.cuda_lib()
buf = C.c_void_p()
ctx.check(lib.wx_device_alloc(ctx._h, 0, bench.WIDTH * bench.HEIGHT * 4, C.byref(buf)))
base = None
for n in a.shards:
    ms = []
    for _ in range(a.frames):
        ctx.render_device(tree, st, bench.WIDTH, bench.HEIGHT, buf.value, shard=(0, n, 8) if n > 1 else None)
        ctx.check(lib.wx_stream_synchronize(ctx._h, 0, None))
        ms.append(ctx.last_render_info().kernel_ms)
    med = sorted(ms[4:])[len(ms[4:]) // 2]
    base = base or med
    print(f"shard 0/{n}: kernel_ms {med:.4f}  speed-up vs whole frame {base / med:.2f}x  launches {ctx.last_render_info().launches}", flush=True)
