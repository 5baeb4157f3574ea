"""Times wx_render (host buffers) for a few chunk counts: wall per call and the device-side split."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
import woxel_b200 as W
from woxel_b200 import _ffi
v, flat, what, prep = bench.build_scene("sphere2048")
import knobs; ctx = knobs.apply_env(W.Context()); tree = ctx.upload(flat); st = bench.make_state("sphere2048", 0)
lib = _ffi.cuda_lib()
nb = bench.WIDTH * bench.HEIGHT * 4
pinned = C.c_void_p(); ctx.check(lib.wx_host_alloc_pinned(nb, C.byref(pinned)))
host = np.frombuffer((C.c_uint8 * nb).from_address(pinned.value), np.uint8).reshape(1, bench.HEIGHT, bench.WIDTH, 4)
for _ in range(3): ctx.render(tree, st, bench.WIDTH, bench.HEIGHT, out=host)
t0 = time.perf_counter()
for _ in range(20): ctx.render(tree, st, bench.WIDTH, bench.HEIGHT, out=host)
dt = (time.perf_counter() - t0) / 20
i = ctx.last_render_info()
print("chunks", os.environ.get("WX_RENDER_CHUNKS", "default"), "wall_ms", round(dt * 1e3, 3), "kernels_ms", round(i.kernel_ms, 3), "total_ms", round(i.total_ms, 3), "launches", i.launches)
