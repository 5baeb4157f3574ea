"""Strong scaling of ONE frame over the GPUs of the box through wx_render with a multi-device context (row bands dealt
round-robin, peer stores into device 0's frame): wall ms per 4K frame for 1..N devices; frames compared with 1 device.
With --orbit: BASELINE config 5, the 64-camera 1080p orbit (elevation 20 deg) over the sphere as ONE wx_render call with 64
states, sharded over the devices and gathered in device 0's frame buffer, then read back (531 MB) to pinned host memory."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
import woxel_b200 as W
from woxel_b200 import _ffi

ORBIT = "--orbit" in sys.argv
n_all = torch.cuda.device_count()
orbit_ref = None


def orbit(ctx, tree, n):
    global orbit_ref
    w, h, cams = 1920, 1080, 64
    sts = [W.ComputeState.build(W.Camera(eye=bench.orbit_eye(k, cams, 2499.0, 20.0), target=(0.5, 0.5, 0.5), aspect=w / h), w, W.RenderMode(0))
           for k in range(cams)]
    nb = cams * w * h * 4
    pinned = C.c_void_p(); ctx.check(lib.wx_host_alloc_pinned(nb, C.byref(pinned)))
    host = np.frombuffer((C.c_uint8 * nb).from_address(pinned.value), np.uint8).reshape(cams, h, w, 4)
    for _ in range(2): ctx.render(tree, sts, w, h, out=host)
    wall, kms, tot = [], [], []
    for _ in range(7):
        t0 = time.perf_counter(); ctx.render(tree, sts, w, h, out=host); wall.append(time.perf_counter() - t0)
        i = ctx.last_render_info(); kms.append(i.kernel_ms); tot.append(i.total_ms)
    if orbit_ref is None: orbit_ref = host.copy()
    dt = float(np.median(wall))
    print(f"orbit 64 x 1080p, devices {n}: wall {dt*1e3:.2f} ms/batch, slowest device's kernels {np.median(kms):.2f} ms, device total {np.median(tot):.2f} ms, "
          f"{cams*w*h/np.median(kms)/1e3:.0f} Mrays/s kernels, {cams*w*h/dt/1e6:.0f} Mrays/s e2e, identical to 1 device: {bool(np.array_equal(host, orbit_ref))}", flush=True)
    lib.wx_host_free_pinned(pinned)

lib = _ffi.cuda_lib()
v = W.VDB345.sphere()
ref = None
for n in [k for k in (1, 2, 4, 8) if k <= n_all]:
    ctx = W.Context(n_devices=n)
    f = v.to_flat(narrow_leaves=False)
    f.compute_sdf_gpu(ctx)
    tree = ctx.upload(f)
    st = bench.make_state("sphere2048", 0)
    if ORBIT:
        orbit(ctx, tree, n)
        tree.free(); ctx.close()
        continue
    nb = bench.WIDTH * bench.HEIGHT * 4
    pinned = C.c_void_p(); ctx.check(lib.wx_host_alloc_pinned(nb, C.byref(pinned)))
    host = np.frombuffer((C.c_uint8 * nb).from_address(pinned.value), np.uint8).reshape(1, bench.HEIGHT, bench.WIDTH, 4)
    for _ in range(3): ctx.render(tree, st, bench.WIDTH, bench.HEIGHT, out=host)
    t0 = time.perf_counter()
    for _ in range(20): ctx.render(tree, st, bench.WIDTH, bench.HEIGHT, out=host)
    dt = (time.perf_counter() - t0) / 20
    i = ctx.last_render_info()
    if ref is None: ref = host.copy()
    print(f"devices {n}: wall {dt*1e3:.3f} ms/frame, slowest kernel {i.kernel_ms:.3f} ms, device total {i.total_ms:.3f} ms, "
          f"{bench.WIDTH*bench.HEIGHT/dt/1e6:.0f} Mrays/s e2e, identical to 1 device: {bool(np.array_equal(host, ref))}")
    lib.wx_host_free_pinned(pinned)
    tree.free(); ctx.close()
