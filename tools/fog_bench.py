"""BASELINE config 4 at full size: the value-noise fog over [-1024, 1024)^3 (2048^3 voxels), generated on the GPU
(tools/scenegen), distances by wx_tree_build, rendered at 3840x2160 from outside and from inside the volume.

  python tools/fog_bench.py [--half 1024] [--tau T | --calibrate] [--devices N]

--calibrate bisects tau for 40 % active voxels (40 +- 5 %, record tau).  With --devices N > 1 the frame is
tile-partitioned over N GPUs by wx_render's multi-device context (row bands dealt round-robin, peer stores into GPU 0)
and compared with the single-device frame.  Prints one JSON object per line."""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenegen
import woxel_b200 as W
from woxel_b200 import _ffi

TAU_2048 = 0.5242  # calibrated by --calibrate at half = 1024: 39.985 % of the 2048^3 voxels active (profiles/r1_fog2048.jsonl)

ap = argparse.ArgumentParser()
ap.add_argument("--half", type=int, default=1024)
ap.add_argument("--tau", type=float, default=None)
ap.add_argument("--calibrate", action="store_true")
ap.add_argument("--devices", type=int, default=1)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
args = ap.parse_args()
half, w, h = args.half, args.width, args.height
lib = _ffi.cuda_lib()


def emit(**kw):
    print(json.dumps(kw), flush=True)


tau = args.tau if args.tau is not None else TAU_2048
if args.calibrate:
    lo, hi = 0.3, 0.7  # occupancy falls as tau rises
    for _ in range(12):
        mid = 0.5 * (lo + hi)
        occ = scenegen.fog_occupancy(half, mid)
        emit(step="calibrate", tau=mid, occupancy=round(occ, 5))
        if occ > 0.40:
            lo = mid
        else:
            hi = mid
    tau = round(0.5 * (lo + hi), 4)

t0 = time.time()
topo = scenegen.fog_topology(half, tau)
t_gen = time.time() - t0
desc = scenegen.desc_of(topo)
emit(step="generate", half=half, tau=tau, occupancy=round(topo["occupancy"], 5), n5=int(desc.n5), n4=int(desc.n4), n3=int(desc.n3),
     seconds=round(t_gen, 2), leaf_fill=round(topo["occupancy"] * (2 * half) ** 3 / (512.0 * max(1, desc.n3)), 4))

cams = {"outside": ((0.5, 0.5, -2.44140625 * half - 0.5), (0.5, 0.5, 0.5)),
        "inside": ((3.5, 2.5, 1.5), (0.78125 * half, 0.46875 * half, 0.625 * half))}
ref_frames = {}
for ndev in sorted({1, args.devices}):
    import knobs
    ctx = knobs.apply_env(W.Context(n_devices=ndev))
    t0 = time.time()
    tree = ctx.build(desc)
    t_build = time.time() - t0
    s = tree.sdf
    emit(step="wx_tree_build", devices=ndev, seconds=round(t_build, 2), sdf_device_ms=round(s.device_ms, 1), rounds=int(s.rounds),
         max_dist=[int(x) for x in s.max_dist], device_MB=round(tree.info.device_bytes / 1e6, 1))
    nb = w * h * 4
    pinned = C.c_void_p()
    ctx.check(lib.wx_host_alloc_pinned(nb, C.byref(pinned)))
    host = np.frombuffer((C.c_uint8 * nb).from_address(pinned.value), np.uint8).reshape(1, h, w, 4)
    for tag, (eye, target) in cams.items():
        for mode in (0, 3):
            st = W.ComputeState.build(W.Camera(eye=eye, target=target, aspect=w / h), w, W.RenderMode(mode))
            for _ in range(3):
                ctx.render(tree, st, w, h, out=host)
            kms, wall = [], []
            for _ in range(9):
                t0 = time.perf_counter()
                ctx.render(tree, st, w, h, out=host)
                wall.append(time.perf_counter() - t0)
                kms.append(ctx.last_render_info().kernel_ms)
            key = (tag, mode)
            same = None
            if ndev == 1:
                ref_frames[key] = host.copy()
            else:
                same = bool(np.array_equal(host, ref_frames[key]))
            extra = {}
            if ndev == 1 and mode == 0:
                _, aov = ctx.render(tree, st, w, h, aov=True)
                it = aov["iters"][0]
                extra = {"hit_fraction": round(float((aov["state"][0] == 0).mean()), 4), "steps_per_ray": round(float(it.mean()), 2),
                         "max_steps": int(it.max())}
            px = (w // 8 * 8) * (h // 4 * 4)
            emit(step="render", devices=ndev, camera=tag, mode=mode, kernel_ms=round(float(np.median(kms)), 4),
                 wall_ms=round(float(np.median(wall)) * 1e3, 4), primary_Mrays_per_s_kernel=round(px / np.median(kms) / 1e3, 1),
                 primary_Mrays_per_s_e2e=round(px / np.median(wall) / 1e6, 1), identical_to_1_device=same, **extra)
    lib.wx_host_free_pinned(pinned)
    tree.free()
    ctx.close()
