#!/usr/bin/env python
"""bench.py -- primary Mrays/s of the raycast hot path on N B200s (one process per GPU) + roofline + CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scene sphere2048|cube|icosahedron|sphere256]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of rays: every rank renders one 3840x2160 frame
(render mode 0 = Gray: exactly one hdda_ray per pixel) of the procedural 2048^3 sphere level set
(BASELINE.json config 3, the configuration the north-star target is quoted on).  Rank r looks from
orbit camera r (camera 0 is config 3's camera; the sphere makes every orbit view the same workload),
so per-GPU work is fixed as N grows ("scaling": "weak", BASELINE config 5's camera sharding).  For
N > 1 every rank's frame is gathered in one frame stack on GPU 0 (mapped into the other processes through
CUDA IPC) INSIDE the timed step: wx_render renders into the rank's own memory in row chunks and each finished
chunk travels over NVLink by DMA while the next chunks render (no collective on the data path; NCCL only
carries the barriers and the timing reductions).  WX_BENCH_GATHER=store selects the fused alternative, the
kernel's own pixel stores going to the peer-mapped slot -- 30 % slower per step on 8 GPUs, see DESIGN.md section 5.

Timing: W warm-up steps, then K steps between barrier+synchronize brackets; the device time of every
step is taken with CUDA events on the launching stream (N > 1: the library's own event pair around the
whole wx_render call), L2 is flushed (256 MiB write) before every timed step outside the event pair, every
rank sums its K step times and the job's time is the MAX over ranks.
`e2e` goes through wx_render with HOST buffers: state H2D + kernel + RGBA D2H into pinned memory.

Beside the contract's keys the line carries: `value_tolerance_mode` (+ `tolerance_mode_vs_oracle`: agreement figures and the
listed mismatches against the oracle at full size, N = 1) for the opt-in WX_OPT_MARCH = 2; `secondary_ray_modes` (modes 3 / 4 in
primary + secondary rays/s, N = 1); at N > 1 `strong_single_frame` (ONE frame tile-partitioned over the ranks through
wx_render_shard, device-timed, bit-checked), `parity_vs_oracle_ranks` (the gathered slots of ranks 0, 1, N-1 against the oracle's
render of each camera) and `gathered_frames_equal_every_ranks_own`; `e2e.host_ingest_GBs` (what the host ingests from N
concurrent frame copies, measured in the run) and `e2e.frac_of_host_ingest`; `roofline.traffic_stale` (the ncu figures of
profiles/traffic.json belong to another build of the kernel).

The oracle (oracle/, a CPU restatement of the reference shader) is used here only for the reported
`cpu_baseline` and for `--impl reference`; the reference itself (Rust + wgpu) cannot run in this image.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

WIDTH, HEIGHT = 3840, 2160
METRIC = "primary Mrays/s per frame"
UNIT = "Mrays/s"


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def orbit_eye(k: int, n: int = 64, radius: float = 2500.0, elevation_deg: float = 0.0):
    """Camera k of an orbit around +y through config 3's eye (k = 0 -> (0.5, 0.5, -2500.5))."""
    th = 2.0 * math.pi * k / n
    el = math.radians(elevation_deg)
    r = radius + 1.0  # centre (0.5,0.5,0.5) + 2501 * direction: k = 0 gives exactly (0.5, 0.5, -2500.5)
    return (0.5 + r * math.cos(el) * math.sin(th), 0.5 + r * math.sin(el), 0.5 - r * math.cos(el) * math.cos(th))


def build_scene(name: str, ctx=None):
    """Returns (VDB345, FlatTree with SDF, description string, timings).  Product code only: with a device context the
    SDF sweep runs on the GPU (wx_compute_sdf, same values as the host sweep), else on the host."""
    import woxel_b200 as W
    import scenes
    t0 = time.time()
    if name == "sphere2048":
        v = W.VDB345.sphere(half=1024, radius=1000.0, band=3.0)
        what = "procedural 2048^3 sphere level set (R=1000, band +-3)"
    elif name == "sphere256":
        v = W.VDB345.sphere(half=128, radius=100.0, band=2.0)
        what = "procedural 256^3 sphere level set (R=100, band +-2) [debug size]"
    elif name in ("cube", "icosahedron"):
        t = scenes.load_topo(name)
        v = None
        what = f"assets/{name}.vdb topology (tests/golden/{name}.topo.npz)"
    else:
        raise SystemExit(f"unknown scene {name}")
    if v is None:
        # rebuild the asset in the product tree from the golden topology
        v = scenes.host_tree_from_scene(_TopoView(t))
    t1 = time.time()
    if ctx is not None:
        flat = v.to_flat(narrow_leaves=False)
        t2 = time.time()
        info = flat.compute_sdf_gpu(ctx)
        return v, flat, what, {"build_s": round(t1 - t0, 2), "flat_s": round(t2 - t1, 2), "sdf_gpu_s": round(time.time() - t2, 3),
                               "sdf_gpu_device_ms": round(info.device_ms, 2)}
    v.compute_sdf()
    t2 = time.time()
    flat = v.to_flat(narrow_leaves=False)
    return v, flat, what, {"build_s": round(t1 - t0, 2), "sdf_s": round(t2 - t1, 2), "flat_s": round(time.time() - t2, 2)}


class _TopoView:
    def __init__(self, t):
        self.origins, self.kids5, self.vals5, self.kids4, self.vals4, self.vals3 = (t[k] for k in ("origins", "kids5", "vals5", "kids4", "vals4", "vals3"))


def camera_for(scene: str, k: int):
    if scene.startswith("sphere2048"):
        return orbit_eye(k), (0.5, 0.5, 0.5)
    if scene == "sphere256":
        return (0.5, 0.5, -300.5), (0.5, 0.5, 0.5)
    return (0.5, 0.5, -500.5), (0.5, 0.5, -498.5)


def make_state(scene: str, k: int):
    import woxel_b200 as W
    eye, target = camera_for(scene, k)
    return W.ComputeState.build(W.Camera(eye=eye, target=target, aspect=WIDTH / HEIGHT), WIDTH, W.RenderMode.Gray)


def workload_config(what: str) -> dict:
    """`config` of the JSON line: identical in both arms (ours and --impl reference), so that the driver's same_config holds."""
    return {
        "workload": f"{what}, {WIDTH}x{HEIGHT}, render mode 0 (Gray, 1 hdda_ray/pixel), one frame per GPU per step (orbit camera = rank; "
                    f"the reference arm renders camera 0 on the host cores)",
        "l2": "GPU arm: flushed (256 MiB fill) before every timed step, outside the timed region; value_warm_l2 is the same loop without the flush",
    }


def sass_md5_of_loaded_kernel():
    """md5 of the SASS of wx::raycast_kernel<0,false> in the library this process loaded (tools/update_traffic.py stores the same
    for the build profiles/traffic.json was captured from).  None when cuobjdump is unavailable."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import update_traffic
        from woxel_b200 import _ffi
        return update_traffic.kernel_md5(_ffi.cuda_lib_path())
    except Exception:
        return None


def oracle_gpudata(flat):
    import oracle_ffi as O
    return O.gpudata_from_tables(flat.origins, flat.kids5, flat.vals5, flat.tab5, flat.kids4, flat.vals4, flat.tab4, flat.vals3, flat.tab3)


def oracle_state(ws):
    import oracle_ffi as O
    return O.State.from_buffer_copy(bytes(ws))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_bytes_per_ray(prof: dict, rays_per_launch: int, hbm_peak_gbs: float):
    """Bytes per ray the kernel moves at each level of the memory hierarchy (one ncu capture of the committed kernel,
    profiles/traffic.json) and the rays/s each level's MEASURED peak bandwidth would allow at that traffic."""
    try:
        if not prof or not prof.get("l2_bytes_per_launch") or not rays_per_launch:
            return None
        l1, l2, dram = (prof.get(k, 0) / rays_per_launch for k in ("l1_bytes_per_launch", "l2_bytes_per_launch", "dram_bytes_per_launch"))
        l2_peak = prof.get("l2_peak_gbs")
        return {"l1": round(l1, 1), "l2": round(l2, 1), "dram": round(dram, 2), "l2_peak_gbs": l2_peak,
                "rays_per_s_roofline_Grays": {"l2": round(l2_peak / l2, 1) if l2_peak and l2 else None,
                                              "dram": round(hbm_peak_gbs / dram, 1) if hbm_peak_gbs and dram else None}}
    except Exception:  # a reporting extra must never cost the bench line
        return None


# ------------------------------------------------------------------------------------------------
# reference arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as g
    g.build()
    _, flat, what, prep = build_scene(args.scene)
    gd = oracle_gpudata(flat)
    st = oracle_state(make_state(args.scene, 0))
    cores = os.cpu_count() or 1
    # bounded sample: one full frame per step (about 1 s on 16 cores); warm-up and step counts are capped so that the
    # arm ends within a few minutes whatever --steps says
    rays_per_step = (WIDTH // 8 * 8) * (HEIGHT // 4 * 4)
    args.steps = max(1, min(args.steps, 200))  # ~1 s per step: the arm ends within a few minutes whatever --steps says

    def step():
        gd.render(st, WIDTH, HEIGHT, aov=False, threads=cores)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = rays_per_step * args.steps / dt / 1e6
    sample = f"one full {WIDTH}x{HEIGHT} frame per step ({rays_per_step} rays/step), rows split over {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(1e3 * dt / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(what),
        "note": "reference arm = CPU restatement (oracle/) of raycast.comp.wgsl on the host cores; the Rust/wgpu reference "
                "cannot be built or run in this image (no rustc, wgpu, Vulkan ICD)",
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--scene", default="sphere2048")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    sys.stdout.flush()
    json_fd = os.dup(1)   # the ONE JSON line goes here; until then fd 1 is stderr (NCCL prints its version banner on fd 1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        g.build()  # no-op when the in-tree .so files are current
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the raycast path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    import woxel_b200 as W
    from woxel_b200 import _ffi
    lib = _ffi.cuda_lib()

    ctx = W.Context()  # current device
    v, flat, what, prep = build_scene(args.scene, ctx)
    tree = ctx.upload(flat)
    state = make_state(args.scene, rank)
    frame_bytes = WIDTH * HEIGHT * 4
    rays_per_frame = (WIDTH // 8 * 8) * (HEIGHT // 4 * 4)

    # ---- frame stack on GPU 0; peers map it through CUDA IPC -----------------------------------
    stack = C.c_void_p()
    if rank == 0:
        ctx.check(lib.wx_device_alloc(ctx._h, 0, frame_bytes * world, C.byref(stack)))
    if world > 1:
        handle = (C.c_uint8 * 64)()
        if rank == 0:
            ctx.check(lib.wx_ipc_export(ctx._h, 0, stack, handle))
        objs = [bytes(handle) if rank == 0 else None]
        dist.broadcast_object_list(objs, src=0)
        if rank != 0:
            h = (C.c_uint8 * 64).from_buffer_copy(objs[0])
            ctx.check(lib.wx_ipc_open(ctx._h, 0, h, C.byref(stack)))
    my_frame = stack.value + rank * frame_bytes
    # How a rank's frame reaches the stack on GPU 0 when N > 1 (WX_BENCH_GATHER):
    #   dma   (default) wx_render with the stack slot as destination: the frame is rendered into the rank's own memory in row
    #         chunks and every finished chunk travels to GPU 0 by DMA over NVLink while the next chunks render;
    #   store wx_render_device with the peer-mapped slot as output: the kernel's own 4-byte pixel stores cross NVLink
    #         (measured 30 % slower per step on 8 GPUs: 16-byte row segments per warp make poor NVLink packets);
    #   none  diagnostic only: the frame stays on the rank's GPU.
    gather = os.environ.get("WX_BENCH_GATHER", "dma") if world > 1 else "none"
    if world > 1 and gather == "none":
        own = C.c_void_p()
        ctx.check(lib.wx_device_alloc(ctx._h, 0, frame_bytes, C.byref(own)))
        my_frame = own.value

    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    step_ms = []          # device time of every timed step of this rank
    launches = [0]
    launches_per_call = [1]

    def step(timed_events=None):
        flush.fill_(1)  # evict L2 (126 MB) between steps; outside the timed region
        if gather == "dma":
            stream.synchronize()  # the library renders on its own streams: the flush must be over
            ctx.render_to(tree, state, WIDTH, HEIGHT, my_frame)  # blocking; timed by the library's events on its launching stream
            if timed_events is not None:
                info = ctx.last_render_info()
                step_ms.append(info.total_ms)  # kernels of all chunks + the DMA of the last chunk
                launches[0] += info.launches
            return
        if timed_events is not None:
            timed_events[0].record(stream)
        ctx.render_device(tree, state, WIDTH, HEIGHT, my_frame, stream=stream.cuda_stream)
        if timed_events is not None:
            timed_events[1].record(stream)
            launches[0] += launches_per_call[0]  # 1, or 2 with the long-tile kernel (read once after the warm-up, below)

    def timed_run(n_warm, n_steps):
        """n_warm untimed + n_steps timed steps between barrier + synchronize brackets.  Returns (ms per step = MAX over ranks of
        the per-rank mean of the device-timed steps, per-rank means, wall seconds)."""
        for _ in range(n_warm):
            step()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        del step_ms[:]
        sync_all()
        t_wall0 = time.perf_counter()
        for k in range(n_steps):
            step(evs[k])
        sync_all()
        wall_s = time.perf_counter() - t_wall0
        # device time of this rank's K steps; the job's time is the MAX over ranks (every rank renders K frames)
        if gather != "dma":
            step_ms.extend(a.elapsed_time(b) for a, b in evs)
        ms = torch.tensor(step_ms, dtype=torch.float64, device="cuda")
        assert ms.numel() == n_steps
        total = ms.sum().reshape(1)
        ranks = [float(total.item()) / n_steps]
        if world > 1:
            gathered = [torch.zeros_like(total) for _ in range(world)]
            dist.all_gather(gathered, total)
            ranks = [float(t.item()) / n_steps for t in gathered]
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        return float(total.item() / n_steps), ranks, wall_s

    for _ in range(args.warmup):
        step()
    sync_all()
    if gather != "dma":
        launches_per_call[0] = int(ctx.last_render_info().launches)  # kernels one wx_render_device call launches in steady state
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ms_per_step, per_rank, wall = timed_run(0, args.steps)
    timed_launches = launches[0]
    value = world * rays_per_frame / (ms_per_step * 1e-3) / 1e6

    # ---- tolerance mode (WX_OPT_MARCH = 2: fused p += t * dir, rays start at the bounding box of the active cells): same loop
    ctx.set_option(_ffi.WX_OPT_MARCH, 2)
    tol_ms, _, _ = timed_run(3, args.steps)
    ctx.set_option(_ffi.WX_OPT_MARCH, 0)
    value_tol = world * rays_per_frame / (tol_ms * 1e-3) / 1e6

    # warm-L2 figure (no flush), for reference only
    if gather == "dma":
        sync_all()
        w = 0.0
        for _ in range(10):
            ctx.render_to(tree, state, WIDTH, HEIGHT, my_frame)
            w += ctx.last_render_info().total_ms
        sync_all()
        warm_ms = torch.tensor([w / 10], dtype=torch.float64, device="cuda")
    else:
        for _ in range(3):
            ctx.render_device(tree, state, WIDTH, HEIGHT, my_frame, stream=stream.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record(stream)
        for _ in range(10):
            ctx.render_device(tree, state, WIDTH, HEIGHT, my_frame, stream=stream.cuda_stream)
        e1.record(stream)
        sync_all()
        warm_ms = torch.tensor([e0.elapsed_time(e1) / 10], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(warm_ms, op=dist.ReduceOp.MAX)
    warm_ms = float(warm_ms.item())

    # ---- e2e: the public host-buffer call (state H2D, kernel, RGBA D2H into pinned memory) ------
    pinned = C.c_void_p()
    ctx.check(lib.wx_host_alloc_pinned(frame_bytes, C.byref(pinned)))
    host_frame = np.frombuffer((C.c_uint8 * frame_bytes).from_address(pinned.value), np.uint8).reshape(1, HEIGHT, WIDTH, 4)
    for _ in range(3):
        ctx.render(tree, state, WIDTH, HEIGHT, out=host_frame)
    sync_all()
    e2e_steps = max(5, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.render(tree, state, WIDTH, HEIGHT, out=host_frame)  # blocking
    e2e_dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_value = world * rays_per_frame * e2e_steps / float(e2e_dt.item()) / 1e6
    e2e_info = ctx.last_render_info()
    e2e_ms = 1e3 * float(e2e_dt.item()) / e2e_steps
    checksum = int(host_frame.view(np.uint32).sum(dtype=np.uint64))
    clocks = sampler.stop() if rank == 0 else None

    # ---- the ceiling of e2e: what the host can ingest.  Every rank copies one finished frame (device -> pinned host) at the
    # same time, nothing else running: aggregate GB/s over the N PCIe links is the most an e2e step could deliver.
    local_frame = C.c_void_p()
    ctx.check(lib.wx_device_alloc(ctx._h, 0, frame_bytes, C.byref(local_frame)))
    ctx.render_device(tree, state, WIDTH, HEIGHT, local_frame.value, stream=stream.cuda_stream)
    for _ in range(2):
        ctx.check(lib.wx_memcpy_d2h(ctx._h, 0, pinned, local_frame, frame_bytes, None))
    sync_all()
    t0 = time.perf_counter()
    ingest_reps = 8
    for _ in range(ingest_reps):
        ctx.check(lib.wx_memcpy_d2h(ctx._h, 0, pinned, local_frame, frame_bytes, None))
        ctx.check(lib.wx_stream_synchronize(ctx._h, 0, None))
    ingest_dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ingest_dt, op=dist.ReduceOp.MAX)
    ingest_gbs = world * frame_bytes * ingest_reps / float(ingest_dt.item()) / 1e9
    e2e_gbs = world * frame_bytes / (e2e_ms * 1e-3) / 1e9

    # ---- N = 1: secondary-ray modes (thesis results.tex:241-255): kernel time of modes 3 and 4, L2 flushed, device-resident output
    mode_ms = {}
    if world == 1:
        for mode in (3, 4):
            st_m = type(state).from_buffer_copy(bytes(state))
            st_m.render_mode[0] = mode
            for _ in range(3):
                ctx.render_device(tree, st_m, WIDTH, HEIGHT, local_frame.value, stream=stream.cuda_stream)
            acc = 0.0
            for _ in range(10):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                ctx.render_device(tree, st_m, WIDTH, HEIGHT, local_frame.value, stream=stream.cuda_stream)
                e1.record(stream)
                torch.cuda.synchronize()
                acc += e0.elapsed_time(e1)
            mode_ms[mode] = acc / 10

    # ---- N > 1: (a) every rank's gathered frame is on GPU 0 -- rank 0 reads back the slots of ranks 0, 1 and N-1 for the oracle
    # check below; (b) strong scaling of ONE frame (BASELINE config 4's tile partition): camera 0's frame in 8-row bands dealt
    # round-robin to the ranks, every rank delivering its rows into the frame on GPU 0 (wx_render_shard), device-timed per rank.
    rank_frames, strong, gather_ok = {}, None, None
    if world > 1:
        sync_all()
        # every rank's own frame (its e2e render of the same camera, in host memory) against what arrived in its slot on GPU 0
        mine = torch.tensor([checksum], dtype=torch.int64, device="cuda")
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        if rank == 0:
            gather_ok = True
            for r in range(world):
                buf = np.empty((HEIGHT, WIDTH, 4), np.uint8)
                ctx.check(lib.wx_memcpy_d2h(ctx._h, 0, buf.ctypes.data, C.c_void_p(stack.value + r * frame_bytes), frame_bytes, None))
                ctx.check(lib.wx_stream_synchronize(ctx._h, 0, None))
                gather_ok = gather_ok and int(buf.view(np.uint32).sum(dtype=np.uint64)) == int(sums[r].item())
                if r in (0, 1, world - 1):
                    rank_frames[r] = buf
        sync_all()
        state0 = make_state(args.scene, 0)
        # one GPU alone: kernel time of the whole frame, same kernel, local output
        for _ in range(3):
            ctx.render_device(tree, state0, WIDTH, HEIGHT, local_frame.value, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        single = 0.0
        for _ in range(10):
            ctx.render_device(tree, state0, WIDTH, HEIGHT, local_frame.value, stream=stream.cuda_stream)
            ctx.check(lib.wx_stream_synchronize(ctx._h, 0, None))
            single += ctx.last_render_info().kernel_ms
        single_ms = torch.tensor([single / 10], dtype=torch.float64, device="cuda")
        dist.all_reduce(single_ms, op=dist.ReduceOp.MAX)
        for _ in range(3):
            ctx.render_shard_to(tree, state0, WIDTH, HEIGHT, (rank, world), stack.value)
        k_ms, t_ms, reps = 0.0, 0.0, 10
        for _ in range(reps):
            sync_all()
            ctx.render_shard_to(tree, state0, WIDTH, HEIGHT, (rank, world), stack.value)
            info = ctx.last_render_info()
            k_ms += info.kernel_ms
            t_ms += info.total_ms
        both = torch.tensor([k_ms / reps, t_ms / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(both, op=dist.ReduceOp.MAX)
        sync_all()
        strong = {"what": f"camera 0's {WIDTH}x{HEIGHT} frame split into 8-row bands dealt round-robin over {world} GPUs (one process each), "
                          "every GPU delivering its rows into the frame on GPU 0 over NVLink (wx_render_shard); device-timed, MAX over ranks",
                  "one_gpu_kernel_ms": round(float(single_ms.item()), 4),
                  "kernel_ms": round(float(both[0].item()), 4), "kernel_plus_delivery_ms": round(float(both[1].item()), 4),
                  "kernel_speedup": round(float(single_ms.item()) / float(both[0].item()), 2),
                  "speedup_incl_delivery": round(float(single_ms.item()) / float(both[1].item()), 2),
                  "Mrays_per_s": round(rays_per_frame / (float(both[1].item()) * 1e-3) / 1e6, 1)}
        if rank == 0:
            buf = np.empty((HEIGHT, WIDTH, 4), np.uint8)
            ctx.check(lib.wx_memcpy_d2h(ctx._h, 0, buf.ctypes.data, stack, frame_bytes, None))
            ctx.check(lib.wx_stream_synchronize(ctx._h, 0, None))
            strong["bit_identical_to_one_gpu_frame"] = bool(np.array_equal(buf, host_frame[0]))  # rank 0's e2e frame is camera 0

    # ---- rank 0: CPU baseline + roofline ----------------------------------------------------------
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
        cpu_baseline, bytes_per_ray, steps_per_ray, parity = None, None, None, None
        parity_ranks, tol_fig, modes = None, None, None
        if not args.no_cpu_baseline:
            # the oracle renders camera 0's frame once: its per-level lookup counters give the algorithmic bytes per ray
            # (roofline, any N); its time is the reported CPU baseline (N = 1 only) and its pixels a full-frame parity check
            gd = oracle_gpudata(flat)
            cores = os.cpu_count() or 1
            t0 = time.perf_counter()
            ref_rgba, _, st = gd.render(oracle_state(state), WIDTH, HEIGHT, aov=False, threads=cores)
            dt = time.perf_counter() - t0
            if world == 1:
                cpu_baseline = {"value": round(st.primary_rays / dt / 1e6, 3), "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"one full {WIDTH}x{HEIGHT} frame of the same scene/camera/mode ({st.primary_rays} rays), "
                                          f"oracle/ C restatement of raycast.comp.wgsl, {cores} threads, {dt:.2f} s"}
            bytes_per_ray = st.primary_alg_bytes / st.primary_rays
            steps_per_ray = sum(st.primary_lookups) / st.primary_rays
            parity = bool(np.array_equal(ref_rgba, host_frame[0]))
            import oracle_ffi as O
            if world > 1:
                # the frames the ranks gathered on GPU 0 inside the timed steps, against the oracle's render of each rank's camera
                parity_ranks = {}
                for r, buf in rank_frames.items():
                    ref_r = ref_rgba if r == 0 else gd.render(oracle_state(make_state(args.scene, r)), WIDTH, HEIGHT, aov=False, threads=cores)[0]
                    parity_ranks[str(r)] = bool(np.array_equal(ref_r, buf))
                if strong is not None:
                    strong["bit_identical_to_oracle_frame"] = bool(strong.get("bit_identical_to_one_gpu_frame")) and parity
            else:
                # tolerance mode against the oracle at full size: the north-star bar, mismatches listed
                import agreement
                ctx.set_option(_ffi.WX_OPT_MARCH, 2)
                t_rgba, t_aov = ctx.render(tree, state, WIDTH, HEIGHT, aov=True)
                ctx.set_option(_ffi.WX_OPT_MARCH, 0)
                _, ref_aov, _ = gd.render(oracle_state(state), WIDTH, HEIGHT, aov=True, threads=cores)
                tol_fig = agreement.compare(t_rgba[0], {k: v[0] for k, v in t_aov.items()}, ref_rgba, ref_aov, max_list=24)
                tol_fig["meets_north_star_bar"] = agreement.meets_bar(tol_fig)
                tol_fig["mean_iterations"] = {"tolerance": round(float(t_aov["iters"][0].mean()), 2), "exact": round(float(ref_aov["iters"].mean()), 2)}
                # modes 3 / 4: (primary + secondary) rays per second; the secondary-ray count is data dependent and comes from
                # the oracle's render of the same frame (the GPU frames are bit-identical to it, tests/test_parity_gpu.py)
                modes = {}
                for mode, ms_m in mode_ms.items():
                    st_m = type(state).from_buffer_copy(bytes(state))
                    st_m.render_mode[0] = mode
                    _, _, st_o = gd.render(oracle_state(st_m), WIDTH, HEIGHT, aov=False, threads=cores)
                    modes[f"mode{mode}"] = {"kernel_ms": round(ms_m, 4), "primary_Mrays_per_s": round(rays_per_frame / ms_m / 1e3, 1),
                                            "rays_incl_secondary": int(st_o.rays),
                                            "primary_plus_secondary_Mrays_per_s": round(st_o.rays / ms_m / 1e3, 1)}
        # per-launch figures of ONE ncu capture (profiles/traffic.json, tools/update_traffic.py).  They describe the build whose SASS
        # hash is stored beside them: a different kernel in the loaded library makes them stale, and the line says so.
        traffic, warp_instr, traffic_stale = None, None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            prof = json.load(open(tpath)).get(args.scene, {})
            traffic, warp_instr = prof.get("dram_bytes_per_launch"), prof.get("warp_instructions_per_launch")
            md5 = sass_md5_of_loaded_kernel()
            traffic_stale = None if (md5 is None or not prof.get("sass_md5")) else bool(md5 != prof["sass_md5"])
        else:
            prof = {}
        roof = None
        if bytes_per_ray is not None:
            achieved = bytes_per_ray * rays_per_frame / (ms_per_step * 1e-3) / 1e9  # per GPU: one launch = one frame
            roof = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                    "traffic": traffic, "traffic_stale": traffic_stale, "traffic_source": prof.get("capture"),
                    "peak_source": peak_src, "kernel": "wx::raycast_kernel<0,false>",
                    "alg_bytes_per_ray": round(bytes_per_ray, 2), "lookups_per_ray": round(steps_per_ray, 2),
                    "rays_per_launch": rays_per_frame,
                    # what actually binds the kernel: issue slots.  Warp instructions per launch are ncu's count
                    # (profiles/traffic.json); the fraction is of 4 schedulers x 1 instruction per clock per SM at the
                    # clock sampled during the run.
                    "issue_slots": None if not (warp_instr and clocks and clocks.get("sm_mhz")) else {
                        "warp_instructions_per_launch": warp_instr,
                        "frac_of_peak": round(warp_instr / (torch.cuda.get_device_properties(0).multi_processor_count * 4 *
                                                             clocks["sm_mhz"] * 1e6 * ms_per_step * 1e-3), 4)},
                    # measured bytes per ray at each level of the hierarchy (one ncu capture, profiles/traffic.json) and the rays/s
                    # each level's measured peak would allow: none of them is what limits the kernel
                    "measured_bytes_per_ray": measured_bytes_per_ray(prof, rays_per_frame, peak),
                    "note": "algorithmic bytes (SURVEY 8d) over the CUDA-event time of the kernel, L2 flushed before each launch; "
                            "the working set is L2-resident so this is a bandwidth-equivalent figure, the kernel is latency/issue bound"}
        out = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(what),
            "gather": {"none": "N = 1: the frame stays on the GPU", "dma": "every frame gathered on GPU0 over NVLink (DMA of row chunks, "
                       "overlapped with the render of the next chunks, inside the timed region)", "store": "every frame "
                       "stored into GPU0 over NVLink by the kernel itself"}[gather if world > 1 else "none"],
            "tree": {"n5": tree.info.n5, "n4": tree.info.n4, "n3": tree.info.n3, "leaf_bits": tree.info.leaf_bits,
                     "device_MB": round(tree.info.device_bytes / 1e6, 1)},
            "prep_s": prep,
            "value_warm_l2": round(world * rays_per_frame / (warm_ms * 1e-3) / 1e6, 1),
            # the opt-in tolerance mode (WX_OPT_MARCH = 2), same timed loop; `value` above is the exact (bit-identical) kernel
            "value_tolerance_mode": round(value_tol, 1), "tolerance_mode_ms_per_step": round(tol_ms, 4),
            "tolerance_mode_vs_oracle": tol_fig, "secondary_ray_modes": modes,
            "strong_single_frame": strong, "parity_vs_oracle_ranks": parity_ranks,
            "gathered_frames_equal_every_ranks_own": gather_ok,  # N > 1: checksum of every slot on GPU 0 == the rank's own host frame
            "wall_ms_per_step_incl_flush": round(1e3 * wall / args.steps, 4),
            "per_rank_ms_per_step": [round(x, 4) for x in per_rank],  # ms_per_step is their maximum
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": 256 * world, "d2h_bytes_per_step": frame_bytes * world,
                    "api": "wx_render (host state in, pinned host RGBA8 out), blocking; read-back pipelined over row chunks",
                    "steps": e2e_steps, "kernel_launches_per_step": int(e2e_info.launches),
                    "last_call_device_ms": {"kernels": round(e2e_info.kernel_ms, 4), "total_incl_readback": round(e2e_info.total_ms, 4)},
                    "ms_per_step": round(e2e_ms, 4), "d2h_GBs": round(e2e_gbs, 1),
                    # measured in this run: all ranks copying one frame each, device -> pinned host, nothing else running
                    "host_ingest_GBs": round(ingest_gbs, 1), "frac_of_host_ingest": round(e2e_gbs / ingest_gbs, 3)},
            "gpu_launches": timed_launches * world,  # raycast kernel launches inside the timed region (rank 0's count x ranks)
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_baseline,
            "parity_vs_oracle_full_frame": parity, "frame_checksum": checksum,
            "parity": "partial: bit-identical to the strict-f32 CPU restatement of the shader (oracle/), which the reference's own tests do "
                      "not pin (it ships no golden image for the raycast) -- 'parity unpinned'",
            "parity_note": "parity_vs_oracle_* compare the GPU frame with the oracle RENDERER run on the tables the product built "
                           "(host tree builder + wx_compute_sdf, whose equality with the oracle's compute_sdf is a test, tests/test_sdf_gpu.py)",
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.barrier()
        if rank != 0:
            lib.wx_ipc_close(ctx._h, 0, stack)
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
