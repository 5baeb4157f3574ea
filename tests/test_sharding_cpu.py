"""Multi-GPU split, exercised without GPUs: world_size-2 (and 3) gloo process groups on CPU.

The raycast path shards with no data-path collective: every rank renders the row bands
(or the cameras) it owns and stores them into one frame on GPU 0.  What can go wrong on the host side is the
dealing itself, so these tests run the product's own dealing function (`wx_shard_rows`, the arithmetic
`wx_render` / `wx_render_device` launch with) and bench.py's camera assignment in real `torch.distributed`
ranks, "render" with the ORACLE on the rows each rank owns, gather on rank 0 like the NVLink store does, and
compare with the unsharded oracle frame.  No compute call of the CUDA library is made.
"""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W_, H_ = 96, 52  # 52 rows: 6 full bands of 8 + one band of 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rows_of(lib, shard_index, shard_count, band_rows, height):
    from woxel_b200 import _ffi
    sh = _ffi.WxShard(shard_index, shard_count, band_rows, 0)
    mask = np.zeros(height, np.uint8)
    rc = lib.wx_shard_rows(height, C.byref(sh), mask.ctypes.data)
    assert rc == 0
    return mask.astype(bool)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import scenes
    from woxel_b200 import _ffi
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = _ffi.cuda_lib()  # loading needs no GPU; only wx_shard_rows (host arithmetic) is called
        s = scenes.get_scene("small_sphere")
        st = scenes.state_for((150.0, 90.0, -170.0), (0.0, 0.0, 0.0), W_, H_, mode=3)
        # --- tile/band sharding of one frame: each rank renders the rows wx_shard_rows gives it -----------------
        mine = _rows_of(lib, rank, world, 8, H_)
        frame = torch.zeros((H_, W_, 4), dtype=torch.uint8)
        owner = torch.full((H_,), -1, dtype=torch.int32)
        for y0 in np.flatnonzero(mine & ~np.roll(mine, 1) | (mine & (np.arange(H_) == 0))):
            y1 = y0
            while y1 < H_ and mine[y1]:
                y1 += 1
            rgba, _, _ = s.gpu.render(st, W_, H_, aov=False, rows=(int(y0), int(y1)))
            frame[y0:y1] = torch.from_numpy(rgba[y0:y1].copy())
            owner[y0:y1] = rank
        # the gather: rank 0's frame receives every peer's rows (stands in for the peer-mapped stores)
        frames = [torch.zeros_like(frame) for _ in range(world)] if rank == 0 else None
        owners = [torch.zeros_like(owner) for _ in range(world)] if rank == 0 else None
        dist.gather(frame, frames, dst=0)
        dist.gather(owner, owners, dst=0)
        # --- camera sharding (bench.py: rank r renders orbit camera r into slot r) -------------------------------
        import bench
        eye = torch.tensor(bench.orbit_eye(rank), dtype=torch.float64)
        eyes = [torch.zeros_like(eye) for _ in range(world)] if rank == 0 else None
        dist.gather(eye, eyes, dst=0)
        # timing reduction used by bench.py: per-step max over ranks
        ms = torch.tensor([1.0 + rank, 5.0 - rank], dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        assert ms.tolist() == [float(world), 5.0]
        if rank == 0:
            own = torch.stack(owners).numpy()
            covered = (own >= 0).sum(0)
            assert (covered == 1).all(), "row bands must partition the frame"
            who = own.max(0)
            assert np.array_equal(who, (np.arange(H_) // 8) % world), "band b belongs to shard b % count"
            full = torch.stack(frames).sum(0).to(torch.uint8).numpy()  # disjoint rows: the sum is the union
            ref, _, _ = s.gpu.render(st, W_, H_, aov=False)
            assert np.array_equal(full, ref), "sharded frame differs from the unsharded one"
            e = torch.stack(eyes).numpy()
            assert np.allclose(e[0], (0.5, 0.5, -2500.5)) and len({tuple(np.round(v, 6)) for v in e}) == world
            open(os.path.join(out_dir, "ok"), "w").write("ok")
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_band_and_camera_sharding_gloo(world, tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").exists()


def test_shard_rows_partition_properties():
    for p in (ROOT,):
        if p not in sys.path:
            sys.path.insert(0, p)
    from woxel_b200 import _ffi
    lib = _ffi.cuda_lib()
    for height in (1, 7, 8, 9, 64, 2160, 1081):
        for count in (1, 2, 4, 8):
            for band in (8, 16, 64):
                total = np.zeros(height, np.int32)
                for idx in range(count):
                    total += _rows_of(lib, idx, count, band, height)
                assert (total == 1).all(), (height, count, band)
    # invalid shards are rejected
    bad = _ffi.WxShard(2, 2, 8, 0)
    assert lib.wx_shard_rows(16, C.byref(bad), np.zeros(16, np.uint8).ctypes.data) == -1
    bad = _ffi.WxShard(0, 2, 12, 0)
    assert lib.wx_shard_rows(16, C.byref(bad), np.zeros(16, np.uint8).ctypes.data) == -1
