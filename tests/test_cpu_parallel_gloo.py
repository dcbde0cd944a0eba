"""world_size-2 gloo test of the window-per-rank sharding + single all_gather of poses (voldor_b200/parallel.py),
with the CPU port standing in for the GPU kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ffi
import oracle_host

HERE = os.path.dirname(os.path.abspath(__file__))


def _window(rank):
    import synth

    # a different image size per rank: the per-pixel RNG streams are re-seeded when the size changes, so a window
    # solved in a fresh worker process equals the same window solved later in the parent process
    win = synth.make_window(48 + 8 * rank, 32, 2, seed=40 + rank)
    boot = (win["Rs"][0], win["ts"][0], synth.noisy_depth(win, 0.05))
    cfg = "--silent --max_iters 2 --no_trunc_iters 1000 --n_poses_to_sample 512"
    ffi.libc_srand(rank)
    return oracle_host.run_window("cpu", win["flows"], win["fx"], win["fy"], win["cx"], win["cy"], config=cfg, boot=boot)


def _worker(rank, world, port, q):
    sys.path.insert(0, ffi.ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from voldor_b200 import parallel

    mine = parallel.shard_windows(world, rank, world)
    assert mine == [rank]
    res = _window(rank)
    allr = parallel.allgather_window_poses(res)
    q.put((rank, [(r["n_registered"], r["poses"].tolist()) for r in allr]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not os.path.exists(oracle_host.CPU), reason="oracle CPU port not built")
def test_two_rank_allgather_of_window_poses():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    got = dict(q.get(timeout=240) for _ in range(world))
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    single = [_window(r) for r in range(world)]
    for rank in range(world):
        for src in range(world):
            n, poses = got[rank][src]
            assert n == single[src]["n_registered"] == 2
            assert np.array_equal(np.asarray(poses, np.float32), single[src]["poses"])


def test_shard_windows_partition():
    from voldor_b200 import parallel

    for n in (1, 7, 8, 9):
        for world in (1, 2, 4, 8):
            parts = [parallel.shard_windows(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))


def test_record_roundtrip():
    from voldor_b200 import parallel

    res = {"n_registered": 3, "poses": np.arange(18, dtype=np.float32).reshape(3, 6),
           "poses_covar": np.arange(108, dtype=np.float32).reshape(3, 6, 6)}
    back = parallel.unpack_record(parallel.pack_record(res))
    assert back["n_registered"] == 3 and np.array_equal(back["poses"], res["poses"])
    assert np.array_equal(back["poses_covar"], res["poses_covar"])
