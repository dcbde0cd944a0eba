"""BASELINE.json configs[3] (window-per-GPU under torchrun): what rank r computes inside the multi-process job is, bit
for bit, what a single process computes for the same window (fresh-process RNG streams on both sides), and the one
collective of the path delivers every rank's poses to every rank."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import ffi

pytestmark = pytest.mark.gpu
HELPER = os.path.join(ffi.ROOT, "tests", "dist_rank_identity.py")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_rank_outputs_equal_single_process_outputs(tmp_path):
    world = 2
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), HELPER, "--out", str(tmp_path)],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    ranks = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    for k in range(world):
        s = subprocess.run([sys.executable, HELPER, "--out", str(tmp_path), "--single", str(k)], capture_output=True,
                           text=True, timeout=900, env=env)
        assert s.returncode == 0, s.stderr[-4000:]
        single = np.load(tmp_path / f"single{k}.npz")
        assert int(ranks[k]["n"]) == int(single["n"]) == 8 and int(ranks[k]["iters"]) == int(single["iters"])
        for key in ("poses", "poses_covar", "depth", "depth_conf"):
            assert ffi.bits_equal(ranks[k][key], single[key]), (k, key)
    # different ranks solved different windows, and everybody received everybody's poses
    assert not np.array_equal(ranks[0]["poses"], ranks[1]["poses"])
    for k in range(world):
        for j in range(world):
            n = int(ranks[k]["gathered_n"][j])
            assert n == int(ranks[j]["n"])
            assert ffi.bits_equal(ranks[k]["gathered_poses"][j][:n], ranks[j]["poses"])
