"""Helper of test_gpu_rank_identity.py (not a test module).  BASELINE.json configs[3] in miniature: every rank solves
its own 640x480x8 window in a fresh process and contributes its poses to one all_gather.

  torchrun --nproc-per-node R dist_rank_identity.py --out DIR      -> DIR/rank{r}.npz (+ what rank r gathered)
  python dist_rank_identity.py --out DIR --single r                -> DIR/single{r}.npz (same window, no torchrun)
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ITERS = 6


def solve(rank):
    import bench
    import voldor_b200

    win, boot = bench.make_inputs(rank)
    cfg = bench.CONFIG.replace(f"--max_iters {bench.EM_ITERS}", f"--max_iters {ITERS}")
    voldor_b200.set_bootstrap_override(*boot)
    r = voldor_b200.voldor_ex(win["flows"], win["fx"], win["fy"], win["cx"], win["cy"], config=cfg)
    voldor_b200.set_bootstrap_override()
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--single", type=int, default=None)
    a = ap.parse_args()
    import torch

    if a.single is not None:
        torch.cuda.set_device(0)
        r = solve(a.single)
        np.savez(os.path.join(a.out, f"single{a.single}.npz"), **{k: r[k] for k in ("poses", "poses_covar", "depth", "depth_conf")},
                 n=r["n_registered"], iters=r["iters"])
        return
    import torch.distributed as dist

    from voldor_b200 import parallel

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ndev = torch.cuda.device_count()
    torch.cuda.set_device(local % ndev)
    nccl = ndev >= world  # ranks sharing one GPU cannot form an NCCL communicator
    dist.init_process_group("nccl" if nccl else "gloo")
    r = solve(rank)
    everyone = parallel.allgather_window_poses(r, device=f"cuda:{local % ndev}" if nccl else "cpu")
    np.savez(os.path.join(a.out, f"rank{rank}.npz"), **{k: r[k] for k in ("poses", "poses_covar", "depth", "depth_conf")},
             n=r["n_registered"], iters=r["iters"],
             gathered_poses=np.stack([np.pad(e["poses"], ((0, 16 - e["n_registered"]), (0, 0))) for e in everyone]),
             gathered_n=np.array([e["n_registered"] for e in everyone]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
