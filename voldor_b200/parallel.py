"""Window-per-GPU sharding of the VO hot path.

Windows are independent units: rank r solves its own window; afterwards every rank contributes a fixed-size
record (n_registered, 16 x 6 poses, 16 x 36 covariances; zero padded) to ONE all_gather so that all ranks — and
the trajectory assembler on the host — see all poses.  The record is ~2.7 KB per rank: the collective is
latency bound, there is nothing to fuse it into (no kernel of the path is followed by an exchange).
Works on any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist

MAX_FRAMES = 16
RECORD = 1 + MAX_FRAMES * 6 + MAX_FRAMES * 36


def pack_record(result, device="cpu"):
    rec = torch.zeros(RECORD, dtype=torch.float32)
    n = int(result["n_registered"])
    rec[0] = n
    if n > 0:
        rec[1:1 + n * 6] = torch.from_numpy(np.ascontiguousarray(result["poses"][:n], np.float32).reshape(-1))
        rec[1 + MAX_FRAMES * 6:1 + MAX_FRAMES * 6 + n * 36] = torch.from_numpy(
            np.ascontiguousarray(result["poses_covar"][:n], np.float32).reshape(-1))
    return rec.to(device)


def unpack_record(rec):
    rec = rec.detach().cpu()
    n = int(rec[0].item())
    poses = rec[1:1 + n * 6].reshape(n, 6).numpy().copy()
    covar = rec[1 + MAX_FRAMES * 6:1 + MAX_FRAMES * 6 + n * 36].reshape(n, 6, 6).numpy().copy()
    return {"n_registered": n, "poses": poses, "poses_covar": covar}


def allgather_window_poses(result, device="cpu", group=None):
    """every rank's window result -> list (by rank) of {n_registered, poses, poses_covar}"""
    rec = pack_record(result, device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [unpack_record(rec)]
    out = [torch.empty_like(rec) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, rec, group=group)
    return [unpack_record(r) for r in out]


def shard_windows(n_windows, rank, world):
    """contiguous block partition of a list of independent windows over the ranks"""
    per = (n_windows + world - 1) // world
    lo = min(n_windows, rank * per)
    return list(range(lo, min(n_windows, lo + per)))
