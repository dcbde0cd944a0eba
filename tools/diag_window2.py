"""Diagnostic: run ONE backend in a fresh process, optionally after polluting freed device memory with NaNs,
and dump the outputs.  usage: python tools/diag_window2.py BACKEND(ref|mine|abi) POLLUTE(0|1|2) OUT.npz [W H N ITERS POSES SEED]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import ffi  # noqa: E402
import oracle_host  # noqa: E402
import synth  # noqa: E402
import voldor_b200  # noqa: E402

backend, pollute, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
w, h, N, iters, poses, seed = [int(v) for v in (sys.argv[4:10] if len(sys.argv) >= 10 else "96 64 3 1 1024 1".split())]
torch.cuda.init()
if pollute:
    # fill a few GB with a NaN / large-int pattern, then give it back to the driver
    fillv = float("nan") if pollute == 1 else 12345.678
    bufs = [torch.full((256 * 1024 * 1024,), fillv, device="cuda", dtype=torch.float32) for _ in range(4)]
    torch.cuda.synchronize()
    del bufs
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
win = synth.make_window(w, h, N, seed=seed)
boot = (win["Rs"][0], win["ts"][0], synth.noisy_depth(win, 0.05))
cfg = f"--silent --max_iters {iters} --no_trunc_iters 1000 --n_poses_to_sample {poses}"
args = (win["flows"], win["fx"], win["fy"], win["cx"], win["cy"])
ffi.libc_srand(1)
if backend == "mine":
    voldor_b200.set_bootstrap_override(*boot)
    r = voldor_b200.voldor_ex(*args, config=cfg)
else:
    r = oracle_host.run_window("ref" if backend == "ref" else "ours_abi", *args, config=cfg, boot=boot)
np.savez(out, **{k: np.asarray(r[k]) for k in ("poses", "poses_covar", "depth", "depth_conf")})
print(backend, pollute, "done", r["poses"].ravel()[:3])
