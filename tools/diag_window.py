"""Diagnostic: one small window through the resident pipeline, the ABI path and the reference kernels; prints
which outputs differ.  usage: python tools/diag_window.py W H N ITERS POSES SEED [order]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import ffi  # noqa: E402
import oracle_host  # noqa: E402
import synth  # noqa: E402
import voldor_b200  # noqa: E402

w, h, N, iters, poses, seed = [int(v) for v in sys.argv[1:7]]
order = sys.argv[7] if len(sys.argv) > 7 else "mine_first"
win = synth.make_window(w, h, N, seed=seed)
boot = (win["Rs"][0], win["ts"][0], synth.noisy_depth(win, 0.05))
cfg = f"--silent --max_iters {iters} --no_trunc_iters 1000 --n_poses_to_sample {poses}"
args = (win["flows"], win["fx"], win["fy"], win["cx"], win["cy"])


def mine_run():
    voldor_b200.set_bootstrap_override(*boot)
    ffi.libc_srand(1)
    r = voldor_b200.voldor_ex(*args, config=cfg)
    voldor_b200.set_bootstrap_override()
    return r


def ref_run(backend="ref"):
    ffi.libc_srand(1)
    return oracle_host.run_window(backend, *args, config=cfg, boot=boot)


if order == "mine_first":
    mine, ref = mine_run(), ref_run()
elif order == "abi":
    mine, ref = ref_run("ours_abi"), ref_run()
else:
    ref, mine = ref_run(), mine_run()
print(sys.argv[1:], "n_reg", mine["n_registered"], ref["n_registered"], "iters", mine["iters"], ref["iters"])
for k in ("poses", "poses_covar", "depth", "depth_conf"):
    a, b = np.asarray(mine[k]), np.asarray(ref[k])
    eq = ffi.bits_equal(a, b)
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    print(f"  {k:12s} bit_equal={eq} max_abs={np.nanmax(d):.3e} n_diff={(a.view(np.uint32) != b.view(np.uint32)).sum()}/{a.size}")
if not ffi.bits_equal(mine["poses"], ref["poses"]):
    print("  mine", mine["poses"].ravel()[:12])
    print("  ref ", ref["poses"].ravel()[:12])
