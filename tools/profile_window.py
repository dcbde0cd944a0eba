"""Profiling driver: one BASELINE configs[1] window (640x480x8) with a few EM iterations through the public
binding; meant to be wrapped by ncu (B200_PROFILING.md).  Numbers printed under a profiler are not bench values."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
import voldor_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--w", type=int, default=640)
ap.add_argument("--h", type=int, default=480)
ap.add_argument("--flows", type=int, default=8)
ap.add_argument("--windows", type=int, default=1)
a = ap.parse_args()
win = synth.make_window(a.w, a.h, a.flows, seed=100)
boot = (win["Rs"][0], win["ts"][0], synth.noisy_depth(win, 0.05, seed=7))
voldor_b200.set_bootstrap_override(*boot)
cfg = f"--silent --max_iters {a.iters} --no_trunc_iters 1000 --n_poses_to_sample 8192"
for _ in range(a.windows):
    r = voldor_b200.voldor_ex(win["flows"], win["fx"], win["fy"], win["cx"], win["cy"], config=cfg)
print("n_registered", r["n_registered"], "iters", r["iters"], "stats_ms", r["stats_ms"])
import ctypes as C  # noqa: E402
from voldor_b200.pyvoldor_vo import load_library  # noqa: E402
cnt = (C.c_longlong * 5)()
load_library().vb_profile_counters(cnt)
print("meanshift runs/iters/trials, robust runs/iters (all windows):", list(cnt))
ph = (C.c_longlong * 24)()
if load_library().vb_debug_pose_mode_phases(ph) == 0:
    tot = sum(ph[:6]) or 1
    print("robust-fit phase cycles (LU, E-step, level1, exchange, level2, M-step):", list(ph[:6]),
          [round(v / tot, 3) for v in ph[:6]])
    tot = sum(ph[8:16]) or 1
    print("mean-shift phase cycles (pool build, staging, weights, level1, exchange, level2, update, tail):",
          list(ph[8:16]), [round(v / tot, 3) for v in ph[8:16]])
