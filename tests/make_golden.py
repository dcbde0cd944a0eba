"""Generate tests/golden/*.npz on the GPU box from the REFERENCE kernels (oracle/_ref): the reference ships no
golden vectors, so the oracle CPU port is pinned against what the reference's own CUDA code
produces on small seeded inputs.  Run:  gpurun -- 'python tests/make_golden.py'  then copy gpurun_out/golden/*
into tests/golden/."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ffi  # noqa: E402
import synth  # noqa: E402

out = os.path.join(ffi.ROOT, "gpurun_out", "golden")
os.makedirs(out, exist_ok=True)
ref = ffi.reference()

# 1. first depth step of a 64x48x3 window
seed, w, h, N = 17, 64, 48, 3
win = synth.make_window(w, h, N, seed=seed)
Rs, ts = synth.perturbed_poses(win)
ones = np.ones((N, h, w), np.float32)
rc, d, r, _ = ref.optimize_depth(w, h, N, flows=list(win["flows"]), rig=list(ones), depth=synth.noisy_depth(win),
                                 K=win["K"], Rs=list(Rs), ts=list(ts))
assert rc == 0
np.savez_compressed(os.path.join(out, "depth_step_64x48x3.npz"), seed=seed, depth=d, rigidness=r)

# 2. pose stage: instance maps of camera 1, hypotheses, pool, mean-shift
rng = np.random.default_rng(5)
rig_in = rng.uniform(0.3, 1.0, (N, h, w)).astype(np.float32)
depth_in = synth.noisy_depth(win, 0.01)
rc, p2, p3 = ref.collect(w, h, N, 1, flows=list(win["flows"]), rig=list(rig_in), depth=depth_in, K=win["K"],
                         Rs=list(win["Rs"]), ts=list(win["ts"]))
assert rc == 0
ok = np.isfinite(p2.sum(-1) + p3.sum(-1))
rc, rv, tv = ref.solve_p3p(p3[ok], p2[ok], win["K"], 1024)
fin = np.isfinite(rv.sum(1) + tv.sum(1))
pool = np.concatenate([rv[fin] * 25.0, tv[fin]], 1).astype(np.float32)
init = np.concatenate([np.zeros(3), win["ts"][1]]).astype(np.float32)
rc, m, conf, used = ref.meanshift(pool, 0.1, init, True)
np.savez_compressed(os.path.join(out, "pose_stage_64x48x3.npz"), seed=seed, rig_in=rig_in, depth_in=depth_in, p2=p2,
                    p3=p3, rvecs=rv, tvecs=tv, pool=pool, ms_init=init, ms_mean=m, ms_conf=conf, ms_iters=used)
print("golden written to", out, "instances", int(ok.sum()), "pool", pool.shape, "ms iters", used)

# 3. minimal solvers alone: 8192 hypotheses of both solvers on the instances of a 160x120x4 window (camera 2), pinned
#    against by tests/test_cpu_p3p_quad.py (host build of the product's solver arithmetic)
seed3, w3, h3, N3 = 23, 160, 120, 4
win3 = synth.make_window(w3, h3, N3, seed=seed3)
rig3 = np.random.default_rng(6).uniform(0.4, 1.0, (N3, h3, w3)).astype(np.float32)
rc, q2, q3 = ref.collect(w3, h3, N3, 2, flows=list(win3["flows"]), rig=list(rig3), depth=synth.noisy_depth(win3, 0.02),
                         K=win3["K"], Rs=list(win3["Rs"]), ts=list(win3["ts"]))
assert rc == 0
ok3 = np.isfinite(q2.sum(-1) + q3.sum(-1))
pts2, pts3 = q2[ok3], q3[ok3]
rc, rv_lt, tv_lt = ref.solve_p3p(pts3, pts2, win3["K"], 8192)
assert rc == 0
rc, rv_ap, tv_ap = ref.solve_p3p(pts3, pts2, win3["K"], 8192, ap3p=True)
assert rc == 0
np.savez_compressed(os.path.join(out, "p3p_hypotheses_160x120.npz"), seed=seed3, K=win3["K"], p2s=pts2, p3s=pts3,
                    rvecs_lambdatwist=rv_lt, tvecs_lambdatwist=tv_lt, rvecs_ap3p=rv_ap, tvecs_ap3p=tv_ap)
print("p3p golden:", pts2.shape[0], "instances; finite lambdatwist", int(np.isfinite(tv_lt.sum(1)).sum()), "ap3p",
      int(np.isfinite(tv_ap.sum(1)).sum()))
