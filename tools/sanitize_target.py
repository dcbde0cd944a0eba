"""Target of the compute-sanitizer runs (memcheck / racecheck / synccheck): every kernel family of voldor_b200 on small
inputs, through this library ONLY (the reference kernels have known latent hazards of their own, reduce_vector_sum.h:3-10,
and would drown the report).

  compute-sanitizer --tool memcheck  --kernel-name kns=vb python tools/sanitize_target.py
  compute-sanitizer --tool racecheck --kernel-name kns=vb python tools/sanitize_target.py
"""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ffi  # noqa: E402
import synth  # noqa: E402
import voldor_b200  # noqa: E402


def window(w, h, N, flags, seed=1, ctx=0):
    voldor_b200.select_context(ctx)
    win = synth.make_window(w, h, N, seed=seed)
    voldor_b200.set_bootstrap_override(win["Rs"][0], win["ts"][0], synth.noisy_depth(win, 0.05))
    r = voldor_b200.voldor_ex(win["flows"], win["fx"], win["fy"], win["cx"], win["cy"],
                              config=f"--silent --no_trunc_iters 1000 {flags}")
    voldor_b200.set_bootstrap_override()
    assert r["n_registered"] == N, (flags, r["n_registered"])
    return r


def main():
    big = "--big" in sys.argv
    # device-resident pipeline: first iteration (start-sample trials), pipelined camera loop, robust fit each iteration
    window(96, 64, 3, "--max_iters 3 --n_poses_to_sample 1024")
    window(96, 64, 3, "--max_iters 2 --n_poses_to_sample 1024 --rg_refine_last_only 0 --lambdatwist 0")
    window(75, 53, 2, "--max_iters 2 --n_poses_to_sample 700 --depth_global_prop_step 1 --depth_local_prop_width 7")
    if big:  # cluster-wide pool compaction (> 16384 hypotheses) and the global-memory partial exchange (> 32768)
        window(96, 64, 3, "--max_iters 2 --n_poses_to_sample 20000")
        window(96, 64, 3, "--max_iters 2 --n_poses_to_sample 40000")
    # two contexts at once
    ts = [threading.Thread(target=window, args=(96, 64, 3, "--max_iters 2 --n_poses_to_sample 1024", 5 + k, k)) for k in (1, 2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    voldor_b200.select_context(0)
    # ABI entry points: depth step with a prior, dense instance maps, batched solvers, pose-mode kernels, gblur, align
    lib = ffi.ours()
    win = synth.make_window(64, 48, 3, seed=3)
    Rs, ts_ = synth.perturbed_poses(win)
    ones = np.ones((3, 48, 64), np.float32)
    prior = synth.noisy_depth(win, 0.02)
    rc, d, r, c = lib.optimize_depth(64, 48, 3, 1, flows=list(win["flows"]), rig=list(ones), depth=synth.noisy_depth(win),
                                     K=win["K"], Rs=list(Rs), ts=list(ts_), priors=[prior], pconfs=[ones[0]], confs=[ones[0]],
                                     dp_Rs=[np.eye(3, dtype=np.float32)], dp_ts=[np.zeros(3, np.float32)], basefocal=30.0)
    assert rc == 0
    rc, p2, p3 = lib.collect(64, 48, 3, 1, flows=list(win["flows"]), rig=list(r), depth=d, K=win["K"], Rs=list(Rs), ts=list(ts_))
    assert rc == 0
    ok = np.isfinite(p2.sum(-1) + p3.sum(-1))
    rc, rv, tv = lib.solve_p3p(p3[ok], p2[ok], win["K"], 512)
    assert rc == 0
    rc, rv2, tv2 = lib.solve_p3p(p3[ok], p2[ok], win["K"], 512, ap3p=True)
    assert rc == 0
    fin = np.isfinite(rv.sum(-1) + tv.sum(-1))
    pool = np.concatenate([rv[fin] * 25, tv[fin]], 1).astype(np.float32)
    rc, mean, conf, used = lib.meanshift(pool, 0.1, np.zeros(6, np.float32), False)
    assert rc == 0
    rc, mean2, conf2, used2 = lib.meanshift(pool, 0.1, mean, True)
    assert rc == 0
    lib.fit_robust_gaussian(pool * 4, mean * 4, np.eye(6, dtype=np.float32) * 16 * 0.1)
    rc, _ = lib.gblur(np.random.default_rng(0).uniform(0, 1, (2, 40, 50)).astype(np.float32), 1.5)
    assert rc == 0
    imgs = np.random.default_rng(1).uniform(0, 1, (2, 48, 64)).astype(np.float32)
    depths = np.stack([win["depth_gt"], win["depth_gt"] * 1.02]).astype(np.float32)
    assert lib.align_init(imgs, depths, np.ones_like(imgs), win["K"], 40.0, 0.5) == 0
    rc, res, jac = lib.align_eval(0, 1, np.zeros(9, np.float32), np.full(9, 0.01, np.float32), 64, 48, True)
    assert rc == 0
    print("sanitize_target: all kernel families executed")


if __name__ == "__main__":
    main()
