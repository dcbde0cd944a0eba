"""GPU parity against the reference STACK: the reference's own host code (voldor.cpp / geometry.cpp / py_export.cpp
compiled unmodified against the OpenCV stand-in, oracle/_ref/libvoldor_host_ref.so) driving the reference's own
kernels (oracle/_ref/libgpu_kernels_ref.so) — nothing restated in between except OpenCV itself.
  * this library's device-resident window (vb_py_voldor_wrapper) must reproduce that stack bit for bit;
  * the reference's host code must produce the same bits when the kernels behind its gpu_kernels.h calls are this
    library's ABI entry points: the drop-in, exercised by the caller it was written for.
Every comparison runs both libraries once per step so that their per-pixel random streams (which continue across
windows) stay at the same point of their history."""
import os

import numpy as np
import pytest

import ffi
import oracle_host
import synth
import voldor_b200

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(oracle_host.REF_HOST), reason="reference host library not built")]
KEYS = ("poses", "poses_covar", "depth", "depth_conf")


def _same(a, b, tag):
    assert a["n_registered"] == b["n_registered"], (tag, a["n_registered"], b["n_registered"])
    for k in KEYS:
        assert ffi.bits_equal(a[k], b[k]), (tag, k, ffi.mismatch_report(a[k], b[k], k))


def _epipolar(win):
    R = np.asarray(win["Rs"][0], np.float32).reshape(3, 3)
    t_pre = (R.T.astype(np.float64) @ np.asarray(win["ts"][0], np.float64)).astype(np.float32)
    return R, t_pre


@pytest.mark.parametrize("w,h,N,flags", [
    (96, 64, 3, "--max_iters 3 --no_trunc_iters 1000"),
    (160, 120, 4, "--max_iters 4"),  # default truncation rules
    (160, 120, 4, "--max_iters 3 --no_trunc_iters 1000 --lambdatwist 0"),
    (128, 96, 3, "--max_iters 3 --no_trunc_iters 1000 --exclusive_gpu_context 0"),
])
def test_mono_window_against_reference_stack(w, h, N, flags):
    win = synth.make_window(w, h, N, seed=13)
    args = (win["flows"], win["fx"], win["fy"], win["cx"], win["cy"])
    epi = _epipolar(win)
    cfg = f"--silent --n_poses_to_sample 2048 {flags}"
    # the state the reference starts the EM from: its own closed-form depth (geometry.cpp:267-287) for the injected pose
    boot = oracle_host.reference_host_bootstrap("ref", *args, epipolar=epi, config=cfg)
    ffi.libc_srand(41)
    ref1 = oracle_host.run_reference_host("ref", *args, config=cfg, epipolar=epi)
    ffi.libc_srand(41)
    abi1 = oracle_host.run_reference_host("ours_abi", *args, config=cfg, epipolar=epi)
    assert ref1["n_registered"] > 0
    _same(abi1, ref1, f"reference host over this library's kernels [{flags}]")
    ffi.libc_srand(42)
    ref2 = oracle_host.run_reference_host("ref", *args, config=cfg, epipolar=epi)
    voldor_b200.set_bootstrap_override(*boot)
    ffi.libc_srand(42)
    mine2 = voldor_b200.voldor_ex(*args, config=cfg)
    voldor_b200.set_bootstrap_override()
    _same(mine2, ref2, f"resident window vs reference stack [{flags}]")


def test_partially_truncated_window_against_reference_stack():
    w, h, N = 128, 96, 5
    win = synth.make_window(w, h, N, seed=5)
    flows = np.array(win["flows"])
    flows[3:] += np.random.default_rng(1).normal(0, 6, flows[3:].shape).astype(np.float32)
    args = (flows, win["fx"], win["fy"], win["cx"], win["cy"])
    epi = _epipolar(win)
    cfg = "--silent --n_poses_to_sample 2048 --no_trunc_iters 1 --trunc_rigidness_density 0.2 --min_iters_after_trunc 2 --max_iters 4"
    boot = oracle_host.reference_host_bootstrap("ref", *args, epipolar=epi, config=cfg)
    ffi.libc_srand(43)
    ref = oracle_host.run_reference_host("ref", *args, config=cfg, epipolar=epi)
    voldor_b200.set_bootstrap_override(*boot)
    ffi.libc_srand(43)
    mine = voldor_b200.voldor_ex(*args, config=cfg)
    voldor_b200.set_bootstrap_override()
    assert 0 < ref["n_registered"] < N
    _same(mine, ref, "partially truncated window")


def test_prior_and_disparity_windows_against_reference_stack():
    w, h, N = 128, 96, 4
    win = synth.make_window(w, h, N, seed=21)
    args = (win["flows"], win["fx"], win["fy"], win["cx"], win["cy"])
    priors = np.stack([synth.noisy_depth(win, 0.02, seed=1), synth.noisy_depth(win, 0.04, seed=2)])
    poses = np.zeros((2, 6), np.float32)
    poses[1] = [0.002, -0.001, 0.0015, 0.01, 0.0, -0.02]
    pconf = np.random.default_rng(0).uniform(0.3, 1, (2, h, w)).astype(np.float32)
    cfg = "--silent --max_iters 3 --n_poses_to_sample 2048"
    kw = dict(depth_priors=priors, depth_prior_poses=poses, depth_prior_pconfs=pconf, config=cfg)
    ffi.libc_srand(5)
    ref = oracle_host.run_reference_host("ref", *args, **kw)
    ffi.libc_srand(5)
    mine = voldor_b200.voldor_ex(*args, **kw)
    assert ref["n_registered"] > 0
    _same(mine, ref, "depth-prior window")

    basefocal = float(0.54 * win["fx"])
    rng = np.random.default_rng(3)
    disp = (basefocal / win["depth_gt"] * (1 + rng.normal(0, 0.02, (h, w)))).astype(np.float32)
    disp[10:14, 20:60] = 0
    kw = dict(basefocal=basefocal, disparity=disp, disparity_pconf=rng.uniform(0.2, 1, (h, w)).astype(np.float32),
              config=cfg)
    ffi.libc_srand(6)
    ref = oracle_host.run_reference_host("ref", *args, **kw)
    ffi.libc_srand(6)
    mine = voldor_b200.voldor_ex(*args, **kw)
    assert ref["n_registered"] > 0
    _same(mine, ref, "disparity window")
