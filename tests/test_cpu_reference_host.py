"""The restated host orchestration (oracle/host_voldor.cpp, the checker of every window-level parity test) against the
reference's OWN host code: voldor/voldor.cpp, geometry.cpp, py_export.cpp (+ utils.h, config.h, voldor.h) compiled
unmodified against a self-written OpenCV stand-in (oracle/ref_shim/cv_min) into oracle/_ref/libvoldor_host_ref.so.
Both drive the same kernels — here the CPU port, so the test runs without a GPU — from the same libc rand() state;
outputs must be identical bit for bit: iteration loop, truncation logic, NULL-cached call patterns, host compaction,
pool scaling and world-scale normalisation of the restatement are the reference's."""
import os

import numpy as np
import pytest

import ffi
import oracle_host
import synth

pytestmark = pytest.mark.skipif(not os.path.exists(oracle_host.REF_HOST), reason="reference host library not built")
KEYS = ("poses", "poses_covar", "depth", "depth_conf")
_n = [0]


def _fresh_pair():
    """two private copies of the CPU kernel library: both sides start with the statics of a fresh process"""
    _n[0] += 1
    return oracle_host.fresh_copy("cpu", f"a{_n[0]}"), oracle_host.fresh_copy("cpu", f"b{_n[0]}")


def _same(a, b, tag):
    assert a["n_registered"] == b["n_registered"], (tag, a["n_registered"], b["n_registered"])
    for k in KEYS:
        assert ffi.bits_equal(a[k], b[k]), (tag, k, ffi.mismatch_report(a[k], b[k], k))


def _mono(seed, w=64, h=48, N=3):
    win = synth.make_window(w, h, N, seed=seed)
    args = (win["flows"], win["fx"], win["fy"], win["cx"], win["cy"])
    # what recoverPose returns: rotation and the translation BEFORE the reference's own "t = R t" (geometry.cpp:331)
    R = np.asarray(win["Rs"][0], np.float32).reshape(3, 3)
    t_pre = (R.T.astype(np.float64) @ np.asarray(win["ts"][0], np.float64)).astype(np.float32)
    return win, args, (R, t_pre)


@pytest.mark.parametrize("flags", [
    "--no_trunc_iters 1000",
    "",  # default truncation rules
    "--no_trunc_iters 1000 --lambdatwist 0",
    "--no_trunc_iters 1000 --rg_refine_last_only 0",
    "--no_trunc_iters 1000 --rg_refine 0",
    "--no_trunc_iters 1000 --norm_world_scale 0",
    "--no_trunc_iters 1000 --exclusive_gpu_context 0",
    "--no_trunc_iters 1000 --optimize_depth 0",
    "--no_trunc_iters 0 --trunc_sample_density 1e9 --min_iters_after_trunc 2",  # every camera truncated at once
    "--no_trunc_iters 1 --trunc_rigidness_density 0.9 --min_iters_after_trunc 1",
])
def test_mono_window_reference_host_equals_restatement(flags):
    win, args, epi = _mono(seed=5)
    cfg = f"--silent --max_iters 3 --n_poses_to_sample 512 {flags}"
    boot = oracle_host.reference_host_bootstrap("cpu", *args, epipolar=epi, config=cfg)
    # the reference's bootstrap: camera 0 = injected rotation, t = R t; depth = its closed form, clamped
    assert np.allclose(boot[1], win["ts"][0], atol=1e-5) and np.isfinite(boot[2]).all() and (boot[2] >= 1e-2).all()
    ka, kb = _fresh_pair()
    ffi.libc_srand(3)
    ref = oracle_host.run_reference_host(ka, *args, config=cfg, epipolar=epi)
    ffi.libc_srand(3)
    mine = oracle_host.run_window(kb, *args, config=cfg, boot=boot)
    _same(mine, ref, f"mono [{flags}]")


def test_depth_prior_window_reference_host_equals_restatement():
    w, h, N = 64, 48, 3
    win = synth.make_window(w, h, N, seed=8)
    priors = np.stack([synth.noisy_depth(win, 0.02, seed=1), synth.noisy_depth(win, 0.04, seed=2)])
    poses = np.zeros((2, 6), np.float32)
    poses[1] = [0.002, -0.001, 0.0015, 0.01, 0.0, -0.02]
    pconf = np.random.default_rng(0).uniform(0.3, 1, (2, h, w)).astype(np.float32)
    args = (win["flows"], win["fx"], win["fy"], win["cx"], win["cy"])
    for kw in (dict(depth_priors=priors, depth_prior_poses=poses),
               dict(depth_priors=priors, depth_prior_poses=poses, depth_prior_pconfs=pconf)):
        cfg = "--silent --max_iters 3 --n_poses_to_sample 512"
        ka, kb = _fresh_pair()
        ffi.libc_srand(4)
        ref = oracle_host.run_reference_host(ka, *args, config=cfg, **kw)
        ffi.libc_srand(4)
        mine = oracle_host.run_window(kb, *args, config=cfg, **kw)
        assert ref["n_registered"] > 0
        _same(mine, ref, "depth priors")


def test_disparity_window_reference_host_equals_restatement():
    w, h, N = 80, 40, 3
    win = synth.make_window(w, h, N, seed=9)
    basefocal = float(0.54 * win["fx"])
    rng = np.random.default_rng(3)
    disp = (basefocal / win["depth_gt"] * (1 + rng.normal(0, 0.02, (h, w)))).astype(np.float32)
    disp[5:9, 10:30] = 0  # missing stereo matches
    pconf = rng.uniform(0.2, 1, (h, w)).astype(np.float32)
    args = (win["flows"], win["fx"], win["fy"], win["cx"], win["cy"])
    for kw in (dict(disparity=disp), dict(disparity=disp, disparity_pconf=pconf)):
        cfg = "--silent --max_iters 3 --n_poses_to_sample 512"
        ka, kb = _fresh_pair()
        ffi.libc_srand(6)
        ref = oracle_host.run_reference_host(ka, *args, basefocal=basefocal, config=cfg, **kw)
        ffi.libc_srand(6)
        mine = oracle_host.run_window(kb, *args, basefocal=basefocal, config=cfg, **kw)
        assert ref["n_registered"] > 0
        _same(mine, ref, "disparity")


def test_partial_truncation_reference_host_equals_restatement():
    """the two last flows are buried in noise: the window is cut at camera 3 (voldor.cpp:189-196) and continues for
    min_iters_after_trunc more iterations on the remaining cameras"""
    win, args, epi = _mono(seed=5, N=5)
    flows = np.array(args[0])
    flows[3:] += np.random.default_rng(1).normal(0, 6, flows[3:].shape).astype(np.float32)
    args = (flows,) + args[1:]
    cfg = "--silent --n_poses_to_sample 512 --no_trunc_iters 1 --trunc_rigidness_density 0.2 --min_iters_after_trunc 2 --max_iters 4"
    boot = oracle_host.reference_host_bootstrap("cpu", *args, epipolar=epi, config=cfg)
    ka, kb = _fresh_pair()
    ffi.libc_srand(3)
    ref = oracle_host.run_reference_host(ka, *args, config=cfg, epipolar=epi)
    ffi.libc_srand(3)
    mine = oracle_host.run_window(kb, *args, config=cfg, boot=boot)
    assert ref["n_registered"] == 3
    _same(mine, ref, "partial truncation")


def test_opencv_stand_in_semantics():
    """the OpenCV behaviours the reference's host code relies on, as implemented by oracle/ref_shim/cv_min: views
    write through, copies of a header are shallow, clone is deep, at<T>() indexes in units of T, row-major comma
    initialiser, 3x3 inverse, scalar assignment in place, type conversion, reductions, s / 0 = 0"""
    import ctypes as C

    L = oracle_host.ref_host_lib()
    L.cvmin_selftest.restype = C.c_int
    out = np.zeros(32, np.float64)
    n = L.cvmin_selftest(out.ctypes.data_as(C.POINTER(C.c_double)))
    K = np.array([[500, 0, 320], [0, 510, 240], [0, 0, 1]], np.float64)
    Ki = np.linalg.inv(K)
    want = [16, 9, 8, -1, -3, 2, 11, 7, 9, 3, 240, Ki[0, 0], Ki[0, 2], 3, 0, 1, 6, np.sqrt(6), 1, 0, 0]
    assert n == len(want)
    assert np.allclose(out[:n], want, rtol=1e-6, atol=1e-7), (out[:n], want)
