"""GPU-less tests of the oracle CPU port and of the host logic (rotation helpers, flag grammar).
They pin the port with closed-form properties of the algorithm and with the golden vectors
produced by the reference kernels on the GPU box (tests/golden/, tests/make_golden.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import ffi
import oracle_host
import synth

CPU = ffi.GpuKernels(oracle_host.CPU, "cpu_") if os.path.exists(oracle_host.CPU) else None
pytestmark = pytest.mark.skipif(CPU is None, reason="oracle/libvoldor_oracle.so not built (make -C oracle)")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_rigidness_of_exact_rigid_scene_is_high():
    win = synth.make_window(64, 48, 3, seed=2, noise_px=0.0, outlier=False)
    ones = np.ones((3, 48, 64), np.float32)
    rc, d, r, _ = CPU.optimize_depth(64, 48, 3, flows=list(win["flows"]), rig=list(ones), depth=win["depth_gt"],
                                     K=win["K"], Rs=list(win["Rs"]), ts=list(win["ts"]), rigidness_only=True)
    assert rc == 0
    assert np.array_equal(d, win["depth_gt"])  # rigidness-only leaves the depth alone
    inner = r[:, 4:-4, 4:-4]
    assert np.median(inner) > 0.55 and inner.min() >= 0 and inner.max() <= 1  # p/(p+mu) at zero residual is ~0.6-0.8


def test_outlier_rectangle_gets_low_rigidness():
    win = synth.make_window(96, 64, 2, seed=3, noise_px=0.05, outlier=True)
    ones = np.ones((2, 64, 96), np.float32)
    rc, _, r, _ = CPU.optimize_depth(96, 64, 2, flows=list(win["flows"]), rig=list(ones), depth=win["depth_gt"],
                                     K=win["K"], Rs=list(win["Rs"]), ts=list(win["ts"]), rigidness_only=True)
    x0, y0 = int(0.55 * 96), int(0.3 * 64)
    inside = r[0, y0 + 3:y0 + 20, x0 + 3:x0 + 30]
    outside = r[0, 5:15, 5:40]
    assert inside.mean() < 0.2 and outside.mean() > 0.6


def test_depth_step_moves_towards_ground_truth():
    win = synth.make_window(64, 48, 3, seed=4, noise_px=0.05, outlier=False)
    depth0 = synth.noisy_depth(win, 0.15)
    ones = np.ones((3, 48, 64), np.float32)
    rc, d, r, _ = CPU.optimize_depth(64, 48, 3, flows=list(win["flows"]), rig=list(ones), depth=depth0, K=win["K"],
                                     Rs=list(win["Rs"]), ts=list(win["ts"]))
    e0 = np.median(np.abs(depth0 / win["depth_gt"] - 1))
    e1 = np.median(np.abs(d / win["depth_gt"] - 1))
    assert e1 < 0.5 * e0


def test_collector_identity_pose_follows_the_flow():
    w, h, N = 48, 32, 2
    win = synth.make_window(w, h, N, seed=5)
    rig = np.ones((N, h, w), np.float32)
    depth = np.full((h, w), 5.0, np.float32)
    I = [np.eye(3, dtype=np.float32)] * N
    z = [np.zeros(3, np.float32)] * N
    rc, p2, p3 = CPU.collect(w, h, N, 0, flows=list(win["flows"]), rig=list(rig), depth=depth, K=win["K"], Rs=I, ts=z)
    assert rc == 0
    ys, xs = np.mgrid[2:h - 2, 2:w - 2]
    ok = np.isfinite(p2[ys, xs, 0])
    assert ok.mean() > 0.8
    # p2 = pixel + bilinear flow at the pixel (active_idx 0: one traced flow), p3 = back-projected pixel
    exp = np.stack([xs, ys], -1) + win["flows"][0][ys, xs]
    assert np.abs(p2[ys, xs][ok] - exp[ok]).max() < 2e-2
    assert np.allclose(p3[ys, xs, 2][ok], 5.0)
    # NaN pattern is all-or-nothing per pixel
    assert np.array_equal(np.isnan(p2[..., 0]), np.isnan(p3[..., 2]))


def test_lambdatwist_recovers_exact_pose():
    rng = np.random.default_rng(0)
    K = np.array([[400, 0, 320], [0, 400, 240], [0, 0, 1]], np.float32)
    R = synth.rodrigues([0.02, -0.01, 0.015])
    t = np.array([0.05, -0.02, 0.3])
    X = np.stack([rng.uniform(-2, 2, 500), rng.uniform(-1.5, 1.5, 500), rng.uniform(4, 9, 500)], -1)
    Xc = X @ R.T + t
    uv = np.stack([400 * Xc[:, 0] / Xc[:, 2] + 320, 400 * Xc[:, 1] / Xc[:, 2] + 240], -1)
    rc, rv, tv = CPU.solve_p3p(X.astype(np.float32), uv.astype(np.float32), K, 256)
    assert rc == 0
    fin = np.isfinite(rv).all(1)
    assert fin.mean() > 0.7
    assert np.median(np.abs(tv[fin] - t).max(1)) < 1e-3
    assert np.median(np.abs(rv[fin] - np.array([0.02, -0.01, 0.015])).max(1)) < 1e-3


def test_meanshift_finds_the_mode_and_robust_fit_the_covariance():
    rng = np.random.default_rng(1)
    mode = np.array([0.5, -0.2, 0.1, 0.3, 0.0, -0.4], np.float32)
    pts = (mode + rng.normal(0, 0.05, (4000, 6))).astype(np.float32)
    pts[:800] = rng.uniform(-3, 3, (800, 6)).astype(np.float32)
    ffi.libc_srand(3)
    rc, m, conf, used = CPU.meanshift(pts, 0.1, mode + 0.2, True)
    assert rc == 0 and 0 < used <= 100
    assert np.abs(m - mode).max() < 0.02 and 0.3 < conf < 1.0
    ffi.libc_srand(3)
    rc, m2, conf2, used2 = CPU.meanshift(pts, 0.1, np.zeros(6, np.float32), False)
    assert np.abs(m2 - mode).max() < 0.02
    cov0 = (np.eye(6) * 0.1).astype(np.float32)
    rc, mean, cov, dens, it = CPU.fit_robust_gaussian(pts, m, cov0)
    assert rc == 0 and it > 0
    # hard truncation at Mahalanobis 3 in 6-D keeps only ~83% of a Gaussian per iteration, so the fixed point of
    # the truncated EM sits below the true sigma (0.05); covariance stays near-isotropic
    sig = np.sqrt(np.diag(cov))
    assert 0.02 < sig.min() and sig.max() < 0.055 and sig.max() / sig.min() < 1.4
    assert 0.2 < dens < 0.9
    rc, mean_b, cov_b, _, _ = CPU.fit_robust_gaussian(pts, m, np.zeros((6, 6), np.float32))
    assert rc == 1 and np.array_equal(cov_b, np.zeros((6, 6), np.float32))  # unreliable: outputs untouched


def test_cpu_window_converges():
    w, h, N = 96, 64, 3
    win = synth.make_window(w, h, N, seed=11)
    boot = (win["Rs"][0], win["ts"][0], synth.noisy_depth(win, 0.05))
    cfg = "--silent --max_iters 3 --no_trunc_iters 1000 --n_poses_to_sample 1024"
    ffi.libc_srand(7)
    r = oracle_host.run_window("cpu", win["flows"], win["fx"], win["fy"], win["cx"], win["cy"], config=cfg, boot=boot)
    assert r["n_registered"] == N and r["iters"] == 3
    for f in range(N):
        t_est, t_gt = r["poses"][f, 3:], win["ts"][f]
        cosang = t_est @ t_gt / np.linalg.norm(t_est) / np.linalg.norm(t_gt)
        assert cosang > 0.99, (f, cosang)
    scale = np.linalg.norm(r["poses"][0, 3:]) / np.linalg.norm(win["ts"][0])
    ratio = r["depth"] / win["depth_gt"] / scale
    assert abs(np.median(ratio) - 1) < 0.05
    assert np.abs(np.mean([np.linalg.norm(p[3:]) for p in r["poses"]]) - 1) < 1e-3  # world scale normalised
    assert r["poses_covar"].shape == (N, 6, 6) and np.isfinite(r["poses_covar"]).all()


def test_rotation_helpers_roundtrip():
    L = oracle_host.lib()
    rng = np.random.default_rng(2)
    for _ in range(50):
        rv = rng.normal(0, 0.5, 3).astype(np.float32)
        R = np.zeros(9, np.float32)
        back = np.zeros(3, np.float32)
        L.oracle_rvec_to_matrix(rv.ctypes.data_as(ffi.FP), R.ctypes.data_as(ffi.FP))
        L.oracle_matrix_to_rvec(R.ctypes.data_as(ffi.FP), back.ctypes.data_as(ffi.FP))
        assert np.allclose(R.reshape(3, 3), synth.rodrigues(rv), atol=2e-7)
        assert np.allclose(back, rv, atol=5e-7)
    R = np.eye(3, dtype=np.float32).reshape(-1)
    back = np.ones(3, np.float32)
    L.oracle_matrix_to_rvec(R.ctypes.data_as(ffi.FP), back.ctypes.data_as(ffi.FP))
    assert np.array_equal(back, np.zeros(3, np.float32))
    # every quadrant of the library's own sin/cos (csrc/host_math.h det::sincos), angles up to ~25 rad
    for k in range(400):
        rv = (rng.normal(0, 1, 3) * (0.02 + 0.06 * k)).astype(np.float32)
        R = np.zeros(9, np.float32)
        L.oracle_rvec_to_matrix(rv.ctypes.data_as(ffi.FP), R.ctypes.data_as(ffi.FP))
        assert np.allclose(R.reshape(3, 3), synth.rodrigues(rv), atol=3e-7), (k, rv)


def test_flag_grammar():
    L = oracle_host.lib()
    out = np.zeros(12, np.float32)
    L.oracle_config_probe(b"", out.ctypes.data_as(ffi.FP))
    assert list(out[:4]) == [5, 2, 8192, 0] and out[5] == 1 and out[7] == 10 and out[9] == 1 and out[10] == 100
    L.oracle_config_probe(b"--silent --max_iters 30.9 --no_trunc_iters 1000 --n_poses_to_sample 4096 --lambda 0.2 "
                          b"--lambdatwist 0 --fb_smooth 0", out.ctypes.data_as(ffi.FP))
    # numeric flags go through stod and are narrowed to the field type (Q16): 30.9 -> 30
    assert list(out[:4]) == [30, 1000, 4096, 1] and abs(out[4] - 0.2) < 1e-7 and out[5] == 0 and out[11] == 0


def _golden(name):
    p = os.path.join(GOLD, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} not generated yet (tests/make_golden.py on the GPU box)")
    return np.load(p)


def test_cpu_port_against_reference_golden_depth_step():
    """first optimize_depth call of a 64x48x3 window: the reference kernels' output (golden) vs the CPU port.
    Same XORWOW streams, same schedule; libm vs libdevice and texture-unit rounding flip a few argmins."""
    g = _golden("depth_step_64x48x3.npz")
    win = synth.make_window(64, 48, 3, seed=int(g["seed"]))
    Rs, ts = synth.perturbed_poses(win)
    ones = np.ones((3, 48, 64), np.float32)
    rc, d, r, _ = CPU.optimize_depth(64, 48, 3, flows=list(win["flows"]), rig=list(ones), depth=synth.noisy_depth(win),
                                     K=win["K"], Rs=list(Rs), ts=list(ts))
    rel_d = np.abs(d - g["depth"]) / np.abs(g["depth"])
    assert (rel_d < 1e-3).mean() > 0.85, (rel_d < 1e-3).mean()
    assert np.abs(r - g["rigidness"]).mean() < 0.03


def test_cpu_port_against_reference_golden_collect_and_pose():
    g = _golden("pose_stage_64x48x3.npz")
    win = synth.make_window(64, 48, 3, seed=int(g["seed"]))
    rc, p2, p3 = CPU.collect(64, 48, 3, 1, flows=list(win["flows"]), rig=list(g["rig_in"]), depth=g["depth_in"],
                             K=win["K"], Rs=list(win["Rs"]), ts=list(win["ts"]))
    same_nan = np.isnan(p2[..., 0]) == np.isnan(g["p2"][..., 0])
    assert same_nan.mean() > 0.995
    both = np.isfinite(p2[..., 0]) & np.isfinite(g["p2"][..., 0])
    assert np.abs(p2[both] - g["p2"][both]).max() < 5e-2 and np.abs(p3[both] - g["p3"][both]).max() < 1e-4
    # mean-shift / robust fit on the reference's own hypothesis pool: identical tree order -> tight agreement
    pool = g["pool"]
    rc, m, conf, used = CPU.meanshift(pool, 0.1, g["ms_init"], True)
    assert used == int(g["ms_iters"]) and np.abs(m - g["ms_mean"]).max() < 1e-5 and abs(conf - float(g["ms_conf"])) < 1e-5


def _cpu_align():
    return ffi.GpuKernels(oracle_host.CPU, "cpu_")


def test_cpu_align_frame_identity_and_analytic_translation():
    """frame-alignment port (oracle/cpu_kernels.cpp, reference gpu-kernels/align_frame.cu): closed-form cases"""
    L = _cpu_align()
    w, h, N = 48, 36, 2
    K = np.array([[40, 0, 24], [0, 40, 18], [0, 0, 1]], np.float32)
    D, vbf = 5.0, 30.0
    depths = np.full((N, h, w), D, np.float32)
    weights = np.full((N, h, w), 0.7, np.float32)
    images = np.zeros((N, h, w), np.float32)
    assert L.align_init(images, depths, weights, K, vbf, 0.0) == 0
    z9 = np.zeros(9, np.float32)
    # identical frames, identity poses: every point lands on itself
    rc, res, jac = L.align_eval(0, 1, z9, z9, w, h, True)
    assert rc == 0 and np.isfinite(res).all() and np.abs(res).max() < 1e-6 and np.abs(jac).max() < 1e-4
    # pure translation delta along the optical axis in front of a fronto-parallel plane at depth D:
    # offset = delta along the normal, residual = sqrt(log(1 + wgt * (vbf / (D (D + delta)))^2 * delta^2 / 2))
    delta = 0.4
    p = z9.copy()
    p[5] = delta
    rc, res, jac = L.align_eval(0, 1, p, z9, w, h, True)
    cy, cx = h // 2, w // 2
    drw = (vbf / (D * (D + delta))) ** 2
    raw = drw * 0.5 * delta ** 2
    assert abs(res[cy, cx] - np.sqrt(np.log1p(0.7 * raw))) < 1e-5
    # Jacobian w.r.t. tz before the loss is drw * delta; the loss multiplies by 0.5 / sqrt(loss) / (1 + r) * wgt
    k = 0.5 / np.sqrt(np.log1p(0.7 * raw)) / (1 + 0.7 * raw) * 0.7
    assert abs(jac[cy, cx, 5] - drw * delta * k) < 1e-4 * abs(drw * delta * k) + 1e-7
    assert abs(jac[cy, cx, 3]) < 1e-6 and abs(jac[cy, cx, 4]) < 1e-6  # symmetric at the principal point
    assert np.all(jac[..., 7:] == 0)  # no photometric term without images
    # un-weighted evaluation differs only by the weight
    rc, res_u, _ = L.align_eval(0, 1, p, z9, w, h, False)
    assert abs(res_u[cy, cx] - np.sqrt(np.log1p(raw))) < 1e-5
    # moving the point behind z = 1 or out of the image invalidates the pixel (NaN), Jacobian stays 0
    p[5] = -4.5
    rc, res, jac = L.align_eval(0, 1, p, z9, w, h, True)
    assert np.isnan(res).all() and np.all(jac == 0)


def test_cpu_align_frame_photometric_term():
    L = _cpu_align()
    w, h, N = 40, 30, 2
    K = np.array([[35, 0, 20], [0, 35, 15], [0, 0, 1]], np.float32)
    depths = np.full((N, h, w), 4.0, np.float32)
    weights = np.ones((N, h, w), np.float32)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = (0.02 * xx + 0.01 * yy).astype(np.float32)
    images = np.stack([img, img])
    crw = 0.5
    assert L.align_init(images, depths, weights, K, 25.0, crw) == 0
    p_ref = np.zeros(9, np.float32)
    p_tar = np.zeros(9, np.float32)
    p_ref[8] = 0.3     # colour offset of the reference frame
    p_tar[7] = 0.2     # colour scale of the target frame
    rc, res, jac = L.align_eval(0, 1, p_ref, p_tar, w, h, False)
    y, x = 12, 17      # interior pixel, identity geometry: only the colour residual is non-zero
    c_ref = img[y, x] + 0.3
    c_tar = img[y, x] * np.exp(0.0 - 0.2)
    raw = crw * 0.5 * (c_ref - c_tar) ** 2
    assert abs(res[y, x] - np.sqrt(np.log1p(raw))) < 1e-5
    k = 0.5 / np.sqrt(np.log1p(raw)) / (1 + raw)
    assert abs(jac[y, x, 8] - crw * (c_ref - c_tar) * k) < 1e-5
    assert abs(jac[y, x, 7] - crw * (c_tar - c_ref) * c_tar * k) < 1e-5
    # image gradient (0.3/0.1/0.1 stencil over 2 pixels -> 0.5 * 2 * slope) enters the translation Jacobian
    gx, gy = 0.02, 0.01
    dc = c_tar - c_ref
    want_tx = crw * dc * (gx * K[0, 0] / 4.0) * k
    want_ty = crw * dc * (gy * K[1, 1] / 4.0) * k
    assert abs(jac[y, x, 3] - want_tx) < 2e-3 * abs(want_tx) + 1e-7
    assert abs(jac[y, x, 4] - want_ty) < 2e-3 * abs(want_ty) + 1e-7
