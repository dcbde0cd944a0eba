"""VO front-end (window chaining of voldor_slam.py:process_vo) with a stub window solver: no GPU needed."""
import numpy as np

import synth
from voldor_b200 import formats, vo_frontend


def _gt_poses6(win):
    return np.stack([np.concatenate([vo_frontend.matrix_to_rvec(win["Rs"][i]), win["ts"][i]]) for i in range(win["N"])])


def test_rvec_matrix_round_trip():
    rng = np.random.default_rng(0)
    for _ in range(50):
        r = rng.normal(size=3) * rng.choice([1e-9, 0.01, 1.0, 3.0])
        R = vo_frontend.rvec_to_matrix(r)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
        assert np.allclose(vo_frontend.rvec_to_matrix(vo_frontend.matrix_to_rvec(R)), R, atol=1e-9)
    # rotation by pi
    R = vo_frontend.rvec_to_matrix(np.array([0, np.pi, 0]))
    assert np.allclose(vo_frontend.rvec_to_matrix(vo_frontend.matrix_to_rvec(R)), R, atol=1e-6)


def test_window_chaining_with_stub_solver():
    F, w, h = 9, 64, 48
    win = synth.make_window(w, h, F, seed=4)
    gt = _gt_poses6(win)
    calls = []

    def stub(flows, fx, fy, cx, cy, basefocal=0, disparity=None, depth_priors=None, depth_prior_poses=None,
             depth_prior_pconfs=None, config=""):
        start = len(calls) and calls[-1]["next"]
        n = flows.shape[0]
        calls.append(dict(start=start, n=n, priors=None if depth_priors is None else depth_priors.shape[0],
                          prior_poses=depth_prior_poses, config=config))
        return {"n_registered": n, "poses": gt[start:start + n].astype(np.float32),
                "poses_covar": np.zeros((n, 6, 6), np.float32), "depth": np.full((h, w), 8.0, np.float32),
                "depth_conf": np.ones((h, w), np.float32)}

    vo = vo_frontend.VisualOdometry(win["fx"], win["fy"], win["cx"], win["cy"], winsize=4, solver=stub)
    orig_step = vo.step

    def step(flows, disparity=None):
        r = orig_step(flows, disparity)
        calls[-1]["next"] = vo.fid_cur
        return r

    vo.step = step
    Tcw = vo.run(list(win["flows"]))
    assert len(Tcw) == F + 1
    want = formats.accumulate_poses(gt)
    for a, b in zip(Tcw, want):
        assert np.allclose(a, b, atol=1e-6)
    # first window has no prior, later ones hand over the temporal keyframe with its pose relative to the current frame
    assert calls[0]["priors"] is None and calls[0]["config"].startswith("--silent --meanshift_kernel_var 0.2")
    assert all(c["priors"] in (1, 2) for c in calls[1:])
    c1 = calls[1]
    T_rel = vo_frontend.T6_to_T44(c1["prior_poses"][0])
    # keyframe = frame 0, current = frame calls[1]['start']: relative pose = inverse of the chained motion
    T_chain = np.eye(4)
    for i in range(c1["start"]):
        T_chain = vo_frontend.T6_to_T44(gt[i]) @ T_chain
    assert np.allclose(T_rel, np.linalg.inv(T_chain), atol=1e-5)


def test_tracking_loss_restarts_without_priors():
    F, w, h = 4, 32, 24
    win = synth.make_window(w, h, F, seed=5)
    gt = _gt_poses6(win)
    state = {"k": 0, "priors": []}

    def stub(flows, *a, depth_priors=None, **kw):
        state["k"] += 1
        state["priors"].append(depth_priors is not None)
        if state["k"] == 2:
            return {"n_registered": 0, "poses": np.zeros((0, 6), np.float32), "poses_covar": np.zeros((0, 6, 6)),
                    "depth": np.zeros((h, w), np.float32), "depth_conf": np.zeros((h, w), np.float32)}
        n = 1
        return {"n_registered": n, "poses": gt[:n].astype(np.float32), "poses_covar": np.zeros((n, 6, 6), np.float32),
                "depth": np.full((h, w), 8.0, np.float32), "depth_conf": np.ones((h, w), np.float32)}

    vo = vo_frontend.VisualOdometry(win["fx"], win["fy"], win["cx"], win["cy"], winsize=1, solver=stub)
    Tcw = vo.run(list(win["flows"]))
    assert len(Tcw) == F + 1 and vo.lost == [1]
    assert state["priors"] == [False, True, False, True]
    assert np.allclose(Tcw[1], Tcw[2])  # the pose is kept across the lost frame


def test_huber_slope_ignores_outliers():
    rng = np.random.default_rng(3)
    x = rng.uniform(1, 10, 2000)
    y = 2.5 * x + rng.normal(0, 0.05, x.size)
    y[:300] += rng.uniform(20, 60, 300)  # 15 % gross outliers
    assert abs(vo_frontend.huber_slope(x, y) - 2.5) < 0.05
    assert abs(np.dot(x, y) / np.dot(x, x) - 2.5) > 0.3  # plain least squares is pulled away


def test_mono_scaled_mode_puts_depth_and_translation_on_the_metric_scale():
    F, w, h = 3, 64, 48
    win = synth.make_window(w, h, F, seed=6)
    gt = _gt_poses6(win)
    basefocal = 0.54 * float(win["fx"])
    true_scale = 3.0  # the monocular solver returns depth / 3, translations / 3
    disparity = (basefocal / win["depth_gt"]).astype(np.float32)

    def stub(flows, *a, **kw):
        n = flows.shape[0]
        p = gt[:n].astype(np.float32).copy()
        p[:, 3:] /= true_scale
        return {"n_registered": n, "poses": p, "poses_covar": np.ones((n, 6, 6), np.float32),
                "depth": (win["depth_gt"] / true_scale).astype(np.float32), "depth_conf": np.ones((h, w), np.float32)}

    vo = vo_frontend.VisualOdometry(win["fx"], win["fy"], win["cx"], win["cy"], basefocal=basefocal, mode="mono-scaled",
                                    winsize=F, solver=stub)
    r = vo.step(win["flows"], disparity=disparity)
    assert abs(r["scale"] - true_scale) < 1e-3
    assert np.allclose(r["depth"], win["depth_gt"], rtol=1e-3)
    assert np.allclose(r["poses"][:, 3:], gt[:F, 3:], rtol=1e-3, atol=1e-5)
    assert np.allclose(r["poses_covar"][0, 0, 0], 1) and np.allclose(r["poses_covar"][0, 4, 4], true_scale ** 2, rtol=1e-3)


def test_resize_flow_scales_vectors_and_matches_bilinear_centres():
    H, W = 6, 8
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    flow = np.stack([2 * xx + 1, 3 * yy - 2], -1)  # linear fields are reproduced exactly away from the border
    out = vo_frontend.resize_flow(flow, 4, 3)
    assert out.shape == (3, 4, 2)
    xs = (np.arange(4) + 0.5) * 2 - 0.5
    ys = (np.arange(3) + 0.5) * 2 - 0.5
    assert np.allclose(out[..., 0], (2 * xs[None, :] + 1) * 0.5 + 0 * ys[:, None])
    assert np.allclose(out[..., 1], (3 * ys[:, None] - 2) * 0.5 + 0 * xs[None, :])
    assert vo_frontend.resize_flow(flow, W, H) is flow


def test_command_line_runs_a_sequence_from_flo_files(tmp_path, monkeypatch):
    F, w, h = 5, 64, 48
    win = synth.make_window(w, h, F, seed=8)
    gt = _gt_poses6(win)
    d = tmp_path / "flows"
    d.mkdir()
    for i in range(F):
        formats.save_flow(str(d / f"{i:06d}.flo"), win["flows"][i])
    seen = {}

    def stub(flows, fx, fy, cx, cy, basefocal=0, config="", **kw):
        start = seen.setdefault("next", 0)
        n = flows.shape[0]
        seen["shape"], seen["fx"], seen["config"], seen["basefocal"] = flows.shape, fx, config, basefocal
        seen["next"] = start + n
        return {"n_registered": n, "poses": gt[start:start + n].astype(np.float32),
                "poses_covar": np.zeros((n, 6, 6), np.float32), "depth": np.full(flows.shape[1:3], 8.0, np.float32),
                "depth_conf": np.ones(flows.shape[1:3], np.float32)}

    monkeypatch.setattr(vo_frontend, "voldor", stub)
    monkeypatch.setattr(vo_frontend.VisualOdometry.__init__, "__defaults__",
                        vo_frontend.VisualOdometry.__init__.__defaults__[:5] + (stub, 1.0))
    out = tmp_path / "poses.txt"
    vo_frontend.main(["--mode", "mono", "--flow_dir", str(d), "--fx", "100", "--fy", "100", "--cx", "64", "--cy", "48",
                      "--resize", "0.5", "--save_poses", str(out)])
    assert seen["shape"][1:] == (24, 32, 2) and seen["fx"] == 50.0 and seen["basefocal"] == 25.0
    assert "--abs_resize_factor 0.5" in seen["config"] and "--pose_sample_max_depth 25.0" in seen["config"]
    assert formats.load_poses_kitti(str(out)).shape == (F + 1, 4, 4)


def test_sequence_with_the_cpu_port_as_solver_follows_the_ground_truth():
    """whole chain (windows, keyframe priors built from own outputs, covisibility steps) with a real solver: the CPU
    port of the hot path under the restated reference orchestration (test infrastructure)"""
    import ffi
    import oracle_host

    w, h, F = 160, 120, 10
    win = synth.make_window(w, h, F, seed=51)
    boot = (win["Rs"][0], win["ts"][0], synth.noisy_depth(win, 0.05))

    def solver(flows, fx, fy, cx, cy, **kw):
        first = kw.get("depth_priors") is None
        return oracle_host.run_window("cpu", flows, fx, fy, cx, cy, boot=boot if first else None, **kw)

    ffi.libc_srand(31)
    vo = vo_frontend.VisualOdometry(win["fx"], win["fy"], win["cx"], win["cy"], winsize=4, solver=solver,
                                    user_config="--no_trunc_iters 1000 --n_poses_to_sample 2048 ")
    T = vo.run(list(win["flows"]))
    assert len(T) == F + 1 and vo.lost == []
    T_gt = formats.accumulate_poses(_gt_poses6(win))
    scale = np.linalg.norm(T[-1][:3, 3]) / np.linalg.norm(T_gt[-1][:3, 3])
    for a, b in zip(T, T_gt):
        assert np.abs(a[:3, :3] - b[:3, :3]).max() < 3e-2
        assert np.linalg.norm(a[:3, 3] - b[:3, 3] * scale) < 0.1 * np.linalg.norm(T[-1][:3, 3])


def test_covisibility_score_is_the_reference_formula():
    """literal transcription of the scoring of slam_py/slam_utils.py:18-53 (visibility, histogram2d coverage, their
    combination) against vo_frontend.eval_covisibility, including a case where the coverage term decides"""
    from voldor_b200 import vo_frontend

    def reference_score(depth, T, K, mask=None, stride=4):
        h, w = depth.shape
        Iy, Ix = np.mgrid[0:h:stride, 0:w:stride]
        c2 = np.stack([Ix, Iy, np.ones_like(Ix)], axis=2).astype(np.float32).reshape(-1, 3)
        c3 = (np.linalg.inv(K) @ c2.T).T * depth[::stride, ::stride].reshape(-1, 1)
        if mask is not None:
            c3 = c3[mask[::stride, ::stride].reshape(-1)]
        c3 = (T[:3, :3] @ c3.T).T + T[:3, 3]
        pr = (K @ c3.T).T
        pr = pr[pr[:, 2] > 0]
        pr = pr[:, :2] / pr[:, 2:3]
        vis = (pr[:, 0] > 0) & (pr[:, 0] < w) & (pr[:, 1] > 0) & (pr[:, 1] < h)
        vis = np.sum(vis) / ((w // stride) * (h // stride))
        cov, _, _ = np.histogram2d(pr[:, 0], pr[:, 1], bins=(w // (2 * stride), h // (2 * stride)), range=((0, w), (0, h)))
        cov = np.sum(cov > 0) / ((w // (2 * stride)) * (h // (2 * stride)))
        return 2 * (vis * cov) / max(vis + cov, 1)

    w, h = 160, 120
    K = np.array([[128.0, 0, 80], [0, 128, 60], [0, 0, 1]])
    rng = np.random.default_rng(0)
    depth = (8 + rng.uniform(-1, 1, (h, w))).astype(np.float32)
    for t, mask in (([0.0, 0, 0], None), ([2.5, 0.5, 1.0], None), ([0, 0, 4.0], depth > 8.0), ([6.0, 0, 0], None)):
        T = np.eye(4)
        T[:3, :3] = vo_frontend.rvec_to_matrix(np.array([0.02, -0.05, 0.01]))
        T[:3, 3] = t
        a, b = vo_frontend.eval_covisibility(depth, T, K, mask), reference_score(depth, T, K, mask)
        assert abs(a - b) < 1e-6, (t, a, b)
    # forward motion keeps everything in view but shrinks the covered area: the score must drop below the visibility
    T = np.eye(4)
    T[2, 3] = -4.0
    s = vo_frontend.eval_covisibility(depth, T, K)
    assert s < 0.9
