"""Deterministic synthetic VO windows: a bumpy slanted plane seen by a smoothly moving
pinhole camera, dense forward flows between successive frames, Gaussian flow noise and one independently
moving rectangle as non-rigid outlier.  Pure numpy; used by tests, smoke() and bench.py."""
import numpy as np


def rodrigues(rvec):
    rvec = np.asarray(rvec, np.float64)
    th = np.linalg.norm(rvec)
    if th < 1e-12:
        return np.eye(3)
    k = rvec / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def _surface_depth0(u, v, fx, fy, cx, cy):
    """analytic depth of the scene in frame 0 at pixel (u, v)"""
    n = np.array([0.05, -0.1, 1.0])
    n /= np.linalg.norm(n)
    d0 = 8.0
    rx, ry = (u - cx) / fx, (v - cy) / fy
    plane = d0 / (n[0] * rx + n[1] * ry + n[2])
    return plane + 0.5 * np.sin(u / 37.0) * np.cos(v / 29.0)


def make_window(w, h, n_flows, seed=0, noise_px=0.15, outlier=True):
    rng = np.random.default_rng(seed)
    fx = fy = 0.8 * w
    cx, cy = w / 2.0, h / 2.0
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    uu, vv = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))

    # relative motions frame f -> f+1:  X_{f+1} = R_f X_f + t_f
    Rs, ts = [], []
    for f in range(n_flows):
        rvec = rng.uniform(-0.01, 0.01, 3)
        t = np.array([0.02, 0.01, 0.25]) + rng.uniform(-0.02, 0.02, 3)
        Rs.append(rodrigues(rvec))
        ts.append(t)

    # cumulative frame 0 -> f
    Rc, tc = [np.eye(3)], [np.zeros(3)]
    for f in range(n_flows):
        Rc.append(Rs[f] @ Rc[f])
        tc.append(Rs[f] @ tc[f] + ts[f])

    def depth_in_frame(f):
        # fixed-point ray/surface intersection of frame-f rays with the frame-0 surface
        rx, ry = (uu - cx) / fx, (vv - cy) / fy
        z = np.full((h, w), 8.0)
        Rinv = Rc[f].T
        for _ in range(12):
            X = np.stack([rx * z, ry * z, z], -1)
            X0 = (X - tc[f]) @ Rinv.T
            u0 = fx * X0[..., 0] / X0[..., 2] + cx
            v0 = fy * X0[..., 1] / X0[..., 2] + cy
            z = z - (X0[..., 2] - _surface_depth0(u0, v0, fx, fy, cx, cy))
        return z

    flows = np.zeros((n_flows, h, w, 2), np.float32)
    depth0 = _surface_depth0(uu, vv, fx, fy, cx, cy)
    for f in range(n_flows):
        z = depth0 if f == 0 else depth_in_frame(f)
        X = np.stack([(uu - cx) / fx * z, (vv - cy) / fy * z, z], -1)
        Xn = X @ Rs[f].T + ts[f]
        un = fx * Xn[..., 0] / Xn[..., 2] + cx
        vn = fy * Xn[..., 1] / Xn[..., 2] + cy
        fl = np.stack([un - uu, vn - vv], -1)
        fl += rng.normal(0, noise_px, fl.shape)
        if outlier:
            x0, y0 = int(0.55 * w), int(0.3 * h)
            x1, y1 = x0 + int(0.39 * w), y0 + int(0.39 * h)
            fl[y0:y1, x0:x1, 0] += 3.0
            fl[y0:y1, x0:x1, 1] += -2.0
        flows[f] = fl.astype(np.float32)

    return dict(
        w=w, h=h, N=n_flows,
        K=K.astype(np.float32),
        fx=np.float32(fx), fy=np.float32(fy), cx=np.float32(cx), cy=np.float32(cy),
        flows=flows,
        depth_gt=depth0.astype(np.float32),
        Rs=np.stack(Rs).astype(np.float32),
        ts=np.stack(ts).astype(np.float32),
        seed=seed,
    )


def noisy_depth(win, rel_sigma=0.05, seed=1):
    rng = np.random.default_rng(seed)
    return (win["depth_gt"] * (1.0 + rng.normal(0, rel_sigma, win["depth_gt"].shape))).astype(np.float32)


def perturbed_poses(win, rot_sigma=0.002, trans_sigma=0.02, seed=2):
    """poses a bootstrap / previous iteration would hand to the depth step (slightly off the truth)"""
    rng = np.random.default_rng(seed)
    Rs, ts = [], []
    for f in range(win["N"]):
        dR = rodrigues(rng.normal(0, rot_sigma, 3))
        Rs.append((dR @ win["Rs"][f].astype(np.float64)).astype(np.float32))
        ts.append((win["ts"][f] + rng.normal(0, trans_sigma, 3)).astype(np.float32))
    return np.stack(Rs), np.stack(ts)
