"""Sliding-window visual odometry front-end: the caller side of the window solver ("callers").

Behavioural source: the VO part of reference slam_py/voldor_slam.py `process_vo` (:417-536) — which windows are
formed, which keyframe depths are handed over as priors and with which relative pose, how far the window start moves
(covisibility, slam_utils.py:18-49) and how window poses are chained into the trajectory (:506-519).  Mapping, loop
closure, frame alignment and pose-graph optimisation of the reference SLAM system are out of scope;
this front-end is what is needed to turn a flow sequence into a trajectory with `voldor_b200.voldor`.

numpy only (no OpenCV): rvec <-> matrix conversions are the Rodrigues formulas, `polish_T44` an SVD projection.
"""
import numpy as np

from . import formats
from .pyvoldor_vo import voldor


def rvec_to_matrix(rvec):
    rvec = np.asarray(rvec, np.float64)
    th = np.linalg.norm(rvec)
    if th < 1e-12:
        return np.eye(3)
    k = rvec / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def matrix_to_rvec(R):
    R = np.asarray(R, np.float64)
    s = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])  # sin(theta) * axis
    sn, c = np.linalg.norm(s), np.clip((np.trace(R) - 1) / 2, -1, 1)
    th = np.arctan2(sn, c)
    if sn > 1e-6:
        return s * (th / sn)
    if c > 0:  # theta -> 0: sin(theta)/theta -> 1
        return s
    A = (R + np.eye(3)) / 2  # theta -> pi: axis from the symmetric part
    ax = np.sqrt(np.maximum(np.diag(A), 0))
    i = int(np.argmax(ax))
    ax = A[i] / ax[i]
    return ax / np.linalg.norm(ax) * th


def T6_to_T44(p):
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = rvec_to_matrix(p[:3]), p[3:6]
    return T


def T44_to_T6(T):
    return np.concatenate([matrix_to_rvec(T[:3, :3]), T[:3, 3]])


def polish_T44(T):
    u, _, vt = np.linalg.svd(T[:3, :3])
    T[:3, :3] = u @ vt


def eval_covisibility(depth, Tc1c2, K, mask=None, stride=4):
    """covisibility score of a depth map after the motion Tc1c2 (slam_utils.py:18-53): harmonic-style combination
    2*v*c / max(v + c, 1) of the fraction v of the (strided) pixels that stay in view and the fraction c of the
    (w/2stride x h/2stride) image cells they cover.  Deviation from the reference, on purpose: K here is the intrinsic
    matrix AFTER the flow rescale (the reference builds it from the unscaled fx, fy, voldor_slam.py:175,495), which only
    matters when rescale != 1."""
    h, w = depth.shape
    Iy, Ix = np.mgrid[0:h:stride, 0:w:stride]
    rays = (np.linalg.inv(K) @ np.stack([Ix, Iy, np.ones_like(Ix)], 2).reshape(-1, 3).astype(np.float64).T).T
    X = rays * depth[::stride, ::stride].reshape(-1, 1)
    if mask is not None:
        X = X[mask[::stride, ::stride].reshape(-1)]
    X = X @ Tc1c2[:3, :3].T + Tc1c2[:3, 3]
    p = X @ np.asarray(K, np.float64).T
    p = p[p[:, 2] > 0]
    p = p[:, :2] / p[:, 2:3]
    vis = (p[:, 0] > 0) & (p[:, 0] < w) & (p[:, 1] > 0) & (p[:, 1] < h)
    visibility = vis.sum() / ((w // stride) * (h // stride))
    bins = (w // (2 * stride), h // (2 * stride))
    cells, _, _ = np.histogram2d(p[:, 0], p[:, 1], bins=bins, range=((0, w), (0, h)))
    coverage = np.sum(cells > 0) / (bins[0] * bins[1])
    return 2 * (visibility * coverage) / max(visibility + coverage, 1)


def huber_slope(x, y, epsilon=1.35, iters=30):
    """slope of y ~ k x without intercept under the Huber loss (iteratively reweighted least squares with a MAD scale);
    stands in for sklearn's HuberRegressor(fit_intercept=False) used by voldor_slam.py:485-487"""
    x, y = np.asarray(x, np.float64).ravel(), np.asarray(y, np.float64).ravel()
    k = float(np.dot(x, y) / max(np.dot(x, x), 1e-300))
    for _ in range(iters):
        r = y - k * x
        sigma = max(1.4826 * np.median(np.abs(r - np.median(r))), 1e-12)
        w = np.minimum(1.0, epsilon * sigma / np.maximum(np.abs(r), 1e-300))
        k_new = float(np.dot(w * x, y) / max(np.dot(w * x, x), 1e-300))
        if abs(k_new - k) <= 1e-12 * max(1.0, abs(k)):
            k = k_new
            break
        k = k_new
    return k


class VisualOdometry:
    """Frame-to-frame trajectory from dense flows.  `flows[i]` maps frame i to frame i+1.

    mode 'mono':        config as voldor_slam.py:153 (`--meanshift_kernel_var 0.2 --delta 1.5 --max_iters 5`)
    mode 'mono-scaled': monocular solve, then depth and translations are put on the metric scale of a disparity map
                        of the window start by a robust regression (voldor_slam.py:148-149,472-489)
    mode 'stereo':      needs `basefocal` and one disparity map per window start (voldor_slam.py:145)
    """

    def __init__(self, fx, fy, cx, cy, basefocal=0.0, mode="mono", winsize=5, user_config="", use_depth_priors=True,
                 solver=voldor, rescale=1.0):
        # voldor_slam.py:193-205 (set_cam_params): intrinsics follow the flow resize; without a baseline a virtual
        # basefocal of half the focal length fixes the depth range in which pose samples are drawn
        fx, fy, cx, cy = fx * rescale, fy * rescale, cx * rescale, cy * rescale
        basefocal = basefocal * rescale if basefocal and basefocal > 0 else (fx + fy) * 0.25
        self.fx, self.fy, self.cx, self.cy, self.basefocal = fx, fy, cx, cy, basefocal
        self.K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
        self.mode, self.winsize, self.use_depth_priors, self.solver = mode, winsize, use_depth_priors, solver
        if mode == "stereo":
            self.config = "--silent --meanshift_kernel_var 0.1 --disp_delta 1 --delta 0.2 --max_iters 4 "
        else:
            self.config = "--silent --meanshift_kernel_var 0.2 --delta 1.5 --max_iters 5 "
        self.config += f"--pose_sample_min_depth {basefocal / 200.0} --pose_sample_max_depth {basefocal / 1.0} "
        self.config += user_config
        self.depth_scaling_max_pixels = 10000  # voldor_slam.py:93-94
        self.depth_scaling_conf_thresh = 0.3
        self._scaling_rng = np.random.default_rng(0)
        self.vostep_visibility_thresh = 0.8   # voldor_slam.py:88-90
        self.spakf_visibility_thresh = 0.8
        self.depth_covis_conf_thresh = 0.1
        # state: world->current transform (the reference calls it Twc_cur), per-frame camera-to-world poses,
        # keyframe depth maps
        self.T_cur = np.eye(4)
        self.Tcw = []          # list of 4x4, one per registered frame
        self.kf_depth = {}     # fid -> (depth, depth_conf)
        self.fid_cur, self.fid_tmpkf, self.fid_spakf = 0, -1, -1
        self.lost = []

    def _priors(self):
        ids = []
        if self.use_depth_priors:
            if self.fid_tmpkf >= 0:
                ids.append(self.fid_tmpkf)
            if self.fid_spakf >= 0 and self.fid_spakf != self.fid_tmpkf:
                ids.append(self.fid_spakf)
        ids = [f for f in ids if f in self.kf_depth]
        if not ids:
            return None, None, None
        dp = np.stack([self.kf_depth[f][0] for f in ids]).astype(np.float32)
        pc = np.stack([self.kf_depth[f][1] for f in ids]).astype(np.float32)
        # pose of the keyframe relative to the current frame (voldor_slam.py:440)
        poses = np.stack([T44_to_T6(np.linalg.inv(self.T_cur @ self.Tcw[f])) for f in ids]).astype(np.float32)
        return dp, pc, poses

    def step(self, flows, disparity=None):
        """solve one window starting at the current frame; returns the solver's dict (plus 'vo_step')"""
        flows = np.ascontiguousarray(flows[: self.winsize], np.float32)
        dp, pc, poses = self._priors()
        r = self.solver(flows, self.fx, self.fy, self.cx, self.cy, basefocal=self.basefocal,
                        disparity=disparity if self.mode == "stereo" else None, depth_priors=dp,
                        depth_prior_pconfs=pc, depth_prior_poses=poses, config=self.config)
        if r["n_registered"] == 0:
            # tracking lost: keep the pose, restart without priors (voldor_slam.py:462-470)
            self.lost.append(self.fid_cur)
            self.Tcw.append(np.linalg.inv(self.T_cur))
            self.fid_tmpkf = self.fid_spakf = -1
            self.fid_cur += 1
            r["vo_step"] = 1
            return r
        if self.mode == "mono-scaled":
            # metric scale from the disparity of the window start (voldor_slam.py:472-489)
            mask = r["depth_conf"] > self.depth_scaling_conf_thresh
            src = self.basefocal / r["depth"][mask]
            dst = np.asarray(disparity, np.float32)[mask]
            if src.size > self.depth_scaling_max_pixels:
                idx = self._scaling_rng.permutation(src.size)[: self.depth_scaling_max_pixels]
                src, dst = src[idx], dst[idx]
            if src.size:
                scale = float(np.clip(1.0 / huber_slope(src, dst), 0.1, 10.0))
                r["depth"] = r["depth"] * scale
                r["poses"] = r["poses"].copy()
                r["poses"][:, 3:6] *= scale
                r["poses_covar"] = r["poses_covar"].copy()
                r["poses_covar"][:, :, 3:6] *= scale
                r["poses_covar"][:, 3:6, :] *= scale
                r["scale"] = scale
        T = [T6_to_T44(p) for p in r["poses"]]
        # how many frames to advance: while the window's depth map stays covisible (voldor_slam.py:496-504)
        vo_step, T_tmp = 0, np.eye(4)
        mask = r["depth_conf"] > self.depth_covis_conf_thresh
        for i in range(r["n_registered"]):
            vo_step += 1
            T_tmp = T[i] @ T_tmp
            if eval_covisibility(r["depth"], T_tmp, self.K, mask) < self.vostep_visibility_thresh:
                break
        for i in range(vo_step):
            self.Tcw.append(np.linalg.inv(self.T_cur))
            if i == 0:
                self.kf_depth[self.fid_cur] = (r["depth"], r["depth_conf"])
            self.T_cur = T[i] @ self.T_cur
            polish_T44(self.T_cur)
        # spatial keyframe: replaced when it is no longer covisible with the current frame (voldor_slam.py:521-531)
        if self.fid_spakf >= 0 and self.fid_spakf in self.kf_depth:
            d, c = self.kf_depth[self.fid_spakf]
            T_spa2cur = self.T_cur @ self.Tcw[self.fid_spakf]
            if eval_covisibility(d, T_spa2cur, self.K, c > self.depth_covis_conf_thresh) < self.spakf_visibility_thresh:
                self.fid_spakf = self.fid_cur
        else:
            self.fid_spakf = self.fid_cur
        self.fid_tmpkf = self.fid_cur
        self.fid_cur += vo_step
        # only the two live keyframes are needed again
        for f in [k for k in self.kf_depth if k not in (self.fid_tmpkf, self.fid_spakf)]:
            del self.kf_depth[f]
        r["vo_step"] = vo_step
        return r

    def run(self, flows, disparities=None):
        """whole sequence: flows [F,H,W,2] (or a list); returns the list of camera-to-world poses (F+1 of them)"""
        n = len(flows)
        while self.fid_cur < n:
            window = np.stack(flows[self.fid_cur:self.fid_cur + self.winsize])
            self.step(window, None if disparities is None else disparities[self.fid_cur])
        self.Tcw.append(np.linalg.inv(self.T_cur))  # last frame (voldor_slam.py:420)
        return self.Tcw

    def save_poses(self, path, format="KITTI"):
        formats.save_poses(path, self.Tcw, format)


def resize_flow(flow, w, h):
    """bilinear resize of a flow map with the vectors rescaled (voldor_slam.py:252-256; half-pixel centres like
    cv2.resize's INTER_LINEAR)"""
    H, W = flow.shape[:2]
    if (W, H) == (w, h):
        return flow
    xs = (np.arange(w) + 0.5) * (W / w) - 0.5
    ys = (np.arange(h) + 0.5) * (H / h) - 0.5
    x0 = np.clip(np.floor(xs).astype(int), 0, W - 1)
    y0 = np.clip(np.floor(ys).astype(int), 0, H - 1)
    x1, y1 = np.minimum(x0 + 1, W - 1), np.minimum(y0 + 1, H - 1)
    ax = np.clip(xs - x0, 0, 1)[None, :, None].astype(np.float32)
    ay = np.clip(ys - y0, 0, 1)[:, None, None].astype(np.float32)
    f = flow.astype(np.float32)
    top = f[y0][:, x0] * (1 - ax) + f[y0][:, x1] * ax
    bot = f[y1][:, x0] * (1 - ax) + f[y1][:, x1] * ax
    out = top * (1 - ay) + bot * ay
    out[..., 0] *= w / W
    out[..., 1] *= h / H
    return np.ascontiguousarray(out, np.float32)


def main(argv=None):
    """command line of the reference demo (demo/demo.py:4-18), restricted to what this front-end does"""
    import argparse
    import os

    ap = argparse.ArgumentParser(description="dense-flow visual odometry with voldor_b200")
    ap.add_argument("--mode", required=True, choices=["mono", "mono-scaled", "stereo"])
    ap.add_argument("--flow_dir", required=True)
    ap.add_argument("--disp_dir")
    ap.add_argument("--fx", type=float, required=True)
    ap.add_argument("--fy", type=float, required=True)
    ap.add_argument("--cx", type=float, required=True)
    ap.add_argument("--cy", type=float, required=True)
    ap.add_argument("--bf", type=float, default=0)
    ap.add_argument("--resize", type=float, default=0.5)
    ap.add_argument("--abs_resize", type=float)
    ap.add_argument("--save_poses")
    ap.add_argument("--pose_format", default="KITTI", choices=["KITTI", "TartanAir"])
    opt = ap.parse_args(argv)
    abs_resize = opt.resize if opt.abs_resize is None else opt.abs_resize
    names = sorted(os.listdir(opt.flow_dir))
    first = formats.load_flow(os.path.join(opt.flow_dir, names[0]))
    h, w = int(first.shape[0] * opt.resize), int(first.shape[1] * opt.resize)
    flows = [resize_flow(formats.load_flow(os.path.join(opt.flow_dir, n)), w, h) for n in names]
    disps = None
    if opt.mode != "mono":
        if not opt.disp_dir:
            ap.error("--disp_dir is required for stereo and mono-scaled")
        disps = []
        for n in sorted(os.listdir(opt.disp_dir)):
            d = formats.load_disparity(os.path.join(opt.disp_dir, n))
            if d.shape != (h, w):  # voldor_slam.py:309-311
                d = resize_flow(np.stack([d, d], -1), w, h)[..., 0]
            disps.append(d)
    vo = VisualOdometry(opt.fx, opt.fy, opt.cx, opt.cy, basefocal=opt.bf, mode=opt.mode, rescale=opt.resize,
                        user_config=f"--abs_resize_factor {abs_resize} ")
    vo.run(flows, disps)
    print(f"{len(vo.Tcw)} poses, {len(vo.lost)} frames lost")
    if opt.save_poses:
        vo.save_poses(opt.save_poses, opt.pose_format)


if __name__ == "__main__":
    main()
