"""Monocular bootstrap (host code of the product library, no GPU): pose of the first frame pair and closed-form
depth from one dense flow.  Functional check against the synthetic ground truth — the reference delegates this
step to OpenCV (voldor/geometry.cpp:267-332), so there is no bit-parity target."""
import ctypes as C
import numpy as np
import synth
from voldor_b200.pyvoldor_vo import load_library


def _boot(flow, K):
    lib = load_library()
    h, w = flow.shape[:2]
    R, t, depth = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros((h, w), np.float32)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    lib.vb_bootstrap_from_flow.restype = C.c_int
    rc = lib.vb_bootstrap_from_flow(fp(np.ascontiguousarray(flow, np.float32)), w, h,
                                    fp(np.ascontiguousarray(K, np.float32).ravel()), fp(R), fp(t), fp(depth))
    return rc, R.reshape(3, 3), t, depth


def test_bootstrap_recovers_pose_and_depth_up_to_scale():
    win = synth.make_window(160, 120, 2, seed=5, noise_px=0.1, outlier=True)
    rc, R, t, depth = _boot(win["flows"][0], win["K"])
    assert rc == 0
    R_gt, t_gt = win["Rs"][0].astype(np.float64), win["ts"][0].astype(np.float64)
    assert np.abs(R - R_gt).max() < 5e-3
    # t := R t (reference quirk, geometry.cpp:330); direction only
    t_dir = R_gt @ t_gt / np.linalg.norm(t_gt)
    assert np.dot(t / np.linalg.norm(t), t_dir) > 0.99
    # depth up to the global scale 1/|t_gt|, on the rigid part of the scene
    d_gt = win["depth_gt"] / np.linalg.norm(t_gt)
    rel = np.abs(depth - d_gt) / d_gt
    assert np.median(rel) < 0.1


def test_bootstrap_rejects_degenerate_flow():
    K = np.array([[100, 0, 32], [0, 100, 24], [0, 0, 1]], np.float32)
    rc, *_ = _boot(np.full((48, 64, 2), np.nan, np.float32), K)
    assert rc == 1
