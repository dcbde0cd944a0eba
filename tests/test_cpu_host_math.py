"""Host-side third-party arithmetic of the path, pinned to OpenCV: the reference converts rotations with cv::Rodrigues
(voldor/utils.h:52-56, geometry.cpp:184,258) — the product re-implements it (csrc/host_math.h: individually rounded
operations so that host and device agree bit for bit).  Golden vectors come from the cv2 wheel (tests/make_golden_cv2.py).
The two implementations are different algorithms, so the bar is closeness (1e-6 absolute on matrix entries, 2e-5 on
rotation vectors of noisy matrices), not bits: README and DESIGN say so."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import ffi

FP = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hm") / "host_math_harness.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17",
                           os.path.join(ffi.ROOT, "tests", "host_math_harness.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.hm_norm3.restype = C.c_double
    return lib


def test_rodrigues_matches_opencv(hm):
    g = np.load(os.path.join(ffi.ROOT, "tests", "golden", "rodrigues_cv2.npz"))
    rvecs, Rs = np.ascontiguousarray(g["rvecs"]), g["Rs"]
    n = rvecs.shape[0]
    R = np.zeros((n, 9), np.float32)
    hm.hm_rvec_to_matrix(rvecs.ctypes.data_as(FP), n, R.ctypes.data_as(FP))
    assert np.abs(R - Rs).max() < 1e-6
    # and back: rotation vectors of slightly non-orthonormal matrices (the pose pipeline's case)
    noisy = np.ascontiguousarray(g["noisy_Rs"])
    rv = np.zeros((n, 3), np.float32)
    hm.hm_matrix_to_rvec(noisy.ctypes.data_as(FP), n, rv.ctypes.data_as(FP))
    gold = g["rvecs_of_noisy"]
    small = np.linalg.norm(gold, axis=1) < 3.0  # away from the pi ambiguity
    # both project the matrix onto SO(3) first (OpenCV by SVD, host_math by a polar iteration): agreement is limited
    # by the 1e-4 perturbation's second-order effect
    assert np.abs(rv[small] - gold[small]).max() < 2e-5
    # round trip through our own pair is tight
    R2 = np.zeros((n, 9), np.float32)
    rv_clean = np.zeros((n, 3), np.float32)
    hm.hm_matrix_to_rvec(np.ascontiguousarray(Rs).ctypes.data_as(FP), n, rv_clean.ctypes.data_as(FP))
    hm.hm_rvec_to_matrix(rv_clean.ctypes.data_as(FP), n, R2.ctypes.data_as(FP))
    assert np.abs(R2 - Rs).max() < 2e-6


def test_norm3_matches_opencv_norm(hm):
    cv2 = pytest.importorskip("cv2")

    rng = np.random.default_rng(1)
    for _ in range(200):
        v = rng.normal(0, 3, 3).astype(np.float32)
        assert abs(hm.hm_norm3(v.ctypes.data_as(FP)) - cv2.norm(v.reshape(3, 1))) < 1e-12 * max(1.0, cv2.norm(v))
