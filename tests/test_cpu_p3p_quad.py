"""The product's minimal-solver arithmetic (voldor_b200/csrc/p3p_twist_quad.cuh) without a GPU.

Those headers are written with individually rounded operations only, so a host build evaluates exactly what the device
evaluates; tests/p3p_host_harness.cpp drives the four lanes of a quad one after the other.  Checked here against
golden hypotheses produced by the REFERENCE kernels on the B200 (tests/make_golden.py -> tests/golden/*.npz):
every finite hypothesis must carry the same translation bits, every failed one must fail here too.  The test also
repeats the search that fixed the undecided contraction sites (p3p_quad_math.cuh)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import ffi
import synth

FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int)
GOLD = os.path.join(ffi.ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("p3p") / "p3p_harness.so")
    cmd = ["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-w", "-DVBQ_SITE_SEARCH",
           "-I/usr/local/cuda/include", os.path.join(ffi.ROOT, "tests", "p3p_host_harness.cpp"), "-o", so]
    if os.path.exists(os.path.join(ffi.ROOT, "voldor_b200", "csrc", "p3p_ap3p_quad.cuh")):
        cmd.insert(-3, "-DHAVE_AP3P_QUAD")
    subprocess.check_call(cmd)
    lib = C.CDLL(so)
    lib.harness_default_sites.restype = C.c_uint
    return lib


def _solve(lib, solver, p2s, p3s, K, n_poses, sites=None):
    p2s = np.ascontiguousarray(np.concatenate([p2s, np.zeros((1, 2), np.float32)]), np.float32)  # padded like the
    p3s = np.ascontiguousarray(np.concatenate([p3s, np.zeros((1, 3), np.float32)]), np.float32)  # device buffers (Q8)
    n = p2s.shape[0] - 1
    idx = np.zeros((n_poses, 4), np.int32)
    lib.harness_indices(n_poses, n, idx.ctypes.data_as(IP))
    R = np.zeros((n_poses, 9), np.float32)
    t = np.zeros((n_poses, 3), np.float32)
    slot = np.zeros(n_poses, np.int32)
    lib.harness_set_sites(solver, lib.harness_default_sites(solver) if sites is None else sites)
    rc = lib.harness_solve(solver, p2s.ctypes.data_as(FP), p3s.ctypes.data_as(FP), idx.ctypes.data_as(IP), n_poses,
                           C.c_float(K[0, 0]), C.c_float(K[1, 1]), C.c_float(K[0, 2]), C.c_float(K[1, 2]),
                           R.ctypes.data_as(FP), t.ctypes.data_as(FP), slot.ctypes.data_as(IP))
    assert rc == 0
    lib.harness_set_sites(solver, lib.harness_default_sites(solver))
    return R, t, slot


def _agreement(t, gold_t):
    finite = np.isfinite(gold_t).all(1)
    same_failures = bool((np.isnan(t).all(1) == ~finite).all())
    exact = (t.view(np.uint32) == gold_t.view(np.uint32)).all(1)
    return int(exact[finite].sum()), int(finite.sum()), same_failures


def _pose_stage_case():
    g = np.load(os.path.join(GOLD, "pose_stage_64x48x3.npz"))
    win = synth.make_window(64, 48, 3, seed=int(g["seed"]))
    ok = np.isfinite(g["p2"].sum(-1) + g["p3"].sum(-1))
    return g["p2"][ok], g["p3"][ok], win["K"], g["tvecs"]


def test_lambdatwist_hypotheses_equal_the_reference_kernels_bit_for_bit(harness):
    p2s, p3s, K, gold_t = _pose_stage_case()
    R, t, slot = _solve(harness, 0, p2s, p3s, K, gold_t.shape[0])
    exact, finite, same_failures = _agreement(t, gold_t)
    assert same_failures and finite > 900
    assert exact == finite, (exact, finite)
    # the typical winner is a proper rotation (degenerate samples are not: the reference re-orthonormalises them
    # afterwards, rodrigues.h:82-113), and more than one lane wins over the batch (the quad is really used)
    ok = slot >= 0
    Rm = R[ok].reshape(-1, 3, 3).astype(np.float64)
    assert np.median(np.abs(Rm @ Rm.transpose(0, 2, 1) - np.eye(3)).max((1, 2))) < 1e-3
    assert len(set(slot[ok].tolist())) >= 3


def test_undecided_contraction_sites_have_exactly_one_consistent_assignment(harness):
    """all 2^9 assignments of the sites p3p_twist_quad.cuh lists: the shipped one reproduces every golden hypothesis,
    no other one does"""
    p2s, p3s, K, gold_t = _pose_stage_case()
    shipped = harness.harness_default_sites(0)
    perfect = []
    for mask in range(512):
        if ((mask >> 2) & 1 and (mask >> 3) & 1) or ((mask >> 5) & 1 and (mask >> 6) & 1):
            continue  # the two ways of folding one site exclude each other
        _, t, _ = _solve(harness, 0, p2s, p3s, K, gold_t.shape[0], sites=mask)
        exact, finite, _ = _agreement(t, gold_t)
        if exact == finite:
            perfect.append(mask)
    assert perfect == [shipped], (perfect, shipped)


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "p3p_hypotheses_160x120.npz")),
                    reason="golden set not generated yet (tests/make_golden.py on the GPU box)")
def test_lambdatwist_8192_hypotheses_golden(harness):
    g = np.load(os.path.join(GOLD, "p3p_hypotheses_160x120.npz"))
    R, t, slot = _solve(harness, 0, g["p2s"], g["p3s"], g["K"], g["tvecs_lambdatwist"].shape[0])
    exact, finite, same_failures = _agreement(t, g["tvecs_lambdatwist"])
    assert same_failures and exact == finite > 7000, (exact, finite, same_failures)


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "p3p_hypotheses_160x120.npz")),
                    reason="golden set not generated yet (tests/make_golden.py on the GPU box)")
def test_ap3p_8192_hypotheses_golden_to_rounding(harness):
    """AP3P calls cbrtf/atan2f/powf/cosf inside Ferrari's formula: the host build uses libm there, the device
    libdevice, so the host agrees with the reference kernels to rounding only (device bits: test_gpu_p3p_sites.py).
    Same failures, and nearly all translations within 1e-3 relative (a last-bit difference in a quartic root can
    flip the 4th-point choice of an ill-conditioned sample)."""
    if not hasattr(harness, "harness_solve"):
        pytest.skip("harness built without AP3P")
    g = np.load(os.path.join(GOLD, "p3p_hypotheses_160x120.npz"))
    gold = g["tvecs_ap3p"]
    R, t, slot = _solve(harness, 1, g["p2s"], g["p3s"], g["K"], gold.shape[0])
    finite = np.isfinite(gold).all(1)
    assert (np.isnan(t).all(1) == ~finite).mean() > 0.999
    both = finite & np.isfinite(t).all(1)
    rel = np.abs(t[both] - gold[both]).max(1) / np.maximum(np.abs(gold[both]).max(1), 1e-6)
    assert (rel < 1e-3).mean() > 0.98, float((rel < 1e-3).mean())
    exact = (t[both].view(np.uint32) == gold[both].view(np.uint32)).all(1).mean()
    assert exact > 0.5, exact  # the arithmetic outside the four library calls is the reference's
