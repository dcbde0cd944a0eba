"""ctypes front-end of oracle/liboracle_host.so: the reference's host orchestration restated over the ABI,
bound at run time to the reference kernels (ref_), to our ABI entry points (vb_) or to the CPU port (cpu_)."""
import ctypes as C
import os

import numpy as np

import ffi

HOST = os.path.join(ffi.ROOT, "oracle", "liboracle_host.so")
CPU = os.path.join(ffi.ROOT, "oracle", "libvoldor_oracle.so")
FP = C.POINTER(C.c_float)
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(HOST)
        _lib.oracle_host_bind.restype = C.c_int
        _lib.oracle_host_bind.argtypes = [C.c_char_p, C.c_char_p]
        _lib.oracle_py_voldor_wrapper.restype = C.c_int
        _lib.oracle_py_voldor_wrapper.argtypes = [FP] * 6 + [C.c_float] * 5 + [C.c_int] * 4 + [C.c_char_p] + \
            [FP] * 3 + [C.POINTER(C.c_int)] + [FP] * 4 + [C.POINTER(C.c_int), FP]
    return _lib


BACKENDS = {"ref": (ffi.REF, b"ref_"), "ours_abi": (ffi.OURS, b"vb_"), "cpu": (CPU, b"cpu_")}


def fresh_copy(backend, tag):
    """A private copy of a kernel library: dlopen of a distinct file gives distinct file statics (device caches,
    per-pixel RNG planes), i.e. the state of a FRESH process inside this process.  Returns a backend for run_window()."""
    import shutil
    import tempfile

    path, prefix = BACKENDS[backend]
    dst = os.path.join(tempfile.gettempdir(), f"{os.path.basename(path)[:-3]}_copy_{os.getpid()}_{tag}.so")
    shutil.copyfile(path, dst)
    return (dst, prefix)


def fresh_reference_copy(tag):
    return fresh_copy("ref", tag)


def _p(a):
    return None if a is None else a.ctypes.data_as(FP)


def _f(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


def run_window(backend, flows, fx, fy, cx, cy, basefocal=0, disparity=None, disparity_pconf=None, depth_priors=None,
               depth_prior_poses=None, depth_prior_pconfs=None, config="", boot=None):
    L = lib()
    # a backend is a name from BACKENDS or an explicit (library path, symbol prefix) pair
    path, prefix = BACKENDS[backend] if isinstance(backend, str) else backend
    assert L.oracle_host_bind(path.encode(), prefix) == 0, f"cannot bind {backend}"
    flows = _f(flows)
    N, h, w = flows.shape[:3]
    disparity, disparity_pconf = _f(disparity), _f(disparity_pconf)
    depth_priors, depth_prior_poses, depth_prior_pconfs = _f(depth_priors), _f(depth_prior_poses), _f(depth_prior_pconfs)
    N_dp = 0 if depth_priors is None else depth_priors.shape[0]
    bR = bt = bd = None
    if boot is not None:
        bR, bt, bd = _f(np.asarray(boot[0]).reshape(9)), _f(np.asarray(boot[1]).reshape(3)), _f(boot[2])
    poses = np.zeros((N, 6), np.float32)
    covar = np.zeros((N, 6, 6), np.float32)
    depth = np.zeros((h, w), np.float32)
    conf = np.zeros((h, w), np.float32)
    n = C.c_int(0)
    iters = C.c_int(0)
    stats = np.zeros(4, np.float32)
    L.oracle_py_voldor_wrapper(_p(flows), _p(disparity), _p(disparity_pconf), _p(depth_priors), _p(depth_prior_poses),
                               _p(depth_prior_pconfs), fx, fy, cx, cy, basefocal, N, N_dp, w, h, config.encode(),
                               _p(bR), _p(bt), _p(bd), C.byref(n), _p(poses), _p(covar), _p(depth), _p(conf),
                               C.byref(iters), _p(stats))
    k = n.value
    return {"n_registered": k, "poses": poses[:k], "poses_covar": covar[:k], "depth": depth, "depth_conf": conf,
            "iters": iters.value, "stats_ms": stats}


# ---- the reference's own host code (voldor.cpp / geometry.cpp / py_export.cpp compiled unmodified against the OpenCV
# stand-in, oracle/ref_shim/cv_min) over a kernel library bound at run time
REF_HOST = os.path.join(ffi.ROOT, "oracle", "_ref", "libvoldor_host_ref.so")
_ref_host = None


def ref_host_lib():
    global _ref_host
    if _ref_host is None:
        L = C.CDLL(REF_HOST)
        L.ref_host_bind.restype = C.c_int
        L.ref_host_bind.argtypes = [C.c_char_p, C.c_char_p]
        L.cvmin_inject_epipolar.restype = None
        L.cvmin_inject_epipolar.argtypes = [FP, FP]
        L.ref_host_py_voldor_wrapper.restype = C.c_int
        L.ref_host_py_voldor_wrapper.argtypes = [FP] * 6 + [C.c_float] * 5 + [C.c_int] * 4 + [C.c_char_p] + \
            [C.POINTER(C.c_int)] + [FP] * 4
        L.ref_host_bootstrap.restype = C.c_int
        L.ref_host_bootstrap.argtypes = [FP] + [C.c_float] * 4 + [C.c_int] * 3 + [C.c_char_p] + [FP] * 3
        _ref_host = L
    return _ref_host


def _bind_ref_host(backend, epipolar):
    L = ref_host_lib()
    path, prefix = BACKENDS[backend] if isinstance(backend, str) else backend
    assert L.ref_host_bind(path.encode(), prefix) == 0, f"cannot bind {backend}"
    if epipolar is None:
        L.cvmin_inject_epipolar(None, None)
    else:
        R, t = _f(np.asarray(epipolar[0]).reshape(9)), _f(np.asarray(epipolar[1]).reshape(3))
        L.cvmin_inject_epipolar(_p(R), _p(t))
    return L


def reference_host_bootstrap(backend, flows, fx, fy, cx, cy, epipolar, config=""):
    """VOLDOR::init + VOLDOR::bootstrap of the reference (its own compiled code) with the essential-matrix pose
    replaced by `epipolar` = (R, t as recoverPose would return them): pose of camera 0 and the closed-form depth."""
    L = _bind_ref_host(backend, epipolar)
    flows = _f(flows)
    N, h, w = flows.shape[:3]
    R, t, depth = np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros((h, w), np.float32)
    L.ref_host_bootstrap(_p(flows), fx, fy, cx, cy, N, w, h, config.encode(), _p(R), _p(t), _p(depth))
    return R.reshape(3, 3), t, depth


def run_reference_host(backend, flows, fx, fy, cx, cy, basefocal=0, disparity=None, disparity_pconf=None,
                       depth_priors=None, depth_prior_poses=None, depth_prior_pconfs=None, config="", epipolar=None):
    """py_voldor_wrapper of the reference (voldor/py_export.cpp, compiled unmodified) over `backend`'s kernels"""
    L = _bind_ref_host(backend, epipolar)
    flows = _f(flows)
    N, h, w = flows.shape[:3]
    disparity, disparity_pconf = _f(disparity), _f(disparity_pconf)
    depth_priors, depth_prior_poses, depth_prior_pconfs = _f(depth_priors), _f(depth_prior_poses), _f(depth_prior_pconfs)
    N_dp = 0 if depth_priors is None else depth_priors.shape[0]
    poses = np.zeros((N, 6), np.float32)
    covar = np.zeros((N, 6, 6), np.float32)
    depth = np.zeros((h, w), np.float32)
    conf = np.zeros((h, w), np.float32)
    n = C.c_int(0)
    L.ref_host_py_voldor_wrapper(_p(flows), _p(disparity), _p(disparity_pconf), _p(depth_priors), _p(depth_prior_poses),
                                 _p(depth_prior_pconfs), fx, fy, cx, cy, basefocal, N, N_dp, w, h, config.encode(),
                                 C.byref(n), _p(poses), _p(covar), _p(depth), _p(conf))
    k = n.value
    return {"n_registered": k, "poses": poses[:k], "poses_covar": covar[:k], "depth": depth, "depth_conf": conf}
