"""Python binding of the VO window, drop-in for the reference's Cython module.

Mirrors reference slam_py/install/pyvoldor_vo.pyx:14-70: same function name, argument names, defaults and
returned dict (`n_registered`, `poses[:n,6]`, `poses_covar[:n,6,6]`, `depth[H,W]`, `depth_conf[H,W]`), so
slam_py/voldor_slam.py:447-457 can `import voldor_b200.pyvoldor_vo as pyvoldor` unchanged.  The binding is a
ctypes front-end over the C ABI in include/voldor_b200.h (vb_py_voldor_wrapper)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvoldor_b200.so")
_lib = None

FP = C.POINTER(C.c_float)


def load_library():
    """dlopen the in-tree CUDA library; fails loudly when it has not been built"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `make lib` (or __graft_entry__.build()); "
            "voldor_b200 has no CPU fallback")
    lib = C.CDLL(_LIB_PATH, mode=C.RTLD_LOCAL)
    lib.vb_py_voldor_wrapper_ex.restype = C.c_int
    lib.vb_py_voldor_wrapper_ex.argtypes = [FP] * 6 + [C.c_float] * 5 + [C.c_int] * 4 + [C.c_char_p] + \
        [C.POINTER(C.c_int)] + [FP] * 4 + [C.POINTER(C.c_int), FP]
    lib.vb_set_bootstrap_override.restype = C.c_int
    lib.vb_set_bootstrap_override.argtypes = [C.c_int, FP, FP, FP, C.c_int, C.c_int]
    lib.vb_version.restype = C.c_char_p
    lib.vb_context_select.restype = C.c_int
    lib.vb_context_select.argtypes = [C.c_int]
    lib.vb_context_srand.argtypes = [C.c_uint]
    _lib = lib
    return lib


def set_device(device):
    """CUDA device of this process' library state; applies to every host thread that calls into the library"""
    rc = load_library().vb_set_device(int(device))
    if rc != 0:
        raise RuntimeError(f"vb_set_device({device}) failed with CUDA error {rc}")


def select_context(ctx):
    """Bind the calling Python thread to execution context `ctx` (0 .. vb_context_max()-1); returns the previous one.
    Contexts are independent copies of the library's device state: windows issued from different threads on
    different contexts run concurrently on the GPU (ctypes releases the GIL for the duration of a call)."""
    prev = load_library().vb_context_select(int(ctx))
    if prev < 0:
        raise ValueError(f"invalid execution context {ctx}")
    return prev


def context_srand(seed):
    """srand() of the current context's start-sample stream (context 0: the process-wide libc rand())"""
    return load_library().vb_context_srand(int(seed))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(FP)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def set_bootstrap_override(R=None, t=None, depth=None):
    """Inject (or clear, with no arguments) the monocular bootstrap pose/depth used by windows without priors."""
    lib = load_library()
    if depth is None:
        return lib.vb_set_bootstrap_override(0, None, None, None, 0, 0)
    R = _f32(np.asarray(R).reshape(9))
    t = _f32(np.asarray(t).reshape(3))
    depth = _f32(depth)
    h, w = depth.shape
    return lib.vb_set_bootstrap_override(1, _ptr(R), _ptr(t), _ptr(depth), w, h)


def voldor_ex(flows, fx, fy, cx, cy, basefocal=0, disparity=None, disparity_pconf=None, depth_priors=None,
              depth_prior_poses=None, depth_prior_pconfs=None, config=""):
    """voldor() plus `iters` (EM iterations executed) and `stats_ms` = [total, cameras, depth, io]"""
    lib = load_library()
    flows = _f32(flows)
    N, h, w = flows.shape[0], flows.shape[1], flows.shape[2]
    disparity, disparity_pconf = _f32(disparity), _f32(disparity_pconf)
    depth_priors, depth_prior_poses = _f32(depth_priors), _f32(depth_prior_poses)
    depth_prior_pconfs = _f32(depth_prior_pconfs)
    N_dp = 0 if depth_priors is None else depth_priors.shape[0]

    poses = np.zeros((N, 6), np.float32)
    poses_covar = np.zeros((N, 6, 6), np.float32)
    depth = np.zeros((h, w), np.float32)
    depth_conf = np.zeros((h, w), np.float32)
    n_registered = C.c_int(0)
    iters = C.c_int(0)
    stats = np.zeros(4, np.float32)
    lib.vb_py_voldor_wrapper_ex(_ptr(flows), _ptr(disparity), _ptr(disparity_pconf), _ptr(depth_priors),
                                _ptr(depth_prior_poses), _ptr(depth_prior_pconfs), fx, fy, cx, cy, basefocal, N,
                                N_dp, w, h, config.encode("ascii"), C.byref(n_registered), _ptr(poses),
                                _ptr(poses_covar), _ptr(depth), _ptr(depth_conf), C.byref(iters), _ptr(stats))
    n = n_registered.value
    return {"n_registered": n, "poses": poses[:n], "poses_covar": poses_covar[:n], "depth": depth,
            "depth_conf": depth_conf, "iters": iters.value, "stats_ms": stats}


def voldor(flows, fx, fy, cx, cy, basefocal=0, disparity=None, disparity_pconf=None, depth_priors=None,
           depth_prior_poses=None, depth_prior_pconfs=None, config=""):
    r = voldor_ex(flows, fx, fy, cx, cy, basefocal, disparity, disparity_pconf, depth_priors, depth_prior_poses,
                  depth_prior_pconfs, config)
    return {k: r[k] for k in ("n_registered", "poses", "poses_covar", "depth", "depth_conf")}
