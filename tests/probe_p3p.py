"""diagnostic (not a test): per-stage bit mismatch counts between the reference's lambdatwist device functions
and ours, evaluated inside one kernel on identical inputs (oracle/ref_shim/p3p_probe.cu)"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ffi
import synth
from test_gpu_abi_parity import _instances

ref = ffi.reference()
win, p2s, p3s = _instances(ref, 160, 120, 3, 7, 1)
lib = C.CDLL(os.path.join(ffi.ROOT, "oracle", "_ref", "libp3p_probe.so"))
mism = (C.c_ulonglong * 8)()
dump = np.zeros((64, 6), np.float32)
K = np.ascontiguousarray(win["K"], np.float32)
n_poses = 4096
cr = np.zeros((n_poses, 6), np.float32)
cm = np.zeros((n_poses, 6), np.float32)
rc = lib.p3p_probe(p3s.ctypes.data_as(ffi.FP), p2s.ctypes.data_as(ffi.FP), K.ctypes.data_as(ffi.FP), p2s.shape[0],
                   n_poses, mism, dump.ctypes.data_as(ffi.FP), cr.ctypes.data_as(ffi.FP), cm.ctypes.data_as(ffi.FP))
mine = ffi.ours()
_, rv_r, tv_r = ref.solve_p3p(p3s, p2s, win["K"], n_poses)
_, rv_m, tv_m = mine.solve_p3p(p3s, p2s, win["K"], n_poses)
lib_r = np.concatenate([rv_r, tv_r], 1)
lib_m = np.concatenate([rv_m, tv_m], 1)


def mm(a, b):
    return int((a.view(np.uint32) != b.view(np.uint32)).any(1).sum())


print("rows differing: clean_ref vs lib_ref", mm(cr, lib_r), "| clean_mine vs lib_mine", mm(cm, lib_m),
      "| clean_ref vs clean_mine", mm(cr, cm), "| lib_ref vs lib_mine", mm(lib_r, lib_m),
      "| clean_mine vs lib_ref", mm(cm, lib_r), "| clean_ref vs lib_mine", mm(cr, lib_m))
names = ["cubic", "eig", "refine", "p3p_valid", "p3p_RT", "p4p_Rt", "rodrigues"]
print("rc", rc, {n: int(mism[i]) for i, n in enumerate(names)}, "of", n_poses)
bad = np.where((dump[:, :3].view(np.uint32) != dump[:, 3:].view(np.uint32)).any(1))[0]
print("rodrigues sample mismatches", bad[:8], dump[bad[:4]])
