"""The lean power function of the Fisk pdf (csrc/residual_model.cuh) is the CUDA math library's powf bit for bit:
every normal positive float as base x every exponent class the residual model produces (-2 and the shape range
-1-c, -c with c = 1 - 0.0022 m, m in [2, 100]), plus the composed pdf on 2^30 random (residual, shape) pairs drawn
over ALL float bit patterns (NaN, inf, denormals, negatives take the library fallback)."""
import ctypes as C
import os

import numpy as np
import pytest

import ffi

pytestmark = pytest.mark.gpu
PROBE = os.path.join(ffi.ROOT, "tests", "_build", "libpow_probe.so")


@pytest.mark.skipif(not os.path.exists(PROBE), reason="tests/_build/libpow_probe.so not built (make probes)")
def test_lean_pow_equals_library_powf_for_every_normal_base():
    lib = C.CDLL(PROBE)
    m = np.array([2.0, 2.5, 3.7, 10.0, 33.3, 64.0, 99.0, 100.0], np.float32)
    c = (m * np.float32(-0.0022) + np.float32(1.0)).astype(np.float32)
    ys = np.concatenate([[np.float32(-2.0)], (np.float32(-1.0) - c), -c]).astype(np.float32)
    mism = np.zeros(len(ys), np.uint64)
    first = np.zeros(len(ys), np.uint32)
    pdf = np.zeros(1, np.uint64)
    rc = lib.pow_probe_run(ys.ctypes.data_as(C.POINTER(C.c_float)), len(ys), mism.ctypes.data_as(C.POINTER(C.c_ulonglong)),
                           first.ctypes.data_as(C.POINTER(C.c_uint)), C.c_ulonglong(1 << 30),
                           pdf.ctypes.data_as(C.POINTER(C.c_ulonglong)))
    assert rc == 0
    os.makedirs(os.path.join(ffi.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ffi.ROOT, "gpurun_out", "pow_exact.txt"), "w") as f:
        f.write(f"bases tested per exponent: {0x7f000000}\nexponents: {ys.tolist()}\nmismatches: {mism.tolist()}\n"
                f"first mismatching base bits: {[hex(int(v)) for v in first]}\npdf samples {1 << 30} mismatches {int(pdf[0])}\n")
    assert mism.sum() == 0, (ys[mism > 0], mism[mism > 0], [hex(int(v)) for v in first[mism > 0]])
    assert pdf[0] == 0
