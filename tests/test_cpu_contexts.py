"""Execution contexts (csrc/context.h), host logic only — no GPU needed: thread-local selection, and the private
start-sample generator of contexts >= 1 reproduces glibc's rand() after srand(seed) draw for draw (a context behaves
like a reference worker process of its own, slam_py/voldor_slam.py:182-187)."""
import ctypes as C
import os
import threading

import pytest

import ffi

pytestmark = pytest.mark.skipif(not os.path.exists(ffi.OURS), reason="libvoldor_b200.so not built (make lib)")


def _lib():
    lib = C.CDLL(ffi.OURS)
    lib.vb_context_srand.argtypes = [C.c_uint]
    return lib


def test_private_stream_equals_glibc_rand():
    lib, libc = _lib(), C.CDLL(None)
    assert lib.vb_context_select(3) == 0
    try:
        # an unseeded process starts from srand(1)
        libc.srand(1)
        assert [lib.vb_context_rand() for _ in range(5)] == [libc.rand() for _ in range(5)]
        for seed in (0, 1, 2, 77, 1000, 2**31 - 1, 2**32 - 5):
            libc.srand(C.c_uint(seed))
            lib.vb_context_srand(seed)
            want = [libc.rand() for _ in range(700)]  # > 2 * 344: wraps the 31-word state many times
            got = [lib.vb_context_rand() for _ in range(700)]
            assert got == want, seed
        # speculation: snapshot, draw 20, rewind, keep 7 -> exactly 7 draws consumed
        libc.srand(5)
        lib.vb_context_srand(5)
        assert lib.vb_debug_rand_speculate(20, 7) == 0
        for _ in range(7):
            libc.rand()
        assert lib.vb_context_rand() == libc.rand()
    finally:
        assert lib.vb_context_select(0) == 3


def test_context_zero_is_the_process_libc_stream():
    lib, libc = _lib(), C.CDLL(None)
    assert lib.vb_context_current() == 0
    lib.vb_context_srand(123)      # == srand(123)
    a = lib.vb_context_rand()      # == rand()
    libc.srand(123)
    assert libc.rand() == a
    libc.srand(9)
    assert lib.vb_debug_rand_speculate(20, 3) == 0
    libc2 = [libc.rand() for _ in range(2)]
    libc.srand(9)
    assert [libc.rand() for _ in range(5)][3:] == libc2


def test_selection_is_per_host_thread_and_bounded():
    lib = _lib()
    assert lib.vb_context_select(-1) == -1 and lib.vb_context_select(lib.vb_context_max()) == -1
    seen = {}

    def worker(k):
        lib.vb_context_select(k)
        seen[k] = lib.vb_context_current()

    ts = [threading.Thread(target=worker, args=(k,)) for k in (1, 2, 5)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert seen == {1: 1, 2: 2, 5: 5}
    assert lib.vb_context_current() == 0  # the main thread's selection is untouched
