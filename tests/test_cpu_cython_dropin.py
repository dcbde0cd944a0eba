"""Drop-in at the Python-binding level: the reference's UNMODIFIED slam_py/install/pyvoldor_vo.pyx
cythonizes, compiles against include/py_export.h and links against libvoldor_b200.so.  Needs the reference checkout
and Cython; skipped where they are absent (the GPU box only uses the module prebuilt by __graft_entry__.build())."""
import os
import subprocess
import sys

import numpy as np
import pytest

import ffi

REF_PYX = os.path.join(os.environ.get("VOLDOR_REFERENCE", "/root/reference"), "slam_py", "install", "pyvoldor_vo.pyx")


@pytest.fixture(scope="module")
def module_path():
    pytest.importorskip("Cython")
    if not os.path.exists(REF_PYX):
        pytest.skip("reference checkout not present")
    sys.path.insert(0, os.path.join(ffi.ROOT, "integration"))
    import build_cython_binding

    return build_cython_binding.build()


def test_unmodified_reference_pyx_builds_against_this_library(module_path):
    assert os.path.exists(module_path)
    deps = subprocess.run(["ldd", module_path], capture_output=True, text=True).stdout
    assert "libvoldor_b200.so" in deps and "not found" not in deps.split("libvoldor_b200.so")[1].split("\n")[0]
    assert "opencv" not in deps.lower() and "ceres" not in deps.lower()
    # the generated C++ is not kept, and nothing of the reference is copied into the tree
    assert not os.path.exists(os.path.join(os.path.dirname(module_path), "pyvoldor_vo.cpp"))


def test_module_imports_under_the_name_the_slam_layer_uses(module_path):
    # separate interpreter: importing must not need a GPU, and must expose voldor() with the reference's keywords
    code = (
        "import sys, inspect; sys.path.insert(0, %r); import pyvoldor_vo; import numpy as np\n"
        "assert callable(pyvoldor_vo.voldor)\n"
        "try:\n"
        "    pyvoldor_vo.voldor(np.zeros((2, 4, 4, 2), np.float64), 1.0, 1.0, 0.0, 0.0)\n"
        "except ValueError as e:\n"
        "    assert 'dtype' in str(e).lower(), e\n"
        "else:\n"
        "    raise SystemExit('float64 flows must be rejected by the typed signature')\n"
        "try:\n"
        "    pyvoldor_vo.voldor(np.zeros((2, 4, 4, 2), np.float32), 1.0, 1.0, 0.0, 0.0, bogus=1)\n"
        "except TypeError:\n"
        "    pass\n"
        "else:\n"
        "    raise SystemExit('unknown keyword must be rejected')\n"
        "print('ok')\n" % os.path.dirname(module_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr + r.stdout
