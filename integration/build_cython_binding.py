"""Build the reference's UNMODIFIED Cython module (slam_py/install/pyvoldor_vo.pyx) against libvoldor_b200.so.

This is the drop-in at the Python-binding level: the reference's own `.pyx` declares
`py_voldor_wrapper` via `cdef extern from "../../voldor/py_export.h"`; here that include resolves to this project's
`include/py_export.h` and the extension links against `voldor_b200/libvoldor_b200.so` instead of the reference's
voldor/*.cpp + libgpu-kernels + OpenCV (slam_py/install/setup_linux_vo.py).

The `.pyx` is read where it lies (reference checkout, default /root/reference, override with VOLDOR_REFERENCE); nothing
of it is copied into this repository: the generated C++ is written without source comments into the git-ignored
`integration/_build/`.  The built module `integration/_build/pyvoldor_vo*.so` is importable as `pyvoldor_vo`, the name
`slam_py/voldor_slam.py:6-16` imports.

usage: python integration/build_cython_binding.py        (prints the path of the built module)
"""
import os
import subprocess
import sys
import sysconfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "integration", "_build")


def build(reference=None, quiet=True):
    reference = reference or os.environ.get("VOLDOR_REFERENCE", "/root/reference")
    pyx = os.path.join(reference, "slam_py", "install", "pyvoldor_vo.pyx")
    lib = os.path.join(ROOT, "voldor_b200", "libvoldor_b200.so")
    if not os.path.exists(pyx):
        raise FileNotFoundError(f"reference binding not found: {pyx}")
    if not os.path.exists(lib):
        raise FileNotFoundError(f"{lib} is missing: run `make lib` first")
    import numpy
    from Cython.Compiler.Main import CompilationOptions, compile as cython_compile

    # the .pyx includes "../../voldor/py_export.h": give it a directory two levels below a `voldor/` that forwards to
    # this project's header
    shim = os.path.join(BUILD, "shim", "a", "b")
    os.makedirs(shim, exist_ok=True)
    os.makedirs(os.path.join(BUILD, "shim", "voldor"), exist_ok=True)
    with open(os.path.join(BUILD, "shim", "voldor", "py_export.h"), "w") as f:
        f.write('#include "%s"  // the boundary header of voldor_b200\n' % os.path.join(ROOT, "include", "py_export.h"))
    cpp = os.path.join(BUILD, "pyvoldor_vo.cpp")
    opts = CompilationOptions(cplus=True, language_level=3, output_file=cpp,
                              compiler_directives={"emit_code_comments": False})
    res = cython_compile(pyx, opts)
    if res.num_errors:
        raise RuntimeError("cythonize failed")
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(BUILD, "pyvoldor_vo" + ext)
    cmd = ["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-w", cpp, "-o", out,
           "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(),
           "-I" + os.path.join(ROOT, "include"), "-I" + shim,
           "-L" + os.path.dirname(lib), "-lvoldor_b200", "-Wl,-rpath,$ORIGIN/../../voldor_b200"]
    subprocess.run(cmd, check=True, capture_output=quiet)
    os.remove(cpp)  # only the extension module is kept
    return out


if __name__ == "__main__":
    print(build(quiet=False))
