"""Builds the native parts IN-TREE (the .so files travel to the GPU box with the
repo snapshot):

  lib/libgapcu.so                     CUDA kernels + host orchestration + C ABI + Fortran-ABI shim
  libgap/libgap.<abi>.so              the f2py extension module "libgap.libgap" (same Python
                                      surface as the reference's, gappy/setup.py:36-45)

nvcc cross-compiles for sm_100a without a GPU.  Usage: python build.py [--force]
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

CU_SOURCES = ["neigh.cu", "centre.cu", "centre_p128.cu", "centre_p256.cu", "centre_p512.cu", "centre_p1024.cu", "gpr.cu", "gather.cu", "halo.cu", "variance.cu",
              "microbench.cu", "context.cu"]
CPP_SOURCES = ["potential.cpp"]
C_SOURCES = ["fortran_shim.c"]
HEADERS = ["domain_host.inc", "device_types.cuh", "centre_impl.cuh", "fastmath.cuh", "geom.cuh", "launch.cuh", "potential.hpp", os.path.join("..", "..", "include", "gapcu.h")]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, **kw):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, **kw)


def build_libgapcu(force=False, verbose_ptxas=False):
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    cmds = []
    for src in CU_SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, src + ".o")
        if force or _newer(o, [s] + hdrs):
            cmd = [NVCC, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off",
                   "-c", s, "-o", o]
            if verbose_ptxas:
                cmd.insert(1, "-Xptxas=-v")
            cmds.append(cmd)
        objs.append(o)
    if cmds:   # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 1)) as ex:
            list(ex.map(_run, cmds))
    for src in CPP_SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, src + ".o")
        if force or _newer(o, [s] + hdrs):
            _run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wall", "-c", s, "-o", o])
        objs.append(o)
    for src in C_SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, src + ".o")
        if force or _newer(o, [s] + hdrs):
            _run(["gcc", "-O2", "-fPIC", "-Wall", "-c", s, "-o", o])
        objs.append(o)
    out = os.path.join(LIB, "libgapcu.so")
    if force or _newer(out, objs):
        _run([NVCC, *ARCH, "-shared", "-o", out, *objs, "-cudart", "static"])
    return out


def build_f2py_module(force=False):
    """numpy.f2py turns f2py/libgap.pyf into libgapmodule.c; gcc links it against
    libgapcu.so, which exports fgap_calc_ / fgap_read_ / fget_bond_ / car2acsf_ /
    write_array_2dim_ (csrc/fortran_shim.c)."""
    import numpy
    import numpy.f2py
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(HERE, "libgap", "libgap" + ext)
    pyf = os.path.join(HERE, "f2py", "libgap.pyf")
    gen = os.path.join(OBJ, "f2py")
    os.makedirs(gen, exist_ok=True)
    lib = os.path.join(LIB, "libgapcu.so")
    if not (force or _newer(out, [pyf, lib])):
        return out
    _run([sys.executable, "-m", "numpy.f2py", pyf, "--build-dir", gen], cwd=gen, stdout=subprocess.DEVNULL)
    modc = os.path.join(gen, "libgapmodule.c")
    if not os.path.exists(modc):  # older/newer f2py drop it in the cwd
        modc = os.path.join(gen, "libgapmodule.c")
    f2py_src = os.path.join(os.path.dirname(numpy.f2py.__file__), "src")
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(), "-I" + f2py_src]
    _run(["gcc", "-O2", "-fPIC", "-shared", "-Wno-unused-function", *inc, modc, os.path.join(f2py_src, "fortranobject.c"),
          "-L" + LIB, "-lgapcu", "-Wl,-rpath,$ORIGIN/../lib", "-o", out])
    return out


def build_all(force=False, verbose_ptxas=False):
    a = build_libgapcu(force, verbose_ptxas)
    b = build_f2py_module(force)
    return a, b


if __name__ == "__main__":
    print(build_all("--force" in sys.argv, "-v" in sys.argv))
