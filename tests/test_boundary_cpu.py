"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/gapcu.h declares, the gap_parameters reader agrees with the
oracle's, the f2py module has the reference's signatures, and without a GPU every
compute entry point fails loudly (no CPU fallback).  No GPU compute here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def lib():
    import gapcu
    return gapcu.lib()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "gapcu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(gapcu_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    for n in sorted(names):
        assert hasattr(lib, n), n
    import gapcu
    assert names == set(gapcu.SYMBOLS)
    for n in gapcu.FORTRAN_SYMBOLS:      # what the f2py module links against
        assert hasattr(lib, n), n


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "calypso-gap_b200")
    for dp, _, files in os.walk(pkg):
        if os.sep + "build" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".c", ".h", ".f90")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "gapo_" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


def test_gapcu_read_matches_oracle_reader(lib, shipped_pot):
    lib.gapcu_read.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                               C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    nsp, dl = C.c_int(), C.c_int()
    theta = np.full(100, -7.0); mm = np.full((4000, 100), -7.0, order="F"); coeff = np.full(4000, -7.0)
    inv = np.ones((50, 50), order="F")
    rc = lib.gapcu_read(os.path.join(GOLDEN, "gap_parameters").encode(), C.byref(nsp), C.byref(dl), theta.ctypes.data, 100,
                        mm.ctypes.data, 4000, 100, inv.ctypes.data, 50, coeff.ctypes.data, 4000)
    assert rc == 0 and (nsp.value, dl.value) == (129, 66)
    assert np.array_equal(theta[:66], shipped_pot.theta) and np.array_equal(mm[:129, :66], shipped_pot.mm)
    assert np.array_equal(coeff[:129], shipped_pot.coeff)
    assert (theta[66:] == -7.0).all() and (mm[129:] == -7.0).all() and not inv.any()
    # missing file: the reference's message (gap_calc.f90:325)
    rc = lib.gapcu_read(b"/nonexistent/gap_parameters", C.byref(nsp), C.byref(dl), theta.ctypes.data, 100, mm.ctypes.data,
                        4000, 100, None, 0, coeff.ctypes.data, 4000)
    assert rc == -1 and b"gap_parameters file does not exist!" in lib.gapcu_last_error()
    # capacity checks of FGAP_READ (gap_calc.f90:341-350)
    rc = lib.gapcu_read(os.path.join(GOLDEN, "gap_parameters").encode(), C.byref(nsp), C.byref(dl), theta.ctypes.data, 100,
                        mm.ctypes.data, 100, 100, None, 0, coeff.ctypes.data, 100)
    assert rc == -4 and b"nsparseX_max" in lib.gapcu_last_error()


def test_gapcu_read_zero_fill_of_invcmm(lib):
    """FGAP_READ sets INVCMM = 0 (gap_calc.f90:361).  The library drops the whole pages of a private
    anonymous allocation with madvise(MADV_DONTNEED) (zero-fill on demand) and memsets everything
    else.  Whatever the history of the buffer, the caller must see zeros everywhere, and nothing
    outside the buffer may change."""
    lib.gapcu_read.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                               C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    nsp, dl = C.c_int(), C.c_int()
    theta = np.zeros(100); mm = np.zeros((4000, 100), order="F"); coeff = np.zeros(4000)
    path = os.path.join(GOLDEN, "gap_parameters").encode()

    def read_into(buf, ld):
        assert lib.gapcu_read(path, C.byref(nsp), C.byref(dl), theta.ctypes.data, 100, mm.ctypes.data, 4000, 100,
                              buf.ctypes.data, ld, coeff.ctypes.data, 4000) == 0

    fresh = np.empty((4000, 4000), order="F")                 # 128 MB, pages never touched
    read_into(fresh, 4000)
    assert not fresh.any()
    dirty = np.full((4000, 4000), 3.5, order="F")
    read_into(dirty, 4000)
    assert not dirty.any()
    partly = np.empty((4000, 4000), order="F")
    partly[17, 5] = 3.0; partly[3999, 3999] = -1.0; partly[0, 0] = 2.0; partly[:, 2000:2003] = 7.0
    read_into(partly, 4000)
    assert not partly.any()
    # a window that starts and ends in the middle of pages of a dirty buffer: guard bytes stay
    big = np.full(2000 * 2000 + 2000, 9.0)
    win = big[777:777 + 2000 * 2000]
    read_into(win, 2000)
    assert not win.any() and (big[:777] == 9.0).all() and (big[777 + 2000 * 2000:] == 9.0).all()
    small = np.full((50, 50), 1.0, order="F")                  # below the lazy threshold: plain memset
    read_into(small, 50)
    assert not small.any()


def test_reader_round_trip_of_written_potential(lib, oracle, tmp_path):
    from structures import write_gap_parameters
    rng = np.random.default_rng(3)
    z = np.array([1, 8, 14], np.int32); w = np.array([0.5, -2.0, 3.25])
    ntype = np.array([1, 3, 2, 4, 2], np.int32); alpha = np.array([0.3, 1.5, 0.02, 0.02, 0.4]); cut = np.array([5.5, 5.5, 4.0, 4.0, 3.0])
    theta = rng.uniform(0.5, 2, 10); mm = rng.normal(size=(7, 10)); coeff = rng.normal(size=7) * 100
    p = str(tmp_path / "gap_parameters")
    write_gap_parameters(p, z, w, ntype, alpha, cut, theta, mm, coeff)
    o = oracle.read(p)
    lib.gapcu_read.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                               C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    nsp, dl = C.c_int(), C.c_int()
    t2 = np.zeros(10); m2 = np.zeros((7, 10), order="F"); c2 = np.zeros(7)
    assert lib.gapcu_read(p.encode(), C.byref(nsp), C.byref(dl), t2.ctypes.data, 10, m2.ctypes.data, 7, 10, None, 0,
                          c2.ctypes.data, 7) == 0
    assert np.array_equal(t2, o.theta) and np.array_equal(m2, o.mm) and np.array_equal(c2, o.coeff)
    assert np.allclose(o.theta, theta, atol=1e-10) and list(o.ntype) == list(ntype)


def test_f2py_module_signatures_and_fgap_read(monkeypatch):
    import libgap.libgap as m
    for name in ("fgap_calc", "fgap_read", "fget_bond", "car2acsf", "write_array_2dim"):
        assert hasattr(m, name)
    assert "ene,force,stress,variance = fgap_calc(species,lat,pos,theta,mm,qmm,coeff,rcut,lgrad,[na,nsparsex,des_len])" in m.fgap_calc.__doc__
    assert "nsparsex,des_len,theta,mm,invcmm,coeff = fgap_read()" in m.fgap_read.__doc__
    assert "min_bond = fget_bond(lat,elements,pos,rcut,[na])" in m.fget_bond.__doc__
    assert "xx,dxdy,strs = car2acsf(nf,pos,neighbor,neighbor_count,lgrad,[na,max_neighbor])" in m.car2acsf.__doc__
    monkeypatch.chdir(GOLDEN)
    nsp, dl, theta, mm, inv, coeff = m.fgap_read()
    assert (nsp, dl) == (129, 66) and theta.shape == (100,) and mm.shape == (4000, 100) and inv.shape == (4000, 4000)
    assert not inv.any() and theta[0] == 2.1695365180


def test_write_array_2dim(tmp_path, monkeypatch):
    import libgap.libgap as m
    monkeypatch.chdir(tmp_path)
    a = np.arange(6, dtype=float).reshape(2, 3) / 7
    m.write_array_2dim(a, "kk.dat")
    got = np.loadtxt(tmp_path / "kk.dat")
    assert np.allclose(got, a, atol=1e-10)
    assert open(tmp_path / "kk.dat").read().split("\n")[0] == "".join("%20.10f" % x for x in a[0])


def test_python_classes_symbol_lookup():
    from libgap._elements import ATOMIC_NUMBER, atomic_numbers
    assert atomic_numbers(["B", "C", "X", "Lr"]) == [5, 6, 0, 103] and len(ATOMIC_NUMBER) == 104
    with pytest.raises(KeyError):
        atomic_numbers(["Qq"])


def test_no_gpu_means_loud_failure():
    """In a process without a visible CUDA device every compute entry point errors."""
    code = ("import sys, os; os.environ['CUDA_VISIBLE_DEVICES']=''; sys.path.insert(0, %r); import gapcu\n"
            "try:\n    gapcu.Context(0); print('CREATED')\nexcept gapcu.GapcuError as e:\n    print('ERR', e.code)\n") % os.path.join(ROOT, "calypso-gap_b200")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120).stdout
    assert "ERR -6" in out and "CREATED" not in out
    # the drop-in batch entry points as well (gapcu_set_devices, gapcu_calc_batch)
    code = ("import sys, os; os.environ['CUDA_VISIBLE_DEVICES']=''; sys.path.insert(0, %r); import gapcu, numpy as np\n"
            "os.chdir(%r)\n"
            "for f in (lambda: gapcu.set_devices([0]), lambda: gapcu.calc_batch([np.array([6, 6], np.int32)], [np.eye(3) * 5], [np.zeros((2, 3))])):\n"
            "    try:\n        f(); print('RAN')\n    except gapcu.GapcuError as e:\n        print('ERR', e.code)\n") % (
                os.path.join(ROOT, "calypso-gap_b200"), GOLDEN)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120).stdout
    assert out.count("ERR -6") == 2 and "RAN" not in out


def test_fastmath_host_versions(tmp_path):
    """csrc/fastmath.cuh compiles for the host too: exp_neg / sincos_0pi / sincos_tab vs libm."""
    src = tmp_path / "fm.cpp"
    src.write_text('#include <cstdio>\n#include <cmath>\n#include "%s"\n'
                   'extern "C" double fm_exp(double x){ static double T[32]; static int i=0; if(!i){gapcu::fill_exp2_table(T);i=1;} return gapcu::exp_neg(x,T);}\n'
                   'extern "C" void fm_sc(double y,double*s,double*c){ gapcu::sincos_0pi(y,s,c);}\n'
                   'extern "C" void fm_sct(double y,double*s,double*c){ static gapcu::SinCosEntry T[gapcu::SINCOS_TAB_N]; static int i=0; if(!i){gapcu::fill_sincos_table(T);i=1;} gapcu::sincos_tab(y,T,s,c);}\n'
                   % os.path.join(ROOT, "calypso-gap_b200", "csrc", "fastmath.cuh"))
    so = tmp_path / "fm.so"
    subprocess.check_call(["g++", "-O2", "-mfma", "-fPIC", "-shared", "-o", str(so), str(src)])
    L = C.CDLL(str(so))
    L.fm_exp.restype = C.c_double; L.fm_exp.argtypes = [C.c_double]
    L.fm_sc.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.fm_sct.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    rng = np.random.default_rng(0)
    xs = -np.concatenate([rng.uniform(0, 700, 20000), rng.uniform(0, 5, 20000), [0.0, 1e-300, 700.0]])
    got = np.array([L.fm_exp(x) for x in xs])
    assert (np.abs(got - np.exp(xs)) <= 4.5e-16 * np.exp(xs)).all()
    # below the normal range the result is a harmless tiny number (<= 2^-1021), never NaN / inf / garbage
    for x in (-709.0, -800.0, -1e4, -1e7, -1e12, -1e300, -np.inf):
        v = L.fm_exp(x)
        assert 0.0 <= v <= 2.0 ** -1020
    s, c = C.c_double(), C.c_double()
    for y in np.concatenate([rng.uniform(0, 3.3, 20000), [0.0, np.pi / 2, 3.141592654]]):
        L.fm_sc(y, C.byref(s), C.byref(c))
        assert abs(s.value - np.sin(y)) < 4e-16 and abs(c.value - np.cos(y)) < 4e-16   # absolute: what fc / fc' need
        L.fm_sct(y, C.byref(s), C.byref(c))                                             # the table version of the hot loops
        assert abs(s.value - np.sin(y)) < 3e-16 and abs(c.value - np.cos(y)) < 3e-16
