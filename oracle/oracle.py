"""ctypes front end of the CPU oracle (oracle/gap_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of gap_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference); never by
the product package.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gap_oracle.c")
OUT = os.path.join(HERE, "_build")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def _cpu_tag():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return hashlib.sha1(line.encode()).hexdigest()[:10]
    except OSError:
        pass
    return "generic"


def build(kind="parity", force=False):
    """Compile the oracle.  kind='parity': -O2 -ffp-contract=off (arbiter);
    kind='fast': -O3 -march=native, keyed by the host CPU's flag set because the
    GPU box's CPU need not be this container's."""
    os.makedirs(OUT, exist_ok=True)
    if kind == "parity":
        path = os.path.join(OUT, "liboracle_parity.so")
        flags = ["-O2", "-ffp-contract=off"]
    else:
        path = os.path.join(OUT, "liboracle_fast_%s.so" % _cpu_tag())
        flags = ["-O3", "-march=native"]
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(SRC):
        tmp = path + ".tmp%d" % os.getpid()
        subprocess.check_call(["gcc", "-std=c99", *flags, "-fPIC", "-shared", "-o", tmp, SRC, "-lm"])
        os.replace(tmp, path)
    return path


class Oracle:
    def __init__(self, kind="parity"):
        self.kind = kind
        self.lib = C.CDLL(build(kind))
        L = self.lib
        L.gapo_read.restype = C.c_void_p
        L.gapo_read.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.gapo_make.restype = C.c_void_p
        L.gapo_make.argtypes = [C.c_int, _ip, _dp, C.c_int, _ip, _dp, _dp, C.c_int, C.c_int, _dp, _dp, _dp]
        L.gapo_free.argtypes = [C.c_void_p]
        L.gapo_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.gapo_get.argtypes = [C.c_void_p, _ip, _dp, _ip, _dp, _dp, _dp, _dp, _dp]
        L.gapo_image_range.argtypes = [_dp, C.c_double, _ip]
        for name in ("gapo_neighbors", "gapo_neighbors_sparse"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_int, _dp, _dp, C.c_double, C.c_int, _ip, _ip, _ip, _dp]
        L.gapo_triplets.restype = C.c_int
        L.gapo_triplets.argtypes = [C.c_int, _dp, _dp, C.c_double, C.c_int, C.c_double, C.c_int, _ip]
        L.gapo_get_bond.restype = C.c_double
        L.gapo_get_bond.argtypes = [C.c_int, _dp, _dp, C.c_double]
        L.gapo_calc_dense.restype = C.c_int
        L.gapo_calc_dense.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_double), _dp, _dp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gapo_calc_sparse.restype = C.c_int
        L.gapo_calc_sparse.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _dp, C.c_double, C.c_int, C.c_int,
                                       C.POINTER(C.c_double), _dp, _dp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gapo_calc_sparse_centres.restype = C.c_int
        L.gapo_calc_sparse_centres.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.POINTER(C.c_double), _dp, _dp]
        L.gapo_car2acsf_dense.restype = C.c_int
        L.gapo_car2acsf_dense.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _dp, C.c_double, C.c_int, _dp, _dp, _dp]

    # ---- potentials -------------------------------------------------------
    def read(self, path):
        err = C.create_string_buffer(256)
        h = self.lib.gapo_read(os.fsencode(path), err, 256)
        if not h:
            raise RuntimeError(err.value.decode())
        return Potential(self, h)

    def make(self, z, w, ntype, alpha, cutoff, theta, mm, coeff):
        z = np.ascontiguousarray(z, np.int32); w = np.ascontiguousarray(w, np.float64)
        ntype = np.ascontiguousarray(ntype, np.int32)
        alpha = np.ascontiguousarray(alpha, np.float64); cutoff = np.ascontiguousarray(cutoff, np.float64)
        theta = np.ascontiguousarray(theta, np.float64); mm = np.ascontiguousarray(mm, np.float64)
        coeff = np.ascontiguousarray(coeff, np.float64)
        h = self.lib.gapo_make(len(z), z, w, len(ntype), ntype, alpha, cutoff, mm.shape[0], mm.shape[1], theta, mm, coeff)
        return Potential(self, h)

    # ---- geometry ---------------------------------------------------------
    def image_range(self, lat, rcut):
        out = np.zeros(3, np.int32)
        self.lib.gapo_image_range(np.ascontiguousarray(lat, np.float64), float(rcut), out)
        return out

    def neighbors(self, lat, pos, rcut, cap=1000, sparse=False):
        """Per-atom neighbour lists in reference order.  Returns count[na],
        idx[na,cap], shift[na,cap,3], dis[na,cap]."""
        lat = np.ascontiguousarray(lat, np.float64); pos = np.ascontiguousarray(pos, np.float64)
        na = pos.shape[0]
        count = np.zeros(na, np.int32); idx = np.zeros((na, cap), np.int32)
        shift = np.zeros((na, cap, 3), np.int32); dis = np.zeros((na, cap))
        f = self.lib.gapo_neighbors_sparse if sparse else self.lib.gapo_neighbors
        mx = f(na, lat, pos, float(rcut), cap, count, idx, shift, dis)
        if mx < 0:
            raise RuntimeError("neighbour overflow (> %d)" % cap)
        return count, idx, shift, dis

    def triplets(self, lat, pos, rcut, centre, cutoff, cap=8192):
        """(slot_j, slot_k) pairs of `centre` that the angular loops keep for a function with this cutoff
        (wacsf.f90:177-244), slots = positions in the reference-order neighbour list.  Returns [n, 2]."""
        lat = np.ascontiguousarray(lat, np.float64); pos = np.ascontiguousarray(pos, np.float64)
        pairs = np.zeros((cap, 2), np.int32)
        n = self.lib.gapo_triplets(pos.shape[0], lat, pos, float(rcut), int(centre), float(cutoff), cap, pairs)
        if n < 0 or n > cap:
            raise RuntimeError("triplet export overflow (%d)" % n)
        return pairs[:n].copy()

    def get_bond(self, lat, pos, rcut):
        lat = np.ascontiguousarray(lat, np.float64); pos = np.ascontiguousarray(pos, np.float64)
        return self.lib.gapo_get_bond(pos.shape[0], lat, pos, float(rcut))


_ERR = {-1: "Atoms neighbor large than max_neighbor", -2: "species not in gap_parameters",
        -3: "des_len != 2*nsf", -4: "out of memory"}


class Potential:
    def __init__(self, oracle, handle):
        self.o = oracle
        self.h = C.c_void_p(handle)
        n = [C.c_int() for _ in range(4)]
        oracle.lib.gapo_info(self.h, *[C.byref(x) for x in n])
        self.nspecies, self.nsf, self.nsparse, self.des_len = [x.value for x in n]
        self.z = np.zeros(self.nspecies, np.int32); self.w = np.zeros(self.nspecies)
        self.ntype = np.zeros(self.nsf, np.int32); self.alpha = np.zeros(self.nsf); self.cutoff = np.zeros(self.nsf)
        self.theta = np.zeros(self.des_len); self.mm = np.zeros((self.nsparse, self.des_len)); self.coeff = np.zeros(self.nsparse)
        oracle.lib.gapo_get(self.h, self.z, self.w, self.ntype, self.alpha, self.cutoff, self.theta, self.mm, self.coeff)

    def __del__(self):
        try:
            self.o.lib.gapo_free(self.h)
        except Exception:
            pass

    def _calc(self, dense, species, lat, pos, rcut, lgrad, extra, want_desc, want_stats):
        species = np.ascontiguousarray(species, np.int32)
        lat = np.ascontiguousarray(lat, np.float64); pos = np.ascontiguousarray(pos, np.float64)
        na = pos.shape[0]
        ene = C.c_double(); force = np.zeros((na, 3)); stress = np.zeros(6)
        if dense:
            c0, c1 = extra
            nc = (na if c1 <= 0 or c1 > na else c1) - max(c0, 0)
        else:
            nc = na
        xx = np.zeros((nc, self.des_len)) if want_desc else None
        dedg = np.zeros((nc, self.des_len)) if want_desc else None
        eat = np.zeros(nc) if want_desc else None
        stats = np.zeros(8) if want_stats else None
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        if dense:
            rc = self.o.lib.gapo_calc_dense(self.h, na, species, lat, pos, float(rcut), int(bool(lgrad)), c0, c1,
                                            C.byref(ene), force, stress, ptr(xx), ptr(dedg), ptr(eat), ptr(stats))
        else:
            rc = self.o.lib.gapo_calc_sparse(self.h, na, species, lat, pos, float(rcut), int(bool(lgrad)), int(extra),
                                             C.byref(ene), force, stress, ptr(xx), ptr(dedg), ptr(eat), ptr(stats))
        if rc:
            raise RuntimeError(_ERR.get(rc, "oracle error %d" % rc))
        out = {"energy": ene.value, "forces": force, "stress": stress}
        if want_desc:
            out.update(xx=xx, dedg=dedg, eatom=eat)
        if want_stats:
            out["stats"] = stats
        return out

    def calc_dense(self, species, lat, pos, rcut=6.0, lgrad=True, centres=(0, 0), desc=False, stats=False):
        """The reference algorithm loop for loop (arbiter / CPU baseline)."""
        return self._calc(True, species, lat, pos, rcut, lgrad, centres, desc, stats)

    def calc_sparse(self, species, lat, pos, rcut=6.0, lgrad=True, max_nb=1000, desc=False, stats=False):
        """Same arithmetic in O(N) memory (validated against calc_dense)."""
        return self._calc(False, species, lat, pos, rcut, lgrad, max_nb, desc, stats)

    def calc_sparse_centres(self, species, lat, pos, rcut, lgrad, c0, c1, max_nb=1000):
        """Centres [c0, c1) of a large structure only (bench.py's bounded CPU sample)."""
        species = np.ascontiguousarray(species, np.int32)
        lat = np.ascontiguousarray(lat, np.float64); pos = np.ascontiguousarray(pos, np.float64)
        ene = C.c_double(); force = np.zeros((pos.shape[0], 3)); stress = np.zeros(6)
        rc = self.o.lib.gapo_calc_sparse_centres(self.h, pos.shape[0], species, lat, pos, float(rcut), int(bool(lgrad)), int(max_nb),
                                                 int(c0), int(c1), C.byref(ene), force, stress)
        if rc:
            raise RuntimeError(_ERR.get(rc, "oracle error %d" % rc))
        return {"energy": ene.value, "forces": force, "stress": stress}

    def variance(self, xx, qmm, delta=1.0):
        """The predictive variance the reference carries commented out (gap_calc.f90:205-210):
        covf(i) = delta - ckm(i,:) . matmul(qmm, ckm(i,:)), VARIANCE = sum(covf) / na, with
        ckm = GET_COV (gap_calc.f90:268-288).  xx: descriptors [na, des_len] of calc_*(desc=True)."""
        d = (np.asarray(xx)[:, None, :] - self.mm[None, :, :]) / self.theta
        ckm = delta * np.exp(-0.5 * (d * d).sum(-1))
        covf = delta - np.einsum("ij,jl,il->i", ckm, np.asarray(qmm, float), ckm)
        return float(covf.sum() / len(covf)), covf

    def car2acsf_dense(self, species, lat, pos, rcut=6.0, lgrad=True):
        species = np.ascontiguousarray(species, np.int32)
        lat = np.ascontiguousarray(lat, np.float64); pos = np.ascontiguousarray(pos, np.float64)
        na, D = pos.shape[0], self.des_len
        xx = np.zeros((na, D)); dxdy = np.zeros((na, na, 3, D)); strs = np.zeros((na, D, 3, 3))
        rc = self.o.lib.gapo_car2acsf_dense(self.h, na, species, lat, pos, float(rcut), int(bool(lgrad)), xx, dxdy, strs)
        if rc:
            raise RuntimeError(_ERR.get(rc, "oracle error %d" % rc))
        return xx, dxdy, strs
