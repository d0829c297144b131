"""ctypes binding of the additive C ABI in include/gapcu.h (persistent context,
batches, device-side timing).  The reference-compatible surface is the ``libgap``
package next to this file; this module is what bench.py and the parity tests use
to reach the same kernels without the per-call file re-read of the Fortran API.

There is no CPU fallback: ``Context()`` raises if lib/libgapcu.so is missing or
no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("GAPCU_LIB", os.path.join(HERE, "lib", "libgapcu.so"))   # GAPCU_LIB: alternative build (A/B timing)
NSTAGE = 8

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_vp = C.c_void_p
_lib = None

# every symbol include/gapcu.h declares (tests check the library exports them all)
SYMBOLS = [
    "gapcu_last_error", "gapcu_calc", "gapcu_read", "gapcu_bond", "gapcu_car2acsf_table", "gapcu_print_last_error",
    "gapcu_set_devices", "gapcu_calc_batch",
    "gapcu_device_count", "gapcu_ctx_create", "gapcu_ctx_destroy", "gapcu_ctx_load_potential",
    "gapcu_ctx_set_potential", "gapcu_ctx_set_pipeline", "gapcu_ctx_set_cluster", "gapcu_nccl_unique_id", "gapcu_ctx_nccl_init",
    "gapcu_ctx_set_domain", "gapcu_ctx_set_drift", "gapcu_ctx_owned", "gapcu_ctx_set_skin", "gapcu_ctx_update_positions",
    "gapcu_group_create", "gapcu_group_destroy", "gapcu_group_size", "gapcu_group_ctx", "gapcu_group_load_potential",
    "gapcu_group_set_skin", "gapcu_group_set_structure", "gapcu_group_update_positions", "gapcu_group_compute", "gapcu_group_fetch",
    "gapcu_ctx_set_structures", "gapcu_ctx_compute", "gapcu_ctx_fetch",
    "gapcu_ctx_fetch_descriptors", "gapcu_ctx_variance", "gapcu_ctx_fetch_neighbors", "gapcu_ctx_debug_triplets", "gapcu_ctx_time_compute", "gapcu_stage_name",
    "gapcu_ctx_work_counters", "gapcu_ctx_balance", "gapcu_fp64_peaks",
]
FORTRAN_SYMBOLS = ["fgap_calc_", "fgap_read_", "fget_bond_", "car2acsf_", "write_array_2dim_"]


class GapcuError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "gapcu error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise ImportError("%s not built: run `python calypso-gap_b200/build.py` (there is no CPU fallback)" % LIBPATH)
        L = C.CDLL(LIBPATH)
        L.gapcu_last_error.restype = C.c_char_p
        L.gapcu_stage_name.restype = C.c_char_p
        L.gapcu_stage_name.argtypes = [C.c_int]
        L.gapcu_ctx_create.restype = _vp
        L.gapcu_ctx_create.argtypes = [C.c_int]
        L.gapcu_ctx_destroy.argtypes = [_vp]
        L.gapcu_ctx_load_potential.argtypes = [_vp, C.c_char_p]
        L.gapcu_ctx_set_potential.argtypes = [_vp, C.c_int, _ip, _dp, C.c_int, _ip, _dp, _dp, C.c_int, C.c_int, _dp, _dp, _dp]
        L.gapcu_ctx_set_structures.argtypes = [_vp, C.c_int, _ip, _ip, _dp, _dp, C.c_double]
        L.gapcu_ctx_compute.argtypes = [_vp, C.c_int]
        L.gapcu_ctx_set_pipeline.argtypes = [_vp, C.c_int]
        L.gapcu_ctx_set_cluster.argtypes = [_vp, C.c_int]
        L.gapcu_nccl_unique_id.argtypes = [C.c_char_p]
        L.gapcu_ctx_nccl_init.argtypes = [_vp, C.c_int, C.c_int, C.c_char_p]
        L.gapcu_ctx_set_domain.argtypes = [_vp] + [C.c_int] * 6
        L.gapcu_ctx_set_drift.argtypes = [_vp, C.c_double]
        L.gapcu_ctx_set_skin.argtypes = [_vp, C.c_double]
        L.gapcu_ctx_owned.argtypes = [_vp, C.POINTER(C.c_int), _vp]
        L.gapcu_ctx_update_positions.argtypes = [_vp, _vp, C.c_int]
        L.gapcu_group_create.restype = _vp
        L.gapcu_group_create.argtypes = [C.c_int, _ip]
        L.gapcu_group_destroy.argtypes = [_vp]
        L.gapcu_group_size.argtypes = [_vp]
        L.gapcu_group_ctx.restype = _vp
        L.gapcu_group_ctx.argtypes = [_vp, C.c_int]
        L.gapcu_group_load_potential.argtypes = [_vp, C.c_char_p]
        L.gapcu_group_set_skin.argtypes = [_vp, C.c_double]
        L.gapcu_group_set_structure.argtypes = [_vp, C.c_int, _ip, _dp, _dp, C.c_double, _vp]
        L.gapcu_group_update_positions.argtypes = [_vp, _dp, C.c_int]
        L.gapcu_group_compute.argtypes = [_vp, C.c_int]
        L.gapcu_group_fetch.argtypes = [_vp, _vp, _vp, _vp]
        L.gapcu_ctx_fetch.argtypes = [_vp, _vp, _vp, _vp]
        L.gapcu_ctx_fetch_descriptors.argtypes = [_vp, _vp, _vp, _vp]
        L.gapcu_ctx_fetch_neighbors.argtypes = [_vp, C.c_int, _ip, _ip, _ip, _dp]
        L.gapcu_ctx_time_compute.argtypes = [_vp, C.c_int, C.c_int, C.c_long, C.POINTER(C.c_double), _vp, C.POINTER(C.c_long)]
        L.gapcu_ctx_work_counters.argtypes = [_vp, _dp, C.c_int]
        L.gapcu_fp64_peaks.argtypes = [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.gapcu_calc.argtypes = [C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, C.c_double, C.c_int,
                                 _vp, _vp, _vp, _vp]
        L.gapcu_set_devices.argtypes = [C.c_int, _vp]
        L.gapcu_calc_batch.argtypes = [C.c_int, _vp, _vp, _vp, _vp, C.c_double, C.c_int, _vp, _vp, _vp]
        L.gapcu_bond.argtypes = [C.c_int, _vp, _vp, _vp, C.c_double, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _check(rc):
    if rc < 0:
        raise GapcuError(rc, lib().gapcu_last_error().decode(errors="replace"))
    return rc


def device_count():
    return lib().gapcu_device_count()


class Context:
    """One GPU, one stream, one potential, one resident batch of structures."""

    def __init__(self, device=0, handle=None):
        self.borrowed = handle is not None      # a group's context: the group destroys it
        h = handle if self.borrowed else lib().gapcu_ctx_create(int(device))
        if not h:
            raise GapcuError(-6, lib().gapcu_last_error().decode(errors="replace"))
        self.h = _vp(h)
        self.natoms = None
        self.des_len = None

    def close(self):
        if self.h and not self.borrowed:
            lib().gapcu_ctx_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_potential(self, path):
        _check(lib().gapcu_ctx_load_potential(self.h, os.fsencode(path)))

    def set_potential(self, z, w, ntype, alpha, cutoff, theta, mm, coeff):
        z = np.ascontiguousarray(z, np.int32); w = np.ascontiguousarray(w, np.float64)
        ntype = np.ascontiguousarray(ntype, np.int32); alpha = np.ascontiguousarray(alpha, np.float64)
        cutoff = np.ascontiguousarray(cutoff, np.float64); theta = np.ascontiguousarray(theta, np.float64)
        mm = np.ascontiguousarray(mm, np.float64); coeff = np.ascontiguousarray(coeff, np.float64)
        _check(lib().gapcu_ctx_set_potential(self.h, len(z), z, w, len(ntype), ntype, alpha, cutoff, mm.shape[0],
                                             mm.shape[1], theta, mm, coeff))
        self.des_len = mm.shape[1]

    def set_structures(self, species_list, lat_list, pos_list, rcut=6.0):
        """Lists of per-structure arrays (or a single structure's arrays)."""
        if np.ndim(lat_list) == 2:
            species_list, lat_list, pos_list = [species_list], [lat_list], [pos_list]
        natoms = np.array([len(p) for p in pos_list], np.int32)
        species = np.ascontiguousarray(np.concatenate([np.asarray(s).ravel() for s in species_list]), np.int32)
        lat = np.ascontiguousarray(np.stack([np.asarray(l, np.float64) for l in lat_list]))
        pos = np.ascontiguousarray(np.concatenate([np.asarray(p, np.float64).reshape(-1, 3) for p in pos_list]))
        _check(lib().gapcu_ctx_set_structures(self.h, len(natoms), natoms, species, lat, pos, float(rcut)))
        self.natoms = natoms

    def set_pipeline(self, mode):
        """'auto' | 'split' (K2 -> DMMA GPR -> K4) | 'fused' (one centre kernel)."""
        _check(lib().gapcu_ctx_set_pipeline(self.h, {"auto": 0, "split": 1, "fused": 2}[mode]))

    def set_cluster(self, ctas_per_centre):
        """CTAs (one thread-block cluster) per centre atom in the fused kernel: 0 = automatic, 1, 2, 4."""
        _check(lib().gapcu_ctx_set_cluster(self.h, int(ctas_per_centre)))

    def nccl_init(self, world, rank, unique_id):
        _check(lib().gapcu_ctx_nccl_init(self.h, int(world), int(rank), unique_id))

    def set_domain(self, grid, mine):
        """Spatial decomposition: this rank owns brick `mine` of `grid` (see domain_grid)."""
        _check(lib().gapcu_ctx_set_domain(self.h, *[int(v) for v in grid], *[int(v) for v in mine]))

    def compute(self, lgrad=True):
        _check(lib().gapcu_ctx_compute(self.h, int(bool(lgrad))))

    def set_skin(self, skin):
        """Verlet skin in Angstrom: candidate lists hold rcut + skin, see update_positions."""
        _check(lib().gapcu_ctx_set_skin(self.h, float(skin)))

    def set_drift(self, drift):
        _check(lib().gapcu_ctx_set_drift(self.h, float(drift)))

    def owned(self):
        """Indices (into the structure last set) of the atoms fetch() returns forces for."""
        n = C.c_int()
        _check(lib().gapcu_ctx_owned(self.h, C.byref(n), None))
        ids = np.zeros(n.value, np.int32)
        _check(lib().gapcu_ctx_owned(self.h, C.byref(n), ids.ctypes.data))
        return ids

    def update_positions(self, pos, reuse_lists=True):
        """Moved positions of the resident atoms ([n, 3]; decomposed runs: the owned atoms in the
        order of owned()).  reuse_lists: re-filter the kept skin lists instead of rebuilding."""
        pos = np.ascontiguousarray(pos, np.float64)
        _check(lib().gapcu_ctx_update_positions(self.h, pos.ctypes.data, int(bool(reuse_lists))))

    def fetch(self):
        ns = len(self.natoms)
        n = C.c_int()
        _check(lib().gapcu_ctx_owned(self.h, C.byref(n), None))
        nt = n.value
        ene = np.zeros(ns); force = np.zeros((nt, 3)); stress = np.zeros((ns, 6))
        _check(lib().gapcu_ctx_fetch(self.h, ene.ctypes.data, force.ctypes.data, stress.ctypes.data))
        return ene, force, stress

    def evaluate(self, species, lat, pos, rcut=6.0, lgrad=True):
        """Single structure convenience: returns dict(energy, forces, stress)."""
        self.set_structures(species, lat, pos, rcut)
        self.compute(lgrad)
        e, f, s = self.fetch()
        return {"energy": float(e[0]), "forces": f, "stress": s[0]}

    def descriptors(self, des_len):
        nt = int(self.natoms.sum())
        xx = np.zeros((nt, des_len)); dedg = np.zeros((nt, des_len)); eat = np.zeros(nt)
        _check(lib().gapcu_ctx_fetch_descriptors(self.h, xx.ctypes.data, dedg.ctypes.data, eat.ctypes.data))
        return xx, dedg, eat

    def variance(self, qmm):
        """Predictive variance of the last compute for a caller-supplied QMM (M x M): returns
        (VARIANCE per structure, covf per atom); formula of gap_calc.f90:207-210."""
        q = np.asfortranarray(qmm, dtype=np.float64)
        nt = int(self.natoms.sum())
        var = np.zeros(len(self.natoms)); covf = np.zeros(nt)
        lib().gapcu_ctx_variance.argtypes = [_vp, _vp, _vp, _vp]
        _check(lib().gapcu_ctx_variance(self.h, q.ctypes.data, var.ctypes.data, covf.ctypes.data))
        return var, covf

    def neighbors(self, cap=1000):
        nt = int(self.natoms.sum())
        count = np.zeros(nt, np.int32); idx = np.zeros((nt, cap), np.int32)
        shift = np.zeros((nt, cap, 3), np.int32); dis = np.zeros((nt, cap))
        _check(lib().gapcu_ctx_fetch_neighbors(self.h, cap, count, idx, shift, dis))
        return count, idx, shift, dis

    def triplets(self, cap=8192):
        """Kept neighbour pairs per atom as the kernel sums them: list of (slot_j, slot_k, nclasses) arrays."""
        nt = int(self.natoms.sum())
        count = np.zeros(nt, np.int32); items = np.zeros((nt, cap), np.uint32)
        lib().gapcu_ctx_debug_triplets.argtypes = [_vp, C.c_int, _ip, _vp]
        mx = _check(lib().gapcu_ctx_debug_triplets(self.h, cap, count, items.ctypes.data))
        if mx > cap:
            raise ValueError("triplet capacity %d too small (need %d)" % (cap, mx))
        out = []
        for i in range(nt):
            it = items[i, :count[i]]
            out.append(np.stack([it & 1023, (it >> 10) & 1023, it >> 20], 1).astype(np.int32))
        return out

    def time_compute(self, steps, lgrad=True, l2_flush_bytes=0, stages=True):
        ms = C.c_double(); launches = C.c_long()
        st = np.zeros(NSTAGE)
        _check(lib().gapcu_ctx_time_compute(self.h, int(bool(lgrad)), int(steps), int(l2_flush_bytes), C.byref(ms),
                                            st.ctypes.data if stages else None, C.byref(launches)))
        names = [lib().gapcu_stage_name(i).decode() for i in range(NSTAGE)]
        return ms.value, {n: v for n, v in zip(names, st) if n}, launches.value

    def work_counters(self):
        out = np.zeros(26)
        _check(lib().gapcu_ctx_work_counters(self.h, out, 26))
        keys = ["atoms", "pairs", "pair_classes", "candidates", "triplets", "triplet_classes", "triplet_sf", "radial_sf", "class_candidates"]
        d = dict(zip(keys, out))
        if out[10:].any():      # GAPCU_VARIANT=16: cycles per phase of the centre kernel (thread 0 of every CTA)
            d["phase_cycles"] = dict(zip(["stage", "radial_fwd", "list_build", "angular_fwd", "reduce_gpr", "radial_bwd", "angular_bwd", "epilogue",
                                         "list_pairs", "list_scan", "list_scatter", "desc_sums", "gpr_dist", "gpr_grad", "bwd_batches", "bwd_merge"], out[10:]))
        return d

    def balance(self):
        out = np.zeros(4)
        lib().gapcu_ctx_balance.argtypes = [_vp, _dp]
        _check(lib().gapcu_ctx_balance(self.h, out))
        return {"ctas": int(out[0]), "span_us": out[1], "first_idle_us": out[2], "busy_frac": out[3]}

    def fp64_peaks(self):
        a = C.c_double(); b = C.c_double()
        _check(lib().gapcu_fp64_peaks(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value


class Group:
    """One structure cut into bricks, one context per brick, all in this process (devices may
    repeat: several bricks on one GPU).  Arrays are those of the whole structure."""

    def __init__(self, devices):
        d = np.ascontiguousarray(list(devices), np.int32)
        h = lib().gapcu_group_create(len(d), d)
        if not h:
            raise GapcuError(-6, lib().gapcu_last_error().decode(errors="replace"))
        self.h = _vp(h)
        self.n = len(d)
        self.na = None

    def close(self):
        if self.h:
            lib().gapcu_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ctx(self, rank):
        return Context(handle=lib().gapcu_group_ctx(self.h, int(rank)))

    def load_potential(self, path):
        _check(lib().gapcu_group_load_potential(self.h, os.fsencode(path)))

    def set_skin(self, skin):
        _check(lib().gapcu_group_set_skin(self.h, float(skin)))

    def set_structure(self, species, lat, pos, rcut=6.0, grid=None):
        species = np.ascontiguousarray(species, np.int32)
        lat = np.ascontiguousarray(lat, np.float64); pos = np.ascontiguousarray(pos, np.float64)
        g = None if grid is None else np.ascontiguousarray(grid, np.int32)
        _check(lib().gapcu_group_set_structure(self.h, len(pos), species, lat, pos, float(rcut),
                                               None if g is None else g.ctypes.data))
        self.na = len(pos)

    def update_positions(self, pos, reuse_lists=True):
        _check(lib().gapcu_group_update_positions(self.h, np.ascontiguousarray(pos, np.float64), int(bool(reuse_lists))))

    def compute(self, lgrad=True):
        _check(lib().gapcu_group_compute(self.h, int(bool(lgrad))))

    def fetch(self):
        ene = np.zeros(1); force = np.zeros((self.na, 3)); stress = np.zeros(6)
        _check(lib().gapcu_group_fetch(self.h, ene.ctypes.data, force.ctypes.data, stress.ctypes.data))
        return float(ene[0]), force, stress

    def evaluate(self, species, lat, pos, rcut=6.0, lgrad=True, grid=None):
        self.set_structure(species, lat, pos, rcut, grid)
        self.compute(lgrad)
        e, f, s = self.fetch()
        return {"energy": e, "forces": f, "stress": s}


def set_devices(devices):
    """Devices of the drop-in entry points (gapcu_calc on the first, gapcu_calc_batch over all)."""
    d = np.ascontiguousarray(list(devices), np.int32)
    _check(lib().gapcu_set_devices(len(d), d.ctypes.data))


def calc_batch(species_list, lat_list, pos_list, rcut=6.0, lgrad=True):
    """Independent structures through the drop-in side channel (./gap_parameters in the working
    directory holds the whole potential), sharded over the devices of set_devices().  Returns
    (energies[ns], list of forces[n_i, 3], stresses[ns, 6])."""
    natoms = np.array([len(p) for p in pos_list], np.int32)
    species = np.ascontiguousarray(np.concatenate([np.asarray(s).ravel() for s in species_list]), np.int32)
    lat = np.ascontiguousarray(np.stack([np.asarray(l, np.float64) for l in lat_list]))
    pos = np.ascontiguousarray(np.concatenate([np.asarray(p, np.float64).reshape(-1, 3) for p in pos_list]))
    ene = np.zeros(len(natoms)); force = np.zeros((int(natoms.sum()), 3)); stress = np.zeros((len(natoms), 6))
    _check(lib().gapcu_calc_batch(len(natoms), natoms.ctypes.data, species.ctypes.data, lat.ctypes.data, pos.ctypes.data,
                                  float(rcut), int(bool(lgrad)), ene.ctypes.data, force.ctypes.data, stress.ctypes.data))
    offs = np.concatenate([[0], np.cumsum(natoms)])
    return ene, [force[offs[k]:offs[k + 1]] for k in range(len(natoms))], stress


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().gapcu_nccl_unique_id(buf))
    return buf.raw


def domain_grid(world, cell, shell=6.5):
    """Brick grid g0 x g1 x g2 = world with the smallest ghost volume for this cell.  `shell` = rcut + skin +
    2 * drift: every brick must be at least that thick, and a rank imports the images within it on ALL six
    sides of its brick (also along undivided directions: those ghosts are the cell's own periodic images)."""
    cell = np.asarray(cell, float)
    inv = np.linalg.inv(cell)
    spacing = 1.0 / np.linalg.norm(inv, axis=0)          # interplanar spacings
    best = None
    for g0 in range(1, world + 1):
        if world % g0:
            continue
        for g1 in range(1, world // g0 + 1):
            if (world // g0) % g1:
                continue
            g = (g0, g1, world // g0 // g1)
            w = [spacing[c] / g[c] for c in range(3)]    # brick widths
            if min(w) < shell:
                continue
            ghost = np.prod([w[c] + 2 * shell for c in range(3)]) - np.prod(w)
            key = (round(float(ghost), 9), g)
            if best is None or key < best:
                best = key
    if best is None:
        raise ValueError("no brick grid: the cell is too small for %d bricks of thickness >= %g" % (world, shell))
    return best[1]


def brick_of(rank, grid):
    return (rank // (grid[1] * grid[2]), (rank // grid[2]) % grid[1], rank % grid[2])


def fortran_calc(species, lat, pos, theta, mm, coeff, rcut, lgrad):
    """gapcu_calc with Fortran-layout buffers built from C-order numpy inputs (the
    marshalling f2py does); reads ./gap_parameters from the CWD like FGAP_CALC."""
    species = np.ascontiguousarray(species, np.int32)
    na = len(species)
    latf = np.asfortranarray(np.asarray(lat, np.float64)); posf = np.asfortranarray(np.asarray(pos, np.float64))
    theta = np.ascontiguousarray(theta, np.float64); mmf = np.asfortranarray(np.asarray(mm, np.float64))
    coeff = np.ascontiguousarray(coeff, np.float64)
    ene = C.c_double(); var = C.c_double()
    force = np.zeros((na, 3), order="F"); stress = np.zeros(6)
    _check(lib().gapcu_calc(na, species.ctypes.data, latf.ctypes.data, posf.ctypes.data, mmf.shape[0], mmf.shape[1],
                            theta.ctypes.data, mmf.ctypes.data, None, coeff.ctypes.data, float(rcut), int(bool(lgrad)),
                            C.addressof(ene), force.ctypes.data, stress.ctypes.data, C.addressof(var)))
    return ene.value, np.ascontiguousarray(force), stress, var.value
