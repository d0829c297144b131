#!/usr/bin/env python
"""bench.py -- atom-steps/s of one energy+force+stress evaluation (BASELINE.json).

Workload: BASELINE config 4 -- ONE synthetic 100,000-atom periodic 3-species supercell
(50x50x40 simple-cubic sites, a = 2.15 A, jitter 0.15 A, seed 4000) with the
synthetic 3-species potential bench_data/gap_parameters_c2 (shipped 33-row SF table, M = 129,
D = 66), rcut = 6.0, lgrad = true.  A step is one E+F+stress evaluation of that cell, neighbour
lists rebuilt from scratch in every step.
N = 1: one context evaluates the whole cell.
N > 1: STRONG scaling of the same cell.  One process per GPU; the cell is cut into N bricks, a
rank keeps only its own atoms, and every step the ranks exchange ghost atoms (grouped
ncclSend/ncclRecv, 26 directions), return the ghost gradients the same way and all-gather one
48-double record (E, stress sums, status flags) -- all on the library's own NCCL communicator
(csrc/halo.cu, csrc/domain_host.inc).  torch.distributed only carries the NCCL id, the barrier
and the max-over-ranks of the times.

  value  whole-job atom-steps/s with inputs resident in HBM, CUDA events on the library's
         stream (gapcu_ctx_time_compute), L2 flushed between steps, max over ranks
  e2e    the same metric with HOST buffers every step.  N = 1: gapcu_calc, the C ABI FGAP_CALC
         binds (stat of ./gap_parameters, H2D of the whole structure, kernels, D2H).  N > 1: the
         distributed persistent API (each rank: H2D of its own atoms' new positions, the
         collective pass, D2H of its own atoms' forces + E + stress)
  roofline  the wACSF centre kernel of rank 0: algorithmic FP64 FLOPs of SURVEY.md 8(d),
         counted by the kernel's own counters, over its CUDA-event time, against the DFMA peak
         measured in the same run (MEASURED_PEAKS.json has no FP64 figure)
  cpu_baseline  the O(N) CPU port of the reference algorithm (oracle, gcc -O3 -march=native) on
         a bounded sample of centres of the same cell, 1 core (the reference is serial; its
         own dense algorithm cannot allocate this cell: gap_calc.f90:123 would need 15.8 TB)
  extra  rank 0, N = 1 only: last round's C2 line, Verlet-skin reuse on C4, C3 and C5

--impl reference times that CPU code with every host core on samples of the same cell.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in ("calypso-gap_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))

from structures import cubic_supercell  # noqa: E402  (pure numpy generators)

POT = os.path.join(ROOT, "bench_data", "gap_parameters_c2")
RCUT = 6.0
METRIC = "atom-steps/s (E+F+stress)"
WORKLOAD = "C4: synthetic 100000-atom periodic 3-species supercell (50x50x40 sites), one E/F/stress evaluation per step"
CONFIG = {"workload": WORKLOAD, "atoms": 100000, "potential": "synthetic 3-species, 33 SF, M=129, D=66", "rcut": RCUT,
          "neighbor_lists": "rebuilt every step",
          "l2": "L2 flushed between timed steps (256 MiB memset outside the timed events); working set 0.6 GB at N=1"}
L2_FLUSH = 256 << 20


def workload():
    return cubic_supercell(50, 50, 40, a=2.15, jitter=0.15, seed=4000)


def survey_flops(w, M, D):
    """algorithmic FP64 work from the kernel's counters (+,-,* = 1; FMA = 2;
    div, sqrt, exp, sin, cos = 1)."""
    w_desc = (13 * w["pairs"] + 6 * w["pair_classes"] + 15 * w["radial_sf"] + 21 * w["pair_classes"] +
              8 * w["class_candidates"] + 35 * w["triplet_classes"] + 33 * w["triplet_sf"] + 23 * w["triplet_classes"])
    w_gpr = w["atoms"] * (4 * M * D + 4 * M + 3 * D)
    return w_desc, w_gpr


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def read_gpr(L, path):
    """FGAP_READ through the C ABI: (theta, mm [M, D], coeff) as the caller of FGAP_CALC holds them."""
    import ctypes as C
    nsp, dl = C.c_int(), C.c_int()
    theta = np.zeros(512); mm = np.zeros((12000, 512), order="F"); coeff = np.zeros(12000)
    L.gapcu_read.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                             C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    rc = L.gapcu_read(os.fsencode(path), C.byref(nsp), C.byref(dl), theta.ctypes.data, 512, mm.ctypes.data, 12000, 512,
                      None, 0, coeff.ctypes.data, 12000)
    assert rc == 0, L.gapcu_last_error()
    M, D = nsp.value, dl.value
    return theta[:D].copy(), mm[:M, :D].copy(), coeff[:M].copy()


class FortranCaller:
    """gapcu_calc exactly as FGAP_CALC's caller drives it: Fortran-layout host buffers, the
    potential side channel ./gap_parameters in the working directory, raw pointers."""

    def __init__(self, L, potfile, z, cell, tag):
        import ctypes as C
        self.C, self.L = C, L
        self.dir = os.path.join("/tmp", "gapcu_bench_%s" % tag)
        os.makedirs(self.dir, exist_ok=True)
        link = os.path.join(self.dir, "gap_parameters")
        if os.path.lexists(link):
            os.remove(link)
        os.symlink(potfile, link)
        th, mmc, co = read_gpr(L, potfile)
        self.M, self.D = mmc.shape
        self.keep = (np.ascontiguousarray(z, np.int32), np.asfortranarray(cell), th, np.asfortranarray(mmc), co)
        self.na = len(z)
        self.f_out = np.zeros((self.na, 3), order="F"); self.s_out = np.zeros(6)
        self.e_out = C.c_double(); self.v_out = C.c_double()
        zi, latf, th, mmf, co = self.keep
        self.args = (zi.ctypes.data, latf.ctypes.data, th.ctypes.data, mmf.ctypes.data, co.ctypes.data)
        self.outs = (C.addressof(self.e_out), self.f_out.ctypes.data, self.s_out.ctypes.data, C.addressof(self.v_out))

    def __call__(self, pos_ptr):
        p_z, p_lat, p_th, p_mm, p_co = self.args
        cwd = os.getcwd()
        os.chdir(self.dir)
        try:
            rc = self.L.gapcu_calc(self.na, p_z, p_lat, pos_ptr, self.M, self.D, p_th, p_mm, None, p_co, RCUT, 1, *self.outs)
        finally:
            os.chdir(cwd)
        assert rc == 0, self.L.gapcu_last_error()
        return self.e_out.value


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gapcu
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)
    if world > 1:
        # stdout carries exactly one JSON line (rank 0): NCCL's banner and INFO lines go to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    os.environ["GAPCU_DEVICE"] = str(local)   # device of the Fortran-style entry points (gapcu_calc) on this rank
    dev = torch.device("cuda", local)
    cell, pos, z = workload()
    natoms = len(pos)
    ctx = gapcu.Context(local)
    ctx.load_potential(POT)
    grid = (1, 1, 1)
    if world > 1:
        obj = [gapcu.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        ctx.nccl_init(world, rank, obj[0])      # the library's own communicator: halo exchange runs on it
        grid = gapcu.domain_grid(world, cell, RCUT + 0.5)
        ctx.set_domain(grid, gapcu.brick_of(rank, grid))
    ctx.set_structures(z, cell, pos, RCUT)
    ctx.compute(True)
    e0, f0, s0 = ctx.fetch()
    ids = ctx.owned()
    dfma, dmma = ctx.fp64_peaks() if rank == 0 else (0.0, 0.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident throughput ------------------------------------------------
    ctx.time_compute(W, True, L2_FLUSH, stages=False)      # warm-up (also settles capacities)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    ms, _, launches = ctx.time_compute(args.steps, True, L2_FLUSH, stages=False)
    barrier()
    ms_max = allmax(ms)
    launches_all = allsum(launches)
    value = natoms * args.steps / (ms_max * 1e-3)
    # the clocks line needs the GPU under load for a few samples: keep stepping briefly (all ranks: the pass is collective)
    for _ in range(3):
        ctx.time_compute(max(4, int(0.3e3 / max(ms / args.steps, 1e-3))), True, 0, stages=False)
    clocks = sampler.stop() if sampler else None
    # per-stage device times (instrumented run) for the roofline
    _, stages, _ = ctx.time_compute(args.steps, True, L2_FLUSH, stages=True)
    work = ctx.work_counters()

    # ---- end to end with host buffers ------------------------------------------------
    rng = np.random.default_rng(5000)
    nsets = min(8, max(args.steps, W))
    L = gapcu.lib()
    if world == 1:
        # fresh host positions every step (an MD-like perturbation), generated before the timed region
        pos_steps = [np.asfortranarray(pos + rng.normal(0.0, 0.01, pos.shape)) for _ in range(nsets)]
        ptrs = [p.ctypes.data for p in pos_steps]
        fc = FortranCaller(L, POT, z, cell, "c4_rank%d" % rank)
        k = [0]

        def e2e_step():
            k[0] += 1
            return fc(ptrs[k[0] % nsets])
        h2d = 24 * natoms + 400   # positions + cell record every step; species weights (8 B/atom) and structure ids (4 B/atom) only travel when they change, i.e. on the first call
        d2h = 24 * natoms + 512 + 256                        # one copy: flags/counters slot, (E, stress, variance) slot, forces
        api = "gapcu_calc (C ABI bound by FGAP_CALC), host buffers, ./gap_parameters side channel"
    else:
        own = [np.ascontiguousarray((pos + rng.normal(0.0, 0.01, pos.shape))[ids]) for _ in range(nsets)]
        k = [0]

        def e2e_step():
            k[0] += 1
            ctx.update_positions(own[k[0] % nsets], False)   # lists rebuilt: same work as the N = 1 call
            ctx.compute(True)
            e, f, s = ctx.fetch()
            return e[0]
        h2d = allsum(24 * len(ids))
        d2h = allsum(24 * len(ids) + 512 + 256)
        api = ("gapcu_ctx_update_positions + gapcu_ctx_compute + gapcu_ctx_fetch on every rank: own atoms' positions in, "
               "own atoms' forces + E + stress out (bytes summed over ranks)")
    for _ in range(W):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e_last = e2e_step()
    torch.cuda.synchronize()
    dt = allmax(time.perf_counter() - t0)
    e2e_value = natoms * args.steps / dt

    # Verlet reuse on the same cell at every N (collective in a decomposed run)
    ctx.set_skin(0.5)
    ctx.set_structures(z, cell, pos, RCUT)
    ctx.compute(True)
    ctx.fetch()
    own_v = np.ascontiguousarray((pos + rng.normal(0.0, 0.01, pos.shape))[ids])
    ctx.update_positions(own_v, True)
    ctx.time_compute(3, True, L2_FLUSH, stages=False)
    ctx.update_positions(own_v, True)
    barrier()
    ms_v, st_v, _ = ctx.time_compute(args.steps, True, L2_FLUSH, stages=(world == 1))
    ms_v = allmax(ms_v)
    verlet = {"value": natoms * args.steps / (ms_v * 1e-3), "unit": "atom-steps/s", "ms_per_step": ms_v / args.steps, "skin": 0.5,
              "note": "skin lists kept, every step re-filters them against rcut with the reference arithmetic "
                      "(neighbour sets stay the reference's); device-resident"}
    if world == 1:
        verlet["stage_ms_per_step"] = {k2: v / args.steps for k2, v in st_v.items()}
    ctx.set_skin(0.0)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (rank 0's centres) ----------------------------------
    th, mmc, co = read_gpr(L, POT)
    M, D = mmc.shape
    w_desc, w_gpr = survey_flops(work, M, D)
    t_desc = (stages["descriptor_forward"] + stages["gpr_dmma"] + stages["descriptor_backward"]) / args.steps * 1e-3
    achieved = (w_desc + w_gpr) / t_desc / 1e12
    traffic, ncu_extra = None, {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True).stdout.strip()
        ncu_extra = {"fp64_pipe_active_pct": tj.get("fp64_pipe_active_pct"), "issue_slots_busy_pct": tj.get("issue_slots_busy_pct"),
                     "source": tj.get("source"), "captured_at_commit": tj.get("commit"), "head": head or None}
        # the capture belongs to a build: a figure from another kernel version is not reported as this one's
        if tj.get("kernel_sha") and tj.get("kernel_sha") == kernel_sha():
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except (OSError, KeyError, ValueError):
        pass
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, KeyError, ValueError):
        hbm_peak, hbm_src = 6456.2, "fallback (B200_PROFILING.md)"
    ncent = max(work["atoms"], 1.0)
    pbar = work["pairs"] / ncent
    # algorithmic bytes: positions, 8-byte list entries (skin list + exact list written, read by the
    # centre kernel and the gather), 24-byte pair gradients, forces
    b_k1 = ncent * (24 + 4 + 2 * 8 * pbar + 8)
    b_k5 = ncent * (8 * pbar + 24 * pbar + 24 + 72 + 8)
    t_k1 = stages["neighbor_build"] / args.steps * 1e-3
    t_k5 = stages["force_gather_reduce"] / args.steps * 1e-3
    hbm_passes = {"peak": hbm_peak, "peak_source": hbm_src, "unit": "GB/s",
                  "neighbor_build": {"bytes": b_k1, "achieved": b_k1 / t_k1 / 1e9, "frac": b_k1 / t_k1 / 1e9 / hbm_peak,
                                     "includes": "halo selection, exchange and unpack when decomposed"},
                  "force_gather_reduce": {"bytes": b_k5, "achieved": b_k5 / t_k5 / 1e9, "frac": b_k5 / t_k5 / 1e9 / hbm_peak,
                                          "includes": "gradient return, record all-gather when decomposed"}}
    roofline = {"bound": "fp64", "kernel": "k_centre<fused> (wACSF forward + in-CTA GPR + backward) on rank 0's %d centres" % int(ncent),
                "achieved": achieved, "peak": dfma, "unit": "TFLOP/s", "frac": achieved / dfma if dfma else None,
                "peak_source": "DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "dmma_peak_tflops": dmma, "traffic": traffic, "traffic_unit": "bytes per launch (dram read+write, ncu)",
                "ncu": ncu_extra,
                "flops_per_step": w_desc + w_gpr, "flops_desc": w_desc, "flops_gpr": w_gpr, "seconds_per_step": t_desc,
                "hbm_passes": hbm_passes, "stage_ms_per_step": {k2: v / args.steps for k2, v in stages.items()}}
    cpu = cpu_reference(1, args.cpu_centres, 1)
    extra = {"md_verlet": verlet}
    if world == 1 and not args.no_extras:
        for name, fn in (("c2", extra_c2), ("c5", extra_c5), ("c3", extra_c3)):
            try:
                extra[name] = fn(args, gapcu, ctx)
            except Exception as ex:   # an extra line never takes the headline down
                extra[name] = {"error": "%s: %s" % (type(ex).__name__, ex)}
    out = {"metric": METRIC, "value": value, "unit": "atom-steps/s", "n_gpus": world, "steps": args.steps,
           "warmup": W, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": CONFIG, "parallelism": ("bricks %dx%dx%d, ghost halo exchange over NCCL" % tuple(grid)) if world > 1 else "single GPU",
           "clocks": clocks, "gpu_launches": int(launches_all),
           "e2e": {"value": e2e_value, "unit": "atom-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "api": api},
           "roofline": roofline, "cpu_baseline": cpu, "check_energy": float(e_last), "energy_resident": float(e0[0]), "extra": extra}
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def kernel_sha():
    """Identity of the centre kernel's sources (ties an ncu capture to a build)."""
    import hashlib
    h = hashlib.sha1()
    for f in ("centre_impl.cuh", "fastmath.cuh", "geom.cuh", "device_types.cuh"):
        h.update(open(os.path.join(ROOT, "calypso-gap_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


# ---- extras (rank 0, one GPU) ------------------------------------------------------------------------
def extra_c2(args, gapcu, _ctx):
    """Round 1's headline, kept for comparison: the 1000-atom cell, device-resident and through gapcu_calc."""
    import torch
    cell, pos, z = cubic_supercell(10, 10, 10, a=2.15, jitter=0.15, seed=1000)
    c = gapcu.Context(int(os.environ.get("LOCAL_RANK", "0")))
    c.load_potential(POT)
    c.set_structures(z, cell, pos, RCUT)
    steps = max(args.steps, 200)
    c.time_compute(20, True, L2_FLUSH, stages=False)
    ms, st, _ = c.time_compute(steps, True, L2_FLUSH, stages=True)
    rng = np.random.default_rng(5001)
    pos_steps = [np.asfortranarray(pos + rng.normal(0.0, 0.01, pos.shape)) for _ in range(8)]
    fc = FortranCaller(gapcu.lib(), POT, z, cell, "c2")
    for k in range(20):
        fc(pos_steps[k % 8].ctypes.data)
    t0 = time.perf_counter()
    for k in range(steps):
        fc(pos_steps[k % 8].ctypes.data)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    c.close()
    return {"workload": "C2: 1000-atom 3-species supercell", "value": 1000 * steps / (ms * 1e-3), "e2e": 1000 * steps / dt,
            "unit": "atom-steps/s", "ms_per_step": ms / steps, "steps": steps, "stage_ms_per_step": {k: v / steps for k, v in st.items()}}


def c5_potential(gapcu, c, M=10000):
    """nsf = 128 (D = 256), M sparse points harvested from sibling structures on the GPU."""
    ntype, alpha, cut = [], [], []
    for a in np.geomspace(1e-3, 2.0, 32): ntype.append(1); alpha.append(a); cut.append(6.0)
    for rs in np.linspace(0.5, 5.5, 32): ntype.append(3); alpha.append(rs); cut.append(6.0)
    for rc in (3.0, 4.0, 5.0, 6.0):
        for a in np.geomspace(2e-3, 0.3, 12): ntype.append(2); alpha.append(a); cut.append(rc)
        for a in np.geomspace(2e-3, 0.3, 12)[::3]: ntype.append(4); alpha.append(a); cut.append(rc)
    ntype = np.array(ntype, np.int32); alpha = np.round(np.array(alpha), 5); cut = np.array(cut)
    D = 2 * len(ntype)
    z3 = np.array([5, 6, 7], np.int32); w3 = np.array([-1.0, 4.0, 2.0])
    c.set_potential(z3, w3, ntype, alpha, cut, np.ones(D), np.zeros((16, D)), np.zeros(16))
    rows, seed = [], 2001
    while sum(len(r) for r in rows) < M:
        cell, pos, z = cubic_supercell(10, 10, 10, seed=seed); seed += 1
        c.evaluate(z, cell, pos, RCUT, False)
        rows.append(c.descriptors(D)[0])
    mm = np.vstack(rows)[:M]
    theta = np.maximum(mm.std(0), 1e-3) * np.sqrt(D)
    coeff = np.random.default_rng(8).normal(size=M) * 50.0
    return z3, w3, ntype, alpha, cut, theta, mm, coeff


def extra_c5(args, gapcu, _ctx):
    """BASELINE config 5: M = 10,000 sparse points, D = 256 on the 1000-atom cell: the DMMA GPR kernel."""
    import torch
    from structures import write_gap_parameters
    c = gapcu.Context(int(os.environ.get("LOCAL_RANK", "0")))
    z3, w3, ntype, alpha, cut, theta, mm, coeff = c5_potential(gapcu, c)
    M, D = mm.shape
    c.set_potential(z3, w3, ntype, alpha, cut, theta, mm, coeff)
    cell, pos, z = cubic_supercell(10, 10, 10, a=2.15, jitter=0.15, seed=1000)
    c.set_structures(z, cell, pos, RCUT)
    steps = max(args.steps, 20)
    c.time_compute(5, True, L2_FLUSH, stages=False)
    ms, st, _ = c.time_compute(steps, True, L2_FLUSH, stages=True)
    _, dmma = c.fp64_peaks()
    flops = 1000 * (4.0 * M * D + 4 * M + 3 * D)
    gpr_t = st["gpr_dmma"] / steps * 1e-3
    # end to end through gapcu_calc: the file carries the SF table and the weights, the arrays carry the GPR block
    d = os.path.join("/tmp", "gapcu_bench_c5")
    os.makedirs(d, exist_ok=True)
    potfile = os.path.join(d, "gap_parameters_c5_sf")
    write_gap_parameters(potfile, z3, w3, ntype, alpha, cut, np.ones(2), np.zeros((1, 2)), np.zeros(1))
    L = gapcu.lib()
    link = os.path.join(d, "gap_parameters")
    if os.path.lexists(link):
        os.remove(link)
    os.symlink(potfile, link)
    import ctypes as C
    zi = np.ascontiguousarray(z, np.int32); latf = np.asfortranarray(cell); mmf = np.asfortranarray(mm)
    f_out = np.zeros((1000, 3), order="F"); s_out = np.zeros(6); e_out = C.c_double(); v_out = C.c_double()
    rng = np.random.default_rng(5005)
    pos_steps = [np.asfortranarray(pos + rng.normal(0.0, 0.01, pos.shape)) for _ in range(8)]
    cwd = os.getcwd()
    os.chdir(d)
    try:
        def call(k):
            rc = L.gapcu_calc(1000, zi.ctypes.data, latf.ctypes.data, pos_steps[k % 8].ctypes.data, M, D, theta.ctypes.data,
                              mmf.ctypes.data, None, coeff.ctypes.data, RCUT, 1, C.addressof(e_out), f_out.ctypes.data,
                              s_out.ctypes.data, C.addressof(v_out))
            assert rc == 0, L.gapcu_last_error()
        for k in range(5):
            call(k)
        t0 = time.perf_counter()
        for k in range(steps):
            call(k)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    finally:
        os.chdir(cwd)
    c.close()
    return {"workload": "C5: 1000 atoms, M=10000 sparse points, D=256", "value": 1000 * steps / (ms * 1e-3), "e2e": 1000 * steps / dt,
            "unit": "atom-steps/s", "ms_per_step": ms / steps, "steps": steps,
            "gpr_dmma": {"achieved": flops / gpr_t / 1e12, "peak": dmma, "unit": "TFLOP/s", "frac": flops / gpr_t / 1e12 / dmma if dmma else None,
                         "kernel": "k_gpr (mma.sync m8n8k4 f64)", "ms": gpr_t * 1e3},
            "stage_ms_per_step": {k: v / steps for k, v in st.items()}}


def extra_c3(args, gapcu, _ctx):
    """BASELINE config 3 as written: 4,096 random candidates (32-128 atoms) through gapcu_calc_batch over
    every visible device, host arrays in and out; the first structures are checked against the oracle."""
    import pickle
    n = args.c3_structs
    out = os.path.join("/tmp", "gapcu_bench_c3_%d.pkl" % n)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_c3.py"), str(n), out])   # child process: no fork from the CUDA process
    with open(out, "rb") as fh:
        structs = pickle.load(fh)
    os.remove(out)
    cells, poss, zs = [s[0] for s in structs], [s[1] for s in structs], [s[2] for s in structs]
    natoms = sum(len(p) for p in poss)
    ndev = gapcu.device_count()
    gapcu.set_devices(range(ndev))
    d = os.path.join("/tmp", "gapcu_bench_c3")
    os.makedirs(d, exist_ok=True)
    link = os.path.join(d, "gap_parameters")
    if os.path.lexists(link):
        os.remove(link)
    os.symlink(POT, link)
    cwd = os.getcwd()
    os.chdir(d)
    try:
        e, f, s = gapcu.calc_batch(zs, cells, poss, RCUT, True)      # warm-up (capacities, potential upload)
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            e, f, s = gapcu.calc_batch(zs, cells, poss, RCUT, True)
        dt = (time.perf_counter() - t0) / reps
    finally:
        os.chdir(cwd)
        gapcu.set_devices([int(os.environ.get("LOCAL_RANK", "0"))])
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle import Oracle
    pot = Oracle("parity").read(POT)
    nchk = min(args.c3_check, n)
    de = df = ds = 0.0
    for k in range(nchk):
        w = pot.calc_sparse(zs[k], cells[k], poss[k], RCUT, True)
        de = max(de, abs(e[k] - w["energy"]) / abs(w["energy"]))
        df = max(df, float(np.abs(f[k] - w["forces"]).max() / max(1.0, np.abs(w["forces"]).max())))
        ds = max(ds, float(np.abs(s[k] - w["stress"]).max() / max(1.0, np.abs(w["stress"]).max())))
    return {"workload": "C3: %d random candidate structures (32-128 atoms, triclinic), %d atoms" % (n, natoms),
            "e2e": natoms / dt, "unit": "atom-steps/s", "seconds_per_pass": dt, "devices": ndev,
            "api": "gapcu_calc_batch (host arrays in, host arrays out, python packing included)",
            "oracle_check": {"structures": nchk, "max_rel_dE": de, "max_rel_dF": df, "max_rel_dS": ds}}


# ---- CPU arm ------------------------------------------------------------------------------------------
_CPU = {}


def cpu_reference(threads, sample_centres, repeats):
    """Times the O(N) CPU port of the reference algorithm (oracle.calc_sparse_centres: same per-centre
    arithmetic as the dense transliteration, validated against it) on `sample_centres` centre atoms
    of the C4 cell per thread; each thread takes a different contiguous range of centres."""
    from concurrent.futures import ThreadPoolExecutor
    if not _CPU:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from oracle import Oracle
        _CPU["pot"] = Oracle("fast").read(POT)
        _CPU["w"] = workload()
    pot = _CPU["pot"]
    cell, pos, z = _CPU["w"]
    n = len(pos)

    def one(t):
        c0 = (t * 7919 * sample_centres) % max(1, n - sample_centres)
        pot.calc_sparse_centres(z, cell, pos, RCUT, True, c0, c0 + sample_centres)

    t0 = time.perf_counter()
    for _ in range(repeats):
        if threads == 1:
            one(0)
        else:
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(one, range(threads)))
    dt = time.perf_counter() - t0
    return {"value": threads * sample_centres * repeats / dt, "unit": "atom-steps/s", "cores": threads, "kind": "port-sparse",
            "sample": "%d of the %d centre atoms of the C4 cell per thread (per-centre sample: energies, the forces those "
                      "centres exert and their stress share; cell list built per call), O(N) C port of the reference "
                      "algorithm, gcc -O3 -march=native; the reference's dense algorithm cannot allocate this cell "
                      "(gap_calc.f90:123: 15.8 TB) and its Fortran cannot be built here (no Fortran compiler)" % (sample_centres, n),
            "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    world = int(os.environ.get("WORLD_SIZE", "1"))
    centres = args.ref_centres
    for _ in range(args.warmup):
        cpu_reference(threads, centres, 1)
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = cpu_reference(threads, centres, 1)
    dt = time.perf_counter() - t0
    value = threads * centres * args.steps / dt
    last["value"] = value
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "atom-steps/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": CONFIG, "parallelism": "%d host threads, each a per-centre sample of the cell" % threads,
           "cpu_baseline": last,
           "e2e": {"value": value, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-centres", type=int, default=8000, help="centre atoms of the cpu_baseline sample (1 core)")
    ap.add_argument("--ref-centres", type=int, default=300, help="centre atoms per thread and step of --impl reference")
    ap.add_argument("--c3-structs", type=int, default=4096)
    ap.add_argument("--c3-check", type=int, default=64)
    ap.add_argument("--no-extras", action="store_true", help="skip the C2 / C3 / C5 lines")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
