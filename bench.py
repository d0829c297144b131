#!/usr/bin/env python
"""bench.py -- atom-steps/s of one energy+force+stress evaluation (BASELINE.json).

Workload (N=1): BASELINE config 2 -- a synthetic 1,000-atom periodic 3-species
supercell (10x10x10 simple-cubic sites, a = 2.15 A, jitter
0.15 A, seed 1000) with the synthetic 3-species potential bench_data/gap_parameters_c2
(shipped 33-row SF table, M = 129, D = 66), rcut = 6.0, lgrad = true.
N>1: one process per GPU, every rank evaluates its own structure of the same shape
(seed 1000+rank): independent structures share only the read-only potential, so
there is no data-path collective ("scaling": "weak"); torch.distributed is used
for the barrier and the max-over-ranks of the device time only.

  value  whole-job atom-steps/s with inputs resident in HBM, timed with CUDA events
         on the library's stream (gapcu_ctx_time_compute), L2 flushed between steps
  e2e    the same metric through the reference-facing C ABI call gapcu_calc (what
         FGAP_CALC binds): host buffers in, host buffers out, every step including
         the stat of ./gap_parameters, H2D, all kernels, D2H
  roofline  the wACSF centre kernel (forward + backward launches): algorithmic FP64
         FLOPs of SURVEY.md 8(d) W_desc, counted on the benchmark structure by the
         kernel's own counters, over the CUDA-event time of those launches, against
         the DFMA peak measured in the same run (MEASURED_PEAKS.json has no FP64 figure)
  cpu_baseline  the dense CPU oracle (the reference's algorithm, gcc -O3 -march=native)
         on the same structure, 1 core (the reference is serial)

--impl reference times that CPU implementation with every host core (independent
structure copies, the only way the reference is ever parallelised: tools/cgg2.py).
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
WORKLOAD = "C2: synthetic 1000-atom periodic 3-species supercell, single-point E/F/stress"
L2_FLUSH = 256 << 20


def workload(rank):
    return cubic_supercell(10, 10, 10, a=2.15, jitter=0.15, seed=1000 + rank)


def survey_flops(w, M, D):
    """algorithmic FP64 work of the whole batch from the kernel's
    counters (+,-,* = 1; FMA = 2; div, sqrt, exp, sin, cos = 1)."""
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


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gapcu
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # stdout carries exactly one JSON line (rank 0).  The image sets NCCL_DEBUG=VERSION, whose banner
        # NCCL prints to stdout and only redirects from level WARN on: same information, sent to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    os.environ["GAPCU_DEVICE"] = str(local)   # device of the Fortran-style entry points (gapcu_calc) on this rank
    cell, pos, z = workload(rank)
    natoms = len(pos)
    ctx = gapcu.Context(local)
    ctx.load_potential(POT)
    ctx.set_structures(z, cell, pos, RCUT)
    dfma, dmma = ctx.fp64_peaks() if rank == 0 else (0.0, 0.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------
    ctx.time_compute(max(args.warmup, 3), True, L2_FLUSH, stages=False)      # warm-up (also settles capacities)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    ms, stages, launches = ctx.time_compute(args.steps, True, L2_FLUSH, stages=False)
    barrier()
    # the clocks line needs the GPU under load for a few samples: keep stepping briefly
    t_end = time.time() + 1.0
    while rank == 0 and time.time() < t_end:
        ctx.time_compute(50, True, 0, stages=False)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=torch.device("cuda", local))
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = natoms * args.steps * world / (ms_max * 1e-3)
    # per-stage device times (instrumented run, rank 0) for the roofline
    _, stages, _ = ctx.time_compute(args.steps, True, L2_FLUSH, stages=True)
    work = ctx.work_counters()

    # ---- end to end through the Fortran-facing C ABI ---------------------------------
    pot_dir = os.path.join("/tmp", "gapcu_bench_rank%d" % rank)
    os.makedirs(pot_dir, exist_ok=True)
    link = os.path.join(pot_dir, "gap_parameters")
    if os.path.lexists(link):
        os.remove(link)
    os.symlink(POT, link)
    cwd = os.getcwd()
    os.chdir(pot_dir)
    import ctypes as C
    L = gapcu.lib()
    nsp, dl = C.c_int(), C.c_int()
    theta = np.zeros(100); mm = np.zeros((4000, 100), order="F"); coeff = np.zeros(4000)
    L.gapcu_read.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                             C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    assert L.gapcu_read(b"gap_parameters", C.byref(nsp), C.byref(dl), theta.ctypes.data, 100, mm.ctypes.data, 4000, 100,
                        None, 0, coeff.ctypes.data, 4000) == 0
    M, D = nsp.value, dl.value
    th = theta[:D].copy(); mmc = mm[:M, :D].copy(); co = coeff[:M].copy()
    rng = np.random.default_rng(5000 + rank)

    # the caller's buffers in the layout FGAP_CALC receives them (Fortran order), made once;
    # only the positions change from step to step
    zi = np.ascontiguousarray(z, np.int32)
    latf = np.asfortranarray(cell); mmf = np.asfortranarray(mmc)
    f_out = np.zeros((natoms, 3), order="F"); s_out = np.zeros(6)
    e_out = C.c_double(); v_out = C.c_double()

    # fresh host positions every step (an MD-like perturbation), generated before the timed region
    pos_steps = [np.asfortranarray(pos + rng.normal(0.0, 0.01, pos.shape)) for _ in range(max(args.steps, args.warmup, 3))]
    step_no = [0]

    # raw addresses taken once: a compiled host driver passes plain pointers, it does not build
    # numpy ctypes views inside its MD loop
    pos_ptrs = [p.ctypes.data for p in pos_steps]
    p_z, p_lat, p_th, p_mm, p_co = zi.ctypes.data, latf.ctypes.data, th.ctypes.data, mmf.ctypes.data, co.ctypes.data
    p_e, p_f, p_s, p_v = C.addressof(e_out), f_out.ctypes.data, s_out.ctypes.data, C.addressof(v_out)
    calc = L.gapcu_calc

    def e2e_step():
        p_pos = pos_ptrs[step_no[0] % len(pos_ptrs)]
        step_no[0] += 1
        rc = calc(natoms, p_z, p_lat, p_pos, M, D, p_th, p_mm, None, p_co, RCUT, 1, p_e, p_f, p_s, p_v)
        assert rc == 0, L.gapcu_last_error()
        return e_out.value, f_out

    for _ in range(max(args.warmup, 3)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e_last = e2e_step()[0]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    os.chdir(cwd)
    t = torch.tensor([dt], dtype=torch.float64, device=torch.device("cuda", local))
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = natoms * args.steps * world / float(t.item())
    h2d = 24 * natoms + 8 * natoms + 4 * natoms + 208   # pos + species weights + structure ids + cell record
    d2h = 24 * natoms + 512 + 256                        # one copy: flags/counters slot, (E, stress, variance) slot, forces

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel ---------------------------------------------
    w_desc, w_gpr = survey_flops(work, M, D)
    # the default pipeline for this potential is the fused centre kernel (GPR inside the
    # CTA); the tiled DMMA GPR kernel is timed separately through the split pipeline
    ctx.set_pipeline("split")
    ms_split, st_split, _ = ctx.time_compute(args.steps, True, L2_FLUSH, stages=True)
    ctx.set_pipeline("auto")
    split_gpr = {"achieved": w_gpr / (st_split["gpr_dmma"] / args.steps * 1e-3) / 1e12, "peak": dmma, "unit": "TFLOP/s",
                 "note": "k_gpr (mma.sync m8n8k4 f64) in the split pipeline; not on the default path at this M*D",
                 "split_pipeline_ms_per_step": ms_split / args.steps,
                 "split_stage_ms_per_step": {k: v / args.steps for k, v in st_split.items()}}
    t_desc = (stages["descriptor_forward"] + stages["gpr_dmma"] + stages["descriptor_backward"]) / args.steps * 1e-3
    achieved = (w_desc + w_gpr) / t_desc / 1e12
    traffic, ncu_extra = None, {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        ncu_extra = {"fp64_pipe_active_pct": tj["fp64_pipe_active_pct"], "issue_slots_busy_pct": tj["issue_slots_busy_pct"],
                     "source": tj["source"]}
    except (OSError, KeyError, ValueError):
        pass
    # the two HBM/latency-bound passes beside it: algorithmic bytes (positions, 8-byte list
    # entries written once / read once, 24-byte pair gradients, forces) over their stage times
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, KeyError, ValueError):
        hbm_peak, hbm_src = 6456.2, "fallback (B200_PROFILING.md)"
    pbar = work["pairs"] / max(work["atoms"], 1.0)
    b_k1 = natoms * (24 + 4 + 8 * pbar + 4)
    b_k5 = natoms * (8 * pbar + 24 * pbar + 24 + 72 + 8)
    t_k1 = stages["neighbor_build"] / args.steps * 1e-3
    t_k5 = stages["force_gather_reduce"] / args.steps * 1e-3
    hbm_passes = {"peak": hbm_peak, "peak_source": hbm_src, "unit": "GB/s",
                  "neighbor_build": {"bytes": b_k1, "achieved": b_k1 / t_k1 / 1e9, "frac": b_k1 / t_k1 / 1e9 / hbm_peak},
                  "force_gather_reduce": {"bytes": b_k5, "achieved": b_k5 / t_k5 / 1e9, "frac": b_k5 / t_k5 / 1e9 / hbm_peak},
                  "note": "both passes are launch/latency bound at 1000 atoms (a few hundred KB per pass); they reach "
                          "their bandwidth regime only on 10^5-atom inputs (BASELINE.md section 5)"}
    roofline = {"bound": "fp64", "kernel": "k_centre<fused> (wACSF forward + in-CTA GPR + backward, one launch per step)",
                "achieved": achieved, "peak": dfma, "unit": "TFLOP/s", "frac": achieved / dfma if dfma else None,
                "peak_source": "DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "dmma_peak_tflops": dmma, "traffic": traffic, "traffic_unit": "bytes per launch (dram read+write, ncu)",
                "ncu": ncu_extra,
                "flops_per_step": w_desc + w_gpr, "flops_desc": w_desc, "flops_gpr": w_gpr, "seconds_per_step": t_desc,
                "gpr_dmma": split_gpr, "hbm_passes": hbm_passes,
                "stage_ms_per_step": {k: v / args.steps for k, v in stages.items()}}
    # ---- CPU baseline: the reference's algorithm (dense oracle) on the same structure ---
    cpu = cpu_reference(1, sample_centres=min(natoms, args.cpu_centres), repeats=1)
    out = {"metric": METRIC, "value": value, "unit": "atom-steps/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD, "atoms_per_gpu": natoms, "potential": "synthetic 3-species, 33 SF, M=129, D=66",
                      "rcut": RCUT, "parallelism": "independent structures, one per GPU" if world > 1 else "single GPU",
                      "l2": "L2 flushed between timed steps (256 MiB memset outside the timed events)"},
           "clocks": clocks, "gpu_launches": launches,
           "e2e": {"value": e2e_value, "unit": "atom-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "api": "gapcu_calc (C ABI bound by FGAP_CALC), host buffers, ./gap_parameters side channel"},
           "roofline": roofline, "cpu_baseline": cpu, "check_energy": e_last}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


_CPU = {}


def cpu_reference(threads, sample_centres, repeats):
    """Times oracle.calc_dense (the reference algorithm loop for loop) on the C2
    structure restricted to `sample_centres` centre atoms per thread."""
    from concurrent.futures import ThreadPoolExecutor
    if not _CPU:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from oracle import Oracle
        _CPU["pot"] = Oracle("fast").read(POT)
        _CPU["w"] = workload(0)
    pot = _CPU["pot"]
    cell, pos, z = _CPU["w"]

    def one(_):
        pot.calc_dense(z, cell, pos, RCUT, True, centres=(0, sample_centres))

    t0 = time.perf_counter()
    for _ in range(repeats):
        if threads == 1:
            one(0)
        else:
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(one, range(threads)))
    dt = time.perf_counter() - t0
    return {"value": threads * sample_centres * repeats / dt, "unit": "atom-steps/s", "cores": threads, "kind": "port",
            "sample": "%d of the %d centre atoms of the C2 structure per thread, dense reference algorithm "
                      "(O(N^2) dxdy for those centres), gcc -O3 -march=native; the reference Fortran cannot be "
                      "built here (no Fortran compiler)" % (sample_centres, len(pos)),
            "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    world = int(os.environ.get("WORLD_SIZE", "1"))
    centres = 125
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
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": WORKLOAD, "atoms_per_gpu": 1000, "potential": "synthetic 3-species, 33 SF, M=129, D=66",
                      "rcut": RCUT, "parallelism": "%d host threads, independent structure copies" % threads},
           "cpu_baseline": last,
           "e2e": {"value": value, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-centres", type=int, default=1000, help="centre atoms of the cpu_baseline sample")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
