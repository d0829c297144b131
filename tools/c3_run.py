"""BASELINE config 3: a CALYPSO-style batch of random candidate structures (32-128 atoms,
triclinic, seeds 3000+i), sharded over the ranks of a node with no data-path collective.
Launch: python tools/c3_run.py [nstruct] [--cpu K] [--check K]            (1 GPU)
        python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
               --master-port P tools/c3_run.py [nstruct] ...                 (N GPUs)
Prints one JSON line: device-resident atom-steps/s (CUDA events, max over ranks), end-to-end
atom-steps/s (host buffers in, host results out, every step), optionally the reference
algorithm's CPU time on the first K structures (--cpu K; oracle, test infrastructure) and a
parity check of the first K structures against the oracle (--check K)."""
import json
import os
import pickle
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "calypso-gap_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
from structures import random_candidate  # noqa: E402


def _make(i):
    return random_candidate(3000 + i)


def opt(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def main():
    import torch
    import torch.distributed as dist
    import batch
    import gapcu

    pos_args = [a for k, a in enumerate(sys.argv[1:], 1) if not a.startswith("--") and not sys.argv[k - 1].startswith("--")]
    nstruct = int(pos_args[0]) if pos_args else 4096
    n_cpu, n_check, steps = opt("--cpu", 0), opt("--check", 0), opt("--steps", 3)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # rank 0 generates the batch on all host cores and hands it to the other ranks of the node
    # (the partition needs every cell and atom count)
    t0 = time.time()
    share = "/dev/shm/c3_structs_%d_%s.pkl" % (nstruct, os.environ.get("MASTER_PORT", "0"))
    if rank == 0:
        with Pool(min(os.cpu_count() or 1, 64)) as pool:
            structs = pool.map(_make, range(nstruct), chunksize=8)
        if world > 1:
            with open(share + ".tmp", "wb") as fh:
                pickle.dump(structs, fh)
            os.replace(share + ".tmp", share)
    if world > 1:
        dist.barrier()
        if rank:
            with open(share, "rb") as fh:
                structs = pickle.load(fh)
        dist.barrier()
        if rank == 0:
            os.remove(share)
    t_gen = time.time() - t0
    costs = [batch.estimate_cost(len(p), abs(np.linalg.det(c))) for c, p, _ in structs]
    mine = batch.partition(costs, world)[rank]
    shard = [structs[i] for i in mine]
    n_atoms_total = sum(len(p) for _, p, _ in structs)
    n_atoms_mine = sum(len(p) for _, p, _ in shard)
    potfile = os.path.join(ROOT, "bench_data", "gap_parameters_c2")
    ctx = gapcu.Context(local)
    ctx.load_potential(potfile)
    zs, cells, poss = [s[2] for s in shard], [s[0] for s in shard], [s[1] for s in shard]
    ctx.set_structures(zs, cells, poss, 6.0)
    ctx.compute(True)
    e, f, s = ctx.fetch()

    def sync_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms, stages, _ = ctx.time_compute(steps, True, 0, stages=True)
    ms = sync_max(ms)
    # end to end: host buffers -> H2D -> kernels -> D2H of E, F, stress, every step
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # (C ABI called on pre-packed host arrays, as a compiled host driver would)
    L = gapcu.lib()
    natoms = np.array([len(p) for p in poss], np.int32)
    species = np.ascontiguousarray(np.concatenate(zs), np.int32)
    lat = np.ascontiguousarray(np.stack(cells)); pos = np.ascontiguousarray(np.concatenate(poss))
    ene = np.zeros(len(shard)); force = np.zeros((n_atoms_mine, 3)); stress = np.zeros((len(shard), 6))
    t0 = time.perf_counter()
    for _ in range(steps):
        rc = L.gapcu_ctx_set_structures(ctx.h, len(natoms), natoms, species, lat, pos, 6.0)
        rc = rc or L.gapcu_ctx_compute(ctx.h, 1)
        rc = rc or L.gapcu_ctx_fetch(ctx.h, ene.ctypes.data, force.ctypes.data, stress.ctypes.data)
        assert rc == 0, gapcu.lib().gapcu_last_error()
    t_e2e = sync_max(time.perf_counter() - t0)
    assert np.array_equal(ene, e)
    out = {"config": "C3", "structures": nstruct, "atoms": n_atoms_total, "n_gpus": world, "steps": steps,
           "ms_per_step": ms / steps, "atom_steps_per_s": n_atoms_total * steps / (ms * 1e-3),
           "e2e_atom_steps_per_s": n_atoms_total * steps / t_e2e, "e2e_ms_per_step": 1e3 * t_e2e / steps,
           "h2d_bytes_per_step_rank0": int(n_atoms_mine * 28 + len(shard) * 76), "d2h_bytes_per_step_rank0": int(n_atoms_mine * 24 + len(shard) * 64),
           "atoms_rank0": n_atoms_mine, "structures_rank0": len(shard), "generate_s": t_gen,
           "stage_ms_rank0": {k: v / steps for k, v in stages.items()}, "balance_rank0": {k: float(v) for k, v in ctx.balance().items()},
           "sum_energy_rank0": float(np.sum(e))}
    if rank == 0 and (n_cpu or n_check):
        from oracle import Oracle
        k = max(n_cpu, n_check)
        sub = [structs[i] for i in range(k)]
        if n_check:
            pot = Oracle("parity").read(potfile)
            c1 = gapcu.Context(local)
            c1.load_potential(potfile)
            worst = [0.0, 0.0, 0.0]
            c1.set_structures([x[2] for x in sub[:n_check]], [x[0] for x in sub[:n_check]], [x[1] for x in sub[:n_check]], 6.0)
            c1.compute(True)
            ge, gf, gs = c1.fetch()
            off = 0
            for j, (cell, pos, z) in enumerate(sub[:n_check]):
                want = pot.calc_sparse(z, cell, pos, 6.0, True)
                n = len(pos)
                worst[0] = max(worst[0], abs(ge[j] - want["energy"]) / abs(want["energy"]))
                worst[1] = max(worst[1], float(np.abs(gf[off:off + n] - want["forces"]).max()))
                worst[2] = max(worst[2], float(np.abs(gs[j] - want["stress"]).max()))
                off += n
            out["parity_first_%d" % n_check] = {"rel_dE": worst[0], "max_dF": worst[1], "max_dS_GPa": worst[2]}
        if n_cpu:
            fast = Oracle("fast").read(potfile)
            t0 = time.perf_counter()
            na = 0
            for cell, pos, z in sub[:n_cpu]:
                fast.calc_dense(z, cell, pos, 6.0, True)
                na += len(pos)
            dt = time.perf_counter() - t0
            out["cpu_reference_algorithm"] = {"structures": n_cpu, "atoms": na, "seconds": dt, "atom_steps_per_s": na / dt, "cores": 1,
                                              "kind": "port (dense oracle, gcc -O3 -march=native)"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
