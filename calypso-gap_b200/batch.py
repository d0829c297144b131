"""Sharding of a batch of independent structures over ranks (BASELINE
config 3).  Structures share only the read-only potential, so there is no data-path
collective: every rank evaluates its shard on its own GPU and the results are
gathered at the end.  Host logic only; the evaluator is passed in (on a GPU box it
is `gapcu.Context` batch evaluation, see `gpu_evaluator`)."""
import math

import numpy as np


def estimate_cost(natoms, volume, rcut=6.0):
    """Relative cost of one structure: N * P^2 with P the mean neighbour count
    (the triplet work dominates and grows with the square of the pair count)."""
    p = 4.0 / 3.0 * math.pi * rcut ** 3 * natoms / max(volume, 1e-30)
    return natoms * p * p


def partition(costs, world):
    """Greedy longest-processing-time assignment.  Returns `world` index lists, each
    in ascending structure order; deterministic (ties by index)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    shards = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += costs[i]
    return [sorted(s) for s in shards]


def evaluate_sharded(structures, evaluate_local, rank=0, world=1, all_gather_object=None, rcut=6.0):
    """structures: list of (cell, pos, z).  evaluate_local(list_of_structures) ->
    list of dict(energy, forces, stress).  Returns the full result list in the input
    order on every rank (all_gather_object is torch.distributed.all_gather_object or
    a stand-in with the same signature; None means single rank)."""
    costs = [estimate_cost(len(p), abs(np.linalg.det(np.asarray(c, float))), rcut) for c, p, _ in structures]
    shards = partition(costs, world)
    mine = shards[rank]
    local = evaluate_local([structures[i] for i in mine]) if mine else []
    if world == 1 or all_gather_object is None:
        gathered = [list(zip(mine, local))]
    else:
        gathered = [None] * world
        all_gather_object(gathered, list(zip(mine, local)))
    out = [None] * len(structures)
    for part in gathered:
        for i, r in part:
            out[i] = r
    return out


def gpu_evaluator(ctx, rcut=6.0, lgrad=True):
    """evaluate_local backed by one gapcu.Context: the whole shard is one batched
    launch sequence (one CTA per centre atom across all structures)."""
    def run(structs):
        ctx.set_structures([s[2] for s in structs], [s[0] for s in structs], [s[1] for s in structs], rcut)
        ctx.compute(lgrad)
        e, f, s = ctx.fetch()
        out, off = [], 0
        for k, (_, pos, _) in enumerate(structs):
            n = len(pos)
            out.append({"energy": float(e[k]), "forces": f[off:off + n].copy(), "stress": s[k].copy()})
            off += n
        return out
    return run
