"""world_size-2 gloo test (CPU) of the batch sharding: structures are partitioned by
estimated cost, each rank evaluates its shard, results come back in input order and
equal the serial evaluation.  The evaluator here is the CPU oracle (this is a test of
the host-side plumbing; on the GPU box bench.py/gpu tests use the CUDA evaluator)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in ("oracle", "tests", "calypso-gap_b200"):
        sys.path.insert(0, os.path.join(ROOT, p))
    import torch.distributed as dist
    from batch import evaluate_sharded
    from oracle import Oracle
    from structures import random_candidate
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pot = Oracle("parity").read(os.path.join(ROOT, "tests", "golden", "gap_parameters"))
    structs = [random_candidate(3200 + i, 8, 20, species=(5, 6)) for i in range(7)]
    calls = []

    def ev(lst):
        calls.append(len(lst))
        return [pot.calc_sparse(z, c, p, 6.0, True) for c, p, z in lst]

    res = evaluate_sharded(structs, ev, rank, world, dist.all_gather_object)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, calls, [(r["energy"], r["forces"].sum(), r["stress"].tolist()) for r in res]))


def test_partition_is_balanced_and_deterministic():
    from batch import estimate_cost, partition
    costs = [estimate_cost(n, n * 10.0) for n in (32, 128, 64, 96, 33, 127, 80, 50)]
    a, b = partition(costs, 3), partition(costs, 3)
    assert a == b and sorted(sum(a, [])) == list(range(8))
    loads = [sum(costs[i] for i in s) for s in a]
    assert max(loads) <= 1.4 * (sum(costs) / 3)
    assert partition(costs, 1) == [list(range(8))]
    assert all(len(s) <= 1 for s in partition(costs[:2], 4))


@pytest.mark.timeout(600)
def test_two_ranks_gloo_equal_serial(shipped_pot):
    import torch.multiprocessing as mp
    from structures import random_candidate
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    structs = [random_candidate(3200 + i, 8, 20, species=(5, 6)) for i in range(7)]
    want = [shipped_pot.calc_sparse(z, c, p, 6.0, True) for c, p, z in structs]
    assert sum(sum(calls) for _, calls, _ in got) == len(structs)      # every structure evaluated exactly once
    for _, _, res in got:                                               # both ranks hold the full, ordered result
        for (e, fs, s), w in zip(res, want):
            assert e == w["energy"] and s == w["stress"].tolist()
