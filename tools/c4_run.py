"""BASELINE config 4: one large supercell, spatial decomposition over the ranks of a node.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
        --master-port P tools/c4_run.py [nx ny nz] [--check]
torch.distributed (NCCL backend) only carries the NCCL unique id and the timing reduction;
the ghost-force return runs inside libgapcu on its own NCCL communicator."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "calypso-gap_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import gapcu  # noqa: E402
from structures import cubic_supercell  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
dims = tuple(int(x) for x in args[:3]) if len(args) >= 3 else (50, 50, 40)
check = "--check" in sys.argv
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cell, pos, z = cubic_supercell(*dims, seed=4000)
ctx = gapcu.Context(local)
ctx.load_potential(os.path.join(ROOT, "bench_data", "gap_parameters_c2"))
if world > 1:
    obj = [gapcu.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    ctx.nccl_init(world, rank, obj[0])
    grid = gapcu.domain_grid(world, cell)
    ctx.set_domain(grid, gapcu.brick_of(rank, grid))
else:
    grid = (1, 1, 1)
ctx.set_structures(z, cell, pos, 6.0)
ctx.compute(True)
e, f, s = ctx.fetch()
steps = 5
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ms, stages, _ = ctx.time_compute(steps, True, 0, stages=True)
t = torch.tensor([ms], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
out = {"config": "C4", "atoms": len(pos), "n_gpus": world, "grid": list(grid), "ms_per_step": float(t.item()) / steps,
       "atom_steps_per_s": len(pos) * steps / (float(t.item()) * 1e-3), "energy": float(e[0]),
       "stage_ms_rank0": {k: v / steps for k, v in stages.items()}}
if check:
    # compare with the single-GPU evaluation of the same structure on this rank's device
    ref = gapcu.Context(local)
    ref.load_potential(os.path.join(ROOT, "bench_data", "gap_parameters_c2"))
    r = ref.evaluate(z, cell, pos, 6.0, True)
    out["rel_dE_vs_1gpu"] = abs(float(e[0]) - r["energy"]) / abs(r["energy"])
    out["max_dF_vs_1gpu"] = float(np.abs(f - r["forces"]).max())
    out["max_dS_vs_1gpu"] = float(np.abs(s[0] - r["stress"]).max())
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
