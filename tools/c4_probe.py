"""Development probe for BASELINE config 4 on ONE GPU: the 100k-atom cell evaluated by a single
context (`single`) or cut into R bricks that all run on device 0 (`group R`: the same halo kernels
and phases as the NCCL transport).  Meant to be run under
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv
to get the per-kernel times of one rank's share (ncu serialises the kernels, so a brick's kernels
are timed as if it had the GPU to itself).  Not part of the product or the test-suite."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "calypso-gap_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import gapcu  # noqa: E402
from structures import cubic_supercell  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "single"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
passes = int(os.environ.get("PASSES", "2"))
dims = tuple(int(x) for x in os.environ.get("DIMS", "50,50,40").split(","))
cell, pos, z = cubic_supercell(*dims, seed=4000)
pot = os.path.join(ROOT, "bench_data", "gap_parameters_c2")
skin = float(os.environ.get("SKIN", "0"))
if mode == "single":
    c = gapcu.Context(0)
    c.load_potential(pot)
    if skin:
        c.set_skin(skin)
    c.set_structures(z, cell, pos, 6.0)
    for k in range(passes):
        if skin and k:
            c.update_positions(pos + np.random.default_rng(k).normal(0, 0.01, pos.shape), True)
        t0 = time.time()
        c.compute(True)
        e, f, s = c.fetch()
        print("pass", k, "E", e[0], "%.2f ms" % ((time.time() - t0) * 1e3), flush=True)
    ms, st, _ = c.time_compute(3, True, 0, stages=True)
    print("single: %.3f ms/step" % (ms / 3), {k: round(v / 3, 3) for k, v in st.items()})
else:
    g = gapcu.Group([0] * R)
    g.load_potential(pot)
    if skin:
        g.set_skin(skin)
    grid = gapcu.domain_grid(R, cell, 6.5 + skin)
    g.set_structure(z, cell, pos, 6.0, grid)
    for k in range(passes):
        if skin and k:
            g.update_positions(pos + np.random.default_rng(k).normal(0, 0.01, pos.shape), True)
        t0 = time.time()
        g.compute(True)
        e, f, s = g.fetch()
        print("pass", k, "grid", grid, "E", e, "%.2f ms" % ((time.time() - t0) * 1e3), flush=True)
    print("owned per brick", [len(g.ctx(r).owned()) for r in range(R)])
