"""Development probe run on the GPU box: parity numbers, FP64 peaks, stage times.
Writes gpurun_out/probe.json.  Not part of the product or the test-suite."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("oracle", "tests", "calypso-gap_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import gapcu  # noqa: E402
from oracle import Oracle  # noqa: E402
from potentials import synthetic_potential  # noqa: E402
from structures import cubic_supercell  # noqa: E402

out = {}
G = os.path.join(ROOT, "tests", "golden")
o = Oracle("parity")
pot = o.read(os.path.join(G, "gap_parameters"))
g = np.load(os.path.join(G, "ase_traj_frames.npz"))
c = gapcu.Context(0)
out["peaks_tflops"] = c.fp64_peaks()
print("FP64 peaks (DFMA, DMMA) TFLOP/s:", out["peaks_tflops"], flush=True)
c.load_potential(os.path.join(G, "gap_parameters"))
res = []
for fr in range(11):
    r = c.evaluate(g["numbers"], g["cell"][fr], g["positions"][fr], 6.0, True)
    res.append((abs(r["energy"] - g["energy"][fr]) / abs(g["energy"][fr]), np.abs(r["forces"] - g["forces"][fr]).max(),
                np.abs(r["stress"] - g["stress"][fr]).max()))
    print(fr, "relE %.2e dF %.2e dS %.2e" % res[-1], flush=True)
out["golden"] = res
want = pot.calc_sparse(g["numbers"], g["cell"][0], g["positions"][0], 6.0, True, desc=True)
xx, dedg, eat = c.descriptors(pot.des_len)
c.evaluate(g["numbers"], g["cell"][0], g["positions"][0], 6.0, True)
xx, dedg, eat = c.descriptors(pot.des_len)
print("xx rel", (np.abs(xx - want["xx"]) / (np.abs(want["xx"]).max(0) + 1e-300)).max(), "eatom", np.abs(eat - want["eatom"]).max(),
      "dedg", np.abs(dedg - want["dedg"]).max() / np.abs(want["dedg"]).max(), flush=True)
print("work", c.work_counters())

tmp = "/tmp/gap_parameters_c2"
pot2 = synthetic_potential(o, os.path.join(G, "gap_parameters"), tmp)
cell, pos, z = cubic_supercell(10, 10, 10)
t = time.time(); want = pot2.calc_sparse(z, cell, pos, 6.0, True); t_cpu = time.time() - t
c2 = gapcu.Context(0)
c2.load_potential(tmp)
r = c2.evaluate(z, cell, pos, 6.0, True)
print("C2 relE %.2e dF %.2e dS %.2e  (oracle sparse %.1fs)" % (abs(r["energy"] - want["energy"]) / abs(want["energy"]),
      np.abs(r["forces"] - want["forces"]).max(), np.abs(r["stress"] - want["stress"]).max(), t_cpu), flush=True)
ms, stages, launches = c2.time_compute(20, True, 256 << 20)
print("C2 device ms/step %.4f launches/step %.1f" % (ms / 20, launches / 20), {k: v / 20 for k, v in stages.items()})
wk = c2.work_counters()
print("work", wk)
out["c2"] = {"ms_per_step": ms / 20, "stages_ms": {k: v / 20 for k, v in stages.items()}, "work": wk}
t = time.time()
for _ in range(20):
    c2.evaluate(z, cell, pos, 6.0, True)
out["c2"]["e2e_ms"] = (time.time() - t) / 20 * 1e3
print("C2 e2e ms/step", out["c2"]["e2e_ms"])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1, default=float)
