"""BASELINE config 1 latency probe (64-atom shipped example): per-call time of the drop-in
entry points (C ABI gapcu_calc, f2py fgap_calc, libgap.GAP.Calculator) and of a persistent
context, with the device stage times.  Development tool; prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "calypso-gap_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import gapcu  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
g = np.load(os.path.join(G, "ase_traj_frames.npz"))
z, cell, pos = g["numbers"], g["cell"][0], g["positions"][0]
out = {"config": "C1", "atoms": len(z)}
os.chdir(G)   # ./gap_parameters side channel of the reference API
c = gapcu.Context(0)
c.load_potential("gap_parameters")
c.set_structures(z, cell, pos, 6.0)
c.time_compute(20, True, 0, stages=False)
ms, st, _ = c.time_compute(200, True, 0, stages=True)
out["device_ms_per_step"] = ms / 200
out["stage_ms"] = {k: v / 200 for k, v in st.items()}
out["balance"] = {k: float(v) for k, v in c.balance().items()}
n = 200
for _ in range(20):
    c.evaluate(z, cell, pos, 6.0, True)
t = time.perf_counter()
for _ in range(n):
    c.evaluate(z, cell, pos, 6.0, True)
out["persistent_context_ms_per_call"] = 1e3 * (time.perf_counter() - t) / n
from libgap import GAP  # noqa: E402
calc = GAP.Calculator(rcut=6.0)
sym = ["C"] * len(z)
for _ in range(5):
    calc.gap_calc(sym, cell, pos, True)
t = time.perf_counter()
for _ in range(n):
    calc.gap_calc(sym, cell, pos, True)
out["libgap_Calculator_gap_calc_ms_per_call"] = 1e3 * (time.perf_counter() - t) / n
t = time.perf_counter()
for _ in range(20):
    GAP.Calculator(rcut=6.0).gap_calc(sym, cell, pos, True)
out["fresh_Calculator_per_step_ms_per_call"] = 1e3 * (time.perf_counter() - t) / 20
for k in ("device_ms_per_step", "persistent_context_ms_per_call", "libgap_Calculator_gap_calc_ms_per_call"):
    out[k.replace("ms_per_step", "atom_steps_per_s").replace("ms_per_call", "atom_steps_per_s")] = len(z) / (out[k] * 1e-3)
print(json.dumps(out))
