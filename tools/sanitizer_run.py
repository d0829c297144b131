"""Small workload for compute-sanitizer (memcheck / racecheck): the 64-atom shipped example through
the fused centre kernel with 1, 2 and 4 CTAs per centre.  Usage on the GPU box:
  compute-sanitizer --tool racecheck python tools/sanitizer_run.py
Development tool (round 1: 0 hazards, 0 memcheck errors)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "calypso-gap_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import gapcu
G = os.path.join(ROOT, "tests", "golden")
g = np.load(os.path.join(G, "ase_traj_frames.npz"))
for cs in (1, 2, 4):
    c = gapcu.Context(0)
    c.set_cluster(cs)
    c.load_potential(os.path.join(G, "gap_parameters"))
    r = c.evaluate(g["numbers"][:64], g["cell"][0], g["positions"][0], 6.0, True)
    print(cs, r["energy"])
    c.close()
