"""Short C2 run for ncu (a few passes of the pipeline; no timing claims)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "calypso-gap_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import gapcu  # noqa: E402
from structures import cubic_supercell  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dims = tuple(int(x) for x in sys.argv[2:5]) if len(sys.argv) >= 5 else (10, 10, 10)
cell, pos, z = cubic_supercell(*dims)
c = gapcu.Context(0)
c.load_potential(os.path.join(ROOT, "bench_data", "gap_parameters_c2"))
c.set_structures(z, cell, pos, 6.0)
for _ in range(steps):
    c.compute(True)
e, f, s = c.fetch()
print("E", e[0])
print("balance", c.balance())
