"""Development probe: how the device-resident rate scales with system size / batch
shape (not part of the product or the test-suite)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "calypso-gap_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import gapcu  # noqa: E402
from structures import cubic_supercell, random_candidate  # noqa: E402

POT = os.path.join(ROOT, "bench_data", "gap_parameters_c2")
c = gapcu.Context(0)
c.load_potential(POT)
for dims in [(10, 10, 10), (20, 20, 20), (30, 30, 30), (50, 50, 40)]:
    cell, pos, z = cubic_supercell(*dims)
    c.set_structures(z, cell, pos, 6.0)
    steps = 20 if len(pos) <= 8000 else 5
    ms, st, _ = c.time_compute(steps, True, 0, stages=True)
    print("N=%6d  %.3f ms/step  %.2f M atom-steps/s  stages(ms) %s" % (len(pos), ms / steps, len(pos) * steps / ms / 1e3,
          {k: round(v / steps, 3) for k, v in st.items()}), flush=True)
# C3-like batch
t = time.time()
structs = [random_candidate(3000 + i) for i in range(512)]
print("generated 512 candidates in %.1fs, atoms %d" % (time.time() - t, sum(len(s[1]) for s in structs)))
c.set_structures([s[2] for s in structs], [s[0] for s in structs], [s[1] for s in structs], 6.0)
ms, st, _ = c.time_compute(5, True, 0, stages=True)
n = sum(len(s[1]) for s in structs)
print("C3 batch 512 structures, %d atoms: %.3f ms/step  %.2f M atom-steps/s  %s" % (n, ms / 5, n * 5 / ms / 1e3,
      {k: round(v / 5, 3) for k, v in st.items()}))
print("work", c.work_counters())
