"""Host-side phase times of the drop-in call: GAPCU_TRACE=1 python tools/trace_run.py prints, per
gapcu_calc call on the C2 structure (or the supercell given as three site counts, e.g. 50 50 40 = C4), the microseconds spent before the first launch, enqueuing, and
waiting for the results (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "calypso-gap_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, gapcu
from structures import cubic_supercell
import shutil, tempfile
d = tempfile.mkdtemp(); shutil.copy(os.path.join(ROOT, "bench_data", "gap_parameters_c2"), os.path.join(d, "gap_parameters")); os.chdir(d)
dims = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (10, 10, 10)
cell, pos, z = cubic_supercell(*dims)
from libgap import GAP
calc = GAP.Calculator(rcut=6.0)
for i in range(12):
    calc.gap_calc(z, cell, pos, True)
