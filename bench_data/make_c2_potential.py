"""Generates bench_data/gap_parameters_c2, the synthetic 3-species potential of
BASELINE config 2.  Run once in the dev container and commit
the output: bench.py then needs no oracle to build its workload.  The descriptor
evaluation of the sibling structure uses the CPU oracle (test infrastructure)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
from oracle import Oracle  # noqa: E402
from potentials import synthetic_potential  # noqa: E402

if __name__ == "__main__":
    out = os.path.join(ROOT, "bench_data", "gap_parameters_c2")
    pot = synthetic_potential(Oracle("parity"), os.path.join(ROOT, "tests", "golden", "gap_parameters"), out)
    print(out, pot.nspecies, pot.nsf, pot.nsparse, pot.des_len)
