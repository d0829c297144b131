"""Generates BASELINE config 3's random candidate structures (seeds 3000+i) on all host cores and
pickles them: python tools/gen_c3.py N OUT.pkl.  A separate process on purpose: the callers (tests,
bench.py) hold a CUDA context and threads, and forking a pool from such a process can deadlock."""
import os
import pickle
import sys
from multiprocessing import Pool

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from structures import random_candidate  # noqa: E402


def make(i):
    return random_candidate(3000 + i)


if __name__ == "__main__":
    n, out = int(sys.argv[1]), sys.argv[2]
    with Pool(min(os.cpu_count() or 1, 64)) as pool:
        structs = pool.map(make, range(n), chunksize=8)
    with open(out + ".tmp", "wb") as fh:
        pickle.dump(structs, fh)
    os.replace(out + ".tmp", out)
