"""Where a centre's time goes: per-phase cycles of the centre kernel (GAPCU_VARIANT=16 makes thread 0 of
every CTA accumulate clock64 differences per phase).  Development tool."""
import os
import sys

os.environ["GAPCU_VARIANT"] = str(int(os.environ.get("GAPCU_VARIANT", "0")) | 16)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "calypso-gap_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import gapcu  # noqa: E402
from structures import cubic_supercell  # noqa: E402

c = gapcu.Context(0)
c.load_potential(os.path.join(ROOT, "bench_data", "gap_parameters_c2"))
for dims in ((30, 30, 30),):
    cell, pos, z = cubic_supercell(*dims, seed=1000)
    c.evaluate(z, cell, pos, 6.0, True)
    c.compute(True); c.fetch()
    w = c.work_counters()
    ph = w.get("phase_cycles", {})
    tot = sum(ph.values()) or 1.0
    print("N=%d: cycles per centre %.0f;" % (len(pos), tot / w["atoms"]), "  ".join("%s %.0f" % (k, v / w["atoms"]) for k, v in ph.items()))
    print("   balance", c.balance())
