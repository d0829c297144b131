"""A/B timing of kernel variants (GAPCU_VARIANT) on the C2 workload: several repeats of
time_compute per variant, interleaved, median reported.  Development tool."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    for p in ("tests", "calypso-gap_b200"):
        sys.path.insert(0, os.path.join(ROOT, p))
    import gapcu
    from structures import cubic_supercell
    cell, pos, z = cubic_supercell(10, 10, 10)
    c = gapcu.Context(0)
    c.load_potential(os.path.join(ROOT, "bench_data", "gap_parameters_c2"))
    c.set_structures(z, cell, pos, 6.0)
    c.time_compute(10, True, 0, stages=False)
    ms, st, _ = c.time_compute(100, True, 0, stages=True)
    print("RESULT %.4f %.4f %.4f %.4f" % (ms / 100, st["descriptor_forward"] / 100, st["neighbor_build"] / 100, st["force_gather_reduce"] / 100))
    sys.exit(0)
variants = sys.argv[1:] or ["0", "1"]
res = {v: [] for v in variants}
for rep in range(3):
    for v in variants:
        env = dict(os.environ)
        if v.startswith("lib:"):
            env["GAPCU_LIB"] = os.path.join(ROOT, v[4:]); env["GAPCU_VARIANT"] = "0"
        else:
            env["GAPCU_VARIANT"] = v
        out = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True).stdout
        for line in out.split("\n"):
            if line.startswith("RESULT"):
                res[v].append(tuple(float(x) for x in line.split()[1:]))
for v in variants:
    steps = sorted(r[0] for r in res[v]); cen = sorted(r[1] for r in res[v])
    nb = sorted(r[2] for r in res[v]); ga = sorted(r[3] for r in res[v])
    print("variant %s: step ms median %.4f (min %.4f)  centre kernel ms median %.4f (min %.4f)  neighbours %.4f  gather %.4f" % (v, steps[len(steps) // 2], steps[0], cen[len(cen) // 2], cen[0], nb[len(nb) // 2], ga[len(ga) // 2]))
