"""A/B timing of environment switches on two workloads (C2: 1000 atoms; a 30x30x30 = 27000-atom cell):
   python tools/ab_env.py "" "GAPCU_NO_SHARE_EXP=1" "GAPCU_LIB=path/to/other/libgapcu.so"
every argument is a space-separated list of VAR=VALUE settings for one variant ("" = defaults); the
variants run interleaved, three repeats, median reported, and their energies are compared.  Development tool."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    for p in ("tests", "calypso-gap_b200"):
        sys.path.insert(0, os.path.join(ROOT, p))
    import gapcu
    from structures import cubic_supercell
    c = gapcu.Context(0)
    c.load_potential(os.path.join(ROOT, "bench_data", "gap_parameters_c2"))
    for dims, steps in (((10, 10, 10), 100), ((30, 30, 30), 10)):
        cell, pos, z = cubic_supercell(*dims, seed=1000)
        r = c.evaluate(z, cell, pos, 6.0, True)
        c.time_compute(3, True, 0, stages=False)
        ms, st, _ = c.time_compute(steps, True, 0, stages=True)
        print("RESULT %d %.5f %.5f %.5f %.5f %.12e %.6e" % (len(pos), ms / steps, st["descriptor_forward"] / steps, st["neighbor_build"] / steps,
                                                      st["force_gather_reduce"] / steps, r["energy"], abs(r["forces"]).max()))
    sys.exit(0)
variants = sys.argv[1:] or [""]
res = {v: {} for v in variants}
for rep in range(3):
    for v in variants:
        env = dict(os.environ)
        for kv in v.split():
            k, val = kv.split("=", 1)
            env[k] = os.path.join(ROOT, val) if k == "GAPCU_LIB" and not os.path.isabs(val) else val
        out = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
        if "RESULT" not in out.stdout:
            print("variant %r failed:\n%s" % (v, out.stderr[-2000:]))
        for line in out.stdout.split("\n"):
            if line.startswith("RESULT"):
                f = line.split()
                res[v].setdefault(int(f[1]), []).append(tuple(float(x) for x in f[2:]))
for v in variants:
    for n, rows in sorted(res[v].items()):
        med = lambda k: sorted(r[k] for r in rows)[len(rows) // 2]
        print("%-40r N=%6d: step %.4f ms  centre %.4f  neighbours %.4f  gather %.4f   E %.12e  maxF %.6e" %
              (v, n, med(0), med(1), med(2), med(3), rows[0][4], rows[0][5]))
