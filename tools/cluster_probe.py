"""A/B of the centre kernel's CTAs-per-centre setting (GAPCU_CLUSTER=1|2|4|0=auto) on the
64-atom shipped example and on sc supercells: results vs CS=1 and stage times.  Development tool."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    for p in ("tests", "calypso-gap_b200"):
        sys.path.insert(0, os.path.join(ROOT, p))
    import gapcu
    from structures import cubic_supercell
    G = os.path.join(ROOT, "tests", "golden")
    g = np.load(os.path.join(G, "ase_traj_frames.npz"))
    out = {}
    c = gapcu.Context(0)
    c.load_potential(os.path.join(G, "gap_parameters"))
    r = c.evaluate(g["numbers"], g["cell"][0], g["positions"][0], 6.0, True)
    out["c1"] = {"relE": abs(r["energy"] - g["energy"][0]) / abs(g["energy"][0]), "dF": float(np.abs(r["forces"] - g["forces"][0]).max()),
                 "dS": float(np.abs(r["stress"] - g["stress"][0]).max()), "E": r["energy"], "F00": float(r["forces"][0, 0])}
    c.time_compute(20, True, 0, stages=False)
    ms, st, _ = c.time_compute(200, True, 0, stages=True)
    out["c1"]["ms"] = ms / 200; out["c1"]["centre_ms"] = st["descriptor_forward"] / 200
    out["c1"]["work"] = [float(x) for x in c.work_counters()[:9]] if not isinstance(c.work_counters(), dict) else {k: float(v) for k, v in c.work_counters().items()}
    for dims in ((4, 4, 4), (6, 6, 6), (10, 10, 10)):
        cell, pos, z = cubic_supercell(*dims)
        c2 = gapcu.Context(0)
        c2.load_potential(os.path.join(ROOT, "bench_data", "gap_parameters_c2"))
        r = c2.evaluate(z, cell, pos, 6.0, True)
        c2.time_compute(10, True, 0, stages=False)
        ms, st, _ = c2.time_compute(100, True, 0, stages=True)
        out["sc%d" % len(z)] = {"E": r["energy"], "Fsum": float(np.abs(r["forces"].sum(0)).max()), "F00": float(r["forces"][0, 0]), "S0": float(r["stress"][0]),
                                "ms": ms / 100, "centre_ms": st["descriptor_forward"] / 100}
    print("RESULT " + json.dumps(out))
    sys.exit(0)
res = {}
for v in sys.argv[1:] or ["1", "2", "4", "0"]:
    env = dict(os.environ); env["GAPCU_CLUSTER"] = v
    try:
        p = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True, timeout=120)
    except subprocess.TimeoutExpired:
        print("variant", v, "TIMEOUT"); continue
    got = [l for l in p.stdout.split("\n") if l.startswith("RESULT")]
    if not got:
        print("variant", v, "FAILED", p.stdout[-500:], p.stderr[-1500:]); continue
    res[v] = json.loads(got[0][7:])
base = res.get("1")
for v, r in res.items():
    for k, x in r.items():
        d = ""
        if base:
            b = base[k]
            d = " | vs CS=1: dE %.1e dF00 %.1e" % (abs(x["E"] - b["E"]) / abs(b["E"]), abs(x["F00"] - b["F00"]))
        print("CS=%s %-7s step %.4f ms centre %.4f ms  E %.10f%s %s" % (v, k, x["ms"], x["centre_ms"], x["E"], d,
              ("golden relE %.1e dF %.1e dS %.1e" % (x["relE"], x["dF"], x["dS"])) if k == "c1" else ""))
print(json.dumps(res))
