"""Development probe for BASELINE config 5: wide descriptor (nsf=128, D=256) and a large
sparse set (M=10,000).  The sparse points are descriptors of sibling structures computed on
the GPU itself; parity of the GPR stage is checked against a float64 numpy evaluation of
GET_COV / dE/dG on a subset of atoms.  Not part of the product or the test-suite."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "calypso-gap_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import gapcu  # noqa: E402
from structures import cubic_supercell  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
# 32 type-1, 32 type-3, 48 type-2 and 16 type-4 over Rc in {3,4,5,6}
ntype, alpha, cut = [], [], []
for a in np.geomspace(1e-3, 2.0, 32): ntype.append(1); alpha.append(a); cut.append(6.0)
for rs in np.linspace(0.5, 5.5, 32): ntype.append(3); alpha.append(rs); cut.append(6.0)
for rc in (3.0, 4.0, 5.0, 6.0):
    for a in np.geomspace(2e-3, 0.3, 12): ntype.append(2); alpha.append(a); cut.append(rc)
    for a in np.geomspace(2e-3, 0.3, 12)[::3]: ntype.append(4); alpha.append(a); cut.append(rc)
ntype = np.array(ntype, np.int32); alpha = np.round(np.array(alpha), 5); cut = np.array(cut)
nsf = len(ntype); D = 2 * nsf
z3 = np.array([5, 6, 7], np.int32); w3 = np.array([-1.0, 4.0, 2.0])
c = gapcu.Context(0)
# descriptors do not depend on the GPR part: first a dummy GPR to harvest sparse points
c.set_potential(z3, w3, ntype, alpha, cut, np.ones(D), np.zeros((16, D)), np.zeros(16))
rows = []
seed = 2001
t0 = time.time()
while sum(len(r) for r in rows) < M:
    cell, pos, z = cubic_supercell(10, 10, 10, seed=seed); seed += 1
    c.evaluate(z, cell, pos, 6.0, False)
    rows.append(c.descriptors(D)[0])
mm = np.vstack(rows)[:M]
print("harvested %d sparse points, D=%d in %.1fs" % (len(mm), D, time.time() - t0), flush=True)
theta = np.maximum(mm.std(0), 1e-3) * np.sqrt(D)
rng = np.random.default_rng(8)
coeff = rng.normal(size=M) * 50.0
c.set_potential(z3, w3, ntype, alpha, cut, theta, mm, coeff)
cell, pos, z = cubic_supercell(10, 10, 10)
r = c.evaluate(z, cell, pos, 6.0, True)
xx, dedg, eat = c.descriptors(D)
# numpy reference of the GPR stage on 24 atoms (gap_calc.f90:268-288, 152-166)
sub = np.arange(0, 1000, 42)
q = (xx[sub][:, None, :] - mm[None, :, :]) / theta
K = np.exp(-0.5 * (q * q).sum(-1))
e_ref = K @ coeff
d_ref = -np.einsum("imk,im->ik", q / theta, K * coeff)
print("GPR parity on %d atoms: max|de| %.2e (|e|~%.2e)  max rel ddedg %.2e" % (len(sub), np.abs(eat[sub] - e_ref).max(),
      np.abs(e_ref).max(), np.abs(dedg[sub] - d_ref).max() / np.abs(d_ref).max()), flush=True)
ms, st, _ = c.time_compute(5, True, 0, stages=True)
flops = 1000 * (4.0 * M * D + 4 * M + 3 * D)
print("C5 (N=1000, M=%d, D=%d): %.3f ms/step; stages %s; GPR %.2f TFLOP/s" % (M, D, ms / 5, {k: round(v / 5, 3) for k, v in st.items()},
      flops / (st["gpr_dmma"] / 5 * 1e-3) / 1e12))
print("peaks", c.fp64_peaks())
