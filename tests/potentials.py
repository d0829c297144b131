"""Synthetic potentials for the BASELINE configs.  Uses the
oracle only to evaluate descriptors of the sibling structure (test/bench
infrastructure; the product never imports this)."""
import numpy as np

from structures import cubic_supercell, write_gap_parameters


def synthetic_potential(oracle, shipped_path, out_path, weights=((5, -1.0), (6, 4.0), (7, 2.0)), nsparse=129,
                        sf=None):
    """C2's potential: 3 species, the shipped 33-row SF table (or `sf` =
    (ntype, alpha, cutoff)), sparse points = descriptors of `nsparse` atoms of a
    seed-1001 sibling structure, theta = max(std,1e-3)*sqrt(D), coeff from a
    ridge solve so it has the large alternating coefficients of a real fit.
    Written in the reference text format and re-read, so everybody sees the same
    rounded values.  Returns the oracle Potential of the written file."""
    shipped = oracle.read(shipped_path)
    ntype, alpha, cutoff = sf if sf is not None else (shipped.ntype, shipped.alpha, shipped.cutoff)
    nsf = len(ntype)
    D = 2 * nsf
    z = np.array([a for a, _ in weights], np.int32)
    w = np.array([b for _, b in weights], float)
    # descriptors do not depend on the GPR part: evaluate them with a dummy one
    dummy = oracle.make(z, w, ntype, alpha, cutoff, np.ones(D), np.zeros((1, D)), np.zeros(1))
    need = nsparse
    rows = []
    seed = 1001
    while need > 0:
        cell, pos, zz = cubic_supercell(6, 6, 6, seed=seed, species=tuple(z), probs=(0.3, 0.4, 0.3)[:len(z)] if len(z) == 3 else None)
        xx = dummy.calc_sparse(zz, cell, pos, 6.0, False, desc=True)["xx"]
        rng = np.random.default_rng(7 + seed)
        take = rng.choice(len(xx), size=min(need, len(xx)), replace=False)
        rows.append(xx[take])
        need -= len(take)
        seed += 1
    mm = np.vstack(rows)
    theta = np.maximum(mm.std(0), 1e-3) * np.sqrt(D)
    d2 = (((mm[:, None, :] - mm[None, :, :]) / theta) ** 2).sum(-1)
    K = np.exp(-0.5 * d2)
    y = -8.0 + 0.5 * np.random.default_rng(8).normal(size=len(mm))
    coeff = np.linalg.solve(K + 1e-8 * np.eye(len(mm)), y)
    write_gap_parameters(out_path, z, w, ntype, alpha, cutoff, theta, mm, coeff)
    return oracle.read(out_path)
