"""Size-independent properties of the CUDA path at BASELINE.json's full shapes, where the
CPU oracle is too slow to serve as the checker (configs 3 and 4): periodic replication,
permutation equivariance, Newton's third law, energy conservation of the force field
(directional finite difference), strain derivative = stress, run-to-run bit equality."""
import os

import numpy as np
import pytest

from structures import cubic_supercell, random_candidate

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POT_C2 = os.path.join(ROOT, "bench_data", "gap_parameters_c2")
GPA2EVPANG = 6.24219e-3   # gap_calc.f90:9


@pytest.fixture(scope="module")
def ctx():
    import gapcu
    c = gapcu.Context(0)
    c.load_potential(POT_C2)
    yield c
    c.close()


def test_replicated_supercell_27k_atoms(ctx):
    """A 3x3x3 replication of the config-2 cell (27,000 atoms) has 27x the energy, the same
    stress and the replicated forces of the 1000-atom cell."""
    cell, pos, z = cubic_supercell(10, 10, 10)
    small = ctx.evaluate(z, cell, pos, 6.0, True)
    shifts = np.array([[a, b, c] for a in range(3) for b in range(3) for c in range(3)], float) @ cell
    big_pos = np.concatenate([pos + s for s in shifts])
    big = ctx.evaluate(np.tile(z, 27), 3.0 * cell, big_pos, 6.0, True)
    assert abs(big["energy"] - 27.0 * small["energy"]) <= 1e-10 * abs(big["energy"])
    assert np.abs(big["forces"] - np.tile(small["forces"], (27, 1))).max() <= 1e-8
    assert np.abs(big["stress"] - small["stress"]).max() <= 1e-7
    assert np.abs(big["forces"].sum(0)).max() <= 1e-6


def test_hundred_thousand_atoms_newton_and_directional_derivative(ctx):
    """BASELINE config 4 shape (50x50x40 sites): net force zero, bit-reproducible, and
    E(x + h d) - E(x - h d) = -2h F.d for a random displacement field d."""
    cell, pos, z = cubic_supercell(50, 50, 40, seed=4000)
    r0 = ctx.evaluate(z, cell, pos, 6.0, True)
    assert np.isfinite(r0["energy"]) and np.abs(r0["forces"].sum(0)).max() <= 1e-5
    again = ctx.evaluate(z, cell, pos, 6.0, True)
    assert again["energy"] == r0["energy"] and np.array_equal(again["forces"], r0["forces"])
    rng = np.random.default_rng(5)
    d = rng.normal(size=pos.shape)
    d /= np.sqrt((d * d).sum())
    h = 1e-4
    ep = ctx.evaluate(z, cell, pos + h * d, 6.0, False)["energy"]
    em = ctx.evaluate(z, cell, pos - h * d, 6.0, False)["energy"]
    want = -(r0["forces"] * d).sum()
    got = (ep - em) / (2 * h)
    assert abs(got - want) <= 1e-6 * max(1.0, abs(want))


def test_permutation_equivariance_and_translation(ctx):
    cell, pos, z = cubic_supercell(12, 11, 10, seed=77)
    r0 = ctx.evaluate(z, cell, pos, 6.0, True)
    perm = np.random.default_rng(1).permutation(len(pos))
    r1 = ctx.evaluate(z[perm], cell, pos[perm], 6.0, True)
    assert abs(r1["energy"] - r0["energy"]) <= 1e-11 * abs(r0["energy"])
    assert np.abs(r1["forces"] - r0["forces"][perm]).max() <= 1e-8
    assert np.abs(r1["stress"] - r0["stress"]).max() <= 1e-7
    # rigid translation, atoms wrapped back into the cell (the reference's +-nabc image window
    # needs wrapped input: gap_calc.f90:86-96)
    t = pos + np.array([3.7, -1.3, 8.1])
    frac = t @ np.linalg.inv(cell)
    t = (frac - np.floor(frac)) @ cell
    r2 = ctx.evaluate(z, cell, t, 6.0, True)
    assert abs(r2["energy"] - r0["energy"]) <= 1e-10 * abs(r0["energy"])
    assert np.abs(r2["forces"] - r0["forces"]).max() <= 1e-8
    assert np.abs(r2["stress"] - r0["stress"]).max() <= 1e-7


def test_stress_is_the_strain_derivative(ctx):
    """sigma_ab = -(1/V) dE/d eps_ab / 6.24219e-3 GPa, output order xx yy zz xy yz xz
    (gap_calc.f90:189-226), checked by central differences on a sheared triclinic cell."""
    cell, pos, z = random_candidate(3207, 96, 128)
    r0 = ctx.evaluate(z, cell, pos, 6.0, True)
    vol = abs(np.linalg.det(cell))
    h = 1e-5
    comps = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2)]
    for q, (a, b) in enumerate(comps):
        eps = np.zeros((3, 3))
        eps[a, b] += 0.5
        eps[b, a] += 0.5
        ep = ctx.evaluate(z, cell @ (np.eye(3) + h * eps), pos @ (np.eye(3) + h * eps), 6.0, False)["energy"]
        em = ctx.evaluate(z, cell @ (np.eye(3) - h * eps), pos @ (np.eye(3) - h * eps), 6.0, False)["energy"]
        want = -(ep - em) / (2 * h) / vol / GPA2EVPANG
        assert abs(r0["stress"][q] - want) <= 1e-5 * max(1.0, np.abs(r0["stress"]).max())


def test_candidate_batch_512_structures(ctx):
    """BASELINE config 3 shape (subset of 512 candidates, 32-128 atoms, triclinic): one batched
    evaluation equals the per-structure evaluations (to rounding: the batch may use another
    list chunking, hence another summation order), and each structure's forces sum to zero."""
    structs = [random_candidate(3000 + i) for i in range(512)]
    ctx.set_structures([s[2] for s in structs], [s[0] for s in structs], [s[1] for s in structs], 6.0)
    ctx.compute(True)
    e, f, s = ctx.fetch()
    off = 0
    worst = 0.0
    for cell, pos, z in structs:
        worst = max(worst, np.abs(f[off:off + len(pos)].sum(0)).max())
        off += len(pos)
    assert off == len(f) and np.isfinite(e).all() and worst <= 1e-7
    off = 0
    for i in (0, 17, 255, 511):
        cell, pos, z = structs[i]
        o = sum(len(t[1]) for t in structs[:i])
        one = ctx.evaluate(z, cell, pos, 6.0, True)
        assert abs(one["energy"] - e[i]) <= 1e-13 * abs(e[i])
        assert np.abs(one["forces"] - f[o:o + len(pos)]).max() <= 1e-11
        assert np.abs(one["stress"] - s[i]).max() <= 1e-10 * max(1.0, np.abs(s[i]).max())
