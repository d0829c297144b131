"""Triplet-set exactness (BASELINE north_star (1)): the neighbour PAIRS
{j, k} the angular symmetry functions of a centre are summed over must be the reference's, i.e. what the
loops of wacsf.f90:177-244 keep -- k_neighbor > j_neighbor in list order and the three tests
rij, rik, rjk .gt. cutoff -> cycle.  The kernel's own list (gapcu_ctx_debug_triplets: what its forward and
backward passes consume) is compared set for set with the oracle's export of those loops, per centre and
per cutoff class, including a lattice whose j-k distances sit exactly on a cutoff."""
import os

import numpy as np
import pytest

from structures import cubic_supercell, random_candidate, sheared

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _classes(pot):
    """Distinct SF cutoffs in descending order (the kernel's classes) and which of them hold angular functions."""
    rcs = sorted({float(c) for t, c in zip(pot.ntype, pot.cutoff) if 1 <= t <= 4}, reverse=True)
    ang = [any(t in (2, 4) and float(c) == rc for t, c in zip(pot.ntype, pot.cutoff)) for rc in rcs]
    return rcs, ang


def _tie_lattice():
    # simple cubic, a = 2: j-k distances of exactly 6.0 (three steps, and (2,2,1) steps) and 4.0
    cell = np.eye(3) * 4.0
    pos = np.array([[0, 0, 0], [2, 0, 0], [0, 2, 0], [2, 2, 0], [0, 0, 2], [2, 0, 2], [0, 2, 2], [2, 2, 2.0]])
    return cell, pos, np.array([6, 5, 6, 6, 5, 6, 6, 5], np.int32)


def _cases(bc):
    cell, pos = sheared(bc["cell"], bc["positions"])
    yield "sheared B8C56", cell, pos, bc["numbers"].astype(np.int32), range(0, 64, 7)
    cell, pos, z = random_candidate(3001, 32, 64, species=(5, 6))          # interplanar spacing < rcut: self images
    yield "self images", cell, pos, z, range(0, len(pos), 5)
    cell, pos, z = _tie_lattice()
    yield "tie lattice", cell, pos, z, range(8)


def test_triplet_sets_equal_the_reference_loops(oracle, shipped_pot, bc_structure):
    import gapcu
    rcs, ang = _classes(shipped_pot)
    c = gapcu.Context(0)
    c.load_potential(os.path.join(GOLDEN, "gap_parameters"))
    for name, cell, pos, z, centres in _cases(bc_structure):
        c.evaluate(z, cell, pos, 6.0, True)
        trip = c.triplets(cap=1 << 16 if name == "tie lattice" else 1 << 15)
        work = c.work_counters()
        tot_kept = tot_tc = 0
        for i in range(len(pos)):
            it = trip[i]
            assert (it[:, 0] < it[:, 1]).all(), name                         # slot_j < slot_k: every pair once
            assert len({(a, b) for a, b, _ in it}) == len(it), name
            tot_kept += len(it)
            tot_tc += sum(int((it[:, 2] > k).sum()) for k in range(len(rcs)) if ang[k])
        # the kernel's own counters (roofline inputs) describe the same lists
        assert work["triplets"] == tot_kept and work["triplet_classes"] == tot_tc, name
        for i in centres:
            it = trip[i]
            for k, rc in enumerate(rcs):
                if not ang[k]:
                    continue
                got = {(int(a), int(b)) for a, b, n in it if n > k}
                want = {(int(a), int(b)) for a, b in oracle.triplets(cell, pos, 6.0, i, rc, cap=1 << 16)}
                assert got == want, "%s: centre %d, cutoff %g: %d vs %d pairs" % (name, i, rc, len(got), len(want))
    c.close()


@pytest.mark.parametrize("pipeline", ["fused", "split"])
def test_tie_lattice_energy_forces_stress(shipped_pot, pipeline):
    """E/F/stress on the lattice whose pair and triplet distances sit exactly on the cutoffs: one pair or
    triplet more or less than the reference would show up far above the gates."""
    import gapcu
    cell, pos, z = _tie_lattice()
    want = shipped_pot.calc_sparse(z, cell, pos, 6.0, True)
    c = gapcu.Context(0)
    c.set_pipeline(pipeline)
    c.load_potential(os.path.join(GOLDEN, "gap_parameters"))
    got = c.evaluate(z, cell, pos, 6.0, True)
    assert abs(got["energy"] - want["energy"]) <= 1e-10 * abs(want["energy"])
    assert np.abs(got["forces"] - want["forces"]).max() <= 1e-8
    assert np.abs(got["stress"] - want["stress"]).max() <= 1e-7
    # run-to-run: same bits (no atomics anywhere on this path; ties make that worth checking)
    for _ in range(3):
        again = c.evaluate(z, cell, pos, 6.0, True)
        assert again["energy"] == got["energy"] and np.array_equal(again["forces"], got["forces"]) and np.array_equal(again["stress"], got["stress"])
    # the same lattice sheared and jittered by one ulp-scale amount: still the oracle's answer
    cell2, pos2 = sheared(cell, pos, ((1, 1e-13, 0), (0, 1, 0), (0, 0, 1)))
    want2 = shipped_pot.calc_sparse(z, cell2, pos2, 6.0, True)
    got2 = c.evaluate(z, cell2, pos2, 6.0, True)
    assert abs(got2["energy"] - want2["energy"]) <= 1e-10 * abs(want2["energy"])
    assert np.abs(got2["forces"] - want2["forces"]).max() <= 1e-8
    c.close()


def test_reproducible_at_the_1024_neighbour_tier(shipped_pot):
    """~840 neighbours per atom: the largest capacity tier (several triplet-list chunks, lists parked in L2).
    Results must not change from run to run there either."""
    import gapcu
    cell = np.eye(3) * 2.05
    pos = np.array([[i, j, k] for i in range(2) for j in range(2) for k in range(2)], float) * 1.02 + 0.05
    pos += np.random.default_rng(9).normal(0, 0.03, pos.shape)
    z = np.array([6, 5, 6, 6, 5, 6, 6, 6], np.int32)
    c = gapcu.Context(0)
    c.load_potential(os.path.join(GOLDEN, "gap_parameters"))
    first = c.evaluate(z, cell, pos, 6.0, True)
    for _ in range(3):
        r = c.evaluate(z, cell, pos, 6.0, True)
        assert r["energy"] == first["energy"] and np.array_equal(r["forces"], first["forces"]) and np.array_equal(r["stress"], first["stress"])
    c.close()
