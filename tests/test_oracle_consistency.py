"""Oracle self-consistency where the golden data is silent:
mixed species weights, triclinic cells, cells smaller than rcut (self images),
unwrapped atoms, stress component order.  The dense transliteration is the
arbiter; the sparse form and finite differences are checked against it."""
import numpy as np
import pytest

from structures import random_candidate, sheared


def _bc(bc_structure, shear=True):
    cell, pos = bc_structure["cell"], bc_structure["positions"]
    if shear:
        cell, pos = sheared(cell, pos)
    return cell, pos, bc_structure["numbers"].astype(np.int32)


def test_survey_crosscheck_values(shipped_pot, bc_structure):
    """Non-authoritative numbers recorded by the survey probe."""
    cell, pos, z = _bc(bc_structure, shear=False)
    r = shipped_pot.calc_dense(z, cell, pos, 6.0, True, stats=True)
    assert abs(r["energy"] - (-452.937805514)) < 1e-8
    assert r["stats"][0] == 64 * 238
    cell, pos, z = _bc(bc_structure, shear=True)
    r = shipped_pot.calc_dense(z, cell, pos, 6.0, True)
    assert abs(r["energy"] - (-442.3872883654)) < 1e-8
    want = np.array([339.558485, 325.789851, 327.983327, -103.038090, -71.148337, 63.245950])
    assert np.abs(r["stress"] - want).max() < 2e-6


def test_sparse_equals_dense_sheared_two_species(shipped_pot, bc_structure):
    cell, pos, z = _bc(bc_structure)
    d = shipped_pot.calc_dense(z, cell, pos, 6.0, True, desc=True)
    s = shipped_pot.calc_sparse(z, cell, pos, 6.0, True, desc=True)
    assert abs(d["energy"] - s["energy"]) <= 1e-13 * abs(d["energy"])
    assert np.array_equal(d["xx"], s["xx"])          # same terms, same order -> same bits
    assert np.abs(d["forces"] - s["forces"]).max() < 1e-11
    assert np.abs(d["stress"] - s["stress"]).max() < 1e-10


@pytest.mark.parametrize("seed", [3000, 3001, 3002])
def test_sparse_equals_dense_small_triclinic_cells(shipped_pot, seed):
    """32-48 atom random cells: interplanar spacings < rcut -> nabc >= 2, self images."""
    cell, pos, z = random_candidate(seed, 32, 48, species=(5, 6))
    d = shipped_pot.calc_dense(z, cell, pos, 6.0, True)
    s = shipped_pot.calc_sparse(z, cell, pos, 6.0, True)
    assert abs(d["energy"] - s["energy"]) <= 1e-13 * abs(d["energy"])
    assert np.abs(d["forces"] - s["forces"]).max() < 1e-10
    assert np.abs(d["stress"] - s["stress"]).max() < 1e-9


def test_neighbor_sets_cell_list_vs_double_loop(oracle, bc_structure):
    cell, pos, _ = _bc(bc_structure)
    a = oracle.neighbors(cell, pos, 6.0, sparse=False)
    b = oracle.neighbors(cell, pos, 6.0, sparse=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_neighbor_sets_unwrapped_atoms(oracle, bc_structure):
    """Atoms displaced by whole lattice vectors: the reference's +-nabc window then
    misses neighbours (SURVEY.md App. A 'known defects'); both oracle forms must
    reproduce exactly that set, not the physically complete one."""
    cell, pos, _ = _bc(bc_structure)
    pos = pos.copy()
    pos[3] += 2 * cell[0] - cell[2]
    pos[17] -= cell[1]
    a = oracle.neighbors(cell, pos, 6.0, sparse=False)
    b = oracle.neighbors(cell, pos, 6.0, sparse=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    full = oracle.neighbors(cell, _bc(bc_structure)[1], 6.0)[0]
    assert a[0].sum() < full.sum()


def test_exact_cutoff_ties_are_inside(oracle):
    """dis == rcut is kept (gap_calc.f90:101 uses .gt.): simple cubic a=2, rcut=6
    has the (3,0,0) shell exactly at the cutoff."""
    cell = np.eye(3) * 4.0
    pos = np.array([[0, 0, 0], [2, 0, 0], [0, 2, 0], [2, 2, 0], [0, 0, 2], [2, 0, 2], [0, 2, 2], [2, 2, 2.0]])
    cnt, idx, sh, dis = oracle.neighbors(cell, pos, 6.0)
    ref = 0
    for i in range(-3, 4):
        for j in range(-3, 4):
            for k in range(-3, 4):
                if (i or j or k) and i * i + j * j + k * k <= 9:
                    ref += 1
    assert (cnt == ref).all()
    assert (dis[0, :cnt[0]] == 6.0).sum() == 6 + 24  # (3,0,0) and (2,2,1) shells
    assert np.array_equal(oracle.neighbors(cell, pos, 6.0, sparse=True)[0], cnt)


def test_image_range_matches_formula(oracle):
    assert list(oracle.image_range(np.eye(3) * 7.1119999886, 6.0)) == [1, 1, 1]
    assert list(oracle.image_range(np.diag([3.0, 6.5, 13.0]), 6.0)) == [2, 1, 1]


def test_finite_difference_forces_and_stress(shipped_pot, bc_structure):
    """F = -dE/dx and sigma = -(1/V) dE/d(eps) / 6.24219e-3 in the order
    (xx yy zz xy yz xz) -- six distinct components on the sheared B8C56 cell."""
    cell, pos, z = _bc(bc_structure)
    r = shipped_pot.calc_sparse(z, cell, pos, 6.0, True)
    h = 1e-5
    for (i, c) in [(0, 0), (5, 1), (40, 2)]:
        p1 = pos.copy(); p1[i, c] += h
        p2 = pos.copy(); p2[i, c] -= h
        fd = -(shipped_pot.calc_sparse(z, cell, p1, 6.0, False)["energy"] -
               shipped_pot.calc_sparse(z, cell, p2, 6.0, False)["energy"]) / (2 * h)
        assert abs(fd - r["forces"][i, c]) < 2e-5 * max(1.0, abs(fd))
    vol = abs(np.linalg.det(cell))
    comps = [(0, 0), (1, 1), (2, 2), (0, 1), (1, 2), (0, 2)]
    for q, (a, b) in enumerate(comps):
        e = np.zeros((3, 3)); e[a, b] += h / 2; e[b, a] += h / 2
        ep = shipped_pot.calc_sparse(z, cell @ (np.eye(3) + e), pos @ (np.eye(3) + e), 6.0, False)["energy"]
        em = shipped_pot.calc_sparse(z, cell @ (np.eye(3) - e), pos @ (np.eye(3) - e), 6.0, False)["energy"]
        sig = -((ep - em) / (2 * h)) / vol / 6.24219e-3
        assert abs(sig - r["stress"][q]) < 1e-4 * max(1.0, abs(sig)), (q, sig, r["stress"][q])
    assert len(set(np.round(r["stress"], 3))) == 6


def test_get_bond(oracle, bc_structure, golden_frames):
    cell, pos, _ = _bc(bc_structure, shear=False)
    cnt, idx, sh, dis = oracle.neighbors(cell, pos, 6.0)
    want = min(dis[i, :cnt[i]].min() for i in range(len(pos)))
    assert oracle.get_bond(cell, pos, 6.0) == want
    assert oracle.get_bond(np.eye(3) * 50.0, np.array([[0., 0, 0], [20., 0, 0]]), 6.0) == 10.0  # get_bond.f90:32


def test_missing_species_is_an_error(shipped_pot, golden_frames):
    g = golden_frames
    z = g["numbers"].copy().astype(np.int32); z[0] = 14
    with pytest.raises(RuntimeError):
        shipped_pot.calc_dense(z, g["cell"][0], g["positions"][0])
