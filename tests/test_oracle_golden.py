"""Pins the CPU oracle against the reference's own shipped output
(gappy/example/ASE-GAPPY/ase.traj -> tests/golden/ase_traj_frames.npz): 11 MD
frames of real libgap E/F/stress with the shipped gap_parameters, rcut 6.0
(run_md.py:10).  Gates from SURVEY.md section 7: E rel <= 2e-12, F <= 1e-10 eV/A,
stress <= 1e-9 GPa (measured: 2e-14 / 5e-13 / 2e-12)."""
import numpy as np
import pytest

E_GOLD = [-571.211650663666, -571.290904684498, -571.454060275248, -571.551537960416,
          -571.527223195738, -571.421512421583, -571.305517514876, -571.232808718274,
          -571.225990557717, -571.281249419903, -571.370274395870]


def test_fixture_matches_published_energies(golden_frames):
    assert np.allclose(golden_frames["energy"], E_GOLD, rtol=0, atol=5e-13)


def test_reader_shipped_file(shipped_pot):
    p = shipped_pot
    assert (p.nspecies, p.nsf, p.nsparse, p.des_len) == (2, 33, 129, 66)
    assert list(p.z) == [5, 6] and list(p.w) == [-1.0, 4.0]
    assert [int((p.ntype == t).sum()) for t in (1, 2, 3, 4)] == [7, 14, 5, 7]
    assert p.theta[0] == 2.1695365180 and p.theta[7] == 45724.7581559285
    assert abs(np.abs(p.coeff).sum() - 362851.52381007315) < 1e-6


@pytest.mark.parametrize("frame", range(11))
def test_dense_oracle_vs_golden(shipped_pot, golden_frames, frame):
    g = golden_frames
    r = shipped_pot.calc_dense(g["numbers"], g["cell"][frame], g["positions"][frame], 6.0, True)
    assert abs(r["energy"] - g["energy"][frame]) <= 2e-12 * abs(g["energy"][frame])
    assert np.abs(r["forces"] - g["forces"][frame]).max() <= 1e-10
    assert np.abs(r["stress"] - g["stress"][frame]).max() <= 1e-9


@pytest.mark.parametrize("frame", [0, 5, 10])
def test_sparse_oracle_vs_golden(shipped_pot, golden_frames, frame):
    g = golden_frames
    r = shipped_pot.calc_sparse(g["numbers"], g["cell"][frame], g["positions"][frame], 6.0, True)
    assert abs(r["energy"] - g["energy"][frame]) <= 2e-12 * abs(g["energy"][frame])
    assert np.abs(r["forces"] - g["forces"][frame]).max() <= 1e-10
    assert np.abs(r["stress"] - g["stress"][frame]).max() <= 1e-9


def test_lgrad_false_gives_zero_force_and_stress(shipped_pot, golden_frames):
    g = golden_frames
    r = shipped_pot.calc_dense(g["numbers"], g["cell"][0], g["positions"][0], 6.0, False)
    assert abs(r["energy"] - g["energy"][0]) <= 2e-12 * abs(g["energy"][0])
    assert not r["forces"].any() and not r["stress"].any()


def test_example_problem_size(shipped_pot, golden_frames):
    """158 neighbours per atom in diamond at rcut 6."""
    g = golden_frames
    r = shipped_pot.calc_dense(g["numbers"], g["cell"][10], g["positions"][10], 6.0, True, stats=True)
    assert r["stats"][0] == 64 * 158
