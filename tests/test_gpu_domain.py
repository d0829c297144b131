"""Spatial decomposition (BASELINE config 4) and Verlet-skin list reuse (8(f) N3)
on ONE GPU: a gapcu group runs 2, 4 or 8 bricks as separate contexts on device 0 -- the same halo
kernels and phases the NCCL transport drives, with device-to-device copies in place of
ncclSend/ncclRecv -- and must reproduce the undecomposed evaluation and the CPU oracle."""
import os

import numpy as np
import pytest

from structures import cubic_supercell, sheared

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POT_C2 = os.path.join(ROOT, "bench_data", "gap_parameters_c2")


@pytest.fixture(scope="module")
def single():
    import gapcu
    c = gapcu.Context(0)
    c.load_potential(POT_C2)
    yield c
    c.close()


def _close(got, want, etol=1e-12, ftol=1e-9, stol=1e-8):
    assert abs(got["energy"] - want["energy"]) <= etol * abs(want["energy"])
    assert np.abs(got["forces"] - want["forces"]).max() <= max(ftol, 1e-12 * np.abs(want["forces"]).max())
    assert np.abs(got["stress"] - want["stress"]).max() <= max(stol, 1e-11 * np.abs(want["stress"]).max())


@pytest.mark.parametrize("dims,grid", [((12, 10, 8), (2, 1, 1)), ((12, 10, 8), (2, 2, 2)), ((16, 8, 8), (4, 1, 1)),
                                       ((12, 10, 8), (1, 2, 2)), ((12, 10, 8), None)])
def test_bricks_on_one_gpu_equal_undecomposed(single, oracle, dims, grid):
    import gapcu
    cell, pos, z = cubic_supercell(*dims, seed=4000)
    n = 8 if grid is None else int(np.prod(grid))
    g = gapcu.Group([0] * n)
    g.load_potential(POT_C2)
    got = g.evaluate(z, cell, pos, 6.0, True, grid)
    want = single.evaluate(z, cell, pos, 6.0, True)
    _close(got, want)
    # every atom is owned by exactly one brick
    ids = np.concatenate([g.ctx(r).owned() for r in range(n)])
    assert np.array_equal(np.sort(ids), np.arange(len(pos)))
    # bit-reproducible: same bricks, same bits
    again = g.evaluate(z, cell, pos, 6.0, True, grid)
    assert again["energy"] == got["energy"] and np.array_equal(again["forces"], got["forces"])
    if dims == (12, 10, 8) and grid == (2, 2, 2):
        ref = oracle.read(POT_C2).calc_sparse(z, cell, pos, 6.0, True)
        _close(got, ref, etol=1e-10, ftol=1e-8, stol=1e-7)          # the north-star gates against the oracle
    g.close()


def test_bricks_triclinic_unwrapped_and_energy_only(single):
    """Sheared cell, atoms shifted out of the cell by whole lattice vectors (the image shifts of the
    ghosts must absorb the wrap offsets), and lgrad = false."""
    import gapcu
    cell, pos, z = cubic_supercell(12, 12, 8, seed=4100)
    cell, pos = sheared(cell, pos)
    rng = np.random.default_rng(5)
    pos = pos + rng.integers(-1, 2, size=(len(pos), 3)).astype(float) @ cell * (rng.random(len(pos)) < 0.1)[:, None]
    g = gapcu.Group([0] * 4)
    g.load_potential(POT_C2)
    got = g.evaluate(z, cell, pos, 6.0, True, (2, 2, 1))
    want = single.evaluate(z, cell, pos, 6.0, True)
    _close(got, want)
    e_only = g.evaluate(z, cell, pos, 6.0, False, (2, 2, 1))
    assert abs(e_only["energy"] - want["energy"]) <= 1e-12 * abs(want["energy"])
    g.close()


def test_brick_too_thin_is_an_error():
    import gapcu
    cell, pos, z = cubic_supercell(10, 10, 10, seed=1000)      # 21.5 A: four bricks would be 5.4 A thick
    g = gapcu.Group([0] * 4)
    g.load_potential(POT_C2)
    with pytest.raises(gapcu.GapcuError) as e:
        g.set_structure(z, cell, pos, 6.0, (4, 1, 1))
    assert e.value.code == -7
    g.close()


def test_verlet_reuse_keeps_the_reference_neighbour_sets(single, oracle):
    """Skin lists kept over MD-like steps: the exact lists (re-filtered each step with the reference's
    arithmetic) equal the oracle's double loop bit for bit, E/F/stress equal a fresh evaluation, and a
    move beyond skin/2 makes the library rebuild on its own."""
    import gapcu
    cell, pos, z = cubic_supercell(8, 8, 8, seed=4200)
    c = gapcu.Context(0)
    c.load_potential(POT_C2)
    c.set_skin(0.5)
    c.evaluate(z, cell, pos, 6.0, True)
    rng = np.random.default_rng(11)
    p = pos.copy()
    for step in range(4):
        p = p + rng.normal(0.0, 0.03, p.shape)
        if step == 3:
            p[17] += np.array([0.4, 0.0, 0.0])                 # > skin/2: stale -> rebuilt inside fetch
        c.update_positions(p, True)
        c.compute(True)
        e, f, s = c.fetch()
        want = single.evaluate(z, cell, p, 6.0, True)
        _close({"energy": e[0], "forces": f, "stress": s[0]}, want)
        cnt, idx, sh, dis = c.neighbors(cap=256)
        ocnt, oidx, osh, odis = oracle.neighbors(cell, p, 6.0, cap=256)
        assert np.array_equal(cnt, ocnt)
        for i in range(len(p)):
            k = cnt[i]
            assert np.array_equal(idx[i, :k], oidx[i, :k]) and np.array_equal(sh[i, :k], osh[i, :k])
            assert np.array_equal(dis[i, :k].view(np.int64), odis[i, :k].view(np.int64))
    c.close()


def test_verlet_reuse_in_a_group(single):
    import gapcu
    cell, pos, z = cubic_supercell(12, 10, 8, seed=4300)
    g = gapcu.Group([0] * 4)
    g.load_potential(POT_C2)
    g.set_skin(0.4)
    g.evaluate(z, cell, pos, 6.0, True, (2, 2, 1))
    rng = np.random.default_rng(12)
    p = pos.copy()
    for step in range(3):
        p = p + rng.normal(0.0, 0.02, p.shape)
        if step == 2:
            p[40] += np.array([0.0, 0.3, 0.0])                 # stale on one brick: all bricks rebuild together
        g.update_positions(p, True)
        g.compute(True)
        e, f, s = g.fetch()
        _close({"energy": e, "forces": f, "stress": s}, single.evaluate(z, cell, p, 6.0, True))
    g.close()


def test_atom_beyond_the_drift_allowance_is_reported():
    import gapcu
    cell, pos, z = cubic_supercell(12, 10, 8, seed=4400)
    g = gapcu.Group([0] * 2)
    g.load_potential(POT_C2)
    g.evaluate(z, cell, pos, 6.0, True, (2, 1, 1))
    p = pos.copy()
    frac = p @ np.linalg.inv(cell)
    i = int(np.argmin(np.abs(frac[:, 0] - 0.25)))              # middle of brick 0
    p[i, 0] += 0.3 * cell[0, 0]                                # now deep inside brick 1
    g.update_positions(p, False)
    g.compute(True)
    with pytest.raises(gapcu.GapcuError) as e:
        g.fetch()
    assert e.value.code == -8
    # setting the structure again re-partitions and works
    r = g.evaluate(z, cell, p, 6.0, True, (2, 1, 1))
    assert np.isfinite(r["energy"])
    g.close()


def test_fgap_calc_on_several_devices_is_a_brick_group(single, monkeypatch):
    """The drop-in call itself: with more than one device set, FGAP_CALC's C entry point cuts a large
    structure into bricks (here two contexts on the one GPU) and must return what a single device returns,
    in the Fortran layouts."""
    import gapcu
    from oracle import Oracle
    pot = Oracle("parity").read(POT_C2)
    monkeypatch.chdir(os.path.join(ROOT, "bench_data"))
    link = os.path.join(ROOT, "bench_data", "gap_parameters")
    made = not os.path.lexists(link)
    if made:
        os.symlink("gap_parameters_c2", link)
    try:
        cell, pos, z = cubic_supercell(12, 10, 8, seed=4500)
        want = single.evaluate(z, cell, pos, 6.0, True)
        monkeypatch.setenv("GAPCU_GROUP_MIN_ATOMS", "500")
        gapcu.set_devices([0, 0])
        e, f, s, v = gapcu.fortran_calc(z, cell, pos, pot.theta, pot.mm, pot.coeff, 6.0, True)
        _close({"energy": e, "forces": f, "stress": s}, want)
        assert v == 0.0
        # a cell too small to cut falls back to one device
        cell2, pos2, z2 = cubic_supercell(5, 5, 5, seed=4501)
        e2, f2, s2, _ = gapcu.fortran_calc(z2, cell2, pos2, pot.theta, pot.mm, pot.coeff, 6.0, True)
        monkeypatch.setenv("GAPCU_GROUP_MIN_ATOMS", "100")
        e3, f3, s3, _ = gapcu.fortran_calc(z2, cell2, pos2, pot.theta, pot.mm, pot.coeff, 6.0, True)
        assert abs(e3 - e2) <= 1e-12 * abs(e2) and np.abs(f3 - f2).max() <= 1e-9
    finally:
        gapcu.set_devices([])
        if made:
            os.remove(link)
