"""Synthetic structure / potential generators shared by tests and bench.py
(SURVEY.md section 8(d) configs C2-C5).  Pure numpy; no oracle, no GPU."""
import numpy as np

SHIPPED_SF = None  # filled lazily from the shipped potential by callers that need it


def min_image_distance_ok(cell, pos, dmin):
    """True if every pair (incl. periodic images in the +-1 shell) is >= dmin apart."""
    n = len(pos)
    shifts = np.array([[a, b, c] for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)], float) @ cell
    for s in shifts:
        d = pos[:, None, :] - (pos[None, :, :] + s)
        r2 = (d * d).sum(-1)
        if not s.any():
            r2[np.arange(n), np.arange(n)] = 1e9
        if r2.min() < dmin * dmin:
            return False
    return True


def cubic_supercell(nx, ny, nz, a=2.15, jitter=0.15, seed=1000, species=(5, 6, 7), probs=(0.3, 0.4, 0.3)):
    """C2 / C4: simple-cubic sites with Gaussian jitter, wrapped into the cell."""
    rng = np.random.default_rng(seed)
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"), -1).reshape(-1, 3)
    cell = np.diag([nx * a, ny * a, nz * a]).astype(float)
    pos = g * a + rng.normal(0.0, jitter, size=g.shape)
    frac = pos @ np.linalg.inv(cell)
    frac -= np.floor(frac)
    pos = frac @ cell
    z = rng.choice(np.array(species), size=len(pos), p=np.array(probs))
    return cell, pos, z.astype(np.int32)


def random_candidate(seed, nmin=32, nmax=128, species=(5, 6, 7), dmin=1.2):
    """C3: one CALYPSO-style random candidate structure (triclinic cell)."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(nmin, nmax + 1))
    while True:
        v = rng.uniform(8.0, 14.0)
        L = (n * v) ** (1.0 / 3.0)
        a, b, c = L * rng.uniform(0.7, 1.3, 3)
        al, be, ga = np.deg2rad(rng.uniform(60.0, 120.0, 3))
        cx = c * np.cos(be)
        cy = c * (np.cos(al) - np.cos(be) * np.cos(ga)) / np.sin(ga)
        cz2 = c * c - cx * cx - cy * cy
        if cz2 <= 0.05 * c * c:
            continue
        cell = np.array([[a, 0, 0], [b * np.cos(ga), b * np.sin(ga), 0], [cx, cy, np.sqrt(cz2)]])
        vol = abs(np.linalg.det(cell))
        if vol < 6.0 * n:
            continue
        # sequential insertion with min-distance rejection
        pos = np.zeros((0, 3))
        tries = 0
        while len(pos) < n and tries < 200 * n:
            tries += 1
            p = rng.uniform(0, 1, 3) @ cell
            trial = np.vstack([pos, p])
            if len(pos) == 0:
                ok = min_image_distance_ok(cell, trial, dmin)
            else:
                ok = True
                for sa in (-1, 0, 1):
                    for sb in (-1, 0, 1):
                        for sc in (-1, 0, 1):
                            s = np.array([sa, sb, sc], float) @ cell
                            d = pos + s - p
                            if ((d * d).sum(-1) < dmin * dmin).any():
                                ok = False
                            if (sa or sb or sc) and (s * s).sum() < dmin * dmin:
                                ok = False
            if ok:
                pos = trial
        if len(pos) == n:
            break
    z = rng.choice(np.array(species), size=n).astype(np.int32)
    return cell, pos, z


def sheared(cell, pos, shear=((1, .07, -.04), (0, 1, .05), (0, 0, 1))):
    s = np.array(shear, float)
    return cell @ s, pos @ s


def write_gap_parameters(path, z, w, ntype, alpha, cutoff, theta, mm, coeff):
    """Write a potential in the reference's text format: values
    with 10 decimals, three skipped records after the size line."""
    with open(path, "w") as f:
        f.write("%5d\n" % len(z))
        for zi, wi in zip(z, w):
            f.write("%5d %14.5f\n" % (zi, wi))
        f.write("%5d\n" % len(ntype))
        for t, a, c in zip(ntype, alpha, cutoff):
            f.write("%3d %9.5f %9.5f\n" % (t, a, c))
        f.write("%12d%12d\n \n \n \n" % (mm.shape[0], mm.shape[1]))
        f.write("".join("%25.10f" % x for x in theta) + "\n")
        for row in mm:
            f.write("".join("%25.10f" % x for x in row) + "\n")
        f.write("".join("%25.10f" % x for x in coeff) + "\n")
