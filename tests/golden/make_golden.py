"""Extract the reference's own golden vectors into small fixtures.

Run HERE (the dev container), where /root/reference exists; the outputs are
committed so the GPU box (which has no /root/reference) can run the parity
tests.  Nothing in here is computed by this repo: every number is read from
files shipped by the reference.

  ase_traj_frames.npz  <- gappy/example/ASE-GAPPY/ase.traj   (11 MD frames of
                          real libgap output: positions, E, F, stress[GPa])
  bc_structure.npz     <- gappy/example/BC/sps_all.xyz        (B8C56 geometry only;
                          its energy/force columns are DFT labels, not libgap)
  poscar_c64.npz       <- gappy/example/ASE-GAPPY/POSCAR, CONTCAR
  gap_parameters       <- gappy/example/ASE-GAPPY/gap_parameters (potential data
                          file, byte-identical copy; the BC copy is identical)
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from ulm import read_ulm  # noqa: E402

REF = "/root/reference/gappy/example"


def read_poscar(path):
    lines = open(path).read().split("\n")
    scale = float(lines[1])
    cell = np.array([[float(x) for x in lines[2 + i].split()] for i in range(3)]) * scale
    symbols = lines[5].split()
    counts = [int(x) for x in lines[6].split()]
    assert lines[7].strip().lower().startswith("d")
    n = sum(counts)
    frac = np.array([[float(x) for x in lines[8 + i].split()[:3]] for i in range(n)])
    return cell, symbols, counts, frac


def main():
    tag, items = read_ulm(os.path.join(REF, "ASE-GAPPY", "ase.traj"))
    assert tag == "ASE-Trajectory" and len(items) == 11
    numbers = items[0]["numbers"]
    pos = np.stack([it["positions"] for it in items])
    cell = np.stack([np.array(it["cell"], dtype=np.float64) for it in items])
    ene = np.array([it["calculator"]["energy"] for it in items])
    frc = np.stack([it["calculator"]["forces"] for it in items])
    sts = np.array([it["calculator"]["stress"] for it in items])
    np.savez(os.path.join(HERE, "ase_traj_frames.npz"), numbers=numbers, positions=pos,
             cell=cell, energy=ene, forces=frc, stress=sts)

    # BC example geometry (extended xyz, one frame)
    lines = open(os.path.join(REF, "BC", "sps_all.xyz")).read().split("\n")
    n = int(lines[0])
    hdr = lines[1]
    lat = hdr.split('Lattice="')[1].split('"')[0].split()
    cellbc = np.array([float(x) for x in lat]).reshape(3, 3)
    zs, xyz = [], []
    for ln in lines[2:2 + n]:
        t = ln.split()
        xyz.append([float(t[1]), float(t[2]), float(t[3])])
        zs.append(int(t[4]))
    np.savez(os.path.join(HERE, "bc_structure.npz"), numbers=np.array(zs), positions=np.array(xyz),
             cell=cellbc)

    c0, sym, cnt, f0 = read_poscar(os.path.join(REF, "ASE-GAPPY", "POSCAR"))
    c1, _, _, f1 = read_poscar(os.path.join(REF, "ASE-GAPPY", "CONTCAR"))
    np.savez(os.path.join(HERE, "poscar_c64.npz"), cell=c0, frac=f0, contcar_cell=c1, contcar_frac=f1)

    shutil.copyfile(os.path.join(REF, "ASE-GAPPY", "gap_parameters"), os.path.join(HERE, "gap_parameters"))
    print("frames", pos.shape, "E0", repr(ene[0]), "BC atoms", n)


if __name__ == "__main__":
    main()
