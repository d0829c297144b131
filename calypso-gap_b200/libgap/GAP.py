"""libgap.GAP.Calculator -- same surface as the reference class
(gappy/libgap/GAP.py:14-61):

    gap = Calculator(rcut=6.0)          # reads ./gap_parameters (fgap_read)
    ene, force, stress, variance = gap.gap_calc(species, lat, pos, lgrad)

``species`` may be chemical symbols or atomic numbers; ``lat`` rows are lattice
vectors; ``stress`` is (xx, yy, zz, xy, yz, zx) in GPa (gappy/README.md:45).
"""
from libgap.libgap import fgap_calc, fgap_read

from ._elements import atomic_numbers


class Calculator(object):
    def __init__(self, rcut=6.0):
        self.gap_read()
        self.rcut = rcut

    def get_elenum(self, x):
        return atomic_numbers(x)

    def gap_read(self):
        (self.nsparseX, self.des_len, self.theta, self.mm, self.invcmm, self.coeff) = fgap_read()

    def gap_calc(self, species, lat, pos, lgrad):
        try:
            numbers = self.get_elenum(species)
        except (KeyError, TypeError):
            numbers = species
        m, d = self.nsparseX, self.des_len
        return fgap_calc(numbers, lat, pos, self.theta[:d], self.mm[:m, :d], self.invcmm[:m, :m],
                         self.coeff[:m], self.rcut, lgrad)
