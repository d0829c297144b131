"""libgap.BOND.Bond -- same surface as gappy/libgap/BOND.py:14-52:
``Bond(rcut=6.0).get_min_bond(lat, elements, pos)`` returns the smallest
interatomic (image) distance within rcut, 10.0 if there is none."""
from libgap.libgap import fget_bond

from ._elements import atomic_numbers


class Bond(object):
    def __init__(self, nf=None, rcut=6.0, lgrad=True):
        self.nf = nf
        self.rcut = rcut
        self.lgrad = lgrad

    def get_elenum(self, x):
        return atomic_numbers(x)

    def get_min_bond(self, lat, elements, pos):
        try:
            numbers = self.get_elenum(elements)
        except (KeyError, TypeError):
            numbers = elements
        return fget_bond(lat, numbers, pos, self.rcut)
