"""Minimal stand-in for the `ase` package (not installed in this image): just enough of
Atoms / Calculator for gappy-style calculators to run in the tests."""
import numpy as np


class Atoms(object):
    def __init__(self, numbers, positions, cell, pbc=True):
        self.numbers = np.asarray(numbers, int)
        self.positions = np.asarray(positions, float)
        self.cell = np.asarray(cell, float)
        self.pbc = pbc
        self.calc = None

    def __len__(self):
        return len(self.numbers)

    def copy(self):
        return Atoms(self.numbers.copy(), self.positions.copy(), self.cell.copy(), self.pbc)

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def set_calculator(self, calc):
        self.calc = calc

    def _get(self, name):
        self.calc.calculate(self, [name])
        return self.calc.results[name]

    def get_potential_energy(self):
        return self._get("energy")

    def get_forces(self):
        return self._get("forces")

    def get_stress(self):
        return self._get("stress")
