"""ASE calculators.

``GAP`` -- same behaviour as the reference's gappy/ASE/gap_calc.py:19-52 (which users
copy into ase/calculators/): energy, free_energy, forces, the stress exactly as libgap
returns it (GPa, order xx yy zz xy yz zx, not converted to ASE units) and 'variance'.
A fresh libgap.GAP.Calculator is built on every step like the reference does
(gap_calc.py:41), so ./gap_parameters is re-read by fgap_read each time (that costs the
128 MB invcmm zero-fill of FGAP_READ per step, gap_calc.f90:313,361).

``GAPPersistent`` -- additive: same results, but one GPU context and
one parse of the potential for the whole run; nothing is re-read or re-allocated per
step.  Use it for MD / relaxations of small cells where the per-step host overheads of
the drop-in path dominate.
"""
import numpy as np
from ase.calculators.calculator import Calculator, all_changes

import libgap.GAP as my_gap


class GAP(Calculator):
    implemented_properties = ['energy', 'forces', 'stress']
    nolabel = True

    def __init__(self, rcut=6.0):
        Calculator.__init__(self)
        self.rcut = rcut

    def calculate(self, atoms=None, properties=['energy'], system_changes=all_changes):
        Calculator.calculate(self, atoms, properties, system_changes)
        gap = my_gap.Calculator(rcut=self.rcut)
        energy, forces, stress, variance = gap.gap_calc(self.atoms.get_atomic_numbers(), np.asarray(self.atoms.cell),
                                                        self.atoms.positions, True)
        self.results['energy'] = energy
        self.results['free_energy'] = energy
        self.results['forces'] = forces
        self.results['stress'] = stress
        self.results['variance'] = variance


class GAPPersistent(Calculator):
    implemented_properties = ['energy', 'forces', 'stress']
    nolabel = True

    def __init__(self, rcut=6.0, potential='gap_parameters', device=0):
        Calculator.__init__(self)
        import gapcu
        self.rcut = rcut
        self.ctx = gapcu.Context(device)
        self.ctx.load_potential(potential)

    def calculate(self, atoms=None, properties=['energy'], system_changes=all_changes):
        Calculator.calculate(self, atoms, properties, system_changes)
        r = self.ctx.evaluate(self.atoms.get_atomic_numbers(), np.asarray(self.atoms.cell), self.atoms.positions,
                              self.rcut, True)
        self.results['energy'] = r['energy']
        self.results['free_energy'] = r['energy']
        self.results['forces'] = r['forces']
        self.results['stress'] = r['stress']
        self.results['variance'] = 0.0
