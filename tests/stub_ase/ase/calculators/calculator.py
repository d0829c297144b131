all_changes = ['positions', 'numbers', 'cell', 'pbc', 'initial_charges', 'initial_magmoms']


class PropertyNotImplementedError(NotImplementedError):
    pass


class Calculator(object):
    implemented_properties = []

    def __init__(self, **kwargs):
        self.atoms = None
        self.results = {}

    def calculate(self, atoms=None, properties=['energy'], system_changes=all_changes):
        if atoms is not None:
            self.atoms = atoms.copy()
