"""Material table and the `Material` record.

API mirror of src/diffelastic/material_model.py:8-26 (`MatSet`, `Material`).  The
reference's `TinyNN` and `LinearElastic` in the same file are unused legacy code
 and are not part of the hot path.
"""


class MatSet:
    """(density, Young's modulus, Poisson's ratio, Rayleigh alpha, Rayleigh beta)."""
    Ceramic = 2700, 7.2E10, 0.19, 6, 1E-7
    Glass = 2600, 6.2E10, 0.20, 1, 1E-7
    Wood = 750, 1.1E10, 0.25, 60, 2E-6
    Plastic = 1070, 1.4E9, 0.35, 30, 1E-6
    Iron = 8000, 2.1E11, 0.28, 10, 1e-7
    Polycarbonate = 1190, 2.4E9, 0.37, 0.5, 4E-7
    Steel = 7850, 2.0E11, 0.29, 20, 3E-8
    Tin = 7265, 5e10, 0.325, 2, 3E-8
    Test = 2700, 6E10, 0.19, 6, 1E-7
    RandomMin = 2700, 1E10, 0.1, 6, 1E-7
    RandomMax = 2700, 1E11, 0.4, 6, 1E-7


class Material(object):
    def __init__(self, material):
        self.density, self.youngs, self.poisson, self.alpha, self.beta = material
