"""Unit conversion and masses: same names and numerical values as the reference's
simulation_utilities/Constants.py (values feed sigma = sqrt(dt/m) and the sample potentials, so
they must agree digit for digit: Constants.py:65-69)."""
from ._isotope_masses import ISOTOPE_MASSES

__all__ = ['Constants', 'get_atomic_num', 'get_atomic_string']

massDict = dict(ISOTOPE_MASSES)
_SYMBOLS = [s for s, _ in ISOTOPE_MASSES]


def get_atomic_num(atms):
    """Position (1-based) of each symbol in the mass table (Constants.py:41-47)."""
    return [_SYMBOLS.index(a) + 1 for a in atms]


def get_atomic_string(atomic_num):
    """Inverse look-up with the reference's indexing (Constants.py:50-57: symbol at index atomic_num)."""
    if type(atomic_num) is not list:
        atomic_num = [atomic_num]
    return [_SYMBOLS[z] for z in atomic_num]


class Constants:
    """convert / mass / reduced_mass with the reference's conversion factors."""
    atomic_units = {
        "wavenumbers": 4.556335281212229e-6,
        "angstroms": 1 / 0.529177,
        "amu": 1.000000000000000000 / 6.02213670000e23 / 9.10938970000e-28,
    }

    @classmethod
    def convert(cls, val, unit, to_AU=True):
        factor = cls.atomic_units[unit]
        return (val * factor) if to_AU else (val / factor)

    @classmethod
    def mass(cls, atom, to_AU=True):
        m = massDict[atom]
        return cls.convert(m, 'amu') if to_AU else m

    @classmethod
    def reduced_mass(cls, atoms, to_AU=True):
        first, second = atoms.split('-')[:2]
        m1, m2 = massDict[first], massDict[second]
        if to_AU:
            m1, m2 = cls.convert(m1, 'amu'), cls.convert(m2, 'amu')
        return m1 * m2 / (m1 + m2)
