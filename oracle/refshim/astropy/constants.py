"""CODATA 2018 / IAU 2015 constants as Quantities (what astropy >= 4 provides)."""
import numpy as np
from . import units as u

h = u.Quantity(6.62607015e-34, u.J * u.s)
k_B = u.Quantity(1.380649e-23, u.J / u.K)
c = u.Quantity(299792458.0, u.m / u.s)
sigma_sb = u.Quantity(2. * np.pi ** 5 * 1.380649e-23 ** 4 / (15. * 6.62607015e-34 ** 3 * 299792458.0 ** 2), u.W / u.m ** 2 / u.K ** 4)
