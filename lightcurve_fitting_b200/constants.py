"""Physical constants in the reference's working units (CODATA 2018 / IAU 2015, what astropy >= 4 uses).

Reference: models.py:10-12, models.py:1101-1102, filters.py:11, bolometric.py:419.
"""
import numpy as np

_h = 6.62607015e-34       # J s
_kB = 1.380649e-23        # J / K
_c = 299792458.0          # m / s
_e = 1.602176634e-19      # J / eV
_sigma_sb = 2. * np.pi ** 5 * _kB ** 4 / (15. * _h ** 3 * _c ** 2)
_Rsun = 6.957e8
_au = 1.495978707e11
_Mpc = 1e6 * (_au * 648000. / np.pi)

k_B = _kB / _e * 1e3                                                         # eV / kK
c3 = (4. * np.pi * (_sigma_sb * 1e7 * _Rsun ** 2 * 1e12)) ** -0.5 / 1000.    # R in 1000 Rsun from L [erg/s], T [kK]
c4 = 1. / (4. * np.pi * _Mpc ** 2.)                                          # L_nu [W/Hz] / d[Mpc]^2 -> F_nu [W m-2 Hz-1]
c1 = _h / _kB * 1e12 / 1e3                                                   # kK / THz
c2 = 8 * np.pi ** 2 * (_h / _c ** 2) * (1000. * _Rsun) ** 2 * 1e36           # W/Hz/(1000 Rsun)^2/THz^3
c_AA_THz = _c * 1e10 / 1e12                                                  # angstrom * THz
c_nm_THz = _c * 1e9 / 1e12                                                   # nm * THz
sigma_sb = _sigma_sb * (1000. * _Rsun) ** 2 * 1e12                           # W/(1000 Rsun)^2/kK^4
