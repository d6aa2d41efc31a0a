"""Filter registry: names/aliases -> photometric system, zero-point flux, packed curve key, unit.

Pure data restating the table at reference filters.py:369-440 (plot colours and
legend offsets, which the hot path never reads, are dropped).  Each row is
``(names, system, fnu [W m-2 Hz-1] or AB default, curve key in data/filter_curves.npz or '',
wavelengths_in_angstrom)``.  Order matters: it defines ``Filter.order`` (decreasing effective
frequency, filters.py:441).
"""
AB = 3.631e-23  # filters.py:117 default zero point


def _jw(name, inst):
    return ((name,), 'JWST ' + inst, AB, 'JWST_%s.%s.dat' % (inst, name), True)


REGISTRY = [
    (('FUV',), 'GALEX', AB, 'GALEX_GALEX.FUV.dat', True),
    (('NUV',), 'GALEX', AB, 'GALEX_GALEX.NUV.dat', True),
    (('UVW2', 'uvw2', 'W2', '2', 'uw2'), 'Swift', 7.379e-24, 'Swift_UVOT.UVW2.dat', True),
    (('UVM2', 'uvm2', 'M2', 'M', 'um2'), 'Swift', 7.656e-24, 'Swift_UVOT.UVM2.dat', True),
    (('UVW1', 'uvw1', 'W1', '1', 'uw1'), 'Swift', 9.036e-24, 'Swift_UVOT.UVW1.dat', True),
    (('u', "u'", 'up', 'uprime'), 'Gunn', AB, 'SLOAN_SDSS.u.dat', True),
    (('U_S', 's', 'us'), 'Swift', 1.419e-23, 'Swift_UVOT.U.dat', True),
    (('U',), 'Johnson', 1.790e-23, 'Generic_Johnson.U.dat', True),
    (('B',), 'Johnson', 4.063e-23, 'Generic_Johnson.B.dat', True),
    (('B_S', 'b', 'bs'), 'Swift', 4.093e-23, 'Swift_UVOT.B.dat', True),
    (('g', "g'", 'gp', 'gprime', 'F475W'), 'Gunn', AB, 'SLOAN_SDSS.g.dat', True),
    (('g-DECam',), 'DECam', AB, 'CTIO_DECam.g.dat', True),
    (('c', 'cyan'), 'ATLAS', AB, 'ATLAS_cyan.txt', False),
    (('V',), 'Johnson', 3.636e-23, 'Generic_Johnson.V.dat', True),
    (('V_S', 'v', 'vs'), 'Swift', 3.664e-23, 'Swift_UVOT.V.dat', True),
    (('Itagaki',), 'Itagaki', AB, 'KAF-1001E.asci', False),
    (('white',), 'MOSFiT', AB, 'white.txt', False),
    (('unfilt.', '0', 'C', 'clear', 'pseudobolometric', 'griz', 'RGB', 'LRGB'), 'MOSFiT', AB,
     'pseudobolometric.txt', False),
    (('G',), 'Gaia', AB, 'GAIA_GAIA0.G.dat', True),
    (('Kepler',), 'Kepler', AB, 'Kepler_Kepler.K.dat', True),
    (('TESS',), 'TESS', AB, 'TESS_TESS.Red.dat', True),
    (('DLT40', 'Open', 'Clear'), 'DLT40', AB, 'QE_E2V_MBBBUV_Broadband.csv', False),
    (('w',), 'Gunn', AB, 'PAN-STARRS_PS1.w.dat', True),
    (('o', 'orange'), 'ATLAS', AB, 'ATLAS_orange.txt', False),
    (('r', "r'", 'rp', 'rprime', 'F625W'), 'Gunn', AB, 'SLOAN_SDSS.r.dat', True),
    (('r-DECam',), 'DECam', AB, 'CTIO_DECam.r.dat', True),
    (('R', 'Rc', 'R_s'), 'Johnson', 3.064e-23, 'Generic_Cousins.R.dat', True),
    (('i', "i'", 'ip', 'iprime', 'F775W'), 'Gunn', AB, 'SLOAN_SDSS.i.dat', True),
    (('i-DECam',), 'DECam', AB, 'CTIO_DECam.i.dat', True),
    (('I', 'Ic'), 'Johnson', 2.416e-23, 'Generic_Cousins.I.dat', True),
    (('z_s', 'zs'), 'Gunn', AB, 'PAN-STARRS_PS1.z.dat', True),
    (('z', "z'", 'zp', 'zprime'), 'Gunn', AB, 'SLOAN_SDSS.z.dat', True),
    (('z-DECam',), 'DECam', AB, 'CTIO_DECam.z.dat', True),
    (('y',), 'Gunn', AB, 'PAN-STARRS_PS1.y.dat', True),
    (('y-DECam',), 'DECam', AB, 'CTIO_DECam.Y.dat', True),
    (('J',), 'UKIRT', 1.589e-23, 'Gemini_Flamingos2.J.dat', True),
    (('H',), 'UKIRT', 1.021e-23, 'Gemini_Flamingos2.H.dat', True),
    (('K', 'Ks'), 'UKIRT', 0.640e-23, 'Gemini_Flamingos2.Ks.dat', True),
    (('L',), 'UKIRT', 0.285e-23, '', False),
] + [_jw(n, 'NIRCam') for n in ('F070W', 'F090W', 'F115W', 'F150W', 'F182M', 'F200W', 'F250M', 'F277W',
                                'F300M', 'F335M', 'F356W', 'F360M', 'F444W')] \
  + [_jw(n, 'MIRI') for n in ('F560W', 'F770W', 'F1000W', 'F1130W', 'F1280W', 'F1500W', 'F1800W',
                              'F2100W', 'F2550W')] + [
    (('pseudobolometric, curve_fit',), None, AB, '', False),
    (('pseudobolometric, MCMC',), None, AB, '', False),
    (('pseudobolometric, integration',), None, AB, '', False),
    (('bolometric, curve_fit',), None, AB, '', False),
    (('bolometric, MCMC',), None, AB, '', False),
    (('unknown', '?'), 'unknown', AB, '', False),
]
