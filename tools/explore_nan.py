#!/usr/bin/env python
"""Which out-of-domain parameter vectors make the reference (oracle) and the device disagree about NaN?  (developer tool)"""
import itertools
import os
import sys
import warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import workloads as W

warnings.simplefilter('ignore')
for name, wl in (('ShockCooling4', W.example_sc4(npoints=40)), ('ShockCooling3', W.synthetic_sc3(npoints=64))):
    wl.priors_spec = [('uniform', -1e9, 1e9)] * wl.ndim
    lp = W.oracle_log_posterior(wl)
    base = 0.5 * (wl.p_lo + wl.p_up)
    tmin, tmax = wl.t.min(), wl.t.max()
    for precision in ('fp64', 'fp32'):
        prob = wl.device_problem(precision)
        rows = []
        nphys = 4
        for signs in itertools.product([1, -1, 0], repeat=nphys):
            for t0 in (base[-1], tmax + 1., 0.5 * (tmin + tmax)):
                p = base.copy()
                p[:nphys] = base[:nphys] * np.array(signs)
                p[-1] = t0
                rows.append(p)
        P = np.array(rows)
        with np.errstate(all='ignore'):
            want = np.array([lp(p) for p in P])
        got = prob.log_posterior(P)
        cls = {}
        for p, a, b in zip(P, want, got):
            key = (tuple(np.sign(p[:nphys]).astype(int)), 'all-before' if p[-1] > tmax else ('some-before' if p[-1] > tmin else 'none-before'))
            kind = ('both-nan' if np.isnan(a) and np.isnan(b) else 'oracle-finite/device-nan' if np.isnan(b) else
                    'oracle-nan/device-finite' if np.isnan(a) else ('agree' if np.isclose(a, b, rtol=1e-4) or (np.isinf(a) and a == b) else 'VALUE-DIFFERS'))
            cls.setdefault(kind, []).append(key)
        print(name, precision, {k: len(v) for k, v in cls.items()})
        for k in ('oracle-finite/device-nan', 'oracle-nan/device-finite', 'VALUE-DIFFERS'):
            for key in sorted(set(cls.get(k, [])))[:40]:
                print('   ', k, key)
