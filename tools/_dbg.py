import sys, os, numpy as np
sys.path.insert(0, '.')
from tests import workloads as W
from lightcurve_fitting_b200._capi import lib, check
wl = W.synthetic_sc4(npoints=1200)
prob = wl.device_problem('fp32')
p0 = wl.start(9001, np.random.default_rng(8))
check(lib().lcf_set_tuning_ex(16, 8, 1))
for n in (5000, 9001, 9000, 8992, 4736, 4752):
    res = {}
    for flat in (0, 1):
        check(lib().lcf_set_tuning_flat(flat))
        res[flat] = prob.log_posterior(p0[:n]); L = prob.last_launch()
        print(n, 'flat', flat, L['grid'], L['groups'], L['sum_units'], L['flat'])
    d = np.flatnonzero(res[0] != res[1])
    print('   diffs', len(d), d[:10].tolist(), 'groups of first diffs', (d[:10] // 16).tolist())
