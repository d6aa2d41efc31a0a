"""Throughput of the segmented-bank fallback (k_pass_seg) on a light curve through the 30 most densely sampled filters of the
registry (> 17 000 transmission samples: larger than shared memory in FP64), next to the FP32 launch of the same problem (fits).
Usage: python tools/bench_segmented_bank.py [walkers] [steps]   (prints one JSON line per precision)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightcurve_fitting_b200.synthetic import Workload   # noqa: E402
from lightcurve_fitting_b200.sampler import EnsembleSampler   # noqa: E402
from lightcurve_fitting_b200 import models as M, filters as F   # noqa: E402

NAMES = ['F2550W', 'F2100W', 'NUV', 'F1800W', 'F444W', 'Itagaki', 'F356W', 'F1500W', 'F277W', 'K', 'H', 'Kepler', 'F200W', 'F1280W',
         'TESS', 'F335M', 'F360M', 'F770W', 'F1000W', 'FUV', 'w', 'F300M', 'F150W', 'UVW1', 'J', 'r-DECam', 'U', 'B', 'g', 'i']


def main():
    nw = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    rng = np.random.default_rng(21)
    n = 300
    t0 = 59000.
    t = np.sort(rng.uniform(t0 + 0.3, t0 + 12., n))
    fn = [NAMES[i % len(NAMES)] for i in range(n)]
    p_true = np.array([1., 1., 1., 3., t0])
    pri = [('uniform', 0., 10.), ('uniform', 0., 10.), ('uniform', 0., 100.), ('uniform', 0., 100.), ('uniform', t0 - 2., t0 + 0.3)]
    lo, hi = p_true * 0.9, p_true * 1.1
    lo[4], hi[4] = t0 - 0.1, t0 + 0.1
    # data from the device model itself (FP32 problem, which fits): the benchmark needs a plausible light curve, not the oracle
    wl0 = Workload('large-bank', 'ShockCooling4', t, fn, np.ones(n), np.ones(n), pri, lo, hi, z=0.005, truth=p_true)
    y = wl0.device_problem('fp32').model_eval(p_true[None, :])[0]
    wl = Workload('large-bank', 'ShockCooling4', t, fn, y, 0.05 * y, pri, lo, hi, z=0.005, truth=p_true)
    for precision in ('fp64', 'fp32'):
        prob = wl.device_problem(precision)
        s = EnsembleSampler(nw, wl.ndim, prob, seed=3)
        st = s.run_mcmc(wl.start(nw, rng), 2, store=False, skip_initial_state_check=True)
        t1 = time.perf_counter()
        s.run_mcmc(None, steps, store=False)
        np.asarray(s._state().log_prob)
        dt = time.perf_counter() - t1
        print(json.dumps({'precision': precision, 'walkers': nw, 'steps': steps, 'points': n, 'filters': len(NAMES),
                          'planck_samples_per_eval': wl.planck_samples_per_eval(), 'launch': prob.last_launch(),
                          'walker_steps_per_s': nw * steps / dt, 'gplanck_samples_per_s': nw * steps * wl.planck_samples_per_eval() / dt / 1e9,
                          'seg_warps_env': os.environ.get('LCF_SEG_WARPS')}), flush=True)


if __name__ == '__main__':
    main()
