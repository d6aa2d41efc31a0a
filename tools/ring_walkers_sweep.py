"""Steps per second of the persistent chain kernel against the number of walkers (developer tool): shows the cost of the 149th and
150th CTA of a look-ahead grid on 148 SMs (profiles/round2_cfg1_lookahead.txt)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightcurve_fitting_b200 import synthetic
from lightcurve_fitting_b200.sampler import EnsembleSampler
rng = np.random.default_rng(0)
for window in ((57468., 57485.), None):
    wl = synthetic.example_sc4(window=window)
    prob = wl.device_problem('fp32')
    for W in (100, 98, 96, 64, 48):
        s = EnsembleSampler(W, wl.ndim, prob, seed=1)
        s.run_mcmc(wl.start(W, rng), 100, store=False)
        best = 1e9
        for _ in range(3):
            s.run_mcmc(None, 200)
            best = min(best, s.last_ms)
        print(len(wl.t), W, prob.last_launch()['kernel'], prob.last_launch()['grid'], 'us per step %.2f' % (best * 1e3 / 200), 'M walker-steps/s %.2f' % (W * 200 / best / 1e3), flush=True)
