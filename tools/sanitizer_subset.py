#!/usr/bin/env python
"""A small pass over every kernel path, sized for compute-sanitizer (memcheck / racecheck / synccheck are 10-50x slower
than a normal run):  compute-sanitizer --tool racecheck python tools/sanitizer_subset.py
(compute-sanitizer is closed on the round-1 GPU pool, so it has only been run plain there.)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import __graft_entry__ as g
    g.build()
    from lightcurve_fitting_b200 import _capi
    from lightcurve_fitting_b200.bolometric import BatchSampler
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    from tests import workloads as W
    L, check = _capi.lib(), _capi.check
    rng = np.random.default_rng(0)
    for wl, prec in ((W.example_sc4(npoints=40), 'fp32'), (W.synthetic_sc3(npoints=64), 'fp32'), (W.synthetic_sc3(npoints=48), 'fp64'),
                     (W.synthetic_cs3(npoints=60), 'fp32')):
        prob = wl.device_problem(prec)
        P = wl.start(24, rng)
        for shape in ((0, 0, 0), (32, 4, 2), (4, 4, 4), (2, 2, 8)):          # plain launches and thread-block clusters (DSMEM reduce)
            check(L.lcf_set_tuning_ex(*shape))
            lp = prob.log_posterior(P)
            assert np.isfinite(lp).any()
            s = EnsembleSampler(24, wl.ndim, prob, seed=1)
            s.run_mcmc(P, 3)
        check(L.lcf_set_tuning_ex(0, 0, 0))
        print('ok', wl.model_name, prec, flush=True)
    # chain kernel: narrow and wide walker groups, diagnostics, batched least squares
    wls = [W.synthetic_sc4(npoints=n, lc_index=i) for i, n in enumerate((30, 17))]
    for nw in (12, 80):
        b = BatchSampler([w.device_problem('fp32') for w in wls], nw, seed=2).run(np.stack([w.start(nw, rng) for w in wls]), 2, 3)
        assert np.all(b.status == 0)
    wl = W.example_sc4(npoints=30)
    s = EnsembleSampler(16, wl.ndim, wl.device_problem('fp32'), seed=3)
    s.run_mcmc(wl.start(16, rng), 12)
    s.get_autocorr_time(tol=0, quiet=True)
    s.get_split_rhat()
    print('ok chain kernel, diagnostics', flush=True)
    # segmented bank (k_pass_seg): slices of ~2 filters staged in turn, ShockCooling3's weight table rebuilt per slice
    os.environ['LCF_SEG_SAMPLES'] = '100'
    os.environ['LCF_RING'] = '0'
    for wl, prec in ((W.synthetic_sc3(npoints=64), 'fp32'), (W.synthetic_cs3(npoints=60), 'fp64')):
        prob = wl.device_problem(prec)
        s = EnsembleSampler(24, wl.ndim, prob, seed=5)
        s.run_mcmc(wl.start(24, rng), 3)
        assert prob.last_launch()['kernel'] == 'k_pass_seg'
    print('ok segmented bank', flush=True)


if __name__ == '__main__':
    main()
