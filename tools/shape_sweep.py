"""The launch-shape cost model (choose_shape, lcf_api.cu) against a sweep of forced shapes: for each workload the time of the model's own
pick, the best forced shape, and their ratio.  Developer tool (GPU):  python tools/shape_sweep.py > profiles/round2_shape_model_vs_sweep.jsonl"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from lightcurve_fitting_b200 import _capi, synthetic
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    L = _capi.lib()
    rng = np.random.default_rng(0)
    cases = [
        ('cfg1 SN 2016bkv ShockCooling4 N=149, 100 walkers', synthetic.example_sc4(), 100, 200,
         [(1, 16, 1, 4), (1, 16, 1, 2), (1, 8, 2, 4), (1, 8, 4, 0), (2, 8, 2, 0), (4, 4, 2, 0), (4, 8, 1, 0), (8, 8, 1, 0), (1, 16, 2, 2)]),
        ('cfg1 SN 2016bkv ShockCooling4 N=758, 100 walkers', synthetic.example_sc4(window=None), 100, 200,
         [(1, 16, 1, 2), (1, 16, 1, 4), (1, 16, 2, 0), (1, 8, 4, 0), (2, 8, 2, 0), (2, 16, 1, 0), (4, 8, 1, 0), (1, 16, 4, 0)]),
        ('cfg4 CompanionShocking3 N=1000, 1e4 walkers', synthetic.synthetic_cs3(bench.kasen_sifto_truth, npoints=1000), 10_000, 30,
         [(4, 8, 1, 0), (2, 8, 1, 0), (8, 8, 1, 0), (8, 16, 1, 0), (16, 16, 1, 0), (32, 16, 1, 0), (32, 16, 2, 0), (4, 16, 1, 0), (4, 4, 1, 0)]),
        ('cfg2 ShockCooling3 N=2000, 1e5 walkers', bench.workload(bench.device_truth, 2000, 'sc3'), 100_000, 5,
         [(32, 16, 1, 0), (32, 8, 1, 0), (16, 16, 1, 0), (16, 8, 1, 0), (8, 8, 1, 0)]),
        ('cfg2 ShockCooling3 N=2000, 2e4 walkers', bench.workload(bench.device_truth, 2000, 'sc3'), 20_000, 10,
         [(32, 16, 1, 0), (32, 8, 1, 0), (16, 16, 1, 0), (16, 8, 1, 0), (8, 8, 1, 0), (32, 16, 2, 0)]),
    ]
    strong = [(32, 16, 1, 0), (16, 16, 1, 0), (8, 16, 1, 0), (8, 8, 1, 0), (4, 16, 1, 0), (4, 8, 1, 0), (2, 8, 1, 0), (1, 8, 1, 0), (32, 16, 2, 0), (16, 16, 2, 0)]
    for nw_s, gpus in ((50_000, 2), (25_000, 4), (12_500, 8)):        # the per-GPU share of the strong-scaling runs of bench.py (10^5 walkers in total)
        cases.append(('cfg2 ShockCooling3 N=2000, %d walkers (strong scaling on %d GPUs)' % (nw_s, gpus), bench.workload(bench.device_truth, 2000, 'sc3'), nw_s, 10, strong))
    only = sys.argv[1] if len(sys.argv) > 1 else ''
    for name, wl, nw, steps, shapes in cases:
        if wl is None or only not in name:
            continue
        prob = wl.device_problem('fp32')
        p0 = wl.start(nw, rng)

        def timed(shape):
            _capi.check(L.lcf_set_tuning_ex(*(shape[:3] if shape else (0, 0, 0))))
            _capi.check(L.lcf_set_tuning_split(shape[3] if shape else 0))
            s = EnsembleSampler(nw, wl.ndim, prob, seed=5)
            s.run_mcmc(p0, 3, skip_initial_state_check=True, store=False)
            best = 1e30
            for _ in range(3):
                s.run_mcmc(None, steps, skip_initial_state_check=True, store=False)
                best = min(best, s.last_ms)
            return best / (2 * steps) * 1e3, prob.last_launch()

        us0, launch = timed(None)
        sweep = {}
        for sh in shapes:
            try:
                sweep['%d,%d,%d,%d' % sh] = round(timed(sh)[0], 2)
            except Exception as exc:                                      # a shape that does not fit / is not allowed
                sweep['%d,%d,%d,%d' % sh] = str(exc)[:40]
        _capi.check(L.lcf_set_tuning_ex(0, 0, 0))
        _capi.check(L.lcf_set_tuning_split(0))
        ok = {k: v for k, v in sweep.items() if isinstance(v, float)}
        kb = min(ok, key=ok.get)
        print(json.dumps({'workload': name, 'model_pick': {k: launch[k] for k in ('walkers_per_cta', 'warps_per_cta', 'cluster', 'grid', 'kernel', 'flat')},
                          'model_pick_us_per_half_step': round(us0, 2), 'sweep_best': kb, 'sweep_best_us': ok[kb],
                          'pick_over_best': round(us0 / ok[kb], 3), 'sweep_us_per_half_step (walkers,warps,cluster,sample chunks)': sweep}), flush=True)


if __name__ == '__main__':
    main()
