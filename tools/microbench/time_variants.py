#!/usr/bin/env python
"""Time experimental builds of liblcf_b200 on the cfg2 half-step kernel (developer tool, not the bench).

    python tools/microbench/time_variants.py [--model 3|4] [--shapes wpb,nw,cluster;...] lib1.so lib2.so ...

Each library runs in its own process (LCF_B200_LIB selects it); prints one line per (library, shape):
kernel ms per half-step, walker-steps/s and Planck samples/clk/SM at the 1965 MHz nominal clock.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(args):
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    from lightcurve_fitting_b200 import _capi, synthetic
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    if args.model == 3:
        wl = bench.workload(bench.device_truth, 2000)
        spe = 97000
    else:
        wl = synthetic.synthetic_sc4(bench.device_truth, npoints=2000, seed=1, filters=['U', 'B', 'V', 'R', 'I', 'g', 'r', 'i'])
        spe = 2 * 97000
    prob = wl.device_problem(args.precision)
    p0 = wl.start(args.walkers, np.random.default_rng(1))
    for shape in args.shapes.split(';'):
        wpb, nw, cl = (int(x) for x in shape.split(','))
        _capi.check(_capi.lib().lcf_set_tuning_ex(wpb, nw, cl))
        s = EnsembleSampler(args.walkers, wl.ndim, prob, seed=5)
        s.run_mcmc(p0, 3, skip_initial_state_check=True, store=False)
        best = 1e30
        for _ in range(3):
            s.run_mcmc(None, args.steps, skip_initial_state_check=True, store=False)
            best = min(best, s.last_ms)
        ms = best / (2 * args.steps)
        extra = {}
        L = _capi.lib()
        if hasattr(L, 'lcf_debug_phase_clocks'):
            import ctypes as C
            buf = (C.c_uint64 * 10)()
            L.lcf_debug_phase_clocks(buf)                      # reset
            s.run_mcmc(None, 1, skip_initial_state_check=True, store=False)
            L.lcf_debug_phase_clocks(buf)
            extra['phase_clk_sums'] = [int(v) for v in buf[:10]]
            if hasattr(L, 'lcf_debug_cta_log'):
                log = (C.c_uint64 * (3 * 4096))()
                L.lcf_debug_cta_log(log)
                a = np.array(log[:], dtype=np.uint64).reshape(4096, 3)
                grid = int(prob.last_launch()['grid'])
                a = a[:min(grid, 4096)].astype(np.int64)
                t0 = a[:, 0].min()
                np.save(os.path.join(ROOT, 'gpurun_out', 'cta_log_flat%s.npy' % os.environ.get('LCF_FLAT', 'x')), a)
                dur = (a[:, 1] - a[:, 0]) / 1e3
                per_sm = np.bincount(a[:, 2], minlength=148)
                extra['cta_log'] = {'grid': grid, 'kernel_us': float((a[:, 1].max() - t0) / 1e3), 'start_spread_us': float((a[:, 0].max() - t0) / 1e3),
                                    'end_min_us': float((a[:, 1].min() - t0) / 1e3), 'dur_us_min_med_max': [float(dur.min()), float(np.median(dur)), float(dur.max())],
                                    'ctas_per_sm_min_max': [int(per_sm.min()), int(per_sm.max())], 'sms_used': int((per_sm > 0).sum())}
        sps = args.walkers / 2 * spe / (ms * 1e-3)
        print(json.dumps({'lib': os.path.basename(os.environ.get('LCF_B200_LIB', 'default')), 'shape': shape, 'model': args.model,
                          'ms_half_step': round(ms, 4), 'walker_steps_per_s': round(args.walkers / (2 * ms * 1e-3)),
                          'samples_per_clk_sm': round(sps / (148 * 1.965e9), 3), 'acc': round(float(s.acceptance_fraction.mean()), 3), **extra}),
              flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--model', type=int, default=3)
    ap.add_argument('--walkers', type=int, default=100000)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--precision', default='fp32')
    ap.add_argument('--shapes', default='0,0,0')
    ap.add_argument('--child', action='store_true')
    ap.add_argument('libs', nargs='*')
    args = ap.parse_args()
    if args.child:
        return child(args)
    for lib in args.libs or ['']:
        env = dict(os.environ)
        if lib:
            env['LCF_B200_LIB'] = os.path.abspath(lib)
        cmd = [sys.executable, os.path.abspath(__file__), '--child', '--model', str(args.model), '--walkers', str(args.walkers),
               '--steps', str(args.steps), '--shapes', args.shapes, '--precision', args.precision]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        sys.stdout.write(r.stdout)
        if r.returncode:
            print(json.dumps({'lib': lib, 'error': r.stderr[-400:]}), flush=True)


if __name__ == '__main__':
    main()
