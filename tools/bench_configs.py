#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (cfg1, cfg3, cfg4, cfg5) on one GPU, one JSON line each.

bench.py measures the headline configuration (cfg2).  This script covers the rest of SURVEY.md section 8(d):
  cfg1  bundled SN 2016bkv light curve, ShockCooling4, 100 walkers, 1000 + 1000 steps   (lightcurve_mcmc)
  cfg3  500 synthetic SED epochs as independent batched ensembles, 10 walkers, 200 + 100 steps (one launch)
  cfg4  CompanionShocking3 on a synthetic 1000-point light curve, 10^4 walkers
  cfg5  survey batch: synthetic light curves x ShockCooling4, 256 walkers each (scaled: --nlc per GPU), one launch
walker-steps/s = walkers x steps / device time (CUDA events inside the library); Planck samples/s beside it.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def device_truth(model_name, t, filter_names, params, z):
    from lightcurve_fitting_b200 import models as M
    from lightcurve_fitting_b200.filters import filtdict
    cls = getattr(M, model_name)
    m = cls(redshift=z)
    f = [filtdict[n] for n in filter_names]
    if model_name == 'BlackbodySED':
        return M.blackbody_to_filters(f, np.full(len(f), params[0]), np.full(len(f), params[1]), z=z)
    return np.asarray(m(np.asarray(t, float), f, *params), float)


def kasen_sifto_truth(t, filter_names, z):
    d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'lightcurve_fitting_b200', 'data', 'sifto.npz'))
    cols, tab = [str(c) for c in d['columns']], d['table'][3:]
    out = np.empty(len(t))
    for i, (ti, fn) in enumerate(zip(t, filter_names)):
        col = tab[:, cols.index(fn)]
        out[i] = 1e21 * np.interp(ti - 58000., tab[:, 0], col / col.max(), left=0., right=0.) + 2e19
    return out


def emit(name, walkers, steps, ms, samples_per_eval, extra=None):
    from lightcurve_fitting_b200 import _capi
    L = _capi.lib()
    if hasattr(L, 'lcf_debug_phase_clocks'):      # experiment builds (-DLCF_X_TIMING): per-phase SM clocks summed over CTAs
        import ctypes as C
        buf = (C.c_uint64 * 10)()
        L.lcf_debug_phase_clocks(buf)
        extra = dict(extra or {}, phase_clk_sum=[int(v) for v in buf[:5]] + [int(buf[5]) & ((1 << 40) - 1)], sub=[int(buf[6]), int(buf[7])], ring_barrier=int(buf[8]), ring_apply=int(buf[9]),
                     lane_tiles={'fast': int(buf[6]), 'clamped': int(buf[7]), 'careful': int(buf[5]) >> 40})
    ws = walkers * steps / (ms * 1e-3)
    out = {'config': name, 'walker_steps_per_s': ws, 'ms': ms, 'walkers': walkers, 'steps': steps,
           'planck_samples_per_eval': samples_per_eval, 'planck_samples_per_s': ws * samples_per_eval}
    out.update(extra or {})
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--precision', default='fp32')
    ap.add_argument('--nlc', type=int, default=1250)
    ap.add_argument('--nepochs', type=int, default=500)
    ap.add_argument('--only', default='')
    ap.add_argument('--tune', default='', help='wpb,warps,cluster[,sample chunks] launch-shape override(s), separated by ;')
    ap.add_argument('--short', action='store_true', help='fewer steps (shape sweeps)')
    args = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    from lightcurve_fitting_b200 import synthetic
    from lightcurve_fitting_b200.bolometric import BatchSampler
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    only = set(args.only.split(',')) if args.only else None
    from lightcurve_fitting_b200 import _capi
    for tune in (args.tune.split(';') if args.tune else ['']):
        if tune:
            v = [int(x) for x in tune.split(',')]
            _capi.check(_capi.lib().lcf_set_tuning_ex(*v[:3]))
            _capi.check(_capi.lib().lcf_set_tuning_split(v[3] if len(v) > 3 else 0))      # optional 4th entry: sample chunks (split-K)
        run_configs(args, only, tune, synthetic, BatchSampler, EnsembleSampler)


def run_configs(args, only, tune, synthetic, BatchSampler, EnsembleSampler):
    rng = np.random.default_rng(0)
    n1 = 100 if args.short else 1000
    tag_t = (' [tune %s]' % tune) if tune else ''

    if not only or 'cfg1' in only:
        for window, tag in (((57468., 57485.), 'N=149 early window'), (None, 'N=758 full')):
            wl = synthetic.example_sc4(window=window)
            prob = wl.device_problem(args.precision)
            # (a) half-step launches (k_pass), as lightcurve_mcmc runs it
            s = EnsembleSampler(100, wl.ndim, prob, seed=1)
            s.run_mcmc(wl.start(100, rng), n1, store=False)
            s.reset()
            s.run_mcmc(None, n1)
            emit('cfg1 %s k_pass%s' % (tag, tag_t), 100, n1, s.last_ms, wl.planck_samples_per_eval(),
                 {'launches': s.last_launches, 'acceptance': float(s.acceptance_fraction.mean())})
            if args.short:
                continue
            # (b) the whole chain in one launch (k_chain)
            b = BatchSampler([prob], 100, seed=1)
            b.run(wl.start(100, rng)[None], 1000, 1000)
            emit('cfg1 %s k_chain (1000+1000 steps, one launch)' % tag, 100, 2000, b.last_ms, wl.planck_samples_per_eval())

    if not only or 'cfg3' in only:
        wls = [synthetic.sed_epoch(device_truth, rng) for _ in range(args.nepochs)]
        probs = [w.device_problem(args.precision) for w in wls]
        for nw in (10, 64):
            b = BatchSampler(probs, nw, seed=2)
            p0 = np.stack([w.start(nw, rng) for w in wls])
            b.run(p0, 200, 100)
            b.run(p0, 200, 100)
            spe = float(np.mean([w.planck_samples_per_eval() for w in wls]))
            emit('cfg3 %d SED epochs, %d walkers, 200+100 steps, one launch' % (args.nepochs, nw), args.nepochs * nw, 300,
                 b.last_ms, spe, {'acceptance': float(b.acceptance_fraction.mean()), 'status_ok': bool(np.all(b.status == 0))})

    if not only or 'cfg4' in only:
        wl = synthetic.synthetic_cs3(kasen_sifto_truth, npoints=1000)
        prob = wl.device_problem(args.precision)
        s = EnsembleSampler(10_000, wl.ndim, prob, seed=3)
        s.run_mcmc(wl.start(10_000, rng), 20, store=False, skip_initial_state_check=True)
        s.run_mcmc(None, 50)
        emit('cfg4 CompanionShocking3 N=1000, 1e4 walkers' + tag_t, 10_000, 50, s.last_ms, wl.planck_samples_per_eval(),
             {'acceptance': float(s.acceptance_fraction.mean())})

    if not only or 'cfg5' in only:
        # survey batch: under torchrun every rank takes its round-robin share of the light curves (parallel.shard_items),
        # no data-path collective; the aggregate is all light curves over the slowest rank's time
        world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
        if world > 1:
            import torch
            import torch.distributed as dist
            from lightcurve_fitting_b200 import _capi
            from lightcurve_fitting_b200.parallel import shard_items
            local = int(os.environ.get('LOCAL_RANK', '0'))
            torch.cuda.set_device(local)
            _capi.check(_capi.lib().lcf_set_device(local))
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
            mine = shard_items(args.nlc * world, rank, world)
        else:
            mine = list(range(args.nlc))
        npts = [int(n) for n in np.random.default_rng(0).integers(100, 301, args.nlc * world)]
        t0 = time.time()
        wls = [synthetic.synthetic_sc4(device_truth, npoints=npts[i], lc_index=i) for i in mine]
        t1 = time.time()                                   # (benchmark's own data generation, not part of the product path)
        probs = [w.device_problem(args.precision) for w in wls]
        prep = time.time() - t1
        b = BatchSampler(probs, 256, seed=4)
        p0 = np.stack([w.start(256, rng) for w in wls])
        b.run(p0, 20, 20)
        ms_short = b.last_ms
        b.run(p0, 200, 200)
        spe = float(np.mean([w.planck_samples_per_eval() for w in wls]))
        ms_all, nlc_all = b.last_ms, len(wls)
        if world > 1:
            t = torch.tensor([b.last_ms, float(len(wls)), spe * len(wls)], device='cuda', dtype=torch.float64)
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms_all, nlc_all, spe = float(tmax[0]), int(t[1]), float(t[2] / t[1])
            dist.destroy_process_group()
            if rank:
                return
        emit('cfg5 %d light curves x ShockCooling4, 256 walkers, 200+200 steps, one launch per GPU, %d GPU(s)' % (nlc_all, world),
             nlc_all * 256, 400, ms_all, spe, {'problem_build_s': prep, 'synthetic_data_s': t1 - t0, 'ms_20+20': ms_short, 'acceptance': float(b.acceptance_fraction.mean()),
                              'status_ok': bool(np.all(b.status == 0))})


if __name__ == '__main__':
    main()
