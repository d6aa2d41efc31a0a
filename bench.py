#!/usr/bin/env python
"""Benchmark of the MCMC hot path (BASELINE.json metric: walker-steps/s = log-posterior evaluations/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config cfg2|cfg5] [--scaling weak|strong] [--model sc3|sc4] [--precision fp32|fp64] [--no-extras]

Main line (what the driver runs): BASELINE.json configs[1] -- ShockCooling3 on a synthetic 2000-point, 8-filter light
curve, 10^5 walkers per GPU, FP32 mode.  A "step" is one stretch-move iteration of the whole ensemble: two fused
half-step kernels, W log-posterior evaluations.  N > 1: ONE ensemble of N x 10^5 walkers shared by the GPUs (weak
scaling); the accept epilogue of the half-step kernel stores accepted walkers into the peers' replicas over NVLink.

`value`     device-resident throughput, CUDA events around exactly K steps, max over ranks.
`e2e`       the same metric through the public Python API with HOST buffers: start positions H2D, initial
            log-probabilities, K steps, every finished step streamed D2H into pinned memory -- all inside the timed region.
`roofline`  the fused kernel against the arithmetic roofline of the path (no dense contraction, ~72 B of HBM per
            walker-step): one MUFU.EX2 per Planck sample is irreducible and the XU pipe issues 16 lanes/clk/SM, so
            peak = 16 samples/clk/SM x 148 SMs x the SM clock sampled during the run.  `traffic` is the DRAM traffic per launch
            taken from the committed ncu capture of the same kernel and shape (profiles/round2_ncu_traffic.json), else null.
`cpu_baseline`  the oracle (numpy restatement of the reference) run the way the reference runs: a serial emcee-order
            stretch-move chain on one host core, bounded to ~15 s.
`extras`    the rest of what BASELINE.json / north_star name, each with its own CUDA-event timing and clock sample:
            fp64 (cfg2 in FP64 mode), sc4 (ShockCooling4 on the cfg2 shape), cfg1 (SN 2016bkv, 100 walkers; FP32 and FP64), cfg3
            (calculate_bolometric end to end, 500 epochs), cfg4 (CompanionShocking3, 10^4 walkers), cfg5 (survey batch);
            N > 1: multigpu_bit_identical (12-step chain against a single-GPU ensemble), strong scaling of cfg2 at 10^5
            walkers in total, cfg5 sharded over the GPUs with no collective.

--impl reference: the reference's own CPU implementation of the path cannot be installed here (astropy / emcee / extinction
are absent, no network), so this arm times the oracle port: a REAL emcee-order stretch-move chain (StretchReplay) whose
log-posterior calls are spread over every host core (emcee's `pool=`), on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WALKERS_PER_GPU = 100_000
NPOINTS = 2000
MUFU_LANES_PER_CLK_SM = 16
FP64_LANES_PER_CLK_SM = 64
FP64_OPS_PER_SAMPLE = 10.75           # FP64-pipe instructions (DFMA+DMUL+DADD) per Planck sample of planck_quad_f64, counted in SASS (profiles/round2_sass_fp64_loop.txt)
BALANCED_SAMPLES_PER_CLK_SM = 13.5    # FMA+MUFU balanced bound of the two-transcendental formulation
SMS = 148
MODEL_NAMES = {'sc3': 'ShockCooling3', 'sc4': 'ShockCooling4'}


def workload_string(model, npoints, walkers):
    return 'cfg2: %s, synthetic %d-point 8-filter light curve, %d walkers per GPU' % (MODEL_NAMES[model], npoints, walkers)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during a timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0, enabled=True):
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled

    def start(self, settle=0.25):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '40'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            # nvidia-smi needs 0.1-1 s to attach on a fresh box: wait for its first line, or a short timed region sees "no samples"
            deadline = time.perf_counter() + 5.
            while not self.rows and time.perf_counter() < deadline and self.proc.poll() is None:
                time.sleep(0.01)
            time.sleep(settle)
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.enabled:
            return None
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (('hw_slowdown', 4), ('hw_thermal_slowdown', 5), ('sw_thermal_slowdown', 6), ('sw_power_cap', 7)):
                if len(r) > col and r[col].lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        # "under load": the upper half of the samples (idle samples before/after the region drag the median down)
        sm_sorted = sorted(sm)
        return {'sm_mhz': float(np.median(sm_sorted[len(sm_sorted) // 2:])), 'sm_max_mhz': float(max(mx)),
                'reasons': sorted(reasons), 'samples': len(sm)}


def device_truth(model_name, t, filter_names, params, z):
    """Noiseless synthetic data from the device model itself (no oracle on the measured arm)."""
    from lightcurve_fitting_b200 import models as M
    from lightcurve_fitting_b200.filters import filtdict
    f = [filtdict[n] for n in filter_names]
    if model_name == 'BlackbodySED':
        return M.blackbody_to_filters(f, np.full(len(f), params[0]), np.full(len(f), params[1]), z=z)
    m = getattr(M, model_name)(redshift=z)
    return np.asarray(m(np.asarray(t, float), f, *params), float)


def oracle_truth(model_name, t, filter_names, params, z):
    from tests import workloads as W
    return W.oracle_truth(model_name, t, filter_names, params, z)


def kasen_sifto_truth(t, filter_names, z):
    d = np.load(os.path.join(ROOT, 'lightcurve_fitting_b200', 'data', 'sifto.npz'))
    cols, tab = [str(c) for c in d['columns']], d['table'][3:]
    out = np.empty(len(t))
    for i, (ti, fn) in enumerate(zip(t, filter_names)):
        col = tab[:, cols.index(fn)]
        out[i] = 1e21 * np.interp(ti - 58000., tab[:, 0], col / col.max(), left=0., right=0.) + 2e19
    return out


MODEL = 'sc3'          # kept as a module global: tools/microbench and tools/check_multigpu.py call workload()


def workload(truth, npoints=NPOINTS, model=None):
    from lightcurve_fitting_b200 import synthetic
    if (model or MODEL) == 'sc4':      # the same shape with the MSW23 model (two blackbody syntheses per point)
        return synthetic.synthetic_sc4(truth, npoints=npoints, seed=1, filters=['U', 'B', 'V', 'R', 'I', 'g', 'r', 'i'])
    return synthetic.synthetic_sc3(truth, npoints=npoints, seed=1)


# ---------------------------------------------------------------------------------------------
# CPU arms (oracle).  Only these functions touch oracle/ (through tests/workloads.py).
# ---------------------------------------------------------------------------------------------
_POOL_LP = None


def _pool_init(npoints, model):
    global _POOL_LP
    from tests import workloads as W
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    _POOL_LP = W.oracle_log_posterior(workload(oracle_truth, npoints, model))


def _pool_eval(p):
    return float(_POOL_LP(np.asarray(p)))


def _stretch_chain(lp_map, lp_one, wl, nwalkers, seed):
    """An emcee-order stretch-move sampler on the oracle (StretchReplay), log-posterior calls through `lp_map`."""
    from oracle import reference_port as rp

    class Chain(rp.StretchReplay):
        def _lnprob(self, q):
            if np.any(np.isinf(q)) or np.any(np.isnan(q)):
                raise ValueError('At least one parameter value was infinite or NaN')
            out = np.array(lp_map(list(q)))
            if np.any(np.isnan(out)):
                raise ValueError('Probability function returned NaN')
            return out

    rs = np.random.RandomState(seed)
    s = Chain(nwalkers, wl.ndim, lp_one, random_state=rs)
    p0 = wl.p_lo + rs.rand(nwalkers, wl.ndim) * (wl.p_up - wl.p_lo)
    return s, p0


def cpu_baseline_1core(wl, budget_s=15.):
    """How the reference runs (fitting.py:130: emcee without a pool): a serial stretch-move chain, one un-vectorised
    log_posterior call per walker with a Python loop over the photometry points, on one host core, for ~budget_s."""
    from tests import workloads as W
    lp = W.oracle_log_posterior(wl)
    nw = 2 * wl.ndim + 2
    s, p0 = _stretch_chain(lambda q: [float(lp(p)) for p in q], lp, wl, nw, 0)
    t0 = time.perf_counter()
    pos, lnp, _ = s.run_mcmc(p0, 0)                       # initial log-probabilities (W evaluations, as run_mcmc does)
    t_init = time.perf_counter() - t0
    steps, t1 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s and steps < 1000:
        pos, lnp, _ = s.run_mcmc(pos, 1, log_prob0=lnp)
        steps += 1
    dt = time.perf_counter() - t1
    value = nw * steps / dt if steps else nw / t_init
    return {'value': value, 'unit': 'walker-steps/s', 'cores': 1, 'kind': 'port',
            'sample': 'serial emcee-order stretch-move chain on the oracle: %d walkers x %d steps of the %d-point %s light curve in %.1f s '
                      '(+ %.1f s for the initial log-probabilities), one core' % (nw, steps, len(wl.t), wl.model_name, dt, t_init)}


def run_reference(args):
    """--impl reference: the oracle port on every host core.  The chain is real (emcee's draw order, red/blue halves, accept
    test); the log-posterior calls of each half-step are mapped over a process pool, emcee's own way of using cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    wl = workload(oracle_truth, args.npoints)
    nw = max(2 * cores, 2 * wl.ndim + 2)
    nw += nw % 2
    ctx = mp.get_context('fork')
    with ctx.Pool(cores, initializer=_pool_init, initargs=(args.npoints, MODEL)) as pool:
        s, p0 = _stretch_chain(lambda q: pool.map(_pool_eval, q, chunksize=1), None, wl, nw, 0)
        pos, lnp, _ = s.run_mcmc(p0, args.warmup)
        t0 = time.perf_counter()
        s.run_mcmc(pos, args.steps, log_prob0=lnp)
        dt = time.perf_counter() - t0
    value = nw * args.steps / dt
    sample = ('stretch-move chain of %d walkers (of %d) x %d steps, log-posterior calls of each half-step over a %d-process pool'
              % (nw, args.walkers, args.steps, cores))
    print(json.dumps({
        'impl': 'reference', 'metric': 'walker-steps/s', 'value': value, 'unit': 'walker-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_string(MODEL, args.npoints, args.walkers)},
        'cpu_baseline': {'value': value, 'unit': 'walker-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'walker-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'acceptance_fraction': float(np.mean(s.acceptance_fraction)), 'gpu_launches': 0}))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide pieces of the GPU arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local))
        from lightcurve_fitting_b200 import _capi
        _capi.check(_capi.lib().lcf_set_device(self.local))
        if args.wpb or args.nw or args.cluster:
            _capi.check(_capi.lib().lcf_set_tuning_ex(args.wpb, args.nw, args.cluster))
        try:
            self.traffic = json.load(open(os.path.join(ROOT, 'profiles', 'round2_ncu_traffic.json')))
        except Exception:
            self.traffic = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            self.peaks = {}

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], device='cuda', dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def clocks(self):
        return ClockSampler(self.local, enabled=self.rank == 0)


def roofline_block(ctx, model, precision, samples_per_eval, value_per_gpu, ms, half_steps, launches, walkers, D, clk):
    sm_mhz = (clk or {}).get('sm_mhz') or 1965.0
    fp64 = precision == 'fp64'
    samples_per_s = value_per_gpu * samples_per_eval
    peak = (FP64_LANES_PER_CLK_SM / FP64_OPS_PER_SAMPLE if fp64 else MUFU_LANES_PER_CLK_SM) * SMS * sm_mhz * 1e6
    kernel = 'lcf::k_pass<%d,%s,5,true,%d>' % (3 if model == 'sc3' else 4, 'double' if fp64 else 'float', 4 if (model == 'sc3' and not fp64) else 2)
    tr = ctx.traffic.get('%s@cfg2' % kernel) if walkers == WALKERS_PER_GPU else None
    return {
        'bound': ('fp64 pipe (no dense contraction, ~72 B of HBM per walker-step)' if fp64 else
                  'sfu (XU pipe: one MUFU.EX2 per Planck sample; no dense contraction, ~72 B of HBM per walker-step)'),
        'kernel': kernel, 'achieved': samples_per_s / 1e9, 'peak': peak / 1e9, 'unit': 'GPlanck-samples/s',
        'frac': samples_per_s / peak,
        'peak_basis': (('%d FP64 lanes/clk/SM / %.2f FP64-pipe instructions per Planck sample of the shipped loop (SASS count, '
                        'profiles/round2_sass_fp64_loop.txt) x %d SMs x %.0f MHz; MEASURED_PEAKS.json has no FP64 entry '
                        '(tools/microbench/pipes.cu: DFMA 57 lanes/clk/SM)' % (FP64_LANES_PER_CLK_SM, FP64_OPS_PER_SAMPLE, SMS, sm_mhz))
                       if fp64 else
                       ('%d MUFU lanes/clk/SM x %d SMs x %.0f MHz (SM clock sampled during the timed region), 1 MUFU.EX2 per Planck '
                        'sample; MEASURED_PEAKS.json has no SFU entry, the pipe rate is confirmed by tools/microbench/loops.cu '
                        '(15.8 lanes/clk/SM)' % (MUFU_LANES_PER_CLK_SM, SMS, sm_mhz))),
        'samples_per_clk_sm': samples_per_s / (SMS * sm_mhz * 1e6),
        'frac_vs_survey_balanced_bound': None if fp64 else samples_per_s / (BALANCED_SAMPLES_PER_CLK_SM * SMS * sm_mhz * 1e6),
        'algorithmic_per_unit': '4 FP32 flops + 2 transcendentals per Planck sample (SURVEY.md 8(d)); %d samples per log-posterior '
                                '(zero-weight end samples of the transmission curves included in the count, skipped by the kernel)'
                                % samples_per_eval,
        'flops_tflops': 4 * samples_per_s / 1e12,
        'kernel_avg_ms': ms / half_steps, 'launches_per_half_step': launches / half_steps,
        'hbm_gbs_chain_writeback': walkers * (D + 1) * 8 * (half_steps / 2) / (ms * 1e-3) / 1e9,
        'hbm_peak_gbs_measured': ctx.peaks.get('hbm_gbs'),
        'traffic': (tr or {}).get('dram_bytes_per_launch'), 'traffic_source': (tr or {}).get('source'),
    }


def measure_cfg2(ctx, model, precision, walkers_per_gpu, steps, warmup, strong=False, e2e=True):
    """Device-resident and end-to-end throughput of one shared ensemble on the cfg2 light curve."""
    torch = ctx.torch
    from lightcurve_fitting_b200.parallel import ShardedEnsemble
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    args = ctx.args
    wl = workload(device_truth, args.npoints, model)
    prob = wl.device_problem(precision)
    W_total = walkers_per_gpu if strong else walkers_per_gpu * ctx.world
    if ctx.world > 1:
        W_total -= W_total % (2 * ctx.world)
    D = wl.ndim
    p0 = wl.start(W_total, np.random.default_rng(1))
    ens = ShardedEnsemble(prob, W_total, seed=1234, rank=ctx.rank, world=ctx.world, exchange=args.exchange)
    ens.set_state(p0)
    # L2 note: a CTA re-reads the 72 KB light curve + 3 KB bank (L2 / shared-memory resident by design); what streams through
    # HBM is the walker state and the chain write-back, once per step: no inter-iteration flush is needed.
    ens.run(warmup, store=False)
    ens.finish()
    ens.reserve(steps)
    ctx.barrier()
    clocks = ctx.clocks().start()
    ctx.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ens.sampler._timing()
    launches0 = ens.sampler.last_launches
    ev0.record()                      # the library launches on torch's current stream (lcf_ensemble_set_stream)
    ens.run(steps, store=True)        # chain write-back to HBM inside the timed region
    ev1.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(ev0.elapsed_time(ev1))
    ens.finish()
    clk = clocks.stop()
    value = W_total * steps / (ms * 1e-3)
    ens.sampler._timing()
    launches = ens.sampler.last_launches - (launches0 if not ens.fused else 0)
    fused = ens.fused
    ens.close()
    del ens
    out = {'value': value, 'ms_per_step': ms / steps, 'steps': steps, 'walkers_total': W_total, 'ndim': D, 'gpu_launches': launches,
           'clocks': clk, 'dtype': 'f32' if precision == 'fp32' else 'f64', 'launch': prob.last_launch() if ctx.rank == 0 else None}
    spe = wl.planck_samples_per_eval()
    out['planck_samples_per_eval'] = spe
    if not strong:
        out['roofline'] = roofline_block(ctx, model, precision, spe, value / ctx.world, ms, 2 * steps, launches, W_total // ctx.world, D, clk)
    if not e2e:
        return out, wl
    # ---- end to end through the public API with host (pinned) buffers ----------------------------------------------
    pin_in = torch.from_numpy(np.ascontiguousarray(p0)).pin_memory()
    if ctx.world == 1:
        pin_chain = torch.empty((steps, W_total, D), dtype=torch.float64).pin_memory()
        pin_lnp = torch.empty((steps, W_total), dtype=torch.float64).pin_memory()
        s = EnsembleSampler(W_total, D, prob, seed=99)
        s.run_mcmc(pin_in.numpy(), 1, skip_initial_state_check=True, store=False)   # warm-up of the API path
        chain, lnp = pin_chain.numpy(), pin_lnp.numpy()
        dts = []
        for _ in range(3):                                # host timing jitters by several per cent: three runs, the median is reported
            s.reset()
            s.reserve(steps)                              # device chain buffer allocated outside the timed region (setup, like the sampler itself)
            ctx.barrier()
            t0 = time.perf_counter()
            s.run_mcmc(pin_in.numpy(), steps, skip_initial_state_check=True, chain_out=chain, log_prob_out=lnp)
            dts.append(time.perf_counter() - t0)
        dt = float(np.median(dts))
        out['e2e_runs_ms'] = [1e3 * x for x in dts]
        assert np.isfinite(lnp).all() and chain.shape == (steps, W_total, D)
        api = 'EnsembleSampler.run_mcmc(host start positions, chain_out=pinned, log_prob_out=pinned): chain streamed to the host (device chain buffer reserved beforehand)'
        h2d = pin_in.numel() * 8
    else:
        e2 = ShardedEnsemble(prob, W_total, seed=99, rank=ctx.rank, world=ctx.world, exchange=args.exchange)
        first, count = e2.own_walkers()
        pin_chain = torch.empty((steps, count, D), dtype=torch.float64).pin_memory()
        pin_lnp = torch.empty((steps, count), dtype=torch.float64).pin_memory()
        e2.set_state(pin_in.numpy())
        e2.run(1, store=False)
        e2.finish()
        dts = []
        for _ in range(3):                    # three runs, the median (of the max over ranks) is reported
            e2.sampler.reset()
            e2.reserve(steps)                 # device chain buffer allocated outside the timed region
            ctx.barrier()
            t0 = time.perf_counter()
            e2.set_state(pin_in.numpy())      # H2D of this rank's walkers, initial log-probabilities, publish to the peers
            t_set = time.perf_counter()
            if e2.fused:                      # this rank's walkers streamed D2H into pinned buffers while the next steps run
                e2.run(steps, store=True, chain_out=pin_chain.numpy(), log_prob_out=pin_lnp.numpy())
                t_run = time.perf_counter()
                e2.finish()
            else:
                e2.run(steps, store=True)
                t_run = time.perf_counter()
                e2.finish()
                e2.get_own_chain(pin_chain.numpy(), pin_lnp.numpy())
            t_end = time.perf_counter()
            ctx.barrier()
            dts.append(ctx.max_over_ranks(time.perf_counter() - t0))
        dt = float(np.median(dts))
        out['e2e_runs_ms'] = [1e3 * x for x in dts]
        assert np.isfinite(pin_lnp.numpy()).all()
        api = ('ShardedEnsemble.set_state(host) + run(store, chain_out=pinned, log_prob_out=pinned) on every rank: each rank uploads and '
               'streams back only its own walkers' if e2.fused else 'ShardedEnsemble.set_state(host) + run(store) + get_own_chain()')
        out['e2e_phases_ms_rank0'] = {'set_state': 1e3 * (t_set - t0), 'run': 1e3 * (t_run - t_set), 'finish': 1e3 * (t_end - t_run)}
        h2d = count * D * 8 * ctx.world
        e2.close()
    out['e2e'] = {'value': W_total * steps / dt, 'unit': 'walker-steps/s', 'h2d_bytes_per_step': int(h2d / steps),
                  'd2h_bytes_per_step': int(W_total * (D + 1) * 8), 'api': api, 'frac_of_device_value': W_total * steps / dt / value}
    out['fused_exchange'] = fused
    return out, wl


def check_multigpu_bit_identical(ctx):
    """12 steps of a small shared ensemble against the same ensemble on ONE GPU (same seed; the RNG is keyed by the global
    walker index): chains and final replicas must be bit-identical.  Every rank runs the single-GPU chain itself."""
    from lightcurve_fitting_b200 import _capi
    from lightcurve_fitting_b200.parallel import ShardedEnsemble
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    L = _capi.lib()
    _capi.check(L.lcf_set_tuning_ex(32, 8, 1))           # pinned shape: FP32 partial sums are taken in the same order
    try:
        wl = workload(device_truth, 200, 'sc3')
        prob = wl.device_problem('fp32')
        W, nsteps = 64 * ctx.world * 4, 12
        p0 = wl.start(W, np.random.default_rng(3))
        single = EnsembleSampler(W, wl.ndim, prob, seed=77)
        single.run_mcmc(p0, nsteps, skip_initial_state_check=True)
        ref, ref_lp = single.get_chain(), single.get_log_prob()
        ens = ShardedEnsemble(prob, W, seed=77, rank=ctx.rank, world=ctx.world, exchange=ctx.args.exchange)
        ens.set_state(p0)
        ens.run(nsteps, store=True)
        ens.finish()
        ch, lp = ens.get_own_chain()
        first, count = ens.own_walkers()
        st = ens.sampler._state()
        ok = (np.array_equal(ch, ref[:, first:first + count]) and np.array_equal(lp, ref_lp[:, first:first + count])
              and np.array_equal(st.coords, ref[-1]) and np.array_equal(st.log_prob, ref_lp[-1]))
        ens.close()
    finally:
        _capi.check(L.lcf_set_tuning_ex(ctx.args.wpb, ctx.args.nw, ctx.args.cluster))
    return ctx.max_over_ranks(0. if ok else 1.) == 0.


def measure_cfg5(ctx, nlc_per_gpu):
    """Survey batch (BASELINE cfg5): light curves x ShockCooling4, 256 walkers, 200 + 200 steps, one launch per GPU; under torchrun
    every rank takes its round-robin share (parallel.shard_items), no data-path collective."""
    from lightcurve_fitting_b200 import synthetic
    from lightcurve_fitting_b200.bolometric import BatchSampler
    from lightcurve_fitting_b200.parallel import shard_items
    mine = shard_items(nlc_per_gpu * ctx.world, ctx.rank, ctx.world)
    npts = [int(n) for n in np.random.default_rng(0).integers(100, 301, nlc_per_gpu * ctx.world)]
    rng = np.random.default_rng(100 + ctx.rank)
    t0 = time.perf_counter()
    wls = [synthetic.synthetic_sc4(device_truth, npoints=npts[i], lc_index=i) for i in mine]
    t1 = time.perf_counter()
    probs = [w.device_problem('fp32') for w in wls]
    t2 = time.perf_counter()
    b = BatchSampler(probs, 256, seed=4)
    p0 = np.stack([w.start(256, rng) for w in wls])
    b.run(p0, 5, 5)                                     # warm-up launch
    pin_chain = ctx.torch.empty((len(wls), 200, 256, wls[0].ndim), dtype=ctx.torch.float64).pin_memory()
    ctx.barrier()
    clocks = ctx.clocks().start()
    ctx.barrier()
    te = time.perf_counter()
    b.run(p0, 200, 200)                                 # H2D start positions + one kernel; device time from the library's CUDA events
    chain = b.get_chain(out=pin_chain.numpy())          # D2H of the stored chain into pinned memory: part of the end-to-end number
    dt_e2e = ctx.max_over_ranks(time.perf_counter() - te)
    ms = ctx.max_over_ranks(b.last_ms)
    clk = clocks.stop()
    ok = bool(np.all(b.status == 0)) and bool(np.isfinite(chain).all())
    nlc = nlc_per_gpu * ctx.world
    spe = float(np.mean([w.planck_samples_per_eval() for w in wls]))
    sm_mhz = (clk or {}).get('sm_mhz') or 1965.0
    value = nlc * 256 * 400 / (ms * 1e-3)
    return {'workload': 'cfg5: %d light curves x ShockCooling4 (100-300 points, 9 filters), 256 walkers, 200+200 steps, one launch per GPU, '
                        '%d GPU(s), light curves dealt round-robin, no collective' % (nlc, ctx.world),
            'value': value, 'unit': 'walker-steps/s', 'kernel_ms': ms, 'gpu_launches': 1,
            'e2e': {'value': nlc * 256 * 400 / dt_e2e, 'unit': 'walker-steps/s', 'includes': 'H2D start positions, the launch, D2H of the stored chain'},
            'planck_samples_per_eval_mean': spe, 'samples_per_clk_sm': value / ctx.world * spe / (SMS * sm_mhz * 1e6),
            'roofline_frac': value / ctx.world * spe / (MUFU_LANES_PER_CLK_SM * SMS * sm_mhz * 1e6),
            'problem_build_s': t2 - t1, 'synthetic_data_s': t1 - t0, 'status_ok': ok,
            'acceptance': float(b.acceptance_fraction.mean()), 'clocks': clk}


def measure_small_configs(ctx):
    """cfg1 / cfg3 / cfg4 on one GPU (own CUDA-event timings inside the library, own clock samples)."""
    from lightcurve_fitting_b200 import synthetic
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    out = {}
    rng = np.random.default_rng(0)
    # cfg1: the reference's default use -- SN 2016bkv, ShockCooling4, 100 walkers x (1000 + 1000) steps
    # (FP32 first; the *_fp64 entries are the same runs in the drop-in's default precision = the reference's arithmetic)
    for window, tag, precision in (((57468., 57485.), 'cfg1_early_window', 'fp32'), (None, 'cfg1_full', 'fp32'),
                                   ((57468., 57485.), 'cfg1_early_window_fp64', 'fp64'), (None, 'cfg1_full_fp64', 'fp64')):
        wl = synthetic.example_sc4(window=window)
        prob = wl.device_problem(precision)
        s = EnsembleSampler(100, wl.ndim, prob, seed=1)
        p0 = wl.start(100, rng)
        s.run_mcmc(p0, 50, store=False)
        clocks = ctx.clocks().start(0.1)
        t0 = time.perf_counter()
        s.run_mcmc(p0, 1000)                       # burn-in
        s.reset()
        s.run_mcmc(None, 1000)
        ms = s.last_ms
        flat = s.flatchain                          # D2H of the chain
        dt = time.perf_counter() - t0
        clk = clocks.stop()
        sm_mhz = (clk or {}).get('sm_mhz') or 1965.0
        spe = wl.planck_samples_per_eval()
        v = 100 * 1000 / (ms * 1e-3)
        out[tag] = {'workload': '%s: SN 2016bkv, ShockCooling4, N=%d points, 100 walkers, 1000+1000 steps, %s' % (tag, len(wl.t), precision),
                    'dtype': 'f64' if precision == 'fp64' else 'f32', 'value': v, 'unit': 'walker-steps/s', 'us_per_half_step': 1e3 * ms / 2000., 'gpu_launches': s.last_launches,
                    'e2e': {'value': 100 * 2000 / dt, 'unit': 'walker-steps/s',
                            'includes': 'lightcurve_mcmc-shaped call: H2D start, burn-in, reset, sampling, flatchain D2H'},
                    'roofline_frac': v * spe / ((FP64_LANES_PER_CLK_SM / FP64_OPS_PER_SAMPLE if precision == 'fp64' else MUFU_LANES_PER_CLK_SM)
                                                * SMS * sm_mhz * 1e6), 'launch': prob.last_launch(),
                    'acceptance': float(s.acceptance_fraction.mean()), 'finite': bool(np.isfinite(flat).all()), 'clocks': clk}
    # cfg4: CompanionShocking3, N = 1000, 10^4 walkers
    wl = synthetic.synthetic_cs3(kasen_sifto_truth, npoints=1000)
    prob = wl.device_problem('fp32')
    s = EnsembleSampler(10_000, wl.ndim, prob, seed=3)
    s.run_mcmc(wl.start(10_000, rng), 20, store=False, skip_initial_state_check=True)
    clocks = ctx.clocks().start(0.1)
    s.run_mcmc(None, 100)
    clk = clocks.stop()
    sm_mhz = (clk or {}).get('sm_mhz') or 1965.0
    v = 10_000 * 100 / (s.last_ms * 1e-3)
    out['cfg4'] = {'workload': 'cfg4: CompanionShocking3 (Kasen + SiFTO), N=1000 points over U,B,V,g,r,i, 10^4 walkers', 'value': v,
                   'unit': 'walker-steps/s', 'kernel_avg_ms': s.last_ms / 200., 'gpu_launches': s.last_launches,
                   'roofline_frac': v * wl.planck_samples_per_eval() / (MUFU_LANES_PER_CLK_SM * SMS * sm_mhz * 1e6),
                   'launch': prob.last_launch(), 'acceptance': float(s.acceptance_fraction.mean()), 'clocks': clk}
    # cfg3 through calculate_bolometric: host table in -> result table out (500 epochs x 3-9 filters, 10 walkers, 200 + 100 steps)
    out['cfg3_calculate_bolometric'] = measure_cfg3(ctx)
    return out


def measure_cfg3(ctx, nepochs=500):
    import warnings
    from lightcurve_fitting_b200 import synthetic
    from lightcurve_fitting_b200.bolometric import calculate_bolometric
    lc = synthetic.sed_table(device_truth, nepochs, seed=2)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        calculate_bolometric(lc.copy(), res=1., seed=2)                       # warm-up (filter packing caches, kernel load)
        clocks = ctx.clocks().start(0.1)                                      # clocks: sampled during one more run of the same call ...
        calculate_bolometric(lc.copy(), res=1., seed=3)
        clk = clocks.stop()
        runs = []
        for rep in range(3):                                                  # ... the timed runs go without the poller: this path is host-bound (hundreds of
            t0 = time.perf_counter()                                          # small CUDA calls in 20 ms) and a 40 ms nvidia-smi loop stretches it two- to four-fold
            t, timing = calculate_bolometric(lc.copy(), res=1., seed=3, return_timing=True)
            runs.append((time.perf_counter() - t0, t, timing))
        e2e_runs = [r[0] for r in runs]
        runs.sort(key=lambda r: r[0])
        dt, t, timing = runs[1]                                               # the median run of three
    n = len(t)
    steps = 300
    return {'workload': 'cfg3: calculate_bolometric on a %d-row table (%d epochs x 3-9 filters), 10 walkers, 200+100 steps per epoch' % (len(lc), nepochs),
            'epochs_fitted': n, 'value': n * 10 * steps / (timing['sampling_ms'] * 1e-3), 'unit': 'walker-steps/s',
            'e2e': {'value': n * 10 * steps / dt, 'unit': 'walker-steps/s', 'seconds': dt, 'host_ms_per_epoch': 1e3 * (dt - timing['device_s']) / max(n, 1),
                    'includes': 'host table in -> result table out: grouping, flux/mag/luminosity conversions, batched least squares, '
                                'batched MCMC (one launch), pseudo-bolometric + Stefan-Boltzmann + percentiles on the device'},
            'phases_ms': timing, 'finite': bool(np.isfinite(t['temp_mcmc'].data).all()), 'clocks': clk, 'e2e_runs_s': e2e_runs,
            'clocks_note': 'sampled during a separate, untimed run of the same call'}


def run_ours(args):
    ctx = Ctx(args)
    torch, dist = ctx.torch, ctx.dist
    extras = {}
    bit_identical = None
    if ctx.world > 1:
        bit_identical = check_multigpu_bit_identical(ctx)
    if args.config == 'cfg5':
        main = measure_cfg5(ctx, args.nlc)
        line = {'metric': 'walker-steps/s', 'value': main['value'], 'unit': 'walker-steps/s', 'n_gpus': ctx.world, 'steps': 400,
                'warmup': 10, 'ms_per_step': main['kernel_ms'] / 400., 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': main['workload']}, 'e2e': main['e2e'],
                'gpu_launches': 1, 'clocks': main['clocks'],
                'roofline': {'bound': 'sfu', 'achieved': main['samples_per_clk_sm'], 'peak': MUFU_LANES_PER_CLK_SM, 'unit': 'Planck samples/clk/SM',
                             'frac': main['roofline_frac'], 'traffic': None},
                'cpu_baseline': None, 'detail': main}
        if ctx.rank == 0:
            print(json.dumps(line))
        if ctx.world > 1:
            dist.destroy_process_group()
        return
    strong = args.scaling == 'strong'
    main, wl = measure_cfg2(ctx, args.model, args.precision, args.walkers, args.steps, args.warmup, strong=strong)
    if not args.no_extras:
        k2 = max(4, args.steps // 2)
        if ctx.world == 1:
            if not (args.model == 'sc3' and args.precision == 'fp64'):
                extras['fp64'], _ = measure_cfg2(ctx, 'sc3', 'fp64', args.walkers, k2, 3)
                extras['fp64']['workload'] = workload_string('sc3', args.npoints, args.walkers) + ', FP64 mode (reference arithmetic)'
            if not (args.model == 'sc4' and args.precision == 'fp32'):
                extras['sc4'], _ = measure_cfg2(ctx, 'sc4', 'fp32', args.walkers, args.steps, 3)
                extras['sc4']['workload'] = workload_string('sc4', args.npoints, args.walkers)
            extras.update(measure_small_configs(ctx))
            extras['cfg5'] = measure_cfg5(ctx, args.nlc)
        else:
            if not strong:
                extras['strong'], _ = measure_cfg2(ctx, args.model, args.precision, WALKERS_PER_GPU, args.steps, 3, strong=True)
                extras['strong']['workload'] = ('cfg2 strong scaling: %s, %d walkers in TOTAL shared by %d GPUs'
                                                % (MODEL_NAMES[args.model], extras['strong']['walkers_total'], ctx.world))
            extras['cfg5'] = measure_cfg5(ctx, args.nlc)
    if ctx.rank != 0:
        if ctx.world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if ctx.world == 1 and not args.no_cpu:
        cpu = cpu_baseline_1core(workload(oracle_truth, args.npoints, args.model), args.cpu_budget)
        if not args.no_extras:        # the CPU reference path on config 1 (a bounded sample of it)
            from lightcurve_fitting_b200 import synthetic
            from tests import workloads as Wk
            extras['cfg1_cpu_baseline'] = cpu_baseline_1core(Wk.example_sc4(window=None), 6.)
            extras['cfg1_cpu_baseline']['note'] = ('full-length cfg1 is 100 walkers x (1000 + 1000) steps = 2.0e5 evaluations; at this rate %.0f s'
                                                   % (2.0e5 / extras['cfg1_cpu_baseline']['value']))
    D = main['ndim']
    out = {
        'metric': 'walker-steps/s', 'value': main['value'], 'unit': 'walker-steps/s', 'n_gpus': ctx.world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': main['ms_per_step'], 'higher_is_better': True, 'scaling': 'strong' if strong else 'weak',
        'vs_baseline': None, 'dtype': main['dtype'], 'data': 'synthetic',
        'config': {'workload': workload_string(args.model, args.npoints, args.walkers) if not strong else
                               'cfg2 strong scaling: %s, %d walkers in TOTAL' % (MODEL_NAMES[args.model], main['walkers_total']),
                   'walkers_total': main['walkers_total'], 'ndim': D, 'planck_samples_per_eval': main['planck_samples_per_eval'],
                   'parallelism': ('one ensemble, half-ensembles split over %d GPU(s); ' % ctx.world) +
                                  ('accepted walkers stored into the peer replicas over NVLink by the half-step kernel itself, '
                                   'device-side half-step flags, no NCCL on the data path' if args.exchange == 'p2p' else
                                   'NCCL all-gather of the colour block per half-step'),
                   'l2': 'working set per CTA (light curve 72 KB + bank 3 KB) is L2 / shared-memory resident by design; the walker state '
                         'streams through HBM once per step; no inter-iteration flush needed (compute-bound: %d Planck samples per 72 B)'
                         % main['planck_samples_per_eval']},
        'log_posterior_evals_per_s': main['value'],
        'e2e': main.get('e2e'), 'gpu_launches': main['gpu_launches'], 'clocks': main['clocks'], 'roofline': main.get('roofline'),
        'cpu_baseline': cpu, 'launch': main['launch'],
    }
    if 'e2e_phases_ms_rank0' in main:
        out['e2e_phases_ms_rank0'] = main['e2e_phases_ms_rank0']
    if bit_identical is not None:
        out['multigpu_bit_identical'] = bool(bit_identical)
    if extras:
        out['extras'] = extras
    print(json.dumps(out))
    if ctx.world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='cfg2', choices=['cfg2', 'cfg5'])
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'], help='cfg2: walkers per GPU fixed (weak) or in total (strong)')
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'fp64'])
    ap.add_argument('--walkers', type=int, default=WALKERS_PER_GPU)
    ap.add_argument('--npoints', type=int, default=NPOINTS)
    ap.add_argument('--nlc', type=int, default=1250, help='cfg5: light curves per GPU')
    ap.add_argument('--model', default='sc3', choices=['sc3', 'sc4'])
    ap.add_argument('--wpb', type=int, default=0)
    ap.add_argument('--cluster', type=int, default=0)
    ap.add_argument('--exchange', default='p2p', choices=['p2p', 'nccl'], help='multi-GPU exchange of the shared ensemble')
    ap.add_argument('--nw', type=int, default=0)
    ap.add_argument('--cpu-budget', type=float, default=15.)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-extras', action='store_true')
    args = ap.parse_args()
    global MODEL
    MODEL = args.model
    if args.impl == 'reference':
        run_reference(args)
    else:
        import __graft_entry__ as g
        g.build()
        run_ours(args)


if __name__ == '__main__':
    main()
