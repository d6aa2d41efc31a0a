#!/usr/bin/env python
"""Benchmark of the MCMC hot path (BASELINE.json metric: walker-steps/s, = log-posterior evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|fp64]

Workload (config.workload): BASELINE.json configs[1] -- ShockCooling3 on a synthetic 2000-point, 8-filter light
curve, 10^5 walkers per GPU.  A "step" is one stretch-move iteration of the whole ensemble (two fused half-step
kernels; W log-posterior evaluations).  N > 1: ONE ensemble of N x 10^5 walkers split across the GPUs with an
all-gather of the updated half-ensemble after every half-step (weak scaling).

`value`    device-resident throughput (walker ensemble, light curve and filter bank already in HBM).
`e2e`      the same metric through the public Python API (`EnsembleSampler.run_mcmc` + `get_chain`) with host
           buffers: the H2D copy of the start positions and the D2H read-back of the chain are inside the timed
           region.
`roofline` the fused kernel against the arithmetic (SFU) roofline: one Planck sample = 1 MUFU.EX2 + 1 MUFU.RCP
           + FMUL/FADD/FFMA; the MUFU pipe issues 16 lanes/clk/SM, so peak = 8 samples/clk/SM x 148 SMs x the SM
           clock measured during the run.  HBM traffic (chain write-back) is reported beside it.
`cpu_baseline` the oracle (numpy port of the reference) timed on this box's host cores on a bounded sample.
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
FP64_OPS_PER_SAMPLE = 17.75
BALANCED_SAMPLES_PER_CLK_SM = 13.5    # FMA+MUFU balanced bound of the Planck-sample formulation
SMS = 148


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
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
    """Noiseless synthetic light curve from the device model itself (no oracle on the measured arm)."""
    from lightcurve_fitting_b200 import models as M
    from lightcurve_fitting_b200.filters import filtdict
    m = getattr(M, model_name)(redshift=z)
    return np.asarray(m(np.asarray(t, float), [filtdict[n] for n in filter_names], *params), float)


def oracle_truth(model_name, t, filter_names, params, z):
    from tests import workloads as W
    return W.oracle_truth(model_name, t, filter_names, params, z)


MODEL = 'sc3'
# dram__bytes_read.sum + dram__bytes_write.sum of one k_pass<3,float> launch of this workload (ncu --set full capture
# summarised in profiles/r02_ncu_sc3_fp32.txt); per launch, like roofline.achieved
NCU_DRAM_BYTES_PER_LAUNCH = 6782208


def workload(truth, npoints=NPOINTS):
    from lightcurve_fitting_b200 import synthetic
    if MODEL == 'sc4':      # the same shape with the MSW23 model (two blackbody syntheses per point)
        return synthetic.synthetic_sc4(truth, npoints=npoints, seed=1, filters=['U', 'B', 'V', 'R', 'I', 'g', 'r', 'i'])
    return synthetic.synthetic_sc3(truth, npoints=npoints, seed=1)


# ---------------------------------------------------------------------------------------------
# CPU arms (oracle): cpu_baseline (1 core, bounded sample) and --impl reference (all cores)
# ---------------------------------------------------------------------------------------------
_POOL_LP = None


def _pool_init(npoints):
    global _POOL_LP
    from tests import workloads as W
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    wl = workload(oracle_truth, npoints)
    _POOL_LP = W.oracle_log_posterior(wl)


def _pool_eval(p):
    return float(_POOL_LP(np.asarray(p)))


def cpu_baseline_1core(wl, budget_s=15.):
    """The faithful oracle (per-point Python loop, one un-vectorised log_posterior call per walker -- how the
    reference runs under emcee, fitting.py:130) on one host core, for ~budget_s seconds."""
    from tests import workloads as W
    lp = W.oracle_log_posterior(wl)
    rng = np.random.default_rng(0)
    P = wl.start(64, rng)
    n, t0 = 0, time.perf_counter()
    while True:
        lp(P[n % len(P)])
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 4096:
            break
    dt = time.perf_counter() - t0
    return {'value': n / dt, 'unit': 'walker-steps/s', 'cores': 1, 'kind': 'port',
            'sample': '%d ShockCooling3 log-posterior evaluations (N=%d points) in %.1f s, serial, one core' % (n, len(wl.t), dt)}


def run_reference(args):
    """--impl reference: the oracle port (the reference itself cannot be imported here: astropy, emcee and
    extinction are absent from the image) with every host core, emcee-style serial chain replaced by a process
    pool over walkers (emcee's `pool=` option; generous to the CPU)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_step = cores * 2                     # walkers evaluated per "step" (bounded sample of the 1e5-walker step)
    rng = np.random.default_rng(0)
    wl = workload(oracle_truth, NPOINTS)
    ctx = mp.get_context('fork')
    with ctx.Pool(cores, initializer=_pool_init, initargs=(NPOINTS,)) as pool:
        for _ in range(args.warmup):
            pool.map(_pool_eval, list(wl.start(cores, rng)))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_pool_eval, list(wl.start(per_step, rng)))
        dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = '%d walkers per step (of %d), %d steps, multiprocessing pool over walkers' % (per_step, WALKERS_PER_GPU, args.steps)
    print(json.dumps({
        'impl': 'reference', 'metric': 'walker-steps/s', 'value': value, 'unit': 'walker-steps/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'cfg2: ShockCooling3, synthetic 2000-point 8-filter light curve, 1e5 walkers',
                   'npoints': NPOINTS, 'walkers_per_step_sampled': per_step},
        'cpu_baseline': {'value': value, 'unit': 'walker-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'walker-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from lightcurve_fitting_b200 import _capi
    from lightcurve_fitting_b200.parallel import ShardedEnsemble
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    _capi.check(_capi.lib().lcf_set_device(local))
    if args.wpb or args.nw or args.cluster:
        _capi.check(_capi.lib().lcf_set_tuning_ex(args.wpb, args.nw, args.cluster))

    wl = workload(device_truth, args.npoints)
    prob = wl.device_problem(args.precision)
    W_total = args.walkers * world
    D = wl.ndim
    rng = np.random.default_rng(1)
    p0 = wl.start(W_total, rng)

    ens = ShardedEnsemble(prob, W_total, seed=1234, rank=rank, world=world, exchange=args.exchange)
    ens.set_state(p0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # L2 note: per step each CTA re-reads the 16 KB light curve + 3 KB bank (L2/L1 resident by design) and the
    # kernel streams 2 x W x D x 8 B of walker state; inputs that matter are register/SMEM resident.
    ens.run(args.warmup, store=False)
    ens.finish()
    ens.reserve(args.steps)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ens.sampler._timing()
    launches0 = ens.sampler.last_launches
    ev0.record()                      # the library launches on torch's current stream (lcf_ensemble_set_stream)
    ens.run(args.steps, store=True)   # chain write-back to HBM inside the timed region
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    ens.finish()
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = W_total * args.steps / (ms * 1e-3)
    ens.sampler._timing()
    # kernels launched in the timed region, counted by the library (a half-step is one launch, or two when the last wave
    # of CTAs is split off onto its own low-priority stream); half_step() accumulates, run() restarts the count
    launches = ens.sampler.last_launches - (launches0 if not ens.fused else 0)
    half_steps = 2 * args.steps

    # ---- e2e through the public API with host (pinned) buffers, rank-local ensemble -------------------------
    pin_in = torch.from_numpy(np.ascontiguousarray(p0)).pin_memory()
    pin_chain = torch.empty((args.steps, W_total, D), dtype=torch.float64).pin_memory()
    pin_lnp = torch.empty((args.steps, W_total), dtype=torch.float64).pin_memory()
    ens.close()
    del ens
    if world == 1:
        s = EnsembleSampler(W_total, D, prob, seed=99)
        s.run_mcmc(pin_in.numpy(), 1, skip_initial_state_check=True, store=False)   # warm-up of the API path
        s.reset()
        barrier()
        t0 = time.perf_counter()
        chain, lnp = pin_chain.numpy(), pin_lnp.numpy()
        # H2D start positions + K steps; every finished step is streamed D2H into the pinned buffers while the next ones run
        s.run_mcmc(pin_in.numpy(), args.steps, skip_initial_state_check=True, chain_out=chain, log_prob_out=lnp)
        dt = time.perf_counter() - t0
        assert np.isfinite(lnp).all() and chain.shape == (args.steps, W_total, D)
        api = 'EnsembleSampler.run_mcmc(host start positions, chain_out=pinned, log_prob_out=pinned): chain streamed to the host'
    else:
        e2 = ShardedEnsemble(prob, W_total, seed=99, rank=rank, world=world, exchange=args.exchange)
        e2.set_state(pin_in.numpy())
        e2.run(1, store=False)
        e2.finish()
        barrier()
        t0 = time.perf_counter()
        e2.set_state(pin_in.numpy())                                                  # H2D start positions
        t_set = time.perf_counter()
        first, count = e2.own_walkers()
        if e2.fused:      # this rank's walkers streamed D2H into pinned buffers while the next steps run
            chain = pin_chain.numpy().reshape(-1)[:args.steps * count * D].reshape(args.steps, count, D)
            lnp = pin_lnp.numpy().reshape(-1)[:args.steps * count].reshape(args.steps, count)
            e2.run(args.steps, store=True, chain_out=chain, log_prob_out=lnp)
            t_run = time.perf_counter()
            e2.finish()
        else:
            e2.run(args.steps, store=True)
            t_run = time.perf_counter()
            e2.finish()
            chain, lnp = e2.get_own_chain(pin_chain.numpy(), pin_lnp.numpy())            # D2H of this rank's walkers
        if os.environ.get('LCF_E2E_TIMING') and rank == 0:
            print('[e2e] set_state %.1f ms, run %.1f ms, finish %.1f ms' % (1e3 * (t_set - t0), 1e3 * (t_run - t_set),
                                                                             1e3 * (time.perf_counter() - t_run)), file=sys.stderr, flush=True)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        api = ('ShardedEnsemble.set_state(host) + run(store, chain_out=pinned, log_prob_out=pinned) on every rank: own walkers '
               'streamed to the host' if e2.fused else
               'ShardedEnsemble.set_state(host) + run(store) + get_own_chain() on every rank')
        e2.close()
    e2e = {'value': W_total * args.steps / dt, 'unit': 'walker-steps/s',
           'h2d_bytes_per_step': int(pin_in.numel() * 8 / args.steps),
           'd2h_bytes_per_step': int(W_total * (D + 1) * 8),
           'api': api}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    samples_per_eval = wl.planck_samples_per_eval()
    samples_per_s_gpu = value * samples_per_eval / world
    sm_mhz = (clk or {}).get('sm_mhz') or 1965.0
    # Roofline: the XU (MUFU) pipe.  One MUFU.EX2 per Planck sample is irreducible (the reciprocal and everything else run
    # on the FMA pipe), and the pipe issues 16 lanes/clk/SM (measured: 15.7, profiles/r02_microbench_loops.txt).
    peak_samples = MUFU_LANES_PER_CLK_SM * SMS * sm_mhz * 1e6
    fp64 = args.precision == 'fp64'
    if fp64:        # FP64 mode has no MUFU for doubles: bound = FP64 pipe, 64 lanes/clk/SM over the loop's FP64 operations per sample
        peak_samples = 64. / FP64_OPS_PER_SAMPLE * SMS * sm_mhz * 1e6
    balanced_samples = BALANCED_SAMPLES_PER_CLK_SM * SMS * sm_mhz * 1e6
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    chain_bytes_per_step = args.walkers * (D + 1) * 8
    hbm_gbs = chain_bytes_per_step * args.steps / (ms * 1e-3) / 1e9
    roofline = {
        'bound': ('fp64 pipe (no dense contraction, ~72 B of HBM per walker-step)' if fp64 else
                  'sfu (XU pipe: one MUFU.EX2 per Planck sample; no dense contraction, ~72 B of HBM per walker-step)'),
        'kernel': 'lcf::k_pass<%d,%s>' % (3 if MODEL == 'sc3' else 4, 'float' if args.precision == 'fp32' else 'double'),
        'achieved': samples_per_s_gpu / 1e9, 'peak': peak_samples / 1e9, 'unit': 'GPlanck-samples/s',
        'frac': samples_per_s_gpu / peak_samples,
        'peak_basis': (('64 FP64 lanes/clk/SM / %.2f FP64 operations per Planck sample of the shipped loop (table exp2 7 + range '
                        'reduction 3 + x 1 + minus-one 1 + quad-shared reciprocal and sums 5.75) x 148 SMs x %.0f MHz; '
                        'MEASURED_PEAKS.json has no FP64 entry (tools/microbench/pipes.cu: DFMA 57 lanes/clk/SM)'
                        % (FP64_OPS_PER_SAMPLE, sm_mhz)) if fp64 else
                       ('16 MUFU lanes/clk/SM x 148 SMs x %.0f MHz (SM clock measured during the timed region), 1 MUFU.EX2 per '
                        'Planck sample; MEASURED_PEAKS.json has no SFU entry, the pipe rate is confirmed by '
                        'tools/microbench/loops.cu (15.7 lanes/clk/SM)' % sm_mhz)),
        'samples_per_clk_sm': samples_per_s_gpu / (SMS * sm_mhz * 1e6),
        'frac_vs_survey_balanced_bound': samples_per_s_gpu / balanced_samples,
        'survey_balanced_bound': '13.5 samples/clk/SM (SURVEY.md 8(d))',
        'inner_loop_in_isolation_samples_per_clk_sm': 14.3,
        'algorithmic_per_unit': '4 FP32 flops + 2 transcendentals per Planck sample (SURVEY.md 8(d)); %d samples per '
                                'log-posterior; shipped loop: 1 MUFU + ~6.3 FP32 lane-ops per sample' % samples_per_eval,
        'fp32_tflops': 4 * samples_per_s_gpu / 1e12,
        'fp32_peak_tflops': 2 * 128 * SMS * sm_mhz * 1e6 / 1e12,
        'kernel_avg_ms': ms / half_steps, 'launches_per_half_step': launches / half_steps,
        'hbm_gbs_chain_writeback': hbm_gbs, 'hbm_peak_gbs_measured': peaks.get('hbm_gbs'),
        'traffic': NCU_DRAM_BYTES_PER_LAUNCH if (MODEL == 'sc3' and args.precision == 'fp32' and args.walkers == WALKERS_PER_GPU) else None,
    }
    cpu = cpu_baseline_1core(workload(oracle_truth, args.npoints), args.cpu_budget) if world == 1 and not args.no_cpu else None
    out = {
        'metric': 'walker-steps/s', 'value': value, 'unit': 'walker-steps/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32' if args.precision == 'fp32' else 'f64', 'data': 'synthetic',
        'config': {'workload': 'cfg2: %s, synthetic %d-point 8-filter light curve, %d walkers per GPU'
                               % ('ShockCooling3' if MODEL == 'sc3' else 'ShockCooling4', args.npoints, args.walkers),
                   'walkers_total': W_total, 'ndim': D, 'planck_samples_per_eval': samples_per_eval,
                   'parallelism': ('one ensemble, half-ensembles split over %d GPU(s); ' % world) +
                                  ('accepted walkers stored into the peer replicas over NVLink by the half-step kernel itself, '
                                   'device-side half-step flags, no NCCL on the data path' if args.exchange == 'p2p' else
                                   'NCCL all-gather of the colour block per half-step'),
                   'l2': 'working set per CTA (light curve 24 KB + bank 3 KB) is L2/SMEM resident by design; walker '
                         'state streamed once per step; no inter-iteration flush needed (compute-bound: %.0f samples '
                         'per 72 B)' % samples_per_eval},
        'log_posterior_evals_per_s': value,
        'e2e': e2e, 'gpu_launches': launches, 'clocks': clk, 'roofline': roofline, 'cpu_baseline': cpu,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=60)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='fp32', choices=['fp32', 'fp64'])
    ap.add_argument('--walkers', type=int, default=WALKERS_PER_GPU)
    ap.add_argument('--npoints', type=int, default=NPOINTS)
    ap.add_argument('--model', default='sc3', choices=['sc3', 'sc4'])
    ap.add_argument('--wpb', type=int, default=0)
    ap.add_argument('--cluster', type=int, default=0)
    ap.add_argument('--exchange', default='p2p', choices=['p2p', 'nccl'], help='multi-GPU exchange of the shared ensemble')
    ap.add_argument('--nw', type=int, default=0)
    ap.add_argument('--cpu-budget', type=float, default=15.)
    ap.add_argument('--no-cpu', action='store_true')
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
