"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): log-posterior rtol 1e-9 in FP64 mode, 1e-4 in FP32 mode; chains driven
with identical stretch-move draws reproduce the oracle's emcee-order chains to the same tolerance.
"""
import numpy as np
import pytest

from tests import workloads as W

pytestmark = pytest.mark.gpu

RTOL = {'fp64': 1e-9, 'fp32': 1e-4}


def _params(wl, n, seed, widen=0.):
    """Parameter sets inside the start box, optionally widened so that some fall outside the priors."""
    rng = np.random.default_rng(seed)
    span = wl.p_up - wl.p_lo
    return (wl.p_lo - widen * span) + rng.random((n, wl.ndim)) * span * (1. + 2. * widen)


def _check_logpost(wl, precision, n=24, seed=0, widen=0.):
    prob = wl.device_problem(precision)
    P = _params(wl, n, seed, widen)
    lp = W.oracle_log_posterior(wl)
    want = np.array([lp(p) for p in P])
    got = prob.log_posterior(P)
    fin = np.isfinite(want)
    assert np.array_equal(np.isneginf(want), np.isneginf(got))
    np.testing.assert_allclose(got[fin], want[fin], rtol=RTOL[precision])
    return got, want


WORKLOADS = {
    'sc4_example': lambda: W.example_sc4(),
    'sc4_example_sigma_rel': lambda: W.example_sc4(npoints=80, use_sigma=True),
    'sc4_example_sigma_abs': lambda: W.example_sc4(npoints=80, use_sigma=True, sigma_type='absolute'),
    'sc3_synth': lambda: W.synthetic_sc3(npoints=160),
    'sc3_synth_sigma': lambda: W.synthetic_sc3(npoints=96, use_sigma=True),
    'cs3_synth': lambda: W.synthetic_cs3(npoints=150),
    'sed': lambda: W.sed_epoch(np.random.default_rng(7)),
    'sed_sigma': lambda: W.sed_epoch(np.random.default_rng(8), use_sigma=True),
}


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
@pytest.mark.parametrize('name', sorted(WORKLOADS))
def test_log_posterior_parity(name, precision):
    _check_logpost(WORKLOADS[name](), precision)


@pytest.mark.parametrize('shape', [(32, 16, 1), (32, 4, 2), (8, 4, 4), (2, 8, 8), (1, 2, 8), (16, 16, 8)])
@pytest.mark.parametrize('name,precision', [('sc4_example', 'fp32'), ('sc3_synth_sigma', 'fp32'), ('cs3_synth', 'fp64'),
                                            ('sc3_synth', 'fp64'), ('sed', 'fp32')])
def test_log_posterior_parity_forced_launch_shapes(name, precision, shape):
    """Every (walkers per CTA, warps per CTA, cluster size) decomposition gives the oracle's log-posterior: partial
    tiles, chunks that straddle filters, clusters larger than the chunk count, DSMEM reduction of the partials."""
    from lightcurve_fitting_b200._capi import lib, check
    check(lib().lcf_set_tuning_ex(*shape))
    try:
        _check_logpost(WORKLOADS[name](), precision, n=37, seed=11, widen=0.1)
    finally:
        check(lib().lcf_set_tuning_ex(0, 0, 0))


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
def test_log_posterior_outside_prior(precision):
    """-inf outside the strict prior bounds (models.py:1055-1059), likelihood skipped (fitting.py:125)."""
    got, want = _check_logpost(W.example_sc4(npoints=40), precision, n=64, seed=3, widen=0.6)
    assert np.isneginf(want).any() and np.isfinite(want).any()


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
def test_points_before_explosion_contribute_zero_model(precision):
    """power() returns 0 for t <= t_exp (models.py:42-48): move t_0 into the middle of the data."""
    wl = W.example_sc4(npoints=60)
    wl.priors_spec[4] = ('uniform', 57460., 57480.)
    wl.p_lo[4], wl.p_up[4] = 57470., 57475.
    _check_logpost(wl, precision, n=16, seed=5)


def _shock_variants():
    from lightcurve_fitting_b200 import synthetic
    base = W.example_sc4(npoints=50)
    out = []
    for model_name, kw, lo, hi in [
        ('ShockCooling', {}, [0.5, 0.1, 0.1, 1., 57468.5], [2., 2., 10., 10., 57468.7]),
        ('ShockCooling', {'n': 3.}, [0.5, 0.1, 0.1, 1., 57468.5], [2., 2., 10., 10., 57468.7]),
        ('ShockCooling', {'RW': True}, [0.5, 0.1, 0.1, 1., 57468.5], [2., 2., 10., 10., 57468.7]),
        ('ShockCooling2', {}, [10., 0.5, 2., 57468.5], [30., 5., 10., 57468.7]),
    ]:
        pri = [('uniform', -1e3, 1e5)] * len(lo)
        out.append(synthetic.Workload('variant-' + model_name, model_name, base.t, base.filter_names, base.y, base.dy, pri,
                                      lo, hi, z=base.z, model_kwargs=kw))
    return out


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
def test_shockcooling_variants(precision):
    for wl in _shock_variants():
        _check_logpost(wl, precision, n=12, seed=11)


def _companion_variants():
    from lightcurve_fitting_b200 import synthetic
    base = W.synthetic_cs3(npoints=90)
    out = []
    tpk = 58000.
    for model_name, lo, hi in [
        ('CompanionShocking', [tpk - 17.5, 0.05, 0.5, tpk - 0.5, 0.9, 0.8, 0.8, 0.8], [tpk - 17., 0.15, 2., tpk + 0.5, 1.1, 1.2, 1.2, 1.2]),
        ('CompanionShocking2', [tpk - 17.5, 0.05, 0.5, tpk - 0.5, 0.9, -0.5, -0.5], [tpk - 17., 0.15, 2., tpk + 0.5, 1.1, 0.5, 0.5]),
    ]:
        pri = [('uniform', -1e6, 1e6)] * len(lo)
        out.append(synthetic.Workload('variant-' + model_name, model_name, base.t, base.filter_names, base.y, base.dy, pri,
                                      lo, hi, z=base.z))
    return out


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
def test_companion_variants(precision):
    for wl in _companion_variants():
        _check_logpost(wl, precision, n=12, seed=13)


@pytest.mark.parametrize('name', ['sc4_example', 'sc3_synth', 'cs3_synth'])
def test_model_call_pointwise_and_grid(name):
    """Model.__call__(t, f, *params): pointwise and grid modes (models.py:1161-1164, fitting.py:350-352)."""
    wl = WORKLOADS[name]()
    model = wl.model('fp64')
    omodel, _, _ = W.oracle_for(wl)
    p = 0.5 * (wl.p_lo + wl.p_up)
    f, of = np.array(wl.filters(), dtype=object), W.oracle_filters(wl.filter_names)
    got = model(wl.t, f, *p)
    want = omodel(wl.t, of, *p)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
    # grid: 3 filters x 7 times x 4 parameter sets
    uf = list(dict.fromkeys(wl.filter_names))[:3]
    tg = np.linspace(wl.t.min(), wl.t.max(), 7)
    ps = np.array([wl.p_lo + (wl.p_up - wl.p_lo) * u for u in (0.1, 0.4, 0.6, 0.9)]).T
    got = model(tg, [f[wl.filter_names.index(n)] for n in uf], *ps)
    want = omodel(tg, W.oracle_filters(uf), *ps)
    assert got.shape == want.shape == (3, 7, 4)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())


def test_log_likelihood_method():
    wl = W.example_sc4(npoints=50, use_sigma=True)
    model = wl.model('fp64')
    omodel, _, _ = W.oracle_for(wl)
    p = 0.5 * (wl.p_lo + wl.p_up)
    got = model.log_likelihood(wl.lc(), p, use_sigma=True, sigma_type='relative')
    want = omodel.log_likelihood(wl.t, W.oracle_filters(wl.filter_names), wl.y, wl.dy, p, use_sigma=True)
    np.testing.assert_allclose(got, want, rtol=1e-9)
    with pytest.raises(Exception, match='sigma_type'):
        model.log_likelihood(wl.lc(), p, use_sigma=True, sigma_type='bogus')


def test_blackbody_to_filters_and_planck():
    from lightcurve_fitting_b200 import models as M
    from lightcurve_fitting_b200.filters import filtdict
    from oracle import reference_port as rp
    names = ['U', 'B', 'g', 'r', 'UVW2', 'F444W']
    T = np.array([5., 9., 14., 22., 40., 3.])
    R = np.array([1., 2., 0.5, 3., 1.5, 10.])
    f = [filtdict[n] for n in names]
    of = W.oracle_filters(names)
    for kw in ({}, {'z': 0.05}, {'z': 0.01, 'cutoff_freq': 700.}, {'ebv': 0.2}):
        np.testing.assert_allclose(M.blackbody_to_filters(f, T, R, **kw), rp.blackbody_to_filters(of, T, R, **kw), rtol=1e-9)
        np.testing.assert_allclose(M.blackbody_to_filters(f[:2], T, R, **kw), rp.blackbody_to_filters(of[:2], T, R, **kw),
                                   rtol=1e-9)
    nu = np.linspace(100., 1500., 57)
    np.testing.assert_allclose(M.planck_fast(nu, 12., 3.), rp.planck_fast(nu, 12., 3.), rtol=1e-9)
    np.testing.assert_allclose(M.planck_fast(nu, T, R, 600.), rp.planck_fast(nu, T, R, 600.), rtol=1e-9)
    # zero / negative temperature: power() semantics give exactly 0
    assert np.all(M.planck_fast(nu, np.array([0., -3.]), np.array([1., 1.])) == 0.)


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
@pytest.mark.parametrize('name', ['sc4_example', 'sc3_synth_sigma', 'sed'])
def test_chain_replay_matches_oracle(name, precision):
    """Identical (split, z, partner, log u) draws => the chain reproduces the oracle's emcee-order chain."""
    from oracle import reference_port as rp
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    wl = WORKLOADS[name]()
    if name == 'sc4_example':
        wl = W.example_sc4(npoints=40)
    nw, nsteps = 2 * wl.ndim + 6, 12
    _, _, lp = W.oracle_for(wl)
    rs = np.random.RandomState(42)
    p0 = wl.p_lo + rs.rand(nw, wl.ndim) * (wl.p_up - wl.p_lo)
    ref = rp.StretchReplay(nw, wl.ndim, lp, random_state=rs)
    ref.run_mcmc(p0, nsteps, record=True)
    s = EnsembleSampler(nw, wl.ndim, wl.device_problem(precision), seed=0)
    state = s.run_replay(p0, ref.draws)
    got, want = s.get_chain(), ref.get_chain()
    assert got.shape == want.shape == (nsteps, nw, wl.ndim)
    if precision == 'fp64':
        np.testing.assert_allclose(got, want, rtol=1e-9)
        np.testing.assert_allclose(s.get_log_prob(), ref.get_log_prob(), rtol=1e-9)
        np.testing.assert_array_equal(s.acceptance_fraction, ref.acceptance_fraction)
    else:
        # FP32 log-posteriors differ from the oracle's at the 1e-4 level, so an accept decision that was a close call may
        # flip, after which the two chains legitimately part.  Criterion: every step before the first divergence matches at
        # 1e-4, and every walker that diverges AT that step was such a close call in the oracle's run:
        # |lnpdiff - ln u| < 1e-3 |lnp'| (the oracle records the margin of every decision).
        close = np.all(np.isclose(got, want, rtol=1e-4, atol=0), axis=2)              # [step, walker]
        bad_steps = np.flatnonzero(~close.all(axis=1))
        if len(bad_steps):
            s0 = bad_steps[0]
            np.testing.assert_allclose(got[:s0], want[:s0], rtol=1e-4)
            np.testing.assert_allclose(s.get_log_prob()[:s0], ref.get_log_prob()[:s0], rtol=1e-4)
            margin, nlp = np.full(nw, np.inf), np.ones(nw)
            for h in ref.draws[s0]['halves']:
                margin[h['walkers']], nlp[h['walkers']] = h['margin'], h['nlp']
            # a walker can also differ at step s0 because its partner (of the other colour) flipped earlier in the same step
            flipped = np.abs(margin) < 1e-3 * np.abs(nlp)
            assert flipped[~close[s0]].any(), 'chains diverge at step %d without a close accept decision' % s0
            first_half = ref.draws[s0]['halves'][0]['walkers']
            assert np.all(flipped[np.intersect1d(np.flatnonzero(~close[s0]), first_half)])
        else:
            np.testing.assert_allclose(s.get_log_prob(), ref.get_log_prob(), rtol=1e-4)
    np.testing.assert_allclose(state.coords, got[-1])
    assert s.chain.shape == (nw, nsteps, wl.ndim) and s.flatchain.shape == (nsteps * nw, wl.ndim)


def test_native_rng_sampler_statistics():
    """Device Philox stretch move: posterior medians / 68 % intervals agree with the oracle sampler (SED case,
    cheap enough for the oracle to run long chains)."""
    from oracle import reference_port as rp
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    wl = W.sed_epoch(np.random.default_rng(21))
    nw, nburn, nsteps = 32, 300, 600
    _, _, lp = W.oracle_for(wl)
    rs = np.random.RandomState(5)
    p0 = wl.p_lo + rs.rand(nw, wl.ndim) * (wl.p_up - wl.p_lo)
    ref = rp.StretchReplay(nw, wl.ndim, lp, random_state=rs)
    pos, lnp, _ = ref.run_mcmc(p0, nburn)
    ref.reset()
    ref.run_mcmc(pos, nsteps, log_prob0=lnp)
    s = EnsembleSampler(nw, wl.ndim, wl.device_problem('fp64'), seed=77)
    s.run_mcmc(p0, nburn)
    s.reset()
    s.run_mcmc(None, nsteps)
    a, b = s.flatchain, ref.flatchain
    qa, qb = np.percentile(a, [16, 50, 84], axis=0), np.percentile(b, [16, 50, 84], axis=0)
    width = qb[2] - qb[0]
    assert np.all(np.abs(qa[1] - qb[1]) < 0.25 * width)          # medians within a quarter of the 68 % width
    assert np.all(np.abs((qa[2] - qa[0]) / width - 1.) < 0.35)    # interval widths agree
    assert 0.1 < s.acceptance_fraction.mean() < 0.9


def test_nan_posterior_raises_like_emcee():
    """A NaN log-probability raises ValueError (emcee) -- negative radius makes L ** 0.5 NaN (models.py:268)."""
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    wl = W.example_sc4(npoints=30)
    wl.priors_spec = [('uniform', -100., 100.)] * 4 + [('uniform', 57460., 57470.)]
    prob = wl.device_problem('fp64')
    rng = np.random.default_rng(0)
    p0 = wl.start(16, rng)
    p0[3, 3] = -2.
    s = EnsembleSampler(16, wl.ndim, prob, seed=0)
    with pytest.raises(ValueError, match='NaN'):
        s.run_mcmc(p0, 2)
    with pytest.raises(RuntimeError, match='fewer walkers'):
        EnsembleSampler(6, wl.ndim, prob, seed=0)


def test_batched_ensembles_match_single():
    """lcf_batch (one CTA per problem, whole chain in one launch) == the per-problem half-step launches."""
    from lightcurve_fitting_b200.bolometric import BatchSampler
    rng = np.random.default_rng(3)
    wls = [W.sed_epoch(rng) for _ in range(7)]
    probs = [w.device_problem('fp64') for w in wls]
    nw, nburn, nsteps = 10, 20, 15
    p0 = np.stack([w.start(nw, rng) for w in wls])
    b = BatchSampler(probs, nw, seed=9).run(p0, nburn, nsteps)
    chain = b.get_chain()
    assert chain.shape == (7, nsteps, nw, 2)
    assert np.all(b.status == 0)
    # every stored position must carry its own log-posterior
    lnp = b.get_log_prob()
    for i, w in enumerate(wls):
        flat = chain[i].reshape(-1, 2)
        np.testing.assert_allclose(probs[i].log_posterior(flat), lnp[i].reshape(-1), rtol=1e-12)
        lp = W.oracle_log_posterior(w)
        np.testing.assert_allclose(lnp[i, -1], [lp(p) for p in chain[i, -1]], rtol=1e-9)
    assert 0.05 < b.acceptance_fraction.mean() < 0.95


@pytest.mark.parametrize('nw', [8, 10, 21, 22])
def test_batched_ensembles_look_ahead_rounds_match_half_steps(nw, monkeypatch):
    """Small batched ensembles (n0 + 2 n1 <= 32 virtual walkers: W <= 21) run every step as ONE log-posterior pass over the proposals
    of the first colour and BOTH candidate proposals of every second-colour walker, followed by the two accept phases (k_chain's
    look-ahead rounds); LCF_CHAIN_LA=0 keeps two dependent half-steps per step.  Same draws, same proposals, same accept test: the
    chains and the acceptance counts must be identical (the stored log-posteriors only to rounding: the two schedules evaluate
    with different walkers-per-CTA shapes), in both precisions; 22 walkers no longer fit and both settings run half-steps."""
    from lightcurve_fitting_b200.bolometric import BatchSampler
    for precision, rtol in (('fp64', 1e-13), ('fp32', 1e-6)):
        rng = np.random.default_rng(3)
        wls = [W.sed_epoch(rng) for _ in range(7)]
        probs = [w.device_problem(precision) for w in wls]
        p0 = np.stack([w.start(nw, rng) for w in wls])
        out = {}
        for la in ('0', '1'):
            monkeypatch.setenv('LCF_CHAIN_LA', la)
            b = BatchSampler(probs, nw, seed=9).run(p0, 20, 15)
            assert np.all(b.status == 0)
            out[la] = (b.get_chain(), b.get_log_prob(), b.acceptance_fraction)
        np.testing.assert_array_equal(out['1'][0], out['0'][0])
        np.testing.assert_array_equal(out['1'][2], out['0'][2])
        np.testing.assert_allclose(out['1'][1], out['0'][1], rtol=rtol)
        assert not np.array_equal(out['1'][0][:, 0], out['1'][0][:, -1])
        if nw == 22:
            np.testing.assert_array_equal(out['1'][1], out['0'][1])
        for i in range(7):                       # every stored position carries its own log-posterior
            np.testing.assert_allclose(probs[i].log_posterior(out['1'][0][i].reshape(-1, 2)), out['1'][1][i].reshape(-1), rtol=1e-12 if precision == 'fp64' else 2e-5)


@pytest.mark.parametrize('nw', [70, 150, 256, 300])
def test_batched_ensembles_wide_walker_groups(nw):
    """Half-ensembles of 35 / 75 / 150 walkers run as ONE wide group (64 / 128 / 256 walkers, several warp columns) per
    half-step in the chain kernel: every stored position carries its own log-posterior, chains move, results match the
    oracle, and the light curves have different lengths (ragged batch)."""
    from lightcurve_fitting_b200.bolometric import BatchSampler
    # (256 walkers x all nine filters is the cfg5 shape: 48.9 KB of dynamic + 0.8 KB of static shared memory, just past
    # the 48 KB a kernel gets without opting in -- it once failed to launch)
    wls = [W.synthetic_sc4(npoints=n, lc_index=i) for i, n in enumerate((40, 200 if nw == 256 else 97, 13))]
    probs = [w.device_problem('fp32' if nw == 256 else 'fp64') for w in wls]
    rng = np.random.default_rng(nw)
    p0 = np.stack([w.start(nw, rng) for w in wls])
    b = BatchSampler(probs, nw, seed=5).run(p0, 6, 5)
    chain, lnp = b.get_chain(), b.get_log_prob()
    assert chain.shape == (3, 5, nw, 5) and np.all(b.status == 0)
    for i, w in enumerate(wls):
        np.testing.assert_allclose(probs[i].log_posterior(chain[i].reshape(-1, 5)), lnp[i].reshape(-1), rtol=1e-12 if nw != 256 else 2e-5)
        lp = W.oracle_log_posterior(w)
        sel = np.random.default_rng(i).choice(nw, 6, replace=False)
        np.testing.assert_allclose(lnp[i, -1, sel], [lp(p) for p in chain[i, -1, sel]], rtol=1e-9 if nw != 256 else 1e-4)
        assert not np.array_equal(chain[i, 0], chain[i, -1])
    assert 0.02 < b.acceptance_fraction.mean() < 0.95


# ---------------------------------------------------------------------------------------------------------
# The CUDA path against golden vectors produced by the reference's OWN code (tests/golden/make_golden.py)
# ---------------------------------------------------------------------------------------------------------
import os  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden.npz'))


def _pf(names):
    from lightcurve_fitting_b200.filters import filtdict
    return np.array([filtdict[str(n)] for n in names], dtype=object)


def _lc_from(t, f, y, dy, q='lum'):
    from lightcurve_fitting_b200.synthetic import _PreparedLC
    return _PreparedLC({'MJD': t, 'filter': f, q: y, 'd' + q: dy})


@pytest.mark.parametrize('tag', ['sc_n15', 'sc_n3', 'sc_rw', 'sc2', 'sc3', 'sc4'])
def test_models_match_reference_golden(tag):
    from lightcurve_fitting_b200 import models as M
    m = {'sc_n15': lambda: M.ShockCooling(redshift=0.002), 'sc_n3': lambda: M.ShockCooling(redshift=0.002, n=3.),
         'sc_rw': lambda: M.ShockCooling(redshift=0.002, RW=True), 'sc2': lambda: M.ShockCooling2(redshift=0.002),
         'sc3': lambda: M.ShockCooling3(redshift=0.005), 'sc4': lambda: M.ShockCooling4(redshift=0.002)}[tag]()
    t, f = GOLD['lc/t'], _pf(GOLD['lc/filters'])
    P, y, dy = GOLD[tag + '/P'], GOLD[tag + '/y'], GOLD[tag + '/dy']
    want = GOLD[tag + '/point']
    got = np.array([m(t, f, *p) for p in P])
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=0)
    assert np.array_equal(got == 0., want == 0.)                     # exact zeros before t_0
    np.testing.assert_allclose(m(GOLD['lc/tgrid'], _pf(GOLD['lc/gridfilters']), *P.T), GOLD[tag + '/grid'], rtol=1e-9)
    lc = _lc_from(t, f, y, dy, m.output_quantity)
    np.testing.assert_allclose([m.log_likelihood(lc, p) for p in P], GOLD[tag + '/loglike'], rtol=1e-9)
    for st in ('relative', 'absolute'):
        Ps = GOLD[tag + '/Psig_' + st]
        np.testing.assert_allclose([m.log_likelihood(lc, p, use_sigma=True, sigma_type=st) for p in Ps],
                                   GOLD[tag + '/loglike_sig_' + st], rtol=1e-9)
    # FP32 throughput mode: 1e-4 on the log-likelihood
    m.precision = 'fp32'
    np.testing.assert_allclose([m.log_likelihood(lc, p) for p in P], GOLD[tag + '/loglike'], rtol=1e-4)


@pytest.mark.parametrize('tag,cls', [('cs1', 'CompanionShocking'), ('cs2', 'CompanionShocking2'), ('cs3', 'CompanionShocking3')])
def test_companion_models_match_reference_golden(tag, cls):
    from lightcurve_fitting_b200 import models as M
    t, f, y, dy = GOLD['cs/t'], _pf(GOLD['cs/filters']), GOLD['cs/y'], GOLD['cs/dy']
    lc = _lc_from(t, f, y, dy)
    m = getattr(M, cls)(lc, redshift=0.01)
    P = GOLD[tag + '/P']
    np.testing.assert_allclose(np.array([m(t, f, *p) for p in P]), GOLD[tag + '/point'], rtol=1e-9)
    np.testing.assert_allclose(m(GOLD['cs/tgrid'], _pf(GOLD['cs/gridfilters']), *P.T), GOLD[tag + '/grid'], rtol=1e-9)
    np.testing.assert_allclose([m.log_likelihood(lc, p) for p in P], GOLD[tag + '/loglike'], rtol=1e-9)
    np.testing.assert_allclose([m.log_likelihood(lc, p, use_sigma=True) for p in GOLD[tag + '/Psig']],
                               GOLD[tag + '/loglike_sig'], rtol=1e-9)
    m.precision = 'fp32'
    np.testing.assert_allclose([m.log_likelihood(lc, p) for p in P], GOLD[tag + '/loglike'], rtol=1e-4)


def test_planck_and_synthesis_match_reference_golden():
    from lightcurve_fitting_b200 import models as M
    nu, T, R = GOLD['planck/nu'], GOLD['planck/T'], GOLD['planck/R']
    np.testing.assert_allclose(M.planck_fast(nu, T, R), GOLD['planck/out'], rtol=1e-9)
    np.testing.assert_allclose(M.planck_fast(nu, T, R, 900.), GOLD['planck/out_cutoff'], rtol=1e-9)
    np.testing.assert_allclose(M.planck_fast(nu, 12., 3.), GOLD['planck/out_scalar'], rtol=1e-9)
    names, Tb, Rb = GOLD['bb/names'], GOLD['bb/T'], GOLD['bb/R']
    for tag, kw in (('plain', {}), ('z', {'z': 0.05}), ('cut', {'z': 0.01, 'cutoff_freq': 700.}), ('ebv', {'ebv': 0.2}),
                    ('zebv', {'z': 0.02, 'ebv': 0.35})):
        np.testing.assert_allclose(M.blackbody_to_filters(_pf(names), Tb, Rb, **kw), GOLD['bb/point_' + tag], rtol=1e-9)
        np.testing.assert_allclose(M.blackbody_to_filters(_pf(names[:3]), Tb, Rb, **kw), GOLD['bb/grid_' + tag], rtol=1e-9)
    np.testing.assert_allclose(M.blackbody_to_filters(_pf(names[:2]), np.outer(Tb[:3], [1., 1.1]), np.outer(Rb[:3], [1., 0.9])),
                               GOLD['bb/grid2d'], rtol=1e-9)
    np.testing.assert_allclose(M.blackbody_to_filters(_pf(names[:3]), Tb[:4], Rb[:4], ebv=np.array([0., 0.1, 0.2, 0.3])),
                               GOLD['bb/ebv_vec'], rtol=1e-9)
    with pytest.raises(Exception, match='same shape'):
        M.blackbody_to_filters(_pf(names[:2]), Tb[:3], Rb[:2])
    f = _pf(['g'])[0]
    np.testing.assert_allclose(f.synthesize(M.planck_fast, 14., 0.5), GOLD['bb/point_plain'][2], rtol=1e-9)
    with pytest.raises(NotImplementedError):
        f.synthesize(lambda nu, T: nu, 1.)


def test_reference_lightcurve_mcmc_chain_reproduced_on_device():
    """The chain the reference's lightcurve_mcmc produced (golden) is reproduced by the device sampler when it is
    driven with the same stretch-move draws (recorded by the oracle under the reference's RNG protocol)."""
    from oracle import reference_port as rp
    from lightcurve_fitting_b200 import models as M
    from lightcurve_fitting_b200.fitting import build_problem
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    t, fn = GOLD['lc/t'], GOLD['lc/filters']
    y, dy = GOLD['sc4/y'], GOLD['sc4/dy']
    nw, nburn, nsteps = 12, 5, 6
    # oracle run, recording draws (CPU test test_driver_chain_matches_reference_lightcurve_mcmc pins it to the golden)
    om = rp.ShockCooling4(redshift=0.002)
    opri = [rp.UniformPrior(0., 10.), rp.UniformPrior(0., 10.), rp.UniformPrior(0., 100.), rp.UniformPrior(0., 100.),
            rp.UniformPrior(57460., 57468.5)]
    lp = rp.make_log_posterior(om, opri, t, W.oracle_filters(fn), y, dy)
    np.random.seed(12345)
    rs = np.random.RandomState()
    rs.set_state(np.random.get_state())
    ref = rp.StretchReplay(nw, 5, lp, random_state=rs)
    start = np.random.rand(nw, 5) * (GOLD['mcmc/p_up'] - GOLD['mcmc/p_lo']) + GOLD['mcmc/p_lo']
    pos, _, _ = ref.run_mcmc(start, nburn, record=True)
    burn_draws = list(ref.draws)
    ref.reset()
    ref.run_mcmc(pos, nsteps, record=True)
    # device run
    m = M.ShockCooling4(redshift=0.002)
    pri = [M.UniformPrior(0., 10.), M.UniformPrior(0., 10.), M.UniformPrior(0., 100.), M.UniformPrior(0., 100.),
           M.UniformPrior(57460., 57468.5)]
    prob = build_problem(_lc_from(t, _pf(fn), y, dy), m, pri)
    s = EnsembleSampler(nw, 5, prob, seed=0)
    s.run_replay(start, burn_draws)
    s.reset()
    s.run_replay(None, ref.draws)
    np.testing.assert_allclose(s.flatchain, GOLD['mcmc/flatchain'], rtol=1e-9)
    np.testing.assert_allclose(s.chain, GOLD['mcmc/chain'], rtol=1e-9)
    np.testing.assert_allclose(s.get_log_prob(), GOLD['mcmc/lnprob'], rtol=1e-9)
    np.testing.assert_array_equal(s.acceptance_fraction, GOLD['mcmc/acceptance'])


def test_lightcurve_mcmc_end_to_end_example():
    """docs/source/usage.rst:174-200 shaped call on the bundled light curve (ShockCooling4, 5 + 1 parameters)."""
    from lightcurve_fitting_b200 import lightcurve_mcmc, LC, models as M
    lc = LC.example().where(MJD_min=57468., MJD_max=57485.)
    lc = lc[~np.asarray(lc['nondet'].data, bool)]
    model = M.ShockCooling4(lc)
    priors = [M.UniformPrior(0., 10.), M.UniformPrior(0., 10.), M.UniformPrior(0., 100.), M.UniformPrior(0., 100.),
              M.UniformPrior(57468., 57468.7), M.GaussianPrior(0., 10.)]
    np.random.seed(1)
    sampler = lightcurve_mcmc(lc, model, priors=priors, p_lo=[0.5, 0.1, 0.1, 1., 57468.5, 0.], p_up=[2., 2., 10., 10., 57468.7, 1.],
                              nwalkers=40, nsteps=60, nsteps_burnin=60, use_sigma=True, sigma_type='relative')
    assert model.nparams == 6 and sampler.flatchain.shape == (60 * 40, 6) and sampler.chain.shape == (40, 60, 6)
    assert np.all(np.isfinite(sampler.get_log_prob())) and 0.02 < sampler.acceptance_fraction.mean() < 0.9
    fc = sampler.flatchain
    assert np.all(fc[:, 4] > 57468.) and np.all(fc[:, 4] < 57468.7) and np.all(fc[:, :4] > 0.)
    # the log-probabilities stored in the chain are the log-posterior of the stored positions
    np.testing.assert_allclose(sampler.problem.log_posterior(fc[-40:]), sampler.get_log_prob()[-1], rtol=1e-12)


def test_pseudo_and_spectrum_mcmc_and_calculate_bolometric(tmp_path):
    """bolometric.py drop-ins: pseudo() vs the oracle; spectrum_mcmc / blackbody_mcmc on one epoch; the batched
    calculate_bolometric on the bundled light curve (79 epochs with >= 3 filters)."""
    import warnings
    from oracle import reference_port as rp
    from lightcurve_fitting_b200 import bolometric as B, models as M, LC
    T = np.array([6., 11., 25.])
    R = np.array([3., 1.5, 0.7])
    np.testing.assert_allclose(B.pseudo(T, R, 0.002), rp.pseudo(T, R, 0.002), rtol=1e-9)
    np.testing.assert_allclose(B.pseudo(T, R, 0.01, cutoff_freq=800.), rp.pseudo(T, R, 0.01, cutoff_freq=800.), rtol=1e-9)
    assert B.blackbody_mcmc is B.spectrum_mcmc

    lc = LC.example()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        t0, batch = B.calculate_bolometric(lc, outpath=str(tmp_path), res=1., nwalkers=10, burnin_steps=200, steps=100,
                                           colors=['B-V', 'g-r'], seed=3, return_sampler=True)
    assert len(t0) == 65        # 91 epochs, 79 with >= 3 filters in the raw table, 65 with >= 3 filters DETECTED (S/N >= 3 after
                                # binning, lightcurve.py:240-251 + bolometric.py:748-751)
    for col in ('MJD', 'temp', 'radius', 'L_bol', 'L', 'temp_mcmc', 'radius_mcmc', 'dtemp_mcmc0', 'dtemp_mcmc1', 'L_bol_mcmc',
                'L_mcmc', 'dL_mcmc0', 'dL_mcmc1', 'L_int', 'npoints', 'B-V', 'd(B-V)', 'filts', 'L_opt', 'lum', 'dtemp0'):
        assert col in t0.colnames
    assert np.all(np.diff(t0['MJD'].data) > 0) and np.all(t0['npoints'].data >= 3)
    ok = np.isfinite(t0['temp'].data)
    assert ok.sum() >= 58
    # the (error-weighted, filter-integrated) MCMC temperatures track the (unweighted, effective-frequency) least-squares ones
    assert np.median(np.abs(np.log(t0['temp_mcmc'].data / t0['temp'].data))[ok]) < 0.3
    assert np.all(t0['dtemp_mcmc0'].data > 0) and np.all(t0['dtemp_mcmc1'].data > 0) and np.all(t0['L_mcmc'].data > 0)
    assert np.all(batch.status == 0) and 0.1 < batch.acceptance_fraction.mean() < 0.9
    chain = batch.get_chain()
    assert chain.shape == (65, 100, 10, 2)
    assert np.all(chain[..., 0] > 1.) and np.all(chain[..., 0] < 100.) and np.all(chain[..., 1] > 0.01)

    # one epoch through spectrum_mcmc, against the oracle's posterior statistics on the same data
    epochs = B.group_by_epoch(lc[np.isfinite(lc['dmag'].data) & (lc['dmag'].data > 0.)], 1.)
    e = max(epochs, key=len)
    e.calcFlux(); e = e.bin(delta=np.inf); e.meta = dict(lc.meta); e.calcMag(); e.calcAbsMag(); e.calcLum()
    priors = [M.UniformPrior(1., 100.), M.LogUniformPrior(0.01, 1000.)]
    rng = np.random.default_rng(0)
    sg = rng.normal(size=(20, 2)) * 0.3 + np.array([9., 3.])
    s = B.spectrum_mcmc(M.planck_fast, e, priors, sg, z=0.002, outpath=str(tmp_path), nwalkers=20, burnin_steps=300, steps=300,
                        save_chains=True, seed=5)
    assert s.flatchain.shape == (6000, 2)
    of = W.oracle_filters([f.name for f in e['filter'].data])
    om = rp.BlackbodySED(redshift=0.002)
    lp = rp.make_log_posterior(om, [rp.UniformPrior(1., 100.), rp.LogUniformPrior(0.01, 1000.)], np.zeros(len(e)), of,
                               e['lum'].data, e['dlum'].data)
    np.testing.assert_allclose(s.problem.log_posterior(s.flatchain[-5:]), [lp(p) for p in s.flatchain[-5:]], rtol=1e-9)
    ref = rp.StretchReplay(20, 2, lp, random_state=np.random.RandomState(1))
    pos, lnp, _ = ref.run_mcmc(sg, 300)
    ref.reset()
    ref.run_mcmc(pos, 300, log_prob0=lnp)
    qa, qb = np.percentile(s.flatchain, [16, 50, 84], axis=0), np.percentile(ref.flatchain, [16, 50, 84], axis=0)
    width = qb[2] - qb[0]
    assert np.all(np.abs(qa[1] - qb[1]) < 0.3 * width)
    with pytest.raises(NotImplementedError):
        B.spectrum_mcmc(lambda nu, T, R: nu, e, priors, sg)


def test_sharded_ensemble_emulated_two_ranks_matches_single():
    """The multi-GPU split (rank/world slicing in the C library, RNG keyed by the global walker) emulated on ONE GPU:
    two rank-local ensembles exchange their colour slices by device copies after every half-step.  The chain must be
    bit-identical to the single-rank run -- this is what `ShardedEnsemble` does with an NCCL all-gather."""
    import ctypes as C
    import torch
    from lightcurve_fitting_b200._capi import lib, check
    from lightcurve_fitting_b200.parallel import _DevArray
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    wl = W.synthetic_sc3(npoints=96)
    prob = wl.device_problem('fp32')
    nw, nsteps = 64, 6
    # pin the launch shape: in the FP32 fast path the two points a lane pairs up share one reciprocal, and which
    # points are paired depends on the tile size (walkers per CTA); with the same shape the chains are bit-identical
    # across GPU counts, otherwise they agree to FP32 rounding (the cluster size fixes the order of the partial sums)
    check(lib().lcf_set_tuning_ex(8, 4, 1))
    p0 = wl.start(nw, np.random.default_rng(2))
    single = EnsembleSampler(nw, wl.ndim, prob, seed=42)
    single.run_mcmc(p0, nsteps, skip_initial_state_check=True)

    ranks = [EnsembleSampler(nw, wl.ndim, prob, seed=42, rank=r, world=2) for r in range(2)]
    views, owns = [], []
    for s in ranks:
        s._set_initial(p0, True)
        dc, dl, st = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n0 = C.c_int64()
        ob, oc = (C.c_int64 * 2)(), (C.c_int64 * 2)()
        check(lib().lcf_ensemble_device_view(s.handle, C.byref(dc), C.byref(dl), C.byref(st), C.byref(n0), ob, oc))
        views.append(torch.as_tensor(_DevArray(dc.value, (nw, wl.ndim)), device='cuda'))
        owns.append(((ob[0], oc[0]), (ob[1], oc[1])))
    n0 = n0.value
    check(lib().lcf_ensemble_reserve(ranks[0].handle, nsteps))
    check(lib().lcf_ensemble_reserve(ranks[1].handle, nsteps))
    for _ in range(nsteps):
        for half in (0, 1):
            for s in ranks:
                check(lib().lcf_ensemble_half_step(s.handle, half, 1))
            for s in ranks:
                check(lib().lcf_ensemble_sync(s.handle))
            base = n0 if half else 0
            for r in range(2):                       # "all-gather": copy each rank's updated slice into the other replica
                b, c = owns[r][half]
                views[1 - r][base + b:base + b + c] = views[r][base + b:base + b + c]
            torch.cuda.synchronize()
        for s in ranks:
            check(lib().lcf_ensemble_end_step(s.handle, 1))
    ref = single.get_chain()
    for r, s in enumerate(ranks):
        (b, c) = owns[r][0]
        first, count = 2 * b, 2 * c
        ch = np.empty((nsteps, count, wl.ndim))
        lp = np.empty((nsteps, count))
        check(lib().lcf_ensemble_get_chain_slice(s.handle, first, count, ch.ctypes.data_as(C.POINTER(C.c_double)),
                                                 lp.ctypes.data_as(C.POINTER(C.c_double))))
        np.testing.assert_array_equal(ch, ref[:, first:first + count])
        np.testing.assert_array_equal(lp, single.get_log_prob()[:, first:first + count])
    check(lib().lcf_set_tuning(0, 0))


@pytest.mark.parametrize('segmented', [False, True])
def test_fused_peer_exchange_two_ranks_one_device_matches_single(segmented, monkeypatch):
    """The fused exchange (accept epilogue stores into the peer replica, device-side half-step flags) on two
    rank-ensembles of one process: chains bit-identical to the single-ensemble run, replicas identical at the end.
    segmented: the same through k_pass_seg (bank streamed in segments, forced with LCF_SEG_SAMPLES), which shares the protocol."""
    import ctypes as C
    from lightcurve_fitting_b200._capi import lib, check
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    wl = W.synthetic_sc3(npoints=96)
    if segmented:
        monkeypatch.setenv('LCF_SEG_SAMPLES', '100')
        monkeypatch.setenv('LCF_RING', '0')
    prob = wl.device_problem('fp32')
    nw, nsteps = 64, 8
    check(lib().lcf_set_tuning_ex(8, 4, 1))
    try:
        p0 = wl.start(nw, np.random.default_rng(2))
        single = EnsembleSampler(nw, wl.ndim, prob, seed=42)
        single.run_mcmc(p0, nsteps, skip_initial_state_check=True)
        assert (prob.last_launch()['kernel'] == 'k_pass_seg') == segmented
        ranks = [EnsembleSampler(nw, wl.ndim, prob, seed=42, rank=r, world=2) for r in range(2)]
        coords, logps, flags = (C.c_void_p * 2)(), (C.c_void_p * 2)(), (C.c_void_p * 2)()
        for r, s in enumerate(ranks):
            s._set_initial(p0, True)
            dc, dl, st, fl = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
            n0 = C.c_int64()
            ob, oc = (C.c_int64 * 2)(), (C.c_int64 * 2)()
            check(lib().lcf_ensemble_device_view(s.handle, C.byref(dc), C.byref(dl), C.byref(st), C.byref(n0), ob, oc))
            check(lib().lcf_ensemble_exchange_view(s.handle, C.byref(fl), None))
            coords[r], logps[r], flags[r] = dc.value, dl.value, fl.value
        for s in ranks:
            check(lib().lcf_ensemble_peers_attach_ptrs(s.handle, coords, logps, flags))
            check(lib().lcf_ensemble_reserve(s.handle, nsteps))
        for _ in range(nsteps):                      # launches interleaved rank by rank; no host synchronisation at all
            for half in (0, 1):
                for s in ranks:
                    check(lib().lcf_ensemble_half_step(s.handle, half, 1))
            for s in ranks:
                check(lib().lcf_ensemble_end_step(s.handle, 1))
        for s in ranks:
            check(lib().lcf_ensemble_sync(s.handle))
        ref, ref_lp = single.get_chain(), single.get_log_prob()
        states = []
        for r, s in enumerate(ranks):
            own = np.zeros(nw, bool)
            own[r * (nw // 2):(r + 1) * (nw // 2)] = True          # walkers 2b..2b+2c of this rank's colour slices
            ch, lp = s.get_chain(), s.get_log_prob()
            np.testing.assert_array_equal(ch[:, own], ref[:, own])
            np.testing.assert_array_equal(lp[:, own], ref_lp[:, own])
            states.append(s._state())
        np.testing.assert_array_equal(states[0].coords, states[1].coords)       # both replicas complete and identical
        np.testing.assert_array_equal(states[0].log_prob, states[1].log_prob)
        np.testing.assert_array_equal(states[0].coords, ref[-1])
    finally:
        check(lib().lcf_set_tuning_ex(0, 0, 0))


def _ragged_workload(filter_names, counts, model='ShockCooling4', seed=5, t_span=(0.4, 14.)):
    """A light curve with a prescribed (ragged) number of points per filter, truth from the oracle."""
    from lightcurve_fitting_b200.synthetic import Workload
    rng = np.random.default_rng(seed)
    fn = [f for f, c in zip(filter_names, counts) for _ in range(c)]
    t0 = 57468.6
    t = t0 + rng.uniform(*t_span, len(fn))
    order = rng.permutation(len(fn))                    # callers do not have to group by filter: the host layer does
    fn, t = [fn[i] for i in order], t[order]
    p_true = np.array([1.2, 0.8, 2., 4., t0])
    ytrue = W.oracle_truth(model, t, fn, p_true, 0.002)
    dy = 0.05 * np.abs(ytrue) + 1e-3 * np.abs(ytrue).max()
    y = ytrue + dy * rng.normal(size=len(fn))
    pri = [('uniform', 0., 10.), ('uniform', 0., 10.), ('uniform', 0., 100.), ('uniform', 0., 100.), ('uniform', t0 - 0.6, t0 + 0.4)]
    return Workload('ragged', model, t, fn, y, dy, pri, [0.5, 0.1, 0.1, 1., t0 - 0.1], [2., 2., 10., 10., t0 + 0.1], z=0.002,
                    truth=p_true)


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
def test_ragged_light_curves(precision):
    """One point in a filter, a single-point light curve, filters with odd / minimal / maximal sample counts
    (unfiltered '0': 4 samples ... F2550W: 1482), per-filter counts that never fill a tile."""
    cases = [(['U', 'B', 'V', 'g', 'r', 'i', 'R', 'I', '0'], [1, 2, 3, 5, 7, 1, 33, 65, 4]),
             (['g'], [1]),
             (['F2550W', 'NUV', 'B', '0'], [3, 2, 9, 1]),
             (['r', 'i'], [64, 63])]
    for names, counts in cases:
        wl = _ragged_workload(names, counts)
        _check_logpost(wl, precision, n=19, seed=4)


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
def test_cold_and_hot_blackbodies_take_every_path(precision):
    """Blackbody SED through UV-to-mid-IR filters from 300 K to 3e6 K: the plain fast path, the clamped fast path (the
    product of four denominators would overflow), the careful path (Rayleigh-Jeans, 2^x - 1 cancels) and the limit in
    which the flux underflows all agree with the oracle."""
    from lightcurve_fitting_b200 import models as M
    from oracle import reference_port as rp
    names = ['NUV', 'U', 'B', 'V', 'R', 'I', 'F444W', 'F2550W']
    T = np.array([0.3, 0.8, 1.5, 2.5, 4., 8., 20., 80., 400., 3000.])          # kK
    R = np.full_like(T, 5.)
    from lightcurve_fitting_b200.filters import filtdict
    f_dev = [filtdict[n] for n in names]
    f_ora = W.oracle_filters(names)
    got = M._sed_eval(f_dev, T, R, 0.01, np.inf, 0., precision=precision)                    # [len(T), nfilters]
    want = np.array([rp.blackbody_to_filters(f_ora, np.full(len(names), t), np.full(len(names), r), z=0.01) for t, r in zip(T, R)])
    scale = np.abs(want).max(axis=1, keepdims=True)
    # relative to the brightest band of each SED: an underflowing / capped Wien tail is allowed to differ by < 1e-9 of it
    np.testing.assert_allclose(got / scale, want / scale, rtol=RTOL[precision], atol=2e-9)


def test_odd_walker_counts_and_partial_groups():
    """Walker counts that are odd and not multiples of the walkers-per-CTA group: stretch move runs, chain complete,
    and a replay with injected draws matches the oracle's emcee-order chain."""
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    wl = W.example_sc4(npoints=30)
    prob = wl.device_problem('fp64')
    for nw in (2 * wl.ndim, 2 * wl.ndim + 1, 37, 67):
        s = EnsembleSampler(nw, wl.ndim, prob, seed=3)
        s.run_mcmc(wl.start(nw, np.random.default_rng(nw)), 7)
        ch, lp = s.get_chain(), s.get_log_prob()
        assert ch.shape == (7, nw, wl.ndim) and np.isfinite(lp).all()
        want = prob.log_posterior(ch[-1])
        np.testing.assert_allclose(lp[-1], want, rtol=1e-12)             # stored log-prob belongs to the stored position


def test_device_convergence_diagnostics_match_oracle():
    """Integrated autocorrelation time (emcee's definition), automatic window and split R-hat computed on the
    device-resident chain equal the oracle's numpy versions on the same chain."""
    from lightcurve_fitting_b200.sampler import EnsembleSampler, AutocorrError
    from oracle import reference_port as rp
    wl = W.example_sc4(npoints=24)
    prob = wl.device_problem('fp32')
    nw, nsteps = 24, 600
    s = EnsembleSampler(nw, wl.ndim, prob, seed=11)
    s.run_mcmc(wl.start(nw, np.random.default_rng(1)), nsteps)
    for discard in (0, 101):
        chain = s.get_chain()[discard:]
        moving = (chain != chain[0]).any(axis=(0, 2))        # a stuck walker makes emcee's estimate NaN; the device leaves it out
        tau_o, win_o = rp.integrated_time(chain[:, moving])
        tau, win = s.get_autocorr_time(discard=discard, tol=0, return_window=True)
        np.testing.assert_array_equal(win, win_o)
        np.testing.assert_allclose(tau, tau_o, rtol=1e-8)
        np.testing.assert_allclose(s.get_split_rhat(discard=discard), rp.split_rhat(chain), rtol=1e-9)
    tau_l, win_l = s.get_autocorr_time(tol=0, max_lag=3, return_window=True)       # window cannot exist within 3 lags
    assert (win_l == -1).all()
    with pytest.raises(AutocorrError):
        s.get_autocorr_time(tol=1e6)


def test_run_streams_chain_to_host_buffers():
    """run_mcmc(chain_out=, log_prob_out=) delivers the same chain as get_chain()/get_log_prob() of an identical run."""
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    wl = W.example_sc4(npoints=30)
    prob = wl.device_problem('fp32')
    nw, nsteps = 32, 9
    p0 = wl.start(nw, np.random.default_rng(5))
    a = EnsembleSampler(nw, wl.ndim, prob, seed=8)
    a.run_mcmc(p0, nsteps)
    b = EnsembleSampler(nw, wl.ndim, prob, seed=8)
    ch, lp = np.full((nsteps, nw, wl.ndim), np.nan), np.full((nsteps, nw), np.nan)
    b.run_mcmc(p0, nsteps, chain_out=ch, log_prob_out=lp)
    np.testing.assert_array_equal(ch, a.get_chain())
    np.testing.assert_array_equal(lp, a.get_log_prob())
    np.testing.assert_array_equal(b.get_chain(), ch)                 # the chain also stays on the device
    with pytest.raises(ValueError):
        b.run_mcmc(None, 3, chain_out=ch, log_prob_out=lp)            # wrong shape


def test_stretch_move_with_cluster_split_matches_unsplit():
    """The stretch move (proposal, accept, in-place update, chain write-back) under a thread-block-cluster split of the
    walker groups gives the chain of the unsplit launch (FP64: the partial sums only differ in grouping)."""
    from lightcurve_fitting_b200._capi import lib, check
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    wl = W.synthetic_sc3(npoints=120)
    prob = wl.device_problem('fp64')
    nw, nsteps = 40, 12
    p0 = wl.start(nw, np.random.default_rng(21))
    chains = {}
    try:
        for shape in ((8, 4, 1), (8, 4, 4), (2, 2, 8), (32, 8, 2)):
            check(lib().lcf_set_tuning_ex(*shape))
            s = EnsembleSampler(nw, wl.ndim, prob, seed=13)
            s.run_mcmc(p0, nsteps)
            chains[shape] = (s.get_chain(), s.get_log_prob(), s.acceptance_fraction)
    finally:
        check(lib().lcf_set_tuning_ex(0, 0, 0))
    ref = chains[(8, 4, 1)]
    assert 0.05 < ref[2].mean() < 0.95
    for shape, (ch, lp, acc) in chains.items():
        np.testing.assert_allclose(ch, ref[0], rtol=1e-9, err_msg=str(shape))
        np.testing.assert_allclose(lp, ref[1], rtol=1e-9, err_msg=str(shape))


def test_batched_blackbody_lstsq_matches_curve_fit():
    """The device's batched bounded Levenberg-Marquardt fits == the reference's per-epoch scipy curve_fit
    (bolometric.py:483-531) to curve_fit's own tolerance: interior optima, optima on the temperature bound, modified
    blackbodies (cutoff), two-point SEDs (covariance undefined -> inf, like curve_fit)."""
    import warnings
    from lightcurve_fitting_b200 import bolometric as B
    from lightcurve_fitting_b200.filters import filtdict
    from lightcurve_fitting_b200.lightcurve import LC
    from oracle import reference_port as rp
    rng = np.random.default_rng(12)
    names = ['U', 'B', 'V', 'R', 'I', 'g', 'r', 'i', 'UVW1', 'z']
    for cutoff in (np.inf, 900.):
        epochs, truths = [], []
        for k in range(60):
            n = int(rng.integers(2, 9))
            fl = [filtdict[x] for x in rng.choice(names, n, replace=False)]
            freq = np.array([f.freq_eff for f in fl])
            T = rng.uniform(4., 60.) if k % 5 else rng.uniform(80., 300.)        # every fifth hotter than the 100 kK bound
            R = 10. ** rng.uniform(0., 1.5)
            lum = rp.planck_fast(freq * 1.003, T, R, cutoff) * (1. + 0.05 * rng.normal(size=n))
            e = LC({'freq': freq, 'lum': lum})
            epochs.append(e)
        temp, radius, dtemp, drad, L, dL, Lopt, status = B.blackbody_lstsq_batch(epochs, 0.003, cutoff_freq=cutoff)
        assert np.all(status == 0)
        nbound = 0
        for i, e in enumerate(epochs):
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                want = rp.blackbody_lstsq(e['freq'].data, e['lum'].data, 0.003, cutoff_freq=cutoff)
            np.testing.assert_allclose([temp[i], radius[i]], want[:2], rtol=1e-5, err_msg='epoch %d' % i)
            np.testing.assert_allclose([L[i], Lopt[i]], [want[4], want[6]], rtol=1e-4)
            if len(e) > 2:
                np.testing.assert_allclose([dtemp[i], drad[i], dL[i]], [want[2], want[3], want[5]], rtol=1e-3)
            else:
                assert np.isinf(dtemp[i]) and np.isinf(want[2])
            nbound += temp[i] > 99.999
        assert nbound >= 3
    one = B.blackbody_lstsq(epochs[0], 0.003, cutoff_freq=900.)
    np.testing.assert_allclose(one, [temp[0], radius[0], dtemp[0], drad[0], L[0], dL[0], Lopt[0]], rtol=1e-12)


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
def test_shockcooling3_with_a_large_filter_bank(precision):
    """ShockCooling3 through space-telescope filters with > 1000 transmission samples each: the per-walker reddened
    weight table no longer fits in shared memory at 32 walkers per CTA, the launch-shape model has to pick fewer walkers
    per CTA (the table is [samples][walkers per CTA]), and the result is still the oracle's."""
    from lightcurve_fitting_b200.synthetic import Workload
    rng = np.random.default_rng(9)
    names = ['NUV', 'F2550W', 'F2100W', 'F444W', 'U', 'B', 'g', 'r']
    n = 48
    t0 = 59000.
    t = np.sort(rng.uniform(t0 + 0.3, t0 + 12., n))
    fn = [names[i % len(names)] for i in range(n)]
    p_true = np.array([1., 1., 1., 3., 20., 0.1, t0])
    ytrue = W.oracle_truth('ShockCooling3', t, fn, p_true, 0.005)
    dy = 0.05 * ytrue
    y = ytrue + dy * rng.normal(size=n)
    pri = [('uniform', 0., 10.), ('uniform', 0., 10.), ('uniform', 0., 100.), ('uniform', 0., 100.), ('uniform', 5., 50.),
           ('uniform', 0., 1.), ('uniform', t0 - 2., t0 + 0.3)]
    lo, hi = p_true * 0.9, p_true * 1.1
    lo[6], hi[6] = t0 - 0.1, t0 + 0.1
    wl = Workload('sc3-large-bank', 'ShockCooling3', t, fn, y, dy, pri, lo, hi, z=0.005, truth=p_true)
    assert wl.planck_samples_per_eval() > 30000
    _check_logpost(wl, precision, n=40, seed=2)
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    s = EnsembleSampler(80, wl.ndim, wl.device_problem(precision), seed=4)
    s.run_mcmc(wl.start(80, rng), 4)
    assert np.isfinite(s.get_log_prob()).all()


# ---- the BASELINE.json shapes themselves, on the kernels the benchmark times -------------------------------------------
def _check_chain_rows_against_oracle(wl, chain, lnp, precision, nrows, seed):
    """Recompute the log-posterior of `nrows` sampled (step, walker) rows of a stored chain with the oracle."""
    lp = W.oracle_log_posterior(wl)
    rng = np.random.default_rng(seed)
    S, Wn = lnp.shape
    pick = rng.choice(S * Wn, nrows, replace=False)
    got = lnp.reshape(-1)[pick]
    want = np.array([lp(p) for p in chain.reshape(S * Wn, -1)[pick]])
    assert np.all(np.isfinite(got))
    np.testing.assert_allclose(got, want, rtol=RTOL[precision])


@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
def test_cfg2_shape_native_moves_on_the_benchmark_kernel(precision):
    """BASELINE cfg2 at its real light-curve shape (ShockCooling3, N = 2000, 8 filters, 10^5 walkers: the benchmark's own ensemble, 1563 CTAs
    of 32 walkers x 16 warps per half-step): five native stretch-move steps must run on
    k_pass<3, real, 5, true> in MODE_MOVE, and the stored log-probabilities must be the oracle's log-posterior of the
    stored positions."""
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    wl = W.synthetic_sc3(npoints=2000)
    prob = wl.device_problem(precision)
    nw = 100_000
    s = EnsembleSampler(nw, wl.ndim, prob, seed=11)
    p0 = wl.start(nw, np.random.default_rng(2))
    s.run_mcmc(p0, 5, skip_initial_state_check=True)
    launch = prob.last_launch()
    assert launch['kernel'] == 'k_pass<32 walkers, plain>' and launch['walkers_per_cta'] == 32 and launch['warps_per_cta'] == 16
    assert launch['groups'] == (nw // 2 + 31) // 32 and launch['cluster'] == 1
    assert launch['sum_units'] == (8 if precision == 'fp64' else 1) and launch['points_per_lane'] == (4 if precision == 'fp32' else 2)
    if launch['flat']:                                           # one CTA per co-resident slot, each with the same share of the work
        assert launch['grid'] % _sm_count() == 0 and launch['grid'] < launch['groups']
    else:
        assert launch['grid'] == launch['groups']
    chain, lnp = s.get_chain(), s.get_log_prob()
    assert chain.shape == (5, nw, wl.ndim)
    acc = s.acceptance_fraction
    assert 0.05 < acc.mean() < 0.95
    moved = np.any(chain[-1] != p0, axis=1)
    assert moved.mean() > 0.5                                   # the ensemble really moved
    _check_chain_rows_against_oracle(wl, chain, lnp, precision, 24, 5)
    # a walker that never accepted still carries its initial log-posterior: initial evaluation (MODE_LOGPOST) and moves agree
    still = np.flatnonzero(~moved)[:4]
    if len(still):
        lp = W.oracle_log_posterior(wl)
        np.testing.assert_allclose(lnp[-1, still], [lp(p) for p in p0[still]], rtol=RTOL[precision])


def test_cfg4_shape_native_moves(precision='fp32'):
    """BASELINE cfg4: CompanionShocking3, N = 1000 over U,B,V,g,r,i, 10^4 walkers (the launch shape the cost model picks for
    it), native moves checked against the oracle."""
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    wl = W.synthetic_cs3(npoints=1000)
    prob = wl.device_problem(precision)
    nw = 10_000
    s = EnsembleSampler(nw, wl.ndim, prob, seed=12)
    s.run_mcmc(wl.start(nw, np.random.default_rng(3)), 4, skip_initial_state_check=True)
    launch = prob.last_launch()
    assert launch['groups'] * launch['walkers_per_cta'] >= nw // 2
    chain, lnp = s.get_chain(), s.get_log_prob()
    assert 0.05 < s.acceptance_fraction.mean() < 0.95
    _check_chain_rows_against_oracle(wl, chain, lnp, precision, 16, 6)


def test_cfg5_shape_batched_chains_256_walkers():
    """BASELINE cfg5 member shape: ShockCooling4 on N = 300 points, 256 walkers, whole chain in one k_chain launch
    (wide 128-walker groups); stored log-probabilities against the oracle."""
    from lightcurve_fitting_b200.bolometric import BatchSampler
    rng = np.random.default_rng(4)
    wls = [W.synthetic_sc4(npoints=n, lc_index=i) for i, n in enumerate((300, 117))]
    probs = [w.device_problem('fp32') for w in wls]
    b = BatchSampler(probs, 256, seed=9)
    b.run(np.stack([w.start(256, rng) for w in wls]), 6, 4)
    assert np.all(b.status == 0)
    chain, lnp = b.get_chain(), b.get_log_prob()
    assert chain.shape == (2, 4, 256, 5)
    for k, w in enumerate(wls):
        _check_chain_rows_against_oracle(w, chain[k], lnp[k], 'fp32', 12, 7 + k)


def test_two_live_problems_with_different_shared_memory_on_one_kernel(monkeypatch):
    """Two problems of the same model and precision share a k_pass instantiation but need different dynamic shared memory (a
    large UV bank next to a small optical one).  Launches of the large one must keep working after the small one was set up
    (the opt-in shared-memory limit of a kernel is process-wide and must only ever be raised)."""
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    big = _ragged_workload(['UVW2', 'UVW1', 'NUV', 'w', 'Kepler', 'TESS', 'Itagaki', 'g'], [40] * 8, model='ShockCooling4')   # 5200 samples: 83 KB in FP64
    small = _ragged_workload(['B', 'V'], [40, 40], model='ShockCooling4')
    from lightcurve_fitting_b200._capi import lib, check
    pb, ps = big.device_problem('fp64'), small.device_problem('fp64')
    rng = np.random.default_rng(0)
    nw = 2048
    monkeypatch.setenv('LCF_RING', '0')                            # half-step launches for both (not the persistent kernel)
    check(lib().lcf_set_tuning_ex(8, 8, 1))                        # one shape for both: the generic k_pass<4, double> instantiation
    try:
        sb = EnsembleSampler(nw, big.ndim, pb, seed=1)
        sb.run_mcmc(big.start(nw, rng), 1, skip_initial_state_check=True)
        ss = EnsembleSampler(nw, small.ndim, ps, seed=2)
        ss.run_mcmc(small.start(nw, rng), 1, skip_initial_state_check=True)
        assert pb.last_launch()['kernel'] == ps.last_launch()['kernel'] == 'k_pass<generic>'
        sb.run_mcmc(None, 2)                                       # cached shape of the large problem, after the small one ran
        ss.run_mcmc(None, 2)
        P = big.start(8, rng)
        lp = W.oracle_log_posterior(big)
        np.testing.assert_allclose(pb.log_posterior(P), [lp(p) for p in P], rtol=1e-9)
    finally:
        check(lib().lcf_set_tuning_ex(0, 0, 0))


# ---- bolometric.py against vectors produced by the REFERENCE's bolometric.py (tests/golden/make_golden.py) ---------------
def test_bolometric_functions_match_reference_golden():
    from lightcurve_fitting_b200 import bolometric as B, LC
    from lightcurve_fitting_b200.filters import filtdict
    T, R = GOLD['bolo/pseudo/T'], GOLD['bolo/pseudo/R']
    np.testing.assert_allclose(B.pseudo(T, R, 0.), GOLD['bolo/pseudo/z0'], rtol=1e-9)
    np.testing.assert_allclose(B.pseudo(T, R, 0.023), GOLD['bolo/pseudo/z'], rtol=1e-9)
    np.testing.assert_allclose(B.pseudo(T, R, 0.01, cutoff_freq=800.), GOLD['bolo/pseudo/cut'], rtol=1e-9)
    np.testing.assert_allclose(B.pseudo(12., 3., 0.002), GOLD['bolo/pseudo/scalar'], rtol=1e-9)
    np.testing.assert_allclose(B.pseudo(T, R, 0.002, filter0=filtdict['V'], filter1=filtdict['B']), GOLD['bolo/pseudo/BtoV'], rtol=1e-9)
    np.testing.assert_allclose(B.sigma_sb, GOLD['bolo/sigma_sb'], rtol=1e-14)
    np.testing.assert_allclose(B.stefan_boltzmann(T, R), GOLD['bolo/sb/lum'], rtol=1e-13)
    lum, dlum = B.stefan_boltzmann(T, R, GOLD['bolo/sb/dT'], GOLD['bolo/sb/dR'], GOLD['bolo/sb/cov'])
    np.testing.assert_allclose(lum, GOLD['bolo/sb/lum2'], rtol=1e-13)
    np.testing.assert_allclose(dlum, GOLD['bolo/sb/dlum'], rtol=1e-13)
    for x, out, pc in ((GOLD['bolo/mu/x1'], GOLD['bolo/mu/out1'], 68.), (GOLD['bolo/mu/x2'], GOLD['bolo/mu/out2'], 68.),
                       (GOLD['bolo/mu/x1'], GOLD['bolo/mu/out1_100'], 100.)):
        np.testing.assert_allclose(np.array(B.median_and_unc(x, pc)), out, rtol=1e-12)
    # least-squares blackbody of every golden SED in one launch (the reference: scipy curve_fit per epoch, stops at 1e-8)
    off, out = GOLD['bolo/lstsq/offsets'], GOLD['bolo/lstsq/out']
    z = float(GOLD['bolo/lstsq/z'])
    for cutoff, tmax in sorted({(float(c), float(t)) for c, t in out[:, 7:9]}):
        ks = [k for k in range(len(out)) if out[k, 7] == cutoff and out[k, 8] == tmax]
        eps = []
        for k in ks:
            names = [str(n) for n in GOLD['bolo/lstsq/filters'][off[k]:off[k + 1]]]
            np.testing.assert_allclose([filtdict[n].freq_eff for n in names], GOLD['bolo/lstsq/freq'][off[k]:off[k + 1]], rtol=1e-12)
            eps.append(LC({'freq': GOLD['bolo/lstsq/freq'][off[k]:off[k + 1]], 'lum': GOLD['bolo/lstsq/lum'][off[k]:off[k + 1]]}))
        res = np.array(B.blackbody_lstsq_batch(eps, z, T_range=(1., tmax), cutoff_freq=cutoff)[:7]).T       # [epoch, 7]
        want = out[ks, :7]
        # scipy's TRF stops when the cost changes by < 1e-8 (ftol), which leaves its parameters good to ~1e-6..1e-5; the device
        # Levenberg-Marquardt iterates to machine precision
        np.testing.assert_allclose(res[:, [0, 1, 4, 6]], want[:, [0, 1, 4, 6]], rtol=2e-5)                    # temp, radius, lum, L_opt
        fin = np.isfinite(want[:, [2, 3, 5]]) & (want[:, [0]] < 0.999 * tmax)    # covariance: undefined for nfilt <= 2 / at a bound
        np.testing.assert_allclose(res[:, [2, 3, 5]][fin], want[:, [2, 3, 5]][fin], rtol=5e-4)


@pytest.mark.parametrize('tag,use_sigma,sigma_type', [('bolo/mcmc', False, 'relative'), ('bolo/mcmc_sigma', True, 'absolute')])
def test_reference_spectrum_mcmc_chain_reproduced_on_device(tag, use_sigma, sigma_type):
    """The chain the reference's spectrum_mcmc produced (golden: its own closure, bolometric.py:154-164, and driver) is
    reproduced by the device SED problem driven with the same stretch-move draws."""
    from tests.test_oracle_golden import _sed_driver
    from lightcurve_fitting_b200 import bolometric as B, models as M, LC
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    ref, burn = _sed_driver(tag, use_sigma, sigma_type, record=True)
    names = [str(n) for n in GOLD[tag + '/filters']]
    ep = LC({'MJD': np.full(len(names), 58000.25), 'filter': np.array(_pf(names), dtype=object), 'lum': GOLD[tag + '/lum'],
             'dlum': GOLD[tag + '/dlum']})
    priors = [M.UniformPrior(1., 100.), M.LogUniformPrior(0.01, 1000.)] + ([M.GaussianPrior(0., 10.)] if use_sigma else [])
    prob = B._sed_problem(ep, priors, 0.002, 0., float(GOLD[tag + '/cutoff']), use_sigma, sigma_type, 'fp64')
    s = EnsembleSampler(10, len(priors), prob, seed=0)
    s.run_replay(GOLD[tag + '/start'], burn)
    s.reset()
    s.run_replay(None, ref.draws)
    np.testing.assert_allclose(s.flatchain, GOLD[tag + '/flatchain'], rtol=1e-9)
    np.testing.assert_allclose(s.get_log_prob(), GOLD[tag + '/lnprob'], rtol=1e-9)
    np.testing.assert_array_equal(s.acceptance_fraction, GOLD[tag + '/acceptance'])


def test_shared_ensemble_start_state_from_own_slices():
    """lcf_ensemble_set_state_slice on two rank-ensembles of one process: each uploads and evaluates only its own walkers and
    stores them into the other replica; both replicas must equal the state a single ensemble gets from the full array,
    and the chains that follow must stay bit-identical."""
    import ctypes as C
    from lightcurve_fitting_b200._capi import lib, check, dptr
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    wl = W.synthetic_sc3(npoints=96)
    prob = wl.device_problem('fp64')
    nw, nsteps = 48, 3
    check(lib().lcf_set_tuning_ex(8, 4, 1))
    try:
        p0 = wl.start(nw, np.random.default_rng(8))
        single = EnsembleSampler(nw, wl.ndim, prob, seed=5)
        single._set_initial(p0, True)
        want = single._state()
        ranks = [EnsembleSampler(nw, wl.ndim, prob, seed=5, rank=r, world=2) for r in range(2)]
        coords, logps, flags = (C.c_void_p * 2)(), (C.c_void_p * 2)(), (C.c_void_p * 2)()
        for r, s in enumerate(ranks):
            dc, dl, st, fl = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
            check(lib().lcf_ensemble_device_view(s.handle, C.byref(dc), C.byref(dl), C.byref(st), None, None, None))
            check(lib().lcf_ensemble_exchange_view(s.handle, C.byref(fl), None))
            coords[r], logps[r], flags[r] = dc.value, dl.value, fl.value
        for s in ranks:
            check(lib().lcf_ensemble_peers_attach_ptrs(s.handle, coords, logps, flags))
        per = nw // 2
        with pytest.raises(ValueError, match='owns'):
            lib_rc = lib().lcf_ensemble_set_state_slice(ranks[0].handle, 2, per, dptr(np.ascontiguousarray(p0[2:2 + per])))
            check(lib_rc)
        for r, s in enumerate(ranks):
            check(lib().lcf_ensemble_set_state_slice(s.handle, r * per, per, dptr(np.ascontiguousarray(p0[r * per:(r + 1) * per]))))
        for s in ranks:
            st = s._state()
            np.testing.assert_array_equal(st.coords, want.coords)
            np.testing.assert_array_equal(st.log_prob, want.log_prob)
        bad = p0.copy()
        bad[3, 1] = np.nan
        with pytest.raises(ValueError, match='NaN'):
            check(lib().lcf_ensemble_set_state_slice(ranks[0].handle, 0, per, dptr(np.ascontiguousarray(bad[:per]))))
    finally:
        check(lib().lcf_set_tuning_ex(0, 0, 0))


def test_calculate_bolometric_device_summaries_match_host_post_processing():
    """calculate_bolometric end to end on a synthetic table: the posterior summaries taken on the device (Stefan-Boltzmann and
    pseudo-bolometric luminosity of every sample, percentiles by an in-kernel sort) equal the reference's post-processing
    (bolometric.py:792-798) applied on the host to the same chains; the least-squares columns equal the per-epoch fits."""
    import warnings
    from lightcurve_fitting_b200 import bolometric as B, synthetic
    import bench
    lc = synthetic.sed_table(bench.device_truth, nepochs=40, seed=11)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        t0, batch, tm = B.calculate_bolometric(lc, nwalkers=12, burnin_steps=60, steps=50, use_sigma=True, sigma_type='absolute',
                                               colors=['B-V'], cutoff_freq=1500., seed=4, return_sampler=True, return_timing=True)
    assert len(t0) == 40 and np.all(batch.status == 0) and tm['sampling_ms'] > 0
    chain = batch.get_chain()                                       # [E, S, W, 3]
    flat = chain.reshape(len(t0), -1, 3)
    for i in range(len(t0)):
        (T, R), (dT0, dR0), (dT1, dR1) = B.median_and_unc(flat[i][:, :2])
        Lb, dLb0, dLb1 = B.median_and_unc(B.stefan_boltzmann(flat[i][:, 0], flat[i][:, 1]))
        Lp, dLp0, dLp1 = B.median_and_unc(B.pseudo(flat[i][:, 0], flat[i][:, 1], 0.002, cutoff_freq=1500.))
        for names, vals in ((('temp_mcmc', 'dtemp_mcmc0', 'dtemp_mcmc1'), (T, dT0, dT1)), (('radius_mcmc', 'dradius_mcmc0', 'dradius_mcmc1'), (R, dR0, dR1)),
                            (('L_bol_mcmc', 'dL_bol_mcmc0', 'dL_bol_mcmc1'), (Lb, dLb0, dLb1)), (('L_mcmc', 'dL_mcmc0', 'dL_mcmc1'), (Lp, dLp0, dLp1))):
            for k, v in zip(names, vals):                              # interval half-widths are differences: tolerance relative to the median
                np.testing.assert_allclose(t0[k][i], v, rtol=1e-9, atol=1e-11 * abs(vals[0]), err_msg=k)
    # the truths are recovered (5 % photometry): medians within a few sigma-widths
    assert np.all(t0['temp_mcmc'].data > 1.) and np.all(np.isfinite(t0['L_mcmc'].data))
    assert np.all(t0['npoints'].data >= 3) and np.all(np.diff(t0['MJD'].data) > 0)
    with pytest.raises(NotImplementedError, match='single detected filter'):
        one = lc[np.asarray(lc['MJD'].data) < 58000.5][:1]
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            B.calculate_bolometric(one, min_nfilt=1)


@pytest.mark.parametrize('shape', [(0, 0, 0), (1, 8, 4), (4, 4, 2), (8, 8, 1)])
def test_persistent_chain_kernel_matches_half_step_launches(shape, monkeypatch):
    """cfg1-sized ensembles (100 walkers) run as ONE cooperative launch of the persistent kernel k_ring, with a device-side
    barrier between half-steps (LCF_RING=1) or -- look-ahead rounds, LCF_RING=2 and the default -- one per step.  The chain,
    the final state and the acceptance counts must be bit-identical to those of one k_pass launch per half-step (LCF_RING=0)
    with the same launch shape, in both precisions, with and without the intrinsic-scatter parameter."""
    from lightcurve_fitting_b200._capi import lib, check
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    check(lib().lcf_set_tuning_ex(*shape))
    try:
        for precision, use_sigma in (('fp32', False), ('fp64', True)):
            wl = W.example_sc4(use_sigma=use_sigma)
            prob = wl.device_problem(precision)
            p0 = wl.start(100, np.random.default_rng(5))
            out = {}
            for ring in ('0', '1', '2'):
                monkeypatch.setenv('LCF_RING', ring)
                s = EnsembleSampler(100, wl.ndim, prob, seed=31)
                s.run_mcmc(p0, 7, store=False)
                s.run_mcmc(None, 9)
                s.run_mcmc(None, 4)                                   # a second stored run appends to the chain
                st = s.run_mcmc(None, 1)                              # a single round: the final state comes from the drain alone
                out[ring] = (s.get_chain(), s.get_log_prob(), s.acceptance_fraction, prob.last_launch(),
                             np.array(st.coords), np.array(st.log_prob))
            assert out['1'][3]['kernel'] == 'k_ring' and out['0'][3]['kernel'].startswith('k_pass')
            # LCF_RING=2: look-ahead rounds (one grid barrier per step, both outcomes of the partner's move evaluated) whenever the
            # grid of n0 + 2 n1 = 150 virtual walkers is co-resident; else one half-step per barrier
            look_ahead = (precision == 'fp32' and shape != (1, 8, 4)) or shape in ((4, 4, 2), (8, 8, 1))
            assert out['2'][3]['kernel'] == ('k_ring<look-ahead>' if look_ahead else out['2'][3]['kernel'])
            assert out['2'][3]['kernel'] in ('k_ring', 'k_ring<look-ahead>')
            assert out['1'][0].shape == (14, 100, wl.ndim)
            for ring in ('1', '2'):
                for k in (0, 1, 2, 4, 5):
                    np.testing.assert_array_equal(out[ring][k], out['0'][k])
        # a large ensemble does not use the persistent kernel (its grid is not co-resident / not launch-bound)
        monkeypatch.delenv('LCF_RING')
        wl = W.synthetic_sc3(npoints=400)
        prob = wl.device_problem('fp32')
        s = EnsembleSampler(20000, wl.ndim, prob, seed=1)
        s.run_mcmc(wl.start(20000, np.random.default_rng(1)), 2, skip_initial_state_check=True)
        assert prob.last_launch()['kernel'].startswith('k_pass')
    finally:
        check(lib().lcf_set_tuning_ex(0, 0, 0))


@pytest.mark.parametrize('case', ['sc3_odd', 'cs3', 'sed', 'sc4_sigma_small'])
def test_look_ahead_rounds_other_models_and_odd_ensembles(case, monkeypatch):
    """The look-ahead rounds of the persistent kernel (LCF_RING=2) against one k_pass launch per half-step (LCF_RING=0) on ensembles
    whose colours differ in size (odd walker counts), on the other model families (per-walker weight table, SiFTO spline, bare SED)
    and with the intrinsic-scatter parameter: chains, log-probabilities, acceptance counts and final states must be bit-identical."""
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    rng = np.random.default_rng(11)
    if case == 'sc3_odd':
        wl, nwalk, precision = W.synthetic_sc3(npoints=120), 51, 'fp32'
    elif case == 'cs3':
        wl, nwalk, precision = W.synthetic_cs3(npoints=90), 40, 'fp32'
    elif case == 'sed':
        wl, nwalk, precision = W.sed_epoch(rng), 33, 'fp64'
    else:
        wl, nwalk, precision = W.example_sc4(use_sigma=True, npoints=60), 17, 'fp32'
    prob = wl.device_problem(precision)
    p0 = wl.start(nwalk, np.random.default_rng(2))
    out = {}
    for ring in ('0', '2'):
        monkeypatch.setenv('LCF_RING', ring)
        s = EnsembleSampler(nwalk, wl.ndim, prob, seed=77)
        s.run_mcmc(p0, 5, store=False, skip_initial_state_check=True)
        st = s.run_mcmc(None, 11)
        out[ring] = (s.get_chain(), s.get_log_prob(), s.acceptance_fraction, np.array(st.coords), np.array(st.log_prob), prob.last_launch()['kernel'])
    assert out['2'][5] == 'k_ring<look-ahead>' and out['0'][5].startswith('k_pass')
    for k in range(5):
        np.testing.assert_array_equal(out['2'][k], out['0'][k])
    assert not np.array_equal(out['2'][0][0], out['2'][0][-1])


def test_native_sampler_posterior_matches_a_long_reference_run():
    """Statistical parity on a light-curve model: the reference's own lightcurve_mcmc (its closure and driver under the emcee
    stand-in; 32 walkers, 600 + 1200 steps, ShockCooling2 on a 45-point synthetic light curve) froze the posterior
    percentiles; the native device sampler (Philox draws, fixed red/blue split) must reproduce medians and 68 % widths within
    the Monte-Carlo error of the reference chain, in both precisions."""
    from lightcurve_fitting_b200 import lightcurve_mcmc, models as M
    lc = _lc_from(GOLD['stats/t'], _pf(GOLD['stats/filters']), GOLD['stats/y'], GOLD['stats/dy'])
    pri = [M.UniformPrior(5., 60.), M.UniformPrior(0.1, 20.), M.UniformPrior(1., 30.), M.UniformPrior(57465., 57468.2)]
    pt = GOLD['stats/p_true']
    p_lo, p_up = pt * [0.9, 0.9, 0.9, 1.] - [0, 0, 0, 0.1], pt * [1.1, 1.1, 1.1, 1.] + [0, 0, 0, 0.1]
    want = GOLD['stats/percentiles']
    width = want[2] - want[0]
    for precision in ('fp64', 'fp32'):
        np.random.seed(7)
        s = lightcurve_mcmc(lc, M.ShockCooling2(redshift=0.002), priors=pri, p_lo=p_lo, p_up=p_up, nwalkers=256, nsteps=800,
                            nsteps_burnin=600, seed=123, precision=precision)
        got = np.percentile(s.flatchain, [16., 50., 84.], axis=0)
        # reference: 38 400 correlated samples (tau ~ 30-40 steps -> ~1000 independent): the median's standard error is ~0.02 widths
        assert np.all(np.abs(got[1] - want[1]) < 0.12 * width), (precision, (got[1] - want[1]) / width)
        assert np.all(np.abs((got[2] - got[0]) / width - 1.) < 0.15), (precision, (got[2] - got[0]) / width)
        assert abs(float(s.acceptance_fraction.mean()) - float(GOLD['stats/acceptance'])) < 0.05


@pytest.mark.parametrize('precision', ['fp64', 'fp32'])
@pytest.mark.parametrize('name', ['ShockCooling4', 'ShockCooling3'])
def test_out_of_domain_parameters_never_give_a_silent_wrong_value(name, precision):
    """Parameter vectors no sensible prior admits (zero / negative v_s, M_env, f_rho M, R; explosion after some or all points):
    the reference's numpy expressions give NaN for most of them and a finite number for some (power() zeroes non-positive
    bases, models.py:42-48).  The device must return NaN whenever the reference does, and where the reference is finite it
    must either agree or report NaN (-> emcee's "Probability function returned NaN") -- never a different finite value.
    All 243 sign patterns x 3 explosion epochs; the classes that differ are listed in DESIGN.md section 2."""
    import itertools
    import warnings
    wl = W.example_sc4(npoints=40) if name == 'ShockCooling4' else W.synthetic_sc3(npoints=64)
    wl.priors_spec = [('uniform', -1e9, 1e9)] * wl.ndim
    lp = W.oracle_log_posterior(wl)
    base = 0.5 * (wl.p_lo + wl.p_up)
    tmin, tmax = wl.t.min(), wl.t.max()
    rows = []
    for signs in itertools.product([1, -1, 0], repeat=4):
        for t0 in (base[-1], tmax + 1., 0.5 * (tmin + tmax)):
            p = base.copy()
            p[:4] = base[:4] * np.array(signs)
            p[-1] = t0
            rows.append(p)
    P = np.array(rows)
    with warnings.catch_warnings(), np.errstate(all='ignore'):
        warnings.simplefilter('ignore')
        want = np.array([lp(p) for p in P])
    got = wl.device_problem(precision).log_posterior(P)
    assert not np.any(np.isnan(want) & ~np.isnan(got)), 'the device is finite where the reference is NaN'
    both = ~np.isnan(want) & ~np.isnan(got)
    fin = both & np.isfinite(want)
    np.testing.assert_allclose(got[fin], want[fin], rtol=RTOL[precision])
    assert np.array_equal(got[both & ~fin], want[both & ~fin])
    inside = np.all(P[:, :4] > 0, axis=1)                                 # everything a prior with positive support can propose
    assert np.array_equal(np.isnan(got[inside]), np.isnan(want[inside]))
    assert both.sum() >= 18


def test_fused_peer_exchange_stress_randomised_launch_order():
    """2 000 steps of the fused exchange on two rank-ensembles of one process, the two ranks' launches of every half-step issued
    in random order (compute-sanitizer is closed on this pool; this is the race hunt for the flag protocol and the in-place
    update): rank chains bit-identical to the single-ensemble run, replicas identical, no exchange timeout."""
    import ctypes as C
    from lightcurve_fitting_b200._capi import lib, check
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    wl = W.sed_epoch(np.random.default_rng(12))
    prob = wl.device_problem('fp32')
    nw, nsteps = 48, 2000
    check(lib().lcf_set_tuning_ex(4, 2, 1))
    try:
        p0 = wl.start(nw, np.random.default_rng(2))
        single = EnsembleSampler(nw, wl.ndim, prob, seed=42)
        single.run_mcmc(p0, nsteps, skip_initial_state_check=True)
        ranks = [EnsembleSampler(nw, wl.ndim, prob, seed=42, rank=r, world=2) for r in range(2)]
        coords, logps, flags = (C.c_void_p * 2)(), (C.c_void_p * 2)(), (C.c_void_p * 2)()
        for r, s in enumerate(ranks):
            s._set_initial(p0, True)
            dc, dl, st, fl = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
            check(lib().lcf_ensemble_device_view(s.handle, C.byref(dc), C.byref(dl), C.byref(st), None, None, None))
            check(lib().lcf_ensemble_exchange_view(s.handle, C.byref(fl), None))
            coords[r], logps[r], flags[r] = dc.value, dl.value, fl.value
        for s in ranks:
            check(lib().lcf_ensemble_peers_attach_ptrs(s.handle, coords, logps, flags))
            check(lib().lcf_ensemble_reserve(s.handle, nsteps))
        order = np.random.default_rng(0).integers(0, 2, size=(nsteps, 2))
        for it in range(nsteps):
            for half in (0, 1):
                first = int(order[it, half])
                for r in (first, 1 - first):
                    check(lib().lcf_ensemble_half_step(ranks[r].handle, half, 1))
            for s in ranks:
                check(lib().lcf_ensemble_end_step(s.handle, 1))
        for s in ranks:
            check(lib().lcf_ensemble_sync(s.handle))              # raises if a wait on the peer's flag timed out
        ref, ref_lp = single.get_chain(), single.get_log_prob()
        states = []
        for r, s in enumerate(ranks):
            own = np.zeros(nw, bool)
            own[r * (nw // 2):(r + 1) * (nw // 2)] = True
            np.testing.assert_array_equal(s.get_chain()[:, own], ref[:, own])
            np.testing.assert_array_equal(s.get_log_prob()[:, own], ref_lp[:, own])
            states.append(s._state())
        np.testing.assert_array_equal(states[0].coords, states[1].coords)
        np.testing.assert_array_equal(states[0].log_prob, states[1].log_prob)
        np.testing.assert_array_equal(states[0].coords, ref[-1])
    finally:
        check(lib().lcf_set_tuning_ex(0, 0, 0))


@pytest.mark.parametrize('shape', [(1, 8, 1, 32), (2, 4, 2, 8), (4, 8, 1, 4), (8, 4, 1, 2), (16, 4, 1, 2), (1, 16, 1, 4)])
@pytest.mark.parametrize('name,precision', [('sc4_example', 'fp32'), ('sc4_example_sigma_abs', 'fp64'), ('sc3_synth_sigma', 'fp32'),
                                            ('cs3_synth', 'fp64'), ('sed', 'fp32'), ('sed_sigma', 'fp64')])
def test_split_k_tiles_match_oracle(name, precision, shape):
    """Split-K tiles (the samples of a (walker, point pair) swept by 2..32 lanes, partial sums combined with warp shuffles): every
    model family, both precisions, with thread-block clusters, against the oracle; and a short native run on the persistent
    kernel equals the half-step launches of the same shape bit for bit."""
    from lightcurve_fitting_b200._capi import lib, check
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    check(lib().lcf_set_tuning_ex(*shape[:3]))
    check(lib().lcf_set_tuning_split(shape[3]))
    try:
        wl = WORKLOADS[name]()
        _check_logpost(wl, precision, n=19, seed=4, widen=0.3)
    finally:
        check(lib().lcf_set_tuning_ex(0, 0, 0))
        check(lib().lcf_set_tuning_split(0))


def test_split_k_persistent_kernel_and_batches(monkeypatch):
    from lightcurve_fitting_b200._capi import lib, check
    from lightcurve_fitting_b200.bolometric import BatchSampler
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    wl = W.example_sc4()
    prob = wl.device_problem('fp32')
    p0 = wl.start(100, np.random.default_rng(5))
    check(lib().lcf_set_tuning_ex(1, 16, 1))
    check(lib().lcf_set_tuning_split(8))
    try:
        out = {}
        for ring in ('0', '1'):
            monkeypatch.setenv('LCF_RING', ring)
            s = EnsembleSampler(100, wl.ndim, prob, seed=31)
            s.run_mcmc(p0, 9)
            out[ring] = (s.get_chain(), s.get_log_prob())
        np.testing.assert_array_equal(out['1'][0], out['0'][0])
        np.testing.assert_array_equal(out['1'][1], out['0'][1])
        _check_chain_rows_against_oracle(wl, out['1'][0], out['1'][1], 'fp32', 12, 3)
    finally:
        check(lib().lcf_set_tuning_ex(0, 0, 0))
        check(lib().lcf_set_tuning_split(0))
        monkeypatch.delenv('LCF_RING')
    # SED batches pick split-K themselves (10 walkers, one point per filter): chains against the oracle
    rng = np.random.default_rng(3)
    wls = [W.sed_epoch(rng) for _ in range(5)]
    b = BatchSampler([w.device_problem('fp64') for w in wls], 10, seed=5)
    b.run(np.stack([w.start(10, rng) for w in wls]), 15, 10)
    chain, lnp = b.get_chain(), b.get_log_prob()
    for k, w in enumerate(wls):
        _check_chain_rows_against_oracle(w, chain[k], lnp[k], 'fp64', 10, k)


def _sm_count():
    import torch
    return torch.cuda.get_device_properties(0).multi_processor_count


@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
@pytest.mark.parametrize('name,nw,shape', [('sc3_synth', 20000, (32, 16, 1)), ('sc4_synth', 9001, (16, 8, 1)), ('cs3_synth', 12000, (32, 8, 1))])
def test_flat_split_chains_are_bit_identical_to_one_group_per_cta(name, nw, shape, precision):
    """Flat split (every CTA of a slots-sized grid takes the same share of the (group, unit) space; groups shared by several CTAs are
    finished by the last one to deliver) against the plain launch of the same shape (one whole group per CTA): the structured sums
    have one order whichever CTA computed which unit, so chains, log-probabilities, acceptance counts and the initial evaluation must be
    bit-identical; odd walker counts leave a partial last group."""
    from lightcurve_fitting_b200._capi import lib, check
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    wl = {'sc3_synth': lambda: W.synthetic_sc3(npoints=1100), 'sc4_synth': lambda: W.synthetic_sc4(npoints=1200),
          'cs3_synth': lambda: W.synthetic_cs3(npoints=1000)}[name]()
    prob = wl.device_problem(precision)
    p0 = wl.start(nw, np.random.default_rng(8))
    out = {}
    check(lib().lcf_set_tuning_ex(*shape))
    check(lib().lcf_set_tuning_split(1))                          # no split-K either: the launch shape is pinned completely
    try:
        for flat in (0, 1):
            check(lib().lcf_set_tuning_flat(flat))
            s = EnsembleSampler(nw, wl.ndim, prob, seed=21)
            s.run_mcmc(p0, 6, skip_initial_state_check=True)
            launch = prob.last_launch()
            assert launch['sum_units'] == 8, launch
            assert launch['flat'] == bool(flat), launch
            lp = prob.log_posterior(p0[:5000])
            out[flat] = (s.get_chain(), s.get_log_prob(), s.acceptance_fraction, lp)
    finally:
        check(lib().lcf_set_tuning_flat(-1))
        check(lib().lcf_set_tuning_split(0))
        check(lib().lcf_set_tuning_ex(0, 0, 0))
    assert 0.02 < out[1][2].mean() < 0.98
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a, b)
    _check_chain_rows_against_oracle(wl, out[1][0], out[1][1], precision, 8, 3)


# ---- filter banks larger than shared memory: streamed in segments of consecutive filters (k_pass_seg) ------------------------------
_SEG_CASES = {
    'sc4_example': ('fp32', 120), 'sc4_example_sigma_abs': ('fp64', 64), 'sc3_synth_sigma': ('fp32', 100), 'sc3_synth': ('fp64', 150),
    'cs3_synth': ('fp64', 90), 'sed': ('fp32', 50),
}


@pytest.mark.parametrize('name', sorted(_SEG_CASES))
def test_segmented_bank_is_bit_identical_to_the_unsegmented_launch(name, monkeypatch):
    """LCF_SEG_SAMPLES forces the fallback of a bank larger than shared memory on an ordinary problem: the bank slices of a few
    filters at a time are staged in turn and every warp still meets its tiles in table order, so log-posteriors, model grids and
    whole chains are those of the unsegmented launch of the same shape bit for bit -- and the oracle's within the tolerance."""
    from lightcurve_fitting_b200._capi import lib, check
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    precision, cap = _SEG_CASES[name]
    wl = WORKLOADS[name]()
    shape = (4, 16, 1) if wl.model_name == 'ShockCooling3' else (8, 16, 1)     # the fixed shape of the segmented launches
    P = _params(wl, 37, 3, widen=0.1)
    p0 = wl.start(50, np.random.default_rng(2))
    monkeypatch.setenv('LCF_RING', '0')                                          # half-step launches on both sides
    out = {}
    for seg in (True, False):
        if seg:
            monkeypatch.setenv('LCF_SEG_SAMPLES', str(cap))
        else:
            monkeypatch.delenv('LCF_SEG_SAMPLES')
            check(lib().lcf_set_tuning_ex(*shape))
            check(lib().lcf_set_tuning_split(1))
        try:
            prob = wl.device_problem(precision)
            lp = prob.log_posterior(P)
            ll = prob.last_launch()
            grid = prob.model_eval(P[:5, :prob.nmodel] if hasattr(prob, 'nmodel') else P[:5])
            s = EnsembleSampler(50, wl.ndim, prob, seed=17)
            s.run_mcmc(p0, 6)
            out[seg] = (lp, grid, s.get_chain(), s.get_log_prob(), s.acceptance_fraction, ll, prob.last_launch())
        finally:
            check(lib().lcf_set_tuning_ex(0, 0, 0))
            check(lib().lcf_set_tuning_split(0))
    assert out[True][5]['kernel'] == 'k_pass_seg' and out[True][6]['kernel'] == 'k_pass_seg'
    assert out[False][5]['kernel'] == 'k_pass<generic>'
    for k in ('walkers_per_cta', 'warps_per_cta', 'cluster', 'sum_units'):
        assert out[True][5][k] == out[False][5][k], k
    for k in range(5):
        np.testing.assert_array_equal(out[True][k], out[False][k])
    lpo = W.oracle_log_posterior(wl)
    want = np.array([lpo(p) for p in P])
    fin = np.isfinite(want)
    assert np.array_equal(np.isneginf(want), np.isneginf(out[True][0]))
    np.testing.assert_allclose(out[True][0][fin], want[fin], rtol=RTOL[precision])


@pytest.mark.parametrize('model', ['ShockCooling4', 'ShockCooling3'])
def test_filter_bank_larger_than_shared_memory_fp64(model):
    """A light curve through the 30 most densely sampled filters of the registry (JWST MIRI / NIRCam, GALEX, Flamingos-2 ...:
    > 17 000 transmission samples, 16 B each in FP64) does not fit in the 227 KB of shared memory of an SM: the reference has
    no such limit (filters.py:296-335 integrates whatever it is given), so the library streams the bank (round 1: LCF_ERR_ARG)."""
    from lightcurve_fitting_b200.synthetic import Workload
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    names = ['F2550W', 'F2100W', 'NUV', 'F1800W', 'F444W', 'Itagaki', 'F356W', 'F1500W', 'F277W', 'K', 'H', 'Kepler', 'F200W', 'F1280W',
             'TESS', 'F335M', 'F360M', 'F770W', 'F1000W', 'FUV', 'w', 'F300M', 'F150W', 'UVW1', 'J', 'r-DECam', 'U', 'B', 'g', 'i']
    rng = np.random.default_rng(21)
    n = 2 * len(names) + 7
    t0 = 59000.
    t = np.sort(rng.uniform(t0 + 0.3, t0 + 12., n))
    fn = [names[i % len(names)] for i in range(n)]
    if model == 'ShockCooling3':
        p_true = np.array([1., 1., 1., 3., 20., 0.1, t0])
        pri = [('uniform', 0., 10.), ('uniform', 0., 10.), ('uniform', 0., 100.), ('uniform', 0., 100.), ('uniform', 5., 50.),
               ('uniform', 0., 1.), ('uniform', t0 - 2., t0 + 0.3)]
        tcol = 6
    else:
        p_true = np.array([1., 1., 1., 3., t0])
        pri = [('uniform', 0., 10.), ('uniform', 0., 10.), ('uniform', 0., 100.), ('uniform', 0., 100.), ('uniform', t0 - 2., t0 + 0.3)]
        tcol = 4
    ytrue = W.oracle_truth(model, t, fn, p_true, 0.005)
    dy = 0.05 * ytrue
    y = ytrue + dy * rng.normal(size=n)
    lo, hi = p_true * 0.9, p_true * 1.1
    lo[tcol], hi[tcol] = t0 - 0.1, t0 + 0.1
    wl = Workload('large-bank-' + model, model, t, fn, y, dy, pri, lo, hi, z=0.005, truth=p_true)
    prob = wl.device_problem('fp64')
    _check_logpost(wl, 'fp64', n=24, seed=6, widen=0.05)
    got = prob.log_posterior(_params(wl, 24, 6, 0.05))
    assert prob.last_launch()['kernel'] == 'k_pass_seg'
    lpo = W.oracle_log_posterior(wl)
    s = EnsembleSampler(64, wl.ndim, prob, seed=4)
    s.run_mcmc(wl.start(64, rng), 5)
    assert prob.last_launch()['kernel'] == 'k_pass_seg'
    chain, lnp = s.get_chain(), s.get_log_prob()
    for (i, j) in ((0, 3), (2, 40), (4, 63), (4, 0)):
        np.testing.assert_allclose(lnp[i, j], lpo(chain[i, j]), rtol=RTOL['fp64'])
    assert (s.acceptance_fraction > 0).any()
