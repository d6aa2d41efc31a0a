"""CPU tests: the oracle (numpy restatement) against golden vectors produced by the REFERENCE's own code
(tests/golden/make_golden.py ran /root/reference's models.py / filters.py / fitting.py), plus analytic identities.
"""
import os

import numpy as np
import pytest

from oracle import reference_port as rp

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden.npz'))
RT = 1e-12


def of(names):
    return np.array([rp.filtdict[str(n)] for n in names], dtype=object)


def test_constants():
    np.testing.assert_allclose([rp.k_B, rp.c1, rp.c2, rp.c3, rp.c4, rp.c_AA_THz], G['const'], rtol=1e-14)
    # values the survey computed with a real astropy install
    np.testing.assert_allclose([rp.k_B, rp.c3, rp.c4, rp.c1, rp.c2, rp.sigma_sb],
                               [0.08617333262145178, 5.38477047522316e-19, 8.357743635931361e-47, 0.04799243073366221,
                                281739904251.4432, 2.744452656619892e28], rtol=1e-14)


def test_filter_curves_match_reference():
    for n in G['filters/names']:
        f = rp.filtdict[str(n)]
        np.testing.assert_allclose(f.trans['freq'], G['filters/%s/freq' % n], rtol=1e-14)
        np.testing.assert_allclose(f.trans['T_norm_per_freq'], G['filters/%s/Tn' % n], rtol=1e-12, atol=0)
        s = G['filters/%s/scalars' % n]
        np.testing.assert_allclose([f.freq_eff, f.dfreq], s[:2], rtol=1e-12)
        np.testing.assert_allclose([f.m0, f.M0], s[3:5], rtol=1e-14)
        assert f.char == str(G['filters/%s/char' % n])


def test_filter_normalisation_identity():
    """integral of T_norm_per_freq over nu is 1 for every filter with a curve."""
    for f in rp.all_filters:
        if f.filename:
            assert abs(rp._trapz(f.trans['T_norm_per_freq'], f.trans['freq']) - 1.) < 1e-12


def test_planck_and_blackbody_to_filters():
    nu, T, R = G['planck/nu'], G['planck/T'], G['planck/R']
    np.testing.assert_allclose(rp.planck_fast(nu, T, R), G['planck/out'], rtol=RT)
    np.testing.assert_allclose(rp.planck_fast(nu, T, R, 900.), G['planck/out_cutoff'], rtol=RT)
    np.testing.assert_allclose(rp.planck_fast(nu, 12., 3.), G['planck/out_scalar'], rtol=RT)
    assert np.all(G['planck/out'][-2:] == 0.)          # T <= 0 -> exactly zero (power() semantics)
    names, Tb, Rb = G['bb/names'], G['bb/T'], G['bb/R']
    for tag, kw in (('plain', {}), ('z', {'z': 0.05}), ('cut', {'z': 0.01, 'cutoff_freq': 700.}), ('ebv', {'ebv': 0.2}),
                    ('zebv', {'z': 0.02, 'ebv': 0.35})):
        np.testing.assert_allclose(rp.blackbody_to_filters(of(names), Tb, Rb, **kw), G['bb/point_' + tag], rtol=RT)
        np.testing.assert_allclose(rp.blackbody_to_filters(of(names[:3]), Tb, Rb, **kw), G['bb/grid_' + tag], rtol=RT)
    np.testing.assert_allclose(rp.blackbody_to_filters(of(names[:2]), np.outer(Tb[:3], [1., 1.1]), np.outer(Rb[:3], [1., 0.9])),
                               G['bb/grid2d'], rtol=RT)
    np.testing.assert_allclose(rp.blackbody_to_filters(of(names[:3]), Tb[:4], Rb[:4], ebv=np.array([0., 0.1, 0.2, 0.3])),
                               G['bb/ebv_vec'], rtol=RT)


def test_planck_limits():
    """Rayleigh-Jeans and Wien limits of planck_fast."""
    T, R = 50., 2.
    nu = np.array([1e-3])
    rj = rp.c2 * R ** 2 * nu ** 2 * T / rp.c1
    np.testing.assert_allclose(rp.planck_fast(nu, T, R), rj[0], rtol=1e-5)
    nu = np.array([3e4])
    wien = rp.c2 * R ** 2 * nu ** 3 * np.exp(-rp.c1 * nu / T)
    np.testing.assert_allclose(rp.planck_fast(nu, T, R), wien[0], rtol=1e-10)


MODELS = {
    'sc_n15': lambda: rp.ShockCooling(redshift=0.002),
    'sc_n3': lambda: rp.ShockCooling(redshift=0.002, n=3.),
    'sc_rw': lambda: rp.ShockCooling(redshift=0.002, RW=True),
    'sc2': lambda: rp.ShockCooling2(redshift=0.002),
    'sc3': lambda: rp.ShockCooling3(redshift=0.005),
    'sc4': lambda: rp.ShockCooling4(redshift=0.002),
}


@pytest.mark.parametrize('tag', sorted(MODELS))
def test_shock_cooling_models_match_reference(tag):
    m = MODELS[tag]()
    t, f = G['lc/t'], of(G['lc/filters'])
    P, y, dy = G[tag + '/P'], G[tag + '/y'], G[tag + '/dy']
    got = np.array([m(t, f, *p) for p in P])
    np.testing.assert_allclose(got, G[tag + '/point'], rtol=RT, atol=0)
    assert (G[tag + '/point'] == 0.).any()      # points before t_0 contribute exactly zero
    np.testing.assert_allclose(m(G['lc/tgrid'], of(G['lc/gridfilters']), *P.T), G[tag + '/grid'], rtol=RT)
    np.testing.assert_allclose([m.log_likelihood(t, f, y, dy, p) for p in P], G[tag + '/loglike'], rtol=RT)
    for st in ('relative', 'absolute'):
        Ps = G[tag + '/Psig_' + st]
        np.testing.assert_allclose([m.log_likelihood(t, f, y, dy, p, use_sigma=True, sigma_type=st) for p in Ps],
                                   G[tag + '/loglike_sig_' + st], rtol=RT)


def test_rw_suppression_is_identity():
    """RW=True => a = 0 => exp(-power(0, alpha)) = 1 (models.py:221-224)."""
    m = rp.ShockCooling(redshift=0., RW=True)
    T1, R1 = m.temperature_radius(np.array([1., 2., 5.]), 1., 1., 1., 3., 0.)
    m2 = rp.ShockCooling(redshift=0., RW=True)
    m2.A = m.A
    L_ratio = (R1 ** 2 * T1 ** 4)
    t = np.array([1., 2., 5.])
    np.testing.assert_allclose(L_ratio / L_ratio[0], (t / t[0]) ** m.epsilon_L, rtol=1e-12)


def test_sc4_precedence_quirk():
    """models.py:586 computes v_s ** (0.58 ** (f_rho_M ** 0.03)), not v_s^0.58 f^0.03."""
    m = rp.ShockCooling4()
    v, f, R = 2., 1.5, 3.
    T_K, _ = m.temperature_radius(np.array([1.]), v, 1., f, R)
    t_br = 0.036 * R ** 1.26 * v ** -1.13 * f ** -0.13
    tt = 1. / t_br
    expect = 8.19 * R ** -0.32 * v ** (0.58 ** (f ** 0.03)) * min(0.97 * tt ** (-1. / 3.), tt ** -0.45) / rp.k_B
    np.testing.assert_allclose(T_K, expect, rtol=1e-13)


@pytest.mark.parametrize('tag,cls', [('cs1', 'CompanionShocking'), ('cs2', 'CompanionShocking2'), ('cs3', 'CompanionShocking3')])
def test_companion_models_match_reference(tag, cls):
    t, f, y, dy = G['cs/t'], of(G['cs/filters']), G['cs/y'], G['cs/dy']
    m = getattr(rp, cls)(f, y, redshift=0.01)
    P = G[tag + '/P']
    np.testing.assert_allclose(np.array([m(t, f, *p) for p in P]), G[tag + '/point'], rtol=1e-11)
    np.testing.assert_allclose(m(G['cs/tgrid'], of(G['cs/gridfilters']), *P.T), G[tag + '/grid'], rtol=1e-11)
    np.testing.assert_allclose([m.log_likelihood(t, f, y, dy, p) for p in P], G[tag + '/loglike'], rtol=1e-11)
    np.testing.assert_allclose([m.log_likelihood(t, f, y, dy, p, use_sigma=True) for p in G[tag + '/Psig']],
                               G[tag + '/loglike_sig'], rtol=1e-11)


def test_priors_match_reference():
    x = G['prior/x']
    np.testing.assert_array_equal([rp.UniformPrior(0., 10.)(v) for v in x], G['prior/uniform'])
    np.testing.assert_allclose([rp.LogUniformPrior(0., 10.)(v) for v in x], G['prior/loguniform'], rtol=1e-15)
    np.testing.assert_allclose([rp.GaussianPrior(0., 10., 2., 1.5)(v) for v in x], G['prior/gaussian'], rtol=1e-15)
    with pytest.raises(ValueError):
        rp.LogUniformPrior(-1., 1.)


def _driver(seed, use_sigma, sigma_type, nwalkers, nsteps, nburn, extra_prior=()):
    """fitting.py:121-145 restated with the oracle pieces, same RNG protocol as the reference run."""
    m = rp.ShockCooling4(redshift=0.002)
    t, f = G['lc/t'], of(G['lc/filters'])
    y, dy = G['sc4/y'], G['sc4/dy']
    priors = [rp.UniformPrior(0., 10.), rp.UniformPrior(0., 10.), rp.UniformPrior(0., 100.), rp.UniformPrior(0., 100.),
              rp.UniformPrior(57460., 57468.5)] + list(extra_prior)
    p_lo = np.append(G['mcmc/p_lo'], [0.] if use_sigma else [])
    p_up = np.append(G['mcmc/p_up'], [2.] if use_sigma else [])
    lp = rp.make_log_posterior(m, priors, t, f, y, dy, use_sigma=use_sigma, sigma_type=sigma_type)
    np.random.seed(seed)
    rs = np.random.RandomState()
    rs.set_state(np.random.get_state())
    s = rp.StretchReplay(nwalkers, len(priors), lp, random_state=rs)
    start = np.random.rand(nwalkers, len(priors)) * (p_up - p_lo) + p_lo
    pos, _, _ = s.run_mcmc(start, nburn)
    s.reset()
    s.run_mcmc(pos, nsteps)
    return s


def test_driver_chain_matches_reference_lightcurve_mcmc():
    s = _driver(12345, False, 'relative', 12, 6, 5)
    np.testing.assert_allclose(s.flatchain, G['mcmc/flatchain'], rtol=1e-12)
    np.testing.assert_allclose(s.chain, G['mcmc/chain'], rtol=1e-12)
    np.testing.assert_allclose(s.get_log_prob(), G['mcmc/lnprob'], rtol=1e-12)
    np.testing.assert_array_equal(s.acceptance_fraction, G['mcmc/acceptance'])
    s2 = _driver(777, True, 'absolute', 14, 5, 4, extra_prior=[rp.GaussianPrior(0., 10.)])
    np.testing.assert_allclose(s2.flatchain, G['mcmc_sigma/flatchain'], rtol=1e-12)
    np.testing.assert_allclose(s2.get_log_prob(), G['mcmc_sigma/lnprob'], rtol=1e-12)
    assert int(G['mcmc_sigma/nparams_after']) == 6     # use_sigma appended '\\sigma' (fitting.py:74-76)


def test_stretch_replay_properties():
    """Detailed balance sanity: sampling a 3-d Gaussian recovers its mean and variance; walkers < 2*ndim raises."""
    rs = np.random.RandomState(3)
    mu, sig = np.array([1., -2., 0.5]), np.array([0.5, 2., 1.])
    lp = lambda p: -0.5 * np.sum(((p - mu) / sig) ** 2)
    s = rp.StretchReplay(24, 3, lp, random_state=rs)
    pos, lnp, _ = s.run_mcmc(mu + rs.randn(24, 3), 300)
    s.reset()
    s.run_mcmc(pos, 1500, log_prob0=lnp)
    fc = s.flatchain
    assert np.all(np.abs(fc.mean(0) - mu) < 0.15 * sig)
    assert np.all(np.abs(fc.std(0) / sig - 1.) < 0.12)
    assert 0.2 < s.acceptance_fraction.mean() < 0.8
    with pytest.raises(RuntimeError):
        rp.StretchReplay(4, 3, lp).run_mcmc(np.zeros((4, 3)), 1)
    with pytest.raises(ValueError, match='NaN'):
        rp.StretchReplay(8, 3, lambda p: np.nan, random_state=rs).run_mcmc(rs.randn(8, 3), 1)


# ---- bolometric.py (reference run through the shim: tests/golden/make_golden.py, section "bolometric.py") ----------------
def test_bolometric_post_processing_matches_reference():
    T, R = G['bolo/pseudo/T'], G['bolo/pseudo/R']
    np.testing.assert_allclose(rp.pseudo(T, R, 0.), G['bolo/pseudo/z0'], rtol=RT)
    np.testing.assert_allclose(rp.pseudo(T, R, 0.023), G['bolo/pseudo/z'], rtol=RT)
    np.testing.assert_allclose(rp.pseudo(T, R, 0.01, cutoff_freq=800.), G['bolo/pseudo/cut'], rtol=RT)
    np.testing.assert_allclose(rp.pseudo(12., 3., 0.002), G['bolo/pseudo/scalar'], rtol=RT)
    np.testing.assert_allclose(rp.pseudo(T, R, 0.002, filter0=rp.filtdict['V'], filter1=rp.filtdict['B']), G['bolo/pseudo/BtoV'], rtol=RT)
    np.testing.assert_allclose(rp.sigma_sb, G['bolo/sigma_sb'], rtol=1e-14)
    np.testing.assert_allclose(rp.stefan_boltzmann(T, R), G['bolo/sb/lum'], rtol=1e-14)
    lum, dlum = rp.stefan_boltzmann(T, R, G['bolo/sb/dT'], G['bolo/sb/dR'], G['bolo/sb/cov'])
    np.testing.assert_allclose(lum, G['bolo/sb/lum2'], rtol=1e-14)
    np.testing.assert_allclose(dlum, G['bolo/sb/dlum'], rtol=1e-14)
    np.testing.assert_allclose(np.array(rp.median_and_unc(G['bolo/mu/x1'])), G['bolo/mu/out1'], rtol=1e-14)
    np.testing.assert_allclose(np.array(rp.median_and_unc(G['bolo/mu/x2'])), G['bolo/mu/out2'], rtol=1e-14)
    np.testing.assert_allclose(np.array(rp.median_and_unc(G['bolo/mu/x1'], 100.)), G['bolo/mu/out1_100'], rtol=1e-14)


def test_blackbody_lstsq_matches_reference():
    off, out = G['bolo/lstsq/offsets'], G['bolo/lstsq/out']
    for k in range(len(off) - 1):
        sl = slice(off[k], off[k + 1])
        # the effective frequencies the reference's Filter objects carried are the oracle's (filters/*/scalars pins the rest)
        np.testing.assert_allclose([float(rp.filtdict[str(n)].freq_eff) for n in G['bolo/lstsq/filters'][sl]], G['bolo/lstsq/freq'][sl], rtol=1e-12)
        got = rp.blackbody_lstsq(G['bolo/lstsq/freq'][sl], G['bolo/lstsq/lum'][sl], float(G['bolo/lstsq/z']), cutoff_freq=out[k, 7],
                                 T_range=(1., out[k, 8]))
        # curve_fit stops at ftol = xtol = 1e-8 with a finite-difference Jacobian: two runs agree to the optimiser's tolerance
        np.testing.assert_allclose(got, out[k, :7], rtol=2e-6)


def _sed_driver(tag, use_sigma, sigma_type, record=False):
    names = [str(n) for n in G[tag + '/filters']]
    model = rp.BlackbodySED(redshift=0.002, cutoff_freq=float(G[tag + '/cutoff']))
    priors = [rp.UniformPrior(1., 100.), rp.LogUniformPrior(0.01, 1000.)] + ([rp.GaussianPrior(0., 10.)] if use_sigma else [])
    lp = rp.make_log_posterior(model, priors, np.zeros(len(names)), of(names), G[tag + '/lum'], G[tag + '/dlum'], use_sigma=use_sigma,
                               sigma_type=sigma_type)
    np.random.seed(4242)                       # bolometric.py:166-174 under the emcee stand-in: private copy of the global RNG
    rs = np.random.RandomState()
    rs.set_state(np.random.get_state())
    s = rp.StretchReplay(10, len(priors), lp, random_state=rs)
    pos, _, _ = s.run_mcmc(G[tag + '/start'], 12, record=record)
    burn = list(s.draws)
    s.reset()
    s.run_mcmc(pos, 9, record=record)
    return s, burn


@pytest.mark.parametrize('tag,use_sigma,sigma_type', [('bolo/mcmc', False, 'relative'), ('bolo/mcmc_sigma', True, 'absolute')])
def test_spectrum_mcmc_chain_matches_reference(tag, use_sigma, sigma_type):
    s, _ = _sed_driver(tag, use_sigma, sigma_type)
    np.testing.assert_allclose(s.flatchain, G[tag + '/flatchain'], rtol=1e-12)
    np.testing.assert_allclose(s.get_log_prob(), G[tag + '/lnprob'], rtol=1e-12)
    np.testing.assert_array_equal(s.acceptance_fraction, G[tag + '/acceptance'])


def test_every_filter_curve_matches_reference_moments():
    """All filters with a transmission curve (not only the 16 whose full curves are frozen): sample count, end frequencies and
    five moments of (freq, T_norm_per_freq) as the reference's read_curve produced them, plus freq_eff / dfreq / wl_eff / zero
    points.  A packing error in data/filter_curves.npz or filter_registry.py, which the oracle and the product share, cannot
    hide behind a common mode."""
    names, mom = G['filters_all/names'], G['filters_all/moments']
    assert len(names) >= 60
    for n, row in zip(names, mom):
        f = rp.filtdict[str(n)]
        nu, tn = np.asarray(f.trans['freq'], float), np.asarray(f.trans['T_norm_per_freq'], float)
        got = [len(nu), nu[0], nu[-1], nu.sum(), tn.sum(), (nu * tn).sum(), (nu * nu * tn).sum(), np.abs(np.diff(tn)).sum(),
               float(f.freq_eff), float(f.dfreq), row[10], f.m0, f.M0]      # (the oracle's filter does not carry wl_eff)
        np.testing.assert_allclose(got, row, rtol=1e-11, err_msg=str(n))
