"""Generate golden vectors by running the REFERENCE's own code (models.py, filters.py, fitting.py, bolometric.py).

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference imports astropy / extinction / emcee / corner / matplotlib, which are not installed here; the
stand-ins under oracle/refshim/ supply the few calls it makes (see oracle/refshim/README.md).  Everything
numerical below is executed by the reference's code: the Planck function, the filter normalisation and
synthesis, every Model.evaluate, Model.log_likelihood, the Prior classes, and -- through lightcurve_mcmc with the
emcee stand-in -- the log_posterior closure and the driver (burn-in, reset, sampling, flatchain layout).

Outputs: tests/golden/reference_golden.npz (inputs + expected outputs, a few hundred kB).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('LCF_REFERENCE', '/root/reference')
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle', 'refshim'), REF]
warnings.simplefilter('ignore')

import lightcurve_fitting.models as RM      # noqa: E402  (the reference)
import lightcurve_fitting.filters as RF     # noqa: E402
import lightcurve_fitting.fitting as RFit   # noqa: E402
import lightcurve_fitting.bolometric as RB  # noqa: E402

# spectrum_mcmc always ends by drawing a corner plot (bolometric.py:183-188): presentation, out of scope;
# everything before it -- the log_posterior closure and the emcee driver -- is the reference's own code
RB.spectrum_corner = lambda *a, **k: None


class Col(np.ndarray):
    def __new__(cls, a):
        return np.asarray(a).view(cls)

    @property
    def data(self):
        return np.asarray(self)


class MiniLC:
    """The slice of the LC interface that reference models.py / fitting.py touch."""

    def __init__(self, **cols):
        self.cols = {k: np.asarray(v) for k, v in cols.items()}
        self.meta = {}

    @property
    def colnames(self):
        return list(self.cols)

    def __getitem__(self, k):
        return Col(self.cols[k])

    def __len__(self):
        return len(next(iter(self.cols.values())))

    def where(self, filter=None):
        f = RF.filtdict[filter] if isinstance(filter, str) else filter
        sel = np.array([x == f for x in self.cols['filter']])
        return MiniLC(**{k: v[sel] for k, v in self.cols.items()})

    def calcFlux(self):
        pass

    def calcAbsMag(self):
        pass

    def calcLum(self):
        pass


def rfilters(names):
    return np.array([RF.filtdict[n] for n in names], dtype=object)


def main():
    G = {}
    rng = np.random.default_rng(20261018)

    # ---- constants -----------------------------------------------------------------------------------
    G['const'] = np.array([RM.k_B, RM.c1, RM.c2, RM.c3, RM.c4, RF.c])

    # ---- filter curves (read_curve) ------------------------------------------------------------------
    fnames = ['U', 'B', 'V', 'R', 'I', 'g', 'r', 'i', '0', 'UVW2', 'DLT40', 'w', 'F444W', 'white', 'z', 'o']
    G['filters/names'] = np.array(fnames)
    for n in fnames:
        f = RF.filtdict[n]
        G['filters/%s/freq' % n] = np.asarray(f.trans['freq'].value)
        G['filters/%s/Tn' % n] = np.asarray(f.trans['T_norm_per_freq'].data)
        G['filters/%s/scalars' % n] = np.array([float(f.freq_eff.value), float(np.asarray(f.dfreq)), float(f.wl_eff.value),
                                                f.m0, f.M0])
        G['filters/%s/char' % n] = np.array(f.char)

    # ---- planck_fast / blackbody_to_filters -----------------------------------------------------------
    nu = np.linspace(80., 2500., 41)
    T = np.array([3., 5., 9., 14., 22., 40., 80., 0., -2.])
    R = np.array([10., 1., 2., 0.5, 3., 1.5, 0.2, 1., 1.])
    G['planck/nu'], G['planck/T'], G['planck/R'] = nu, T, R
    G['planck/out'] = RM.planck_fast(nu, T, R)
    G['planck/out_cutoff'] = RM.planck_fast(nu, T, R, 900.)
    G['planck/out_scalar'] = RM.planck_fast(nu, 12., 3.)
    bb_names = ['U', 'B', 'g', 'r', 'UVW2', 'F444W', 'DLT40']
    Tb = np.array([5., 9., 14., 22., 40., 3., 11.])
    Rb = np.array([1., 2., 0.5, 3., 1.5, 10., 2.5])
    G['bb/names'], G['bb/T'], G['bb/R'] = np.array(bb_names), Tb, Rb
    for tag, kw in (('plain', {}), ('z', {'z': 0.05}), ('cut', {'z': 0.01, 'cutoff_freq': 700.}), ('ebv', {'ebv': 0.2}),
                    ('zebv', {'z': 0.02, 'ebv': 0.35})):
        G['bb/point_' + tag] = RM.blackbody_to_filters(rfilters(bb_names), Tb, Rb, **kw)
        G['bb/grid_' + tag] = RM.blackbody_to_filters(rfilters(bb_names[:3]), Tb, Rb, **kw)
    G['bb/grid2d'] = RM.blackbody_to_filters(rfilters(bb_names[:2]), np.outer(Tb[:3], [1., 1.1]), np.outer(Rb[:3], [1., 0.9]))
    G['bb/ebv_vec'] = RM.blackbody_to_filters(rfilters(bb_names[:3]), Tb[:4], Rb[:4], ebv=np.array([0., 0.1, 0.2, 0.3]))

    # ---- light-curve models -----------------------------------------------------------------------------
    names9 = ['U', 'B', 'V', 'R', 'I', 'g', 'r', 'i', '0']
    N = 45
    t = np.sort(rng.uniform(57468.2, 57482., N))
    fn = [names9[i % 9] for i in rng.permutation(N)]
    G['lc/t'], G['lc/filters'] = t, np.array(fn)
    f = rfilters(fn)
    tg = np.linspace(57468.5, 57480., 6)
    gf = ['U', 'g', 'i']

    def run_model(tag, model, P, ygen):
        """P: [nsets, nparams]; stores pointwise values, a grid evaluation, and log-likelihoods."""
        y = ygen
        dy = np.abs(y) * rng.uniform(0.02, 0.08, N) + 1e-3 * np.abs(y).max()
        lc = MiniLC(MJD=t, filter=f, **{model.output_quantity: y, 'd' + model.output_quantity: dy})
        G[tag + '/P'], G[tag + '/y'], G[tag + '/dy'] = P, y, dy
        G[tag + '/point'] = np.array([model(t, f, *p) for p in P])
        G[tag + '/grid'] = model(tg, rfilters(gf), *P.T)
        G[tag + '/loglike'] = np.array([model.log_likelihood(lc, p) for p in P])
        for st in ('relative', 'absolute'):
            Ps = np.column_stack([P, rng.uniform(0.1, 3., len(P))])
            G[tag + '/Psig_' + st] = Ps
            G[tag + '/loglike_sig_' + st] = np.array([model.log_likelihood(lc, p, use_sigma=True, sigma_type=st) for p in Ps])
        return lc

    def box(lo, hi, n=6):
        lo, hi = np.asarray(lo, float), np.asarray(hi, float)
        return lo + rng.random((n, len(lo))) * (hi - lo)

    sc_lo, sc_hi = [0.5, 0.1, 0.1, 1., 57467.], [2., 2., 10., 10., 57470.]     # some points precede t_0
    G['lc/tgrid'], G['lc/gridfilters'] = tg, np.array(gf)
    for tag, kw in (('sc_n15', {}), ('sc_n3', {'n': 3.}), ('sc_rw', {'RW': True})):
        m = RM.ShockCooling(redshift=0.002, **kw)
        P = box(sc_lo, sc_hi)
        run_model(tag, m, P, m(t, f, *P[0]) * (1. + 0.05 * rng.normal(size=N)))
    m = RM.ShockCooling2(redshift=0.002)
    P = box([10., 0.5, 2., 57467.], [30., 5., 10., 57470.])
    run_model('sc2', m, P, m(t, f, *P[0]) * (1. + 0.05 * rng.normal(size=N)))
    m = RM.ShockCooling3(redshift=0.005)
    P = box([0.5, 0.1, 0.1, 1., 10., 0., 57467.], [2., 2., 10., 10., 40., 0.5, 57470.])
    run_model('sc3', m, P, m(t, f, *P[0]) * (1. + 0.05 * rng.normal(size=N)))
    m = RM.ShockCooling4(redshift=0.002)
    P = box(sc_lo, sc_hi)
    lc4 = run_model('sc4', m, P, m(t, f, *P[0]) * (1. + 0.05 * rng.normal(size=N)))

    # CompanionShocking family: the SiFTO templates are scaled to the light curve's peak luminosities
    names6 = ['U', 'B', 'V', 'g', 'r', 'i']
    tpk = 58000.
    tc = np.sort(rng.uniform(tpk - 19., tpk + 40., N))
    fnc = [names6[i % 6] for i in rng.permutation(N)]
    fc = rfilters(fnc)
    yc = 1e21 * np.exp(-0.5 * ((tc - tpk) / 12.) ** 2) * rng.uniform(0.8, 1.2, N) + 1e19
    dyc = 0.04 * yc
    lcc = MiniLC(MJD=tc, filter=fc, lum=yc, dlum=dyc)
    G['cs/t'], G['cs/filters'], G['cs/y'], G['cs/dy'] = tc, np.array(fnc), yc, dyc
    tgc = np.linspace(tpk - 18., tpk + 30., 6)
    G['cs/tgrid'], G['cs/gridfilters'] = tgc, np.array(gf)
    for tag, cls, lo, hi in (
            ('cs1', RM.CompanionShocking, [tpk - 18., 0.05, 0.5, tpk - 1., 0.8, 0.7, 0.7, 0.7], [tpk - 16., 0.3, 2., tpk + 1., 1.2, 1.3, 1.3, 1.3]),
            ('cs2', RM.CompanionShocking2, [tpk - 18., 0.05, 0.5, tpk - 1., 0.8, -1., -1.], [tpk - 16., 0.3, 2., tpk + 1., 1.2, 1., 1.]),
            ('cs3', RM.CompanionShocking3, [tpk - 18., 0.05, 0., tpk - 1., 0.8, -1., -1.], [tpk - 16., 0.3, 180., tpk + 1., 1.2, 1., 1.])):
        m = cls(lcc, redshift=0.01)
        P = box(lo, hi)
        G[tag + '/P'] = P
        G[tag + '/point'] = np.array([m(tc, fc, *p) for p in P])
        G[tag + '/grid'] = m(tgc, rfilters(gf), *P.T)
        G[tag + '/loglike'] = np.array([m.log_likelihood(lcc, p) for p in P])
        Ps = np.column_stack([P, rng.uniform(0.1, 3., len(P))])
        G[tag + '/Psig'] = Ps
        G[tag + '/loglike_sig'] = np.array([m.log_likelihood(lcc, p, use_sigma=True) for p in Ps])

    # ---- priors ---------------------------------------------------------------------------------------
    x = np.array([-1., 0., 1e-3, 0.5, 1., 2., 9.999, 10., 11.])
    G['prior/x'] = x
    G['prior/uniform'] = np.array([RM.UniformPrior(0., 10.)(v) for v in x])
    G['prior/loguniform'] = np.array([RM.LogUniformPrior(0., 10.)(v) for v in x])
    G['prior/gaussian'] = np.array([RM.GaussianPrior(0., 10., 2., 1.5)(v) for v in x])

    # ---- the driver: reference lightcurve_mcmc (closure + burn-in/reset/sampling) on the emcee stand-in ------
    m = RM.ShockCooling4(redshift=0.002)
    priors = [RM.UniformPrior(0., 10.), RM.UniformPrior(0., 10.), RM.UniformPrior(0., 100.), RM.UniformPrior(0., 100.),
              RM.UniformPrior(57460., 57468.5)]
    p_lo, p_up = [0.5, 0.1, 0.1, 1., 57467.5], [2., 2., 10., 10., 57468.2]
    np.random.seed(12345)
    sampler = RFit.lightcurve_mcmc(lc4, m, priors=priors, p_lo=p_lo, p_up=p_up, nwalkers=12, nsteps=6, nsteps_burnin=5)
    G['mcmc/p_lo'], G['mcmc/p_up'] = np.array(p_lo), np.array(p_up)
    G['mcmc/flatchain'] = sampler.flatchain
    G['mcmc/chain'] = sampler.chain
    G['mcmc/lnprob'] = sampler.get_log_prob()
    G['mcmc/acceptance'] = sampler.acceptance_fraction
    # with the intrinsic-scatter parameter and a Gaussian prior on it
    m2 = RM.ShockCooling4(redshift=0.002)
    RM.ShockCooling4.input_names = RM.ShockCooling4.input_names[:5]          # undo the class-level append (fitting.py:74-76)
    RM.ShockCooling4.units = RM.ShockCooling4.units[:5]
    np.random.seed(777)
    s2 = RFit.lightcurve_mcmc(lc4, m2, priors=priors + [RM.GaussianPrior(0., 10.)], p_lo=p_lo + [0.], p_up=p_up + [2.],
                              nwalkers=14, nsteps=5, nsteps_burnin=4, use_sigma=True, sigma_type='absolute')
    G['mcmc_sigma/flatchain'] = s2.flatchain
    G['mcmc_sigma/lnprob'] = s2.get_log_prob()
    G['mcmc_sigma/nparams_after'] = np.array(m2.nparams)

    # ---- bolometric.py: pseudo, stefan_boltzmann, median_and_unc, blackbody_lstsq, spectrum_mcmc ----------------------
    Tp = np.array([4., 6., 9., 11., 14., 18., 25., 33., 47., 60., 85., 2.])
    Rp = np.array([10., 3., 2.5, 1.5, 1.2, 1., 0.7, 0.5, 0.4, 0.3, 0.2, 30.])
    G['bolo/pseudo/T'], G['bolo/pseudo/R'] = Tp, Rp
    G['bolo/pseudo/z0'] = RB.pseudo(Tp, Rp, 0.)
    G['bolo/pseudo/z'] = RB.pseudo(Tp, Rp, 0.023)
    G['bolo/pseudo/cut'] = RB.pseudo(Tp, Rp, 0.01, cutoff_freq=800.)
    G['bolo/pseudo/scalar'] = np.array(RB.pseudo(12., 3., 0.002))
    G['bolo/pseudo/BtoV'] = RB.pseudo(Tp, Rp, 0.002, filter0=RF.filtdict['V'], filter1=RF.filtdict['B'])
    G['bolo/sigma_sb'] = np.array(float(RB.sigma_sb))
    dT, dR, cTR = 0.07 * Tp, 0.05 * Rp, -0.6 * (0.07 * Tp) * (0.05 * Rp)
    G['bolo/sb/dT'], G['bolo/sb/dR'], G['bolo/sb/cov'] = dT, dR, cTR
    G['bolo/sb/lum'] = RB.stefan_boltzmann(Tp, Rp)
    lum, dlum = RB.stefan_boltzmann(Tp, Rp, dT, dR, cTR)
    G['bolo/sb/lum2'], G['bolo/sb/dlum'] = lum, dlum
    x1 = rng.lognormal(1., 0.6, 1000)
    x2 = rng.normal(size=(777, 3)) * np.array([1., 5., 0.1]) + np.array([10., -3., 0.])
    G['bolo/mu/x1'], G['bolo/mu/x2'] = x1, x2
    G['bolo/mu/out1'] = np.array(RB.median_and_unc(x1))
    G['bolo/mu/out2'] = np.array(RB.median_and_unc(x2))
    G['bolo/mu/out1_100'] = np.array(RB.median_and_unc(x1, 100.))

    pool = ['U', 'B', 'V', 'g', 'r', 'i', 'R', 'I', '0', 'UVW1', 'z']
    z_b = 0.002

    def sed(nf, T, R, err, cutoff=np.inf):
        names = list(rng.choice(pool, nf, replace=False))
        filt = rfilters(names)
        lum = RM.blackbody_to_filters(filt, np.full(nf, T), np.full(nf, R), z=z_b, cutoff_freq=cutoff)
        lum = lum * (1. + err * rng.normal(size=nf))
        freq = np.array([float(f.freq_eff.value) for f in filt])
        return names, MiniLC(MJD=np.full(nf, 58000.25), filter=filt, lum=lum, dlum=err * np.abs(lum), freq=freq)

    off, names_all, freq_all, lum_all, res = [0], [], [], [], []
    for k in range(24):
        nf = int(rng.integers(2, 10))
        T, R = rng.uniform(4., 40.), np.exp(rng.uniform(np.log(0.3), np.log(30.)))
        cutoff = 900. if k % 6 == 5 else np.inf
        names, ep = sed(nf, T, R, 0.05, cutoff)
        kw = {'cutoff_freq': cutoff}
        if k % 8 == 7:
            kw.update(T_range=(1., 12.), R_range=(0.01, 1000.))        # a bound-hitting fit
        out = RB.blackbody_lstsq(ep, z_b, **kw)
        names_all += names
        freq_all += list(ep['freq'].data)
        lum_all += list(ep['lum'].data)
        off.append(off[-1] + nf)
        res.append(list(out) + [cutoff, kw.get('T_range', (1., 100.))[1]])
    G['bolo/lstsq/offsets'], G['bolo/lstsq/filters'] = np.array(off), np.array(names_all)
    G['bolo/lstsq/freq'], G['bolo/lstsq/lum'] = np.array(freq_all), np.array(lum_all)
    G['bolo/lstsq/out'] = np.array(res)          # temp, radius, dtemp, drad, lum, dlum, L_opt, cutoff_freq, T_max
    G['bolo/lstsq/z'] = np.array(z_b)

    # spectrum_mcmc: the reference's closure (bolometric.py:154-164) and driver (:166-174) on the emcee stand-in
    for tag, use_sigma, sigma_type, cutoff, nf in (('bolo/mcmc', False, 'relative', np.inf, 6),
                                                   ('bolo/mcmc_sigma', True, 'absolute', 1000., 4)):
        names, ep = sed(nf, 12., 2.5, 0.06, cutoff)
        priors = [RM.UniformPrior(1., 100.), RM.LogUniformPrior(0.01, 1000.)] + ([RM.GaussianPrior(0., 10.)] if use_sigma else [])
        sg = rng.normal(size=(10, 2)) * 0.5 + np.array([12., 2.5])
        if use_sigma:
            sg = np.append(sg, np.abs(rng.normal(size=(10, 1))), axis=1)
        np.random.seed(4242)
        smp = RB.spectrum_mcmc(RM.planck_fast, ep, priors, sg, z=z_b, spectrum_kwargs={'cutoff_freq': cutoff}, outpath='/tmp/lcf_golden_out',
                               nwalkers=10, burnin_steps=12, steps=9, use_sigma=use_sigma, sigma_type=sigma_type)
        G[tag + '/filters'], G[tag + '/lum'], G[tag + '/dlum'] = np.array(names), ep['lum'].data, ep['dlum'].data
        G[tag + '/start'], G[tag + '/cutoff'] = sg, np.array(cutoff)
        G[tag + '/flatchain'] = smp.flatchain
        G[tag + '/lnprob'] = smp.get_log_prob()
        G[tag + '/acceptance'] = smp.acceptance_fraction

    # ---- posterior statistics of a long reference run (statistical parity of the native device sampler) ---------------------
    # ShockCooling2 (T_1, L_1, t_tr, t_0): a well-constrained, unimodal posterior, truth inside the priors
    m3 = RM.ShockCooling2(redshift=0.002)
    p_true = np.array([20., 2., 6., 57467.8])
    srng = np.random.default_rng(99)
    ts = np.sort(srng.uniform(57468.3, 57482., N))
    fs = rfilters([names9[i % 8] for i in srng.permutation(N)])
    ys = m3(ts, fs, *p_true)
    dys = 0.05 * np.abs(ys)
    ys = ys + dys * srng.normal(size=N)
    lcs = MiniLC(MJD=ts, filter=fs, lum=ys, dlum=dys)
    pri3 = [RM.UniformPrior(5., 60.), RM.UniformPrior(0.1, 20.), RM.UniformPrior(1., 30.), RM.UniformPrior(57465., 57468.2)]
    np.random.seed(2024)
    s3 = RFit.lightcurve_mcmc(lcs, m3, priors=pri3, p_lo=p_true * [0.9, 0.9, 0.9, 1.] - [0, 0, 0, 0.1], p_up=p_true * [1.1, 1.1, 1.1, 1.] + [0, 0, 0, 0.1],
                              nwalkers=32, nsteps=1200, nsteps_burnin=600)
    G['stats/t'], G['stats/filters'], G['stats/y'], G['stats/dy'], G['stats/p_true'] = ts, np.array([f.name for f in fs]), ys, dys, p_true
    G['stats/percentiles'] = np.percentile(s3.flatchain, [16., 50., 84.], axis=0)
    G['stats/mean'], G['stats/cov'] = s3.flatchain.mean(axis=0), np.cov(s3.flatchain.T)
    G['stats/acceptance'] = np.array(s3.acceptance_fraction.mean())
    G['stats/setup'] = np.array([32, 600, 1200])

    # ---- every filter with a transmission curve: moments of the normalised curve as read by the reference (filters.py:170-230) --
    names_all, rows_all = [], []
    for f in RF.all_filters:
        if not f.filename:
            continue
        try:
            nu = np.asarray(f.trans['freq'].value, float)
            tn = np.asarray(f.trans['T_norm_per_freq'].data, float)
        except Exception:
            continue
        names_all.append(f.name)
        rows_all.append([len(nu), nu[0], nu[-1], nu.sum(), tn.sum(), (nu * tn).sum(), (nu * nu * tn).sum(), np.abs(np.diff(tn)).sum(),
                         float(f.freq_eff.value), float(np.asarray(f.dfreq)), float(f.wl_eff.value), f.m0, f.M0])
    G['filters_all/names'] = np.array(names_all)
    G['filters_all/moments'] = np.array(rows_all)

    out = os.path.join(HERE, 'reference_golden.npz')
    np.savez_compressed(out, **G)
    print('wrote', out, '%d arrays, %.0f kB' % (len(G), os.path.getsize(out) / 1e3))


if __name__ == '__main__':
    main()
