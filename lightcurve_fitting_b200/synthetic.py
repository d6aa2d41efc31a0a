"""Synthetic workloads of the named BASELINE.json shapes, shared by tests and bench.

A workload is data + model + priors + start box.  The noiseless "truth" light curve is produced by a
caller-supplied ``truth(model_name, t, filter_names, params, z)`` callable: ``bench.py`` passes the device
model, the CPU tests pass the oracle, so this module depends on neither.
"""
import numpy as np

from .filters import filtdict
from .lightcurve import LC
from . import models as M

EXAMPLE_FILTERS = ['U', 'B', 'V', 'R', 'I', 'g', 'r', 'i', '0']


class Workload:
    """One fitting problem in reference terms (what ``lightcurve_mcmc`` receives)."""

    def __init__(self, name, model_name, t, filter_names, y, dy, priors, p_lo, p_up, z=0., use_sigma=False,
                 sigma_type='relative', model_kwargs=None, truth=None):
        self.name, self.model_name = name, model_name
        self.t = np.asarray(t, float)
        self.filter_names = list(filter_names)
        self.y, self.dy = np.asarray(y, float), np.asarray(dy, float)
        self.priors_spec = priors            # list of (kind, *args): ('uniform', lo, hi), ('loguniform', lo, hi), ('gaussian', lo, hi, mean, std)
        self.p_lo, self.p_up = np.asarray(p_lo, float), np.asarray(p_up, float)
        self.z, self.use_sigma, self.sigma_type = z, use_sigma, sigma_type
        self.model_kwargs = dict(model_kwargs or {})
        self.truth = None if truth is None else np.asarray(truth, float)
        self.ndim = len(priors)

    # ---- product-side objects --------------------------------------------------------------
    def filters(self):
        return [filtdict[n] for n in self.filter_names]

    def lc(self):
        """An LC stand-in that already carries the quantity the model fits (so calcAbsMag/calcLum are no-ops)."""
        q = 'flux' if self.model_name == 'ShockCooling3' else 'lum'
        lc = _PreparedLC({'MJD': self.t, 'filter': np.array(self.filters(), dtype=object), q: self.y, 'd' + q: self.dy})
        lc.meta['redshift'] = self.z
        return lc

    def model(self, precision='fp64'):
        cls = getattr(M, self.model_name)
        if self.model_name.startswith('CompanionShocking'):
            m = cls(self.lc(), redshift=self.z)
        elif self.model_name == 'ShockCooling4':
            m = cls(redshift=self.z)
        else:
            m = cls(redshift=self.z, **self.model_kwargs)
        m.precision = precision
        return m

    def priors(self, ns=M):
        out = []
        for spec in self.priors_spec:
            kind, args = spec[0], spec[1:]
            out.append({'uniform': ns.UniformPrior, 'loguniform': ns.LogUniformPrior, 'gaussian': ns.GaussianPrior}[kind](*args))
        return out

    def device_problem(self, precision='fp64'):
        from .fitting import build_problem
        model = self.model(precision)
        return build_problem(self.lc(), model, self.priors(), use_sigma=self.use_sigma, sigma_type=self.sigma_type,
                             precision=precision)

    def start(self, nwalkers, rng):
        return self.p_lo + rng.random((nwalkers, self.ndim)) * (self.p_up - self.p_lo)

    # samples of one log-posterior evaluation (the roofline unit)
    def planck_samples_per_eval(self):
        k = np.array([len(f.trans['freq']) for f in self.filters()])
        return int(k.sum()) * (2 if self.model_name == 'ShockCooling4' else 1)


class _PreparedLC(LC):
    def calcFlux(self, *a, **k):
        if 'flux' not in self.colnames:
            super().calcFlux(*a, **k)

    def calcAbsMag(self, *a, **k):
        if 'lum' not in self.colnames:
            super().calcAbsMag(*a, **k)

    def calcLum(self, *a, **k):
        if 'lum' not in self.colnames:
            super().calcLum(*a, **k)


def _noisy(ytrue, rng, rel_err):
    dy = np.abs(ytrue) * rel_err
    dy = np.where(dy > 0, dy, np.median(dy[dy > 0]) if np.any(dy > 0) else 1.)
    return ytrue + rng.normal(size=len(ytrue)) * dy, dy


def example_sc4(npoints=None, window=(57468., 57485.), use_sigma=False, sigma_type='relative'):
    """cfg1: the bundled SN 2016bkv light curve, ShockCooling4, priors of SURVEY.md 8(d)."""
    lc = LC.example()
    lc = lc.where(MJD_min=window[0], MJD_max=window[1]) if window else lc
    lc = lc[~np.asarray(lc['nondet'].data, bool)]
    if npoints is not None:
        lc = lc[np.linspace(0, len(lc) - 1, npoints).astype(int)]
    lc.calcAbsMag()
    lc.calcLum()
    pri = [('uniform', 0., 10.), ('uniform', 0., 10.), ('uniform', 0., 100.), ('uniform', 0., 100.), ('uniform', 57468., 57468.7)]
    p_lo, p_up = [0.5, 0.1, 0.1, 1., 57468.5], [2., 2., 10., 10., 57468.7]
    if use_sigma:
        pri.append(('gaussian', 0., 10., 0., 1.))
        p_lo, p_up = p_lo + [0.], p_up + [2.]
    return Workload('cfg1-SN2016bkv-ShockCooling4', 'ShockCooling4', lc['MJD'].data, [f.name for f in lc['filter'].data],
                    lc['lum'].data, lc['dlum'].data, pri, p_lo, p_up, z=lc.meta['redshift'], use_sigma=use_sigma,
                    sigma_type=sigma_type)


def synthetic_sc3(truth, npoints=2000, seed=1, use_sigma=False):
    """cfg2: ShockCooling3 on a synthetic 8-filter light curve (U,B,V,R,I,g,r,i round-robin)."""
    rng = np.random.default_rng(seed)
    names = ['U', 'B', 'V', 'R', 'I', 'g', 'r', 'i']
    t0 = 59000.
    t = np.sort(rng.uniform(t0 + 0.3, t0 + 12., npoints))
    fn = [names[i % 8] for i in range(npoints)]
    z = 0.005
    p_true = np.array([1., 1., 1., 3., 20., 0.1, t0])
    ytrue = truth('ShockCooling3', t, fn, p_true, z)
    dmag = np.clip(rng.lognormal(np.log(0.05), 0.5, npoints), 0.01, 0.3)
    dy = ytrue * dmag * np.log(10) / 2.5
    y = ytrue + rng.normal(size=npoints) * dy
    pri = [('uniform', 0., 10.), ('uniform', 0., 10.), ('uniform', 0., 100.), ('uniform', 0., 100.), ('uniform', 5., 50.),
           ('uniform', 0., 1.), ('uniform', t0 - 2., t0 + 0.3)]
    lo = p_true * 0.9
    hi = p_true * 1.1
    lo[6], hi[6] = t0 - 0.1, t0 + 0.1
    if use_sigma:
        pri.append(('gaussian', 0., 10., 0., 1.))
        lo, hi = np.append(lo, 0.), np.append(hi, 1.)
    return Workload('cfg2-synthetic-ShockCooling3', 'ShockCooling3', t, fn, y, dy, pri, lo, hi, z=z, use_sigma=use_sigma,
                    truth=p_true)


def synthetic_sc4(truth, npoints=200, seed=4, filters=None, lc_index=0):
    """cfg5 member: ShockCooling4 on a synthetic light curve over the 9 example filters."""
    rng = np.random.default_rng([seed, lc_index])
    names = filters or EXAMPLE_FILTERS
    t0 = 57468.6
    t = np.sort(rng.uniform(t0 + 0.5, t0 + 15., npoints))
    fn = [names[i % len(names)] for i in rng.permutation(npoints)]
    z = 0.002
    p_true = np.array([rng.uniform(0.5, 2.), rng.uniform(0.1, 2.), rng.uniform(0.1, 10.), rng.uniform(1., 10.), t0])
    ytrue = truth('ShockCooling4', t, fn, p_true, z)
    y, dy = _noisy(ytrue, rng, 0.05)
    pri = [('uniform', 0., 10.), ('uniform', 0., 10.), ('uniform', 0., 100.), ('uniform', 0., 100.), ('uniform', t0 - 0.6, t0 + 0.4)]
    return Workload('cfg5-synthetic-ShockCooling4-%d' % lc_index, 'ShockCooling4', t, fn, y, dy, pri,
                    [0.5, 0.1, 0.1, 1., t0 - 0.1], [2., 2., 10., 10., t0 + 0.1], z=z, truth=p_true)


def synthetic_cs3(truth_kasen_sifto, npoints=1000, seed=3):
    """cfg4: CompanionShocking3 (Kasen + SiFTO) over U,B,V,g,r,i.

    ``truth_kasen_sifto(t, filter_names, z)`` returns a plausible positive light curve used both to scale the
    SiFTO templates and as data (any smooth positive curve works: the benchmark measures throughput).
    """
    rng = np.random.default_rng(seed)
    names = ['U', 'B', 'V', 'g', 'r', 'i']
    t_peak = 58000.
    t = np.sort(rng.uniform(t_peak - 17., t_peak + 30., npoints))
    fn = [names[i % 6] for i in range(npoints)]
    z = 0.01
    ytrue = truth_kasen_sifto(t, fn, z)
    y, dy = _noisy(ytrue, rng, 0.03)
    p_true = np.array([t_peak - 17., 0.1, 30., t_peak, 1., 0., 0.])
    pri = [('uniform', t_peak - 25., t_peak - 16.), ('uniform', 0., 1.), ('uniform', 0., 180.), ('uniform', t_peak - 5., t_peak + 5.),
           ('uniform', 0.5, 2.), ('uniform', -1., 1.), ('uniform', -1., 1.)]
    lo = [t_peak - 17.5, 0.05, 20., t_peak - 0.5, 0.9, -0.1, -0.1]
    hi = [t_peak - 17.0, 0.15, 40., t_peak + 0.5, 1.1, 0.1, 0.1]
    return Workload('cfg4-synthetic-CompanionShocking3', 'CompanionShocking3', t, fn, y, dy, pri, lo, hi, z=z, truth=p_true)


def sed_epoch(truth, rng, z=0.002, use_sigma=False):
    """cfg3 member: one SED epoch with 3-9 distinct filters, 5 % errors, default priors of bolometric.py:728-731."""
    pool = ['U', 'B', 'V', 'g', 'r', 'i', 'R', 'I', '0']
    nf = int(rng.integers(3, 10))
    fn = list(rng.choice(pool, nf, replace=False))
    T, R = rng.uniform(5., 30.), np.exp(rng.uniform(np.log(1.), np.log(30.)))
    ytrue = truth('BlackbodySED', np.zeros(nf), fn, np.array([T, R]), z)
    y, dy = _noisy(ytrue, rng, 0.05)
    pri = [('uniform', 1., 100.), ('loguniform', 0.01, 1000.)]
    lo, hi = [max(1.5, T - 2.), max(0.05, R * 0.8)], [T + 2., R * 1.2]
    if use_sigma:
        pri.append(('gaussian', 0., 10., 0., 1.))
        lo, hi = lo + [0.], hi + [1.]
    return Workload('cfg3-sed', 'BlackbodySED', np.zeros(nf), fn, y, dy, pri, lo, hi, z=z, use_sigma=use_sigma,
                    truth=np.array([T, R]))


def sed_table(truth, nepochs=500, seed=2, z=0.002, dm=32.5):
    """cfg3 as the table ``calculate_bolometric`` receives: ``nepochs`` nightly epochs, each observed in 3-9 distinct filters of
    {U,B,V,g,r,i,R,I,unfiltered}, blackbody truth T ~ U(5, 30) kK, R ~ logU(1, 30) kR_sun, 5 % photometry, magnitudes."""
    rng = np.random.default_rng(seed)
    pool = ['U', 'B', 'V', 'g', 'r', 'i', 'R', 'I', '0']
    mjd, names, mag, dmag = [], [], [], []
    for e in range(nepochs):
        nf = int(rng.integers(3, 10))
        fn = list(rng.choice(pool, nf, replace=False))
        T, R = rng.uniform(5., 30.), np.exp(rng.uniform(np.log(1.), np.log(30.)))
        lum = np.asarray(truth('BlackbodySED', np.zeros(nf), fn, np.array([T, R]), z), float)
        lum = lum * (1. + 0.05 * rng.normal(size=nf))
        zp = np.array([filtdict[n].m0 for n in fn]) + 90.19            # lightcurve.py:358
        mag += list(zp - 2.5 * np.log10(lum) + dm)
        dmag += [2.5 / np.log(10.) * 0.05] * nf
        mjd += list(58000. + e + rng.uniform(-0.2, 0.2, nf))
        names += fn
    lc = LC({'MJD': np.array(mjd), 'mag': np.array(mag), 'dmag': np.array(dmag), 'filter': np.array(names),
             'nondet': np.zeros(len(mjd), bool)})
    lc.meta.update(dm=dm, redshift=z, extinction={}, hostext={})
    return lc
