"""Bolometric light curves from batched per-epoch blackbody MCMC fits -- drop-ins for the MCMC branch of the
reference's ``bolometric.py`` (``spectrum_mcmc`` :87-190, ``calculate_bolometric`` :648-832, ``pseudo`` :32-59,
``stefan_boltzmann`` :422-453, ``median_and_unc`` :456-480, ``group_by_epoch`` :383-416, ``blackbody_lstsq``
:483-531 (here a batched device kernel, ``blackbody_lstsq_batch``), ``integrate_sed`` :537-557, ``calc_colors`` :560-607).

B200 design: the reference loops over epochs serially, creating one emcee sampler per epoch.  Here every
epoch becomes one independent ensemble and ALL epochs run in a single kernel launch (one CTA per epoch, the
whole burn-in + sampling chain inside the kernel; ``lcf_batch_*``).  ``pseudo()`` is the same Planck-sum kernel
applied to a 1-THz frequency comb.
"""
import ctypes as C
import os
import warnings
import numpy as np

from . import constants as K
from ._capi import lib, check, dptr, MODEL_IDS
from .filters import filtdict
from .lightcurve import LC
from .models import planck_fast, UniformPrior, LogUniformPrior, GaussianPrior
from .problem import DeviceProblem
from .sampler import EnsembleSampler

sigma_sb = K.sigma_sb

DEPRECATED_BOLOMETRIC_COLNAMES = [('L_opt', 'L'), ('lum', 'L_bol'), ('dlum', 'dL_bol'), ('dtemp0', 'dtemp_mcmc0'),
                                  ('dtemp1', 'dtemp_mcmc1'), ('dradius0', 'dradius_mcmc0'), ('dradius1', 'dradius_mcmc1')]


def _comb_problem(freq0, freq1, z, cutoff_freq, precision='fp64'):
    """Device problem whose single 'filter' is the 1-THz comb of ``pseudo`` with np.trapz(dx=1) weights."""
    x = np.arange(freq0, freq1) * (1. + z)
    tw = np.ones(len(x))
    tw[0] = tw[-1] = 0.5
    w = K.c2 * x ** 3 * np.minimum(1., cutoff_freq / x) * tw * 1e12
    bank = (np.array([0, len(x)], np.int32), K.c1 * x, w, np.zeros(len(x)))
    return DeviceProblem(MODEL_IDS['BlackbodySED'], [0.], [0], [1.], [1.], ndim=2, bank=bank, precision=precision)


def pseudo(temp, radius, z, filter0=filtdict['I'], filter1=filtdict['U'], cutoff_freq=np.inf):
    """Pseudobolometric luminosity [W]: a blackbody integrated between two filters (bolometric.py:32-59)."""
    freq0 = filter0.freq_eff - filter0.dfreq / 2.
    freq1 = filter1.freq_eff + filter1.dfreq / 2.
    temp = np.asarray(temp, float)
    radius = np.broadcast_to(np.asarray(radius, float), temp.shape)
    prob = _comb_problem(freq0, freq1, z, cutoff_freq)
    out = prob.model_eval(np.stack([temp.ravel(), radius.ravel()], axis=1))[:, 0]
    return out.reshape(temp.shape) if temp.ndim else float(out[0])


def stefan_boltzmann(temp, radius, dtemp=None, drad=None, covTR=None):
    """bolometric.py:422-453"""
    lum = 4 * np.pi * radius ** 2 * sigma_sb * temp ** 4
    if dtemp is None or drad is None or covTR is None:
        return lum
    dlum = 8 * np.pi * sigma_sb * (radius ** 2 * temp ** 8 * drad ** 2 + 4 * radius ** 4 * temp ** 6 * dtemp ** 2
                                   + 4 * radius ** 3 * temp ** 7 * covTR) ** 0.5
    return lum, dlum


def median_and_unc(x, perc_contained=68.):
    """bolometric.py:456-480"""
    q = 50. + np.array([-perc_contained / 2., 0., perc_contained / 2.])
    percentiles = np.percentile(x, q, axis=0)
    median = percentiles[1]
    lower, upper = np.diff(percentiles, axis=0)
    return median, lower, upper


def group_by_epoch(lc, res=1., also_group_by=()):
    """Group a light curve into single-epoch SEDs (bolometric.py:383-416)."""
    mjd = np.asarray(lc['MJD'].data, float)
    if 'epoch' in lc.colnames:
        epochs = np.asarray(lc['epoch'].data, float).copy()
        missing = ~np.isfinite(epochs)
    else:
        epochs = np.full(len(lc), np.nan)
        missing = np.ones(len(lc), bool)
    if missing.any():
        x = mjd[missing] / res
        frac = np.median(x - np.trunc(x))
        epochs[missing] = np.round(x - frac + np.round(frac)) * res
    lc['epoch'] = epochs
    keys = [epochs] + [np.asarray(lc[c].data) for c in also_group_by]
    labels = {}
    for i in range(len(lc)):
        labels.setdefault(tuple(k[i] for k in keys), []).append(i)
    groups = [lc[np.array(idx)] for idx in labels.values()]
    mjdavg = [np.median(g['MJD'].data) for g in groups]
    return [groups[i] for i in np.argsort(mjdavg, kind='stable')]


def blackbody_lstsq_batch(epochs, z, p0=None, T_range=(1., 100.), R_range=(0.01, 1000.), cutoff_freq=np.inf):
    """Chi-square blackbody fits of MANY single-epoch SEDs in one kernel launch (one thread per epoch).

    Same model, bounds, starting point and covariance as the reference's per-epoch ``curve_fit`` call
    (bolometric.py:483-531).  Returns arrays ``temp, radius, dtemp, drad, lum, dlum, L_opt, status`` of length
    ``len(epochs)``; ``status != 0`` marks an epoch whose fit did not converge (the reference catches the
    ``RuntimeError`` curve_fit raises in that case and stores NaN).
    """
    if p0 is None:
        p0 = [10., 10.]
    n = len(epochs)
    off = np.zeros(n + 1, np.int32)
    nus, lums = [], []
    for i, e in enumerate(epochs):
        nus.append(np.asarray(e['freq'].data, float) * (1. + z))
        lums.append(np.asarray(e['lum'].data, float))
        off[i + 1] = off[i] + len(nus[-1])
    nu = np.ascontiguousarray(np.concatenate(nus)) if n else np.zeros(0)
    lum_obs = np.ascontiguousarray(np.concatenate(lums)) if n else np.zeros(0)
    popt, pcov = np.empty((n, 2)), np.empty((n, 2, 2))
    status = np.zeros(n, np.int32)
    lo = np.array([T_range[0], R_range[0]], float)
    hi = np.array([T_range[1], R_range[1]], float)
    if n:
        check(lib().lcf_blackbody_lstsq_batch(n, off.ctypes.data_as(C.POINTER(C.c_int32)), dptr(nu), dptr(lum_obs), K.c1, K.c2,
                                              float(cutoff_freq), dptr(np.asarray(p0, float)), dptr(lo), dptr(hi), dptr(popt),
                                              dptr(pcov), status.ctypes.data_as(C.POINTER(C.c_int32))))
    temp, radius = popt[:, 0].copy(), popt[:, 1].copy()
    with np.errstate(invalid='ignore'):
        dtemp, drad = np.sqrt(pcov[:, 0, 0]), np.sqrt(pcov[:, 1, 1])
        lum, dlum = stefan_boltzmann(temp, radius, dtemp, drad, pcov[:, 0, 1])
    L_opt = pseudo(temp, radius, z, cutoff_freq=cutoff_freq) if n else np.zeros(0)
    return temp, radius, dtemp, drad, lum, dlum, L_opt, status


def blackbody_lstsq(epoch1, z, p0=None, T_range=(1., 100.), R_range=(0.01, 1000.), cutoff_freq=np.inf):
    """Chi-square blackbody fit at the effective frequencies (bolometric.py:483-531): a batch of one."""
    temp, radius, dtemp, drad, lum, dlum, L_opt, status = blackbody_lstsq_batch([epoch1], z, p0, T_range, R_range, cutoff_freq)
    if status[0]:
        raise RuntimeError('Optimal parameters not found: the least-squares fit did not converge')   # what curve_fit raises
    return temp[0], radius[0], dtemp[0], drad[0], lum[0], dlum[0], L_opt[0]


def integrate_sed(epoch1):
    """Trapezoidal integration of the observed SED [W] (bolometric.py:537-557)."""
    order = np.argsort(epoch1['freq'].data, kind='stable')
    freq = np.asarray(epoch1['freq'].data, float)[order]
    dfreq = np.asarray(epoch1['dfreq'].data, float)[order]
    lum = np.asarray(epoch1['lum'].data, float)[order]
    freqs = np.concatenate([[freq[0] - dfreq[0]], freq, [freq[-1] + dfreq[-1]]])
    lums = np.concatenate([[0.], lum, [0.]])
    trapz = getattr(np, 'trapezoid', None) or np.trapz
    return trapz(lums, freqs) * 1e12      # W/Hz * THz -> W


def calc_colors(epoch1, colors):
    """bolometric.py:560-607"""
    mags, dmags, lolims, uplims = [], [], [], []
    filt = list(epoch1['filter'].data)
    for color in colors:
        f0, f1 = [filtdict[f] for f in color.split('-')]
        if f0 in filt and f1 in filt:
            i0, i1 = filt.index(f0), filt.index(f1)
            m0, dm0, n0 = epoch1['absmag'][i0], epoch1['dmag'][i0], bool(epoch1['nondet'][i0])
            m1, dm1, n1 = epoch1['absmag'][i1], epoch1['dmag'][i1], bool(epoch1['nondet'][i1])
            mags.append(np.nan if (n0 and n1) else m0 - m1)
            dmags.append((dm0 ** 2. + dm1 ** 2.) ** 0.5)
            lolims.append(n0)
            uplims.append(n1)
        else:
            mags.append(np.nan)
            dmags.append(np.nan)
            lolims.append(True)
            uplims.append(True)
    return mags, dmags, lolims, uplims


# -------------------------------------------------------------------------------------------
# SED problems and samplers
# -------------------------------------------------------------------------------------------
def _sed_problem(epoch1, priors, z, ebv, cutoff_freq, use_sigma, sigma_type, precision):
    ndim = len(priors)
    if ndim != 2 + (1 if use_sigma else 0):
        raise ValueError('planck_fast takes (T, R)%s: expected %d priors' % (' + sigma' if use_sigma else '',
                                                                              2 + (1 if use_sigma else 0)))
    n = len(epoch1)
    return DeviceProblem(MODEL_IDS['BlackbodySED'], np.zeros(n), list(epoch1['filter'].data), epoch1['lum'].data,
                         epoch1['dlum'].data, ndim=ndim, use_sigma=use_sigma, sigma_type=sigma_type, priors=priors,
                         z=z, cutoff_freq=cutoff_freq, ebv=ebv, precision=precision)


class BatchSampler:
    """Many independent ensembles (one per problem), sampled in one kernel launch (``lcf_batch_*``)."""

    def __init__(self, problems, nwalkers, seed=None):
        self.problems = list(problems)
        self.nwalkers = int(nwalkers)
        self.ndim = self.problems[0].ndim
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        arr = (C.c_void_p * len(self.problems))(*[p.handle for p in self.problems])
        h = C.c_void_p()
        check(lib().lcf_batch_create(len(self.problems), arr, self.nwalkers, C.c_uint64(int(seed)), C.byref(h)))
        self.handle = h
        self.nsteps = 0
        self.last_ms = 0.

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h is not None:
            try:
                lib().lcf_batch_destroy(h)
            except Exception:
                pass
            self.handle = None

    def run(self, starting_guesses, nburn, nsteps):
        """starting_guesses [nproblems, nwalkers, ndim]; ``None`` continues from the current positions."""
        if starting_guesses is not None:
            sg = np.ascontiguousarray(starting_guesses, float)
            if sg.shape != (len(self.problems), self.nwalkers, self.ndim):
                raise ValueError('incompatible input dimensions')
            check(lib().lcf_batch_set_state(self.handle, dptr(sg)))
        check(lib().lcf_batch_run(self.handle, int(nburn), int(nsteps)))
        self.nsteps = int(nsteps)
        ms = C.c_double(0.)
        n = C.c_int64(0)
        check(lib().lcf_batch_last_timing(self.handle, C.byref(ms), C.byref(n)))
        self.last_ms = ms.value
        return self

    def get_chain(self):
        """[nproblems, nsteps, nwalkers, ndim]"""
        out = np.empty((len(self.problems), self.nsteps, self.nwalkers, self.ndim))
        check(lib().lcf_batch_get_chain(self.handle, dptr(out)))
        return out

    def get_log_prob(self):
        out = np.empty((len(self.problems), self.nsteps, self.nwalkers))
        check(lib().lcf_batch_get_log_prob(self.handle, dptr(out)))
        return out

    @property
    def acceptance_fraction(self):
        acc = np.empty((len(self.problems), self.nwalkers), np.int64)
        check(lib().lcf_batch_get_accepted(self.handle, acc.ctypes.data_as(C.POINTER(C.c_int64))))
        return acc / max(self.nsteps, 1)

    @property
    def status(self):
        st = np.empty(len(self.problems), np.int32)
        check(lib().lcf_batch_get_status(self.handle, st.ctypes.data_as(C.POINTER(C.c_int32))))
        return st


def spectrum_mcmc(spectrum, epoch1, priors, starting_guesses, z=0., ebv=0., spectrum_kwargs=None, show=False,
                  outpath='.', nwalkers=10, burnin_steps=200, steps=100, save_chains=False, use_sigma=False,
                  sigma_type='relative', labels=None, freq_min=100., freq_max=1000., precision='fp64', seed=None):
    """Fit a spectral energy distribution to one epoch of photometry (bolometric.py:87-190).

    Only the built-in ``planck_fast`` spectrum runs on the device (an arbitrary Python callable cannot; there is
    no CPU fallback).  Returns a sampler with ``chain`` / ``flatchain`` like the reference's emcee sampler.
    """
    if spectrum is not planck_fast:
        raise NotImplementedError('only spectrum=planck_fast is available on the device')
    if sigma_type not in ('relative', 'absolute'):
        raise Exception('sigma_type must either be "relative" or "absolute"')
    spectrum_kwargs = spectrum_kwargs or {}
    cutoff = spectrum_kwargs.get('cutoff_freq', np.inf)
    mjdavg = np.median(epoch1['MJD'].data)
    prob = _sed_problem(epoch1, priors, z, ebv, cutoff, use_sigma, sigma_type, precision)
    ndim = len(priors)
    sampler = EnsembleSampler(nwalkers, ndim, prob, seed=seed)
    sampler.run_mcmc(starting_guesses, burnin_steps)
    sampler.reset()
    sampler.run_mcmc(None, steps)
    os.makedirs(outpath, exist_ok=True)
    if save_chains:
        np.save(os.path.join(outpath, f'{mjdavg:.3f}.npy'), sampler.flatchain)
    return sampler


blackbody_mcmc = spectrum_mcmc   # pre-v0.7.0 name (docs/source/release-history.rst:69)


def calculate_bolometric(lc, z=0., outpath='.', res=1., nwalkers=10, burnin_steps=200, steps=100, priors=None,
                         save_table_as=None, min_nfilt=3, cutoff_freq=np.inf, show=False, colors=None, do_mcmc=True,
                         save_chains=False, use_sigma=False, sigma_type='relative', also_group_by=(), precision='fp64',
                         seed=None, return_sampler=False):
    """Bolometric light curve from a table of broadband photometry (bolometric.py:648-832).

    Same inputs and output columns as the reference.  Differences, all on the B200 side: the per-epoch MCMC
    fits run as ONE batched launch after the (cheap, host-side) per-epoch preparation, and no corner-plot PDFs
    are produced.
    """
    if z:
        warnings.warn('The z keyword is deprecated. Include the redshift in `lc.meta["redshift"]` instead.')
    z = lc.meta.get('redshift', z)
    if colors is None:
        colors = []
    use_src = 'source' in lc.colnames
    if priors is None:
        priors = [UniformPrior(1., 100.), LogUniformPrior(0.01, 1000.)]
        if use_sigma:
            priors.append(GaussianPrior(0., 10.))

    dmag = np.asarray(lc['dmag'].data, float)
    lc = lc[np.isfinite(dmag) & (dmag > 0.)]
    rows, epochs, guesses = [], [], []
    rng = np.random.default_rng(seed)
    for epoch1 in group_by_epoch(lc, res, also_group_by):
        epoch1.calcFlux()
        epoch1 = epoch1.bin(delta=np.inf)
        epoch1.meta = dict(lc.meta)
        epoch1.calcMag()
        epoch1.calcAbsMag()
        epoch1.calcLum()
        epoch1['freq'] = np.array([f.freq_eff for f in epoch1['filter'].data])
        epoch1['dfreq'] = np.array([f.dfreq for f in epoch1['filter'].data])
        det = ~np.asarray(epoch1['nondet'].data, bool)
        filts = set(np.asarray(epoch1['filter'].data, object)[det])
        nfilt = len(filts)
        if nfilt < min_nfilt or nfilt <= 1:
            continue                        # single-filter epochs need the KDE prior (bolometric.py:753-759)
        mjdavg, dmjd0, dmjd1 = median_and_unc(epoch1['MJD'].data, 100.)
        filtstr = ''.join([f.char for f in sorted(filts)])
        L_int = integrate_sed(epoch1)
        color_mags, color_dmags, color_lolims, color_uplims = calc_colors(epoch1, colors)
        rows.append(dict(MJD=mjdavg, dMJD0=dmjd0, dMJD1=dmjd1, L_int=L_int, npoints=nfilt, filts=filtstr,
                         colors=(color_mags, color_dmags, color_lolims, color_uplims),
                         source=epoch1['source'][0] if use_src and 'source' in epoch1.colnames else None))
        epochs.append(epoch1)

    # least-squares blackbody of every epoch in ONE launch (the reference: one curve_fit per epoch, bolometric.py:768)
    T_range = (priors[0].p_min, priors[0].p_max)
    R_range = (priors[1].p_min, priors[1].p_max)
    temp, radius, dtemp, drad, L_bol, dL_bol, L, fit_status = blackbody_lstsq_batch(epochs, z, [10., 10.], T_range, R_range,
                                                                                    cutoff_freq)
    for i, row in enumerate(rows):
        p0 = np.array([10., 10.])
        if fit_status[i]:                                  # bolometric.py:769-770: RuntimeError -> NaN, default start
            row.update(temp=np.nan, radius=np.nan, dtemp=np.nan, dradius=np.nan, L_bol=np.nan, dL_bol=np.nan, L=np.nan)
        else:
            row.update(temp=temp[i], radius=radius[i], dtemp=dtemp[i], dradius=drad[i], L_bol=L_bol[i], dL_bol=dL_bol[i], L=L[i])
            p0 = np.array([temp[i], radius[i]])
        sg = rng.normal(size=(nwalkers, 2)) + p0
        sg[sg <= 0.] = 1.
        if use_sigma:
            sg = np.append(sg, np.abs(rng.normal(size=(nwalkers, 1))), axis=1)
        guesses.append(sg)

    mc_cols = ['temp_mcmc', 'radius_mcmc', 'dtemp_mcmc0', 'dtemp_mcmc1', 'dradius_mcmc0', 'dradius_mcmc1',
               'L_bol_mcmc', 'dL_bol_mcmc0', 'dL_bol_mcmc1', 'L_mcmc', 'dL_mcmc0', 'dL_mcmc1']
    batch = None
    if do_mcmc and epochs:
        problems = [_sed_problem(e, priors, z, 0., cutoff_freq, use_sigma, sigma_type, precision) for e in epochs]
        batch = BatchSampler(problems, nwalkers, seed=None if seed is None else seed + 1)
        batch.run(np.stack(guesses), burnin_steps, steps)
        chain = batch.get_chain()                                    # [E, S, W, D]
        status = batch.status
        flat = chain.reshape(len(epochs), -1, chain.shape[-1])
        # pseudo-bolometric luminosity of every posterior sample of every epoch: one launch
        L_samples = pseudo(flat[:, :, 0], flat[:, :, 1], z, cutoff_freq=cutoff_freq)
        Lbol_samples = stefan_boltzmann(flat[:, :, 0], flat[:, :, 1])
        os.makedirs(outpath, exist_ok=True)
        for i, row in enumerate(rows):
            if status[i] != 0:
                print('Probability function returned NaN')       # bolometric.py:800-803
                row.update({c: np.nan for c in mc_cols})
                continue
            (T_m, R_m), (dT0, dR0), (dT1, dR1) = median_and_unc(flat[i][:, :2])
            Lb, dLb0, dLb1 = median_and_unc(Lbol_samples[i])
            Lp, dLp0, dLp1 = median_and_unc(L_samples[i])
            row.update(temp_mcmc=T_m, radius_mcmc=R_m, dtemp_mcmc0=dT0, dtemp_mcmc1=dT1, dradius_mcmc0=dR0,
                       dradius_mcmc1=dR1, L_bol_mcmc=Lb, dL_bol_mcmc0=dLb0, dL_bol_mcmc1=dLb1, L_mcmc=Lp,
                       dL_mcmc0=dLp0, dL_mcmc1=dLp1)
            if save_chains:
                np.save(os.path.join(outpath, f'{row["MJD"]:.3f}.npy'), flat[i])
    else:
        for row in rows:
            row.update({c: np.nan for c in mc_cols})

    names = ['MJD', 'dMJD0', 'dMJD1', 'temp', 'radius', 'dtemp', 'dradius', 'L_bol', 'dL_bol', 'L'] + mc_cols + \
            ['L_int', 'npoints']
    t0 = LC()
    for nm in names:
        t0[nm] = np.array([r[nm] for r in rows], dtype=int if nm == 'npoints' else float)
    for j, c in enumerate(colors):
        t0[c] = np.array([r['colors'][0][j] for r in rows], float)
        t0['d({})'.format(c)] = np.array([r['colors'][1][j] for r in rows], float)
        t0['lolims({})'.format(c)] = np.array([r['colors'][2][j] for r in rows], bool)
        t0['uplims({})'.format(c)] = np.array([r['colors'][3][j] for r in rows], bool)
    t0['filts'] = np.array([r['filts'] for r in rows], dtype='U16')
    if use_src:
        t0['source'] = np.array([r['source'] for r in rows])
    for old, new in DEPRECATED_BOLOMETRIC_COLNAMES:
        t0[old] = t0[new]
    warnings.warn('Some column names in the output table have changed (see documentation). Please update your code!')
    if save_table_as is not None and len(t0):
        with open(save_table_as, 'w') as fh:
            cols = t0.colnames
            fh.write(' '.join(cols) + '\n')
            for i in range(len(t0)):
                fh.write(' '.join(str(t0[c][i]) for c in cols) + '\n')
    if return_sampler:
        return t0, batch
    return t0
