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
from .filters import filtdict, pack_bank
from .lightcurve import LC, mag2flux, flux2mag
from .models import planck_fast, UniformPrior, LogUniformPrior, GaussianPrior
from .problem import DeviceProblem
from .sampler import EnsembleSampler, seed_from_global_rng

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


def _lstsq_flat(offsets, nu, lum_obs, z, p0=None, T_range=(1., 100.), R_range=(0.01, 1000.), cutoff_freq=np.inf):
    """Batched least squares on flat arrays: epoch ``e`` owns ``[offsets[e], offsets[e+1])``; ``nu`` already in the rest frame."""
    if p0 is None:
        p0 = [10., 10.]
    off = np.ascontiguousarray(offsets, np.int32)
    n = len(off) - 1
    nu, lum_obs = np.ascontiguousarray(nu, float), np.ascontiguousarray(lum_obs, float)
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


def blackbody_lstsq_batch(epochs, z, p0=None, T_range=(1., 100.), R_range=(0.01, 1000.), cutoff_freq=np.inf):
    """Chi-square blackbody fits of MANY single-epoch SEDs in one kernel launch (one thread per epoch).

    Same model, bounds, starting point and covariance as the reference's per-epoch ``curve_fit`` call
    (bolometric.py:483-531).  Returns arrays ``temp, radius, dtemp, drad, lum, dlum, L_opt, status`` of length
    ``len(epochs)``; ``status != 0`` marks an epoch whose fit did not converge (the reference catches the
    ``RuntimeError`` curve_fit raises in that case and stores NaN).
    """
    off = np.zeros(len(epochs) + 1, np.int32)
    nus, lums = [], []
    for i, e in enumerate(epochs):
        nus.append(np.asarray(e['freq'].data, float) * (1. + z))
        lums.append(np.asarray(e['lum'].data, float))
        off[i + 1] = off[i] + len(nus[-1])
    nu = np.concatenate(nus) if len(epochs) else np.zeros(0)
    lum_obs = np.concatenate(lums) if len(epochs) else np.zeros(0)
    return _lstsq_flat(off, nu, lum_obs, z, p0, T_range, R_range, cutoff_freq)


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

    def __init__(self, problems, nwalkers, seed=None, _handle=None, _nproblems=None, _ndim=None):
        self.nwalkers = int(nwalkers)
        if _handle is not None:                      # created in C from flat arrays (lcf_sed_batch_create)
            self.problems, self.nproblems, self.ndim, self.handle = None, int(_nproblems), int(_ndim), _handle
        else:
            self.problems = list(problems)
            self.nproblems = len(self.problems)
            self.ndim = self.problems[0].ndim
            if seed is None:
                seed = seed_from_global_rng()
            arr = (C.c_void_p * len(self.problems))(*[p.handle for p in self.problems])
            h = C.c_void_p()
            check(lib().lcf_batch_create(len(self.problems), arr, self.nwalkers, C.c_uint64(int(seed)), C.byref(h)))
            self.handle = h
        self.nsteps = 0
        self.last_ms = 0.

    @classmethod
    def from_sed_table(cls, offsets, point_filter, lum, dlum, bank, priors, nwalkers, use_sigma=False, sigma_type='relative',
                       precision='fp64', seed=None):
        """Every SED epoch of a table as ONE batch, built in C from flat arrays (``lcf_sed_batch_create``): epoch ``e`` owns the
        points ``[offsets[e], offsets[e+1])``; ``point_filter`` indexes ``bank`` = ``filters.pack_bank(unique filters, z, cutoff)``."""
        from .problem import _prior_arrays
        from . import _capi
        ndim = len(priors)
        if ndim != 2 + (1 if use_sigma else 0):
            raise ValueError('planck_fast takes (T, R)%s: expected %d priors' % (' + sigma' if use_sigma else '', 2 + (1 if use_sigma else 0)))
        if sigma_type not in ('relative', 'absolute'):
            raise Exception('sigma_type must either be "relative" or "absolute"')
        kind, pmin, pmax, mean, std = _prior_arrays(priors, ndim)
        boff, alpha, w, _ = bank
        offsets, point_filter = np.ascontiguousarray(offsets, np.int32), np.ascontiguousarray(point_filter, np.int32)
        lum, dlum = np.ascontiguousarray(lum, float), np.ascontiguousarray(dlum, float)
        boff, alpha, w = np.ascontiguousarray(boff, np.int32), np.ascontiguousarray(alpha, float), np.ascontiguousarray(w, float)
        if seed is None:
            seed = seed_from_global_rng()
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        h = C.c_void_p()
        check(lib().lcf_sed_batch_create(len(offsets) - 1, ip(offsets), ip(point_filter), dptr(lum), dptr(dlum), len(boff) - 1, ip(boff),
                                         dptr(alpha), dptr(w), ndim, 1 if use_sigma else 0, 0 if sigma_type == 'relative' else 1,
                                         ip(kind), dptr(pmin), dptr(pmax), dptr(mean), dptr(std), _capi.PRECISIONS[precision],
                                         int(nwalkers), C.c_uint64(int(seed)), C.byref(h)))
        return cls(None, nwalkers, _handle=h, _nproblems=len(offsets) - 1, _ndim=ndim)

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
            if sg.shape != (self.nproblems, self.nwalkers, self.ndim):
                raise ValueError('incompatible input dimensions')
            check(lib().lcf_batch_set_state(self.handle, dptr(sg)))
        check(lib().lcf_batch_run(self.handle, int(nburn), int(nsteps)))
        self.nsteps = int(nsteps)
        ms = C.c_double(0.)
        n = C.c_int64(0)
        check(lib().lcf_batch_last_timing(self.handle, C.byref(ms), C.byref(n)))
        self.last_ms = ms.value
        return self

    def get_chain(self, out=None):
        """[nproblems, nsteps, nwalkers, ndim]; ``out``: a caller-provided C-contiguous float64 buffer (e.g. page-locked memory,
        which makes the device-to-host copy of a survey batch's chains several times faster)."""
        shape = (self.nproblems, self.nsteps, self.nwalkers, self.ndim)
        if out is None:
            out = np.empty(shape)
        else:
            out = out.reshape(-1)[:int(np.prod(shape))].reshape(shape)
            if out.dtype != np.float64 or not out.flags.c_contiguous:
                raise ValueError('out must be C-contiguous float64')
        check(lib().lcf_batch_get_chain(self.handle, dptr(out)))
        return out

    def get_log_prob(self):
        out = np.empty((self.nproblems, self.nsteps, self.nwalkers))
        check(lib().lcf_batch_get_log_prob(self.handle, dptr(out)))
        return out

    def summary(self, z, cutoff_freq=np.inf, perc_contained=68., filter0=filtdict['I'], filter1=filtdict['U']):
        """Posterior summaries of every epoch, computed on the HBM-resident chain (bolometric.py:792-798): dict of
        ``temp, radius, L_bol, L`` -> array ``[nproblems, 3]`` = (median, median - lower, upper - median) = ``median_and_unc`` of
        T, R, ``stefan_boltzmann(T, R)`` and ``pseudo(T, R, z, cutoff_freq)`` over the flat chain."""
        freq0 = filter0.freq_eff - filter0.dfreq / 2.
        freq1 = filter1.freq_eff + filter1.dfreq / 2.
        comb = _comb_problem(freq0, freq1, z, cutoff_freq)
        out = np.empty((self.nproblems, 4, 3))
        check(lib().lcf_batch_summary(self.handle, comb.handle, float(sigma_sb), float(perc_contained), dptr(out)))
        return {'temp': out[:, 0], 'radius': out[:, 1], 'L_bol': out[:, 2], 'L': out[:, 3]}

    @property
    def acceptance_fraction(self):
        acc = np.empty((self.nproblems, self.nwalkers), np.int64)
        check(lib().lcf_batch_get_accepted(self.handle, acc.ctypes.data_as(C.POINTER(C.c_int64))))
        return acc / max(self.nsteps, 1)

    @property
    def status(self):
        st = np.empty(self.nproblems, np.int32)
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


# -------------------------------------------------------------------------------------------
# calculate_bolometric: whole-table (vectorised) preparation + batched device fits
# -------------------------------------------------------------------------------------------
def _segment_ids(*keys):
    """Dense ids of the distinct rows of the key columns (any dtypes), numbered in order of first appearance."""
    gid = np.zeros(len(keys[0]), np.int64)
    for k in keys:
        _, inv = np.unique(np.asarray(k), return_inverse=True)
        gid = gid * (int(inv.max()) + 1 if len(inv) else 1) + inv
    _, first, inv = np.unique(gid, return_index=True, return_inverse=True)
    rank = np.empty(len(first), np.int64)
    rank[np.argsort(first, kind='stable')] = np.arange(len(first))
    return rank[inv], len(first)


def _segment_median(seg, nseg, x):
    """Median, minimum and maximum of ``x`` within each segment id (np.median semantics)."""
    order = np.lexsort((x, seg))
    xs, counts = x[order], np.bincount(seg, minlength=nseg)
    start = np.concatenate([[0], np.cumsum(counts)[:-1]])
    lo, hi = start + (counts - 1) // 2, start + counts // 2
    return 0.5 * (xs[lo] + xs[hi]), xs[start], xs[start + counts - 1]


class EpochTable:
    """Every single-epoch SED of a light curve as flat, epoch-major arrays: what the reference builds one epoch at a time with
    ``group_by_epoch`` + ``calcFlux`` + ``bin(delta=inf)`` + ``calcMag`` + ``calcAbsMag`` + ``calcLum`` (bolometric.py:735-746,
    lightcurve.py:189-359), done with whole-table numpy operations so that the host cost per epoch is microseconds."""

    def __init__(self, lc, res=1., also_group_by=()):
        n = len(lc)
        mjd = np.asarray(lc['MJD'].data, float)
        filt = np.asarray(lc['filter'].data, object)
        names = np.array([f.name for f in filt]) if n else np.zeros(0, 'U1')
        uniq, fidx = np.unique(names, return_inverse=True)
        self.filters = [filtdict[str(u)] for u in uniq]
        m0 = np.array([f.m0 for f in self.filters])
        nondet_in = np.asarray(lc['nondet'].data, bool) if 'nondet' in lc.colnames else np.zeros(n, bool)
        sig = float(getattr(lc, 'nondetSigmas', 3.))
        # group_by_epoch (bolometric.py:383-416): bins of `res` days, phased by the median fractional part
        if 'epoch' in lc.colnames:
            epochs = np.asarray(lc['epoch'].data, float).copy()
            missing = ~np.isfinite(epochs)
        else:
            epochs, missing = np.full(n, np.nan), np.ones(n, bool)
        if missing.any():
            x = mjd[missing] / res
            frac = np.median(x - np.trunc(x))
            epochs[missing] = np.round(x - frac + np.round(frac)) * res
        lc['epoch'] = epochs
        eid, ne = _segment_ids(epochs, *[np.asarray(lc[c].data) for c in also_group_by])
        med, _, _ = _segment_median(eid, ne, mjd)
        erank = np.empty(ne, np.int64)
        erank[np.argsort(med, kind='stable')] = np.arange(ne)              # epochs sorted by median MJD
        eid = erank[eid]
        # calcFlux (lightcurve.py:189-204) on every row
        flux, dflux = mag2flux(np.asarray(lc['mag'].data, float), np.asarray(lc['dmag'].data, float), m0[fidx], nondet_in, sig)
        # bin(delta=inf) (lightcurve.py:206-238, 944-1000): one bin per (epoch, filter[, source])
        has_src = 'source' in lc.colnames
        src = np.asarray(lc['source'].data) if has_src else None
        srank = np.unique(src.astype(str), return_inverse=True)[1] if has_src else np.zeros(n, np.int64)
        frank = np.argsort(np.argsort(np.array([str(f) for f in self.filters]), kind='stable'), kind='stable')[fidx]
        bid, nb = _segment_ids(eid, frank, srank)
        # order of the binned rows: epoch, then str(filter), then str(source)
        first = np.zeros(nb, np.int64)
        first[bid[::-1]] = np.arange(n)[::-1]                               # a representative row of every bin
        border = np.lexsort((srank[first], frank[first], eid[first]))
        brank = np.empty(nb, np.int64)
        brank[border] = np.arange(nb)
        bid = brank[bid]
        first = first[border]
        zeros = (dflux == 0) | (dflux == 999) | (dflux == 9999) | (dflux == -1) | np.isnan(dflux)
        anyzero = np.bincount(bid, weights=zeros, minlength=nb) > 0
        cnt = np.bincount(bid, minlength=nb)
        with np.errstate(all='ignore'):
            wgt = np.where(zeros, 0., dflux ** -2.)
            sw = np.bincount(bid, weights=wgt, minlength=nb)
            cnt_w = np.bincount(bid, weights=~zeros, minlength=nb)
            t_bin = np.where(anyzero, np.bincount(bid, weights=mjd, minlength=nb) / cnt,
                             np.bincount(bid, weights=np.where(zeros, 0., mjd), minlength=nb) / cnt_w)
            f_bin = np.where(anyzero, np.bincount(bid, weights=flux, minlength=nb) / cnt,
                             np.bincount(bid, weights=np.where(zeros, 0., flux * wgt), minlength=nb) / sw)
            d_bin = np.where(anyzero, 0., sw ** -0.5)
        self.n_epochs = ne
        self.epoch = eid[first]
        self.fidx = fidx[first]
        self.filter = np.array(self.filters, dtype=object)[self.fidx] if nb else np.zeros(0, object)
        self.source = src[first] if has_src else None
        self.MJD, self.flux, self.dflux = t_bin, f_bin, d_bin
        # calcMag (lightcurve.py:240-269), calcAbsMag (:271-345), calcLum (:347-359)
        zp = m0[self.fidx]
        self.nondet = self.flux < sig * self.dflux
        self.mag, self.dmag = flux2mag(self.flux, self.dflux, zp, self.nondet, sig)
        probe = LC({'mag': np.zeros(len(self.filters)), 'filter': np.array(self.filters, dtype=object)}, meta=lc.meta)
        probe.calcAbsMag()                                                  # per-filter offset: -dm - extinction - host extinction
        lc.meta.update(probe.meta)
        self.absmag = self.mag + probe['absmag'].data[self.fidx]
        self.lum, self.dlum = mag2flux(self.absmag, self.dmag, zp + 90.19, self.nondet, sig)
        self.freq = np.array([f.freq_eff for f in self.filters], float)[self.fidx]
        self.dfreq = np.array([f.dfreq for f in self.filters], float)[self.fidx]
        self.offsets = np.concatenate([[0], np.cumsum(np.bincount(self.epoch, minlength=ne))]).astype(np.int64)
        # distinct DETECTED filters per epoch (bolometric.py:748-749)
        det = ~self.nondet
        pair = np.unique(self.epoch[det] * len(self.filters) + self.fidx[det])
        self.nfilt = np.bincount(pair // max(len(self.filters), 1), minlength=ne)
        order_rank = np.array([Filter_order(f) for f in self.filters])
        chars = np.array([f.char for f in self.filters])
        pe, pf = pair // max(len(self.filters), 1), pair % max(len(self.filters), 1)
        o = np.lexsort((order_rank[pf], pe))
        cuts = np.cumsum(np.bincount(pe, minlength=ne))[:-1]
        self.filtstr = [''.join(c) for c in np.split(chars[pf[o]], cuts)]
        self.mjd_med, self.mjd_min, self.mjd_max = _segment_median(self.epoch, ne, self.MJD)

    def take(self, keep):
        """Row mask and new offsets of the epochs selected by the boolean array ``keep``."""
        rows = keep[self.epoch]
        counts = np.diff(self.offsets)[keep]
        return rows, np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)

    def epoch_lc(self, e):
        """Epoch ``e`` as an LC table (what the reference's loop body sees)."""
        sl = slice(self.offsets[e], self.offsets[e + 1])
        cols = {k: getattr(self, k)[sl] for k in ('MJD', 'flux', 'dflux', 'filter', 'nondet', 'mag', 'dmag', 'absmag', 'lum', 'dlum', 'freq', 'dfreq')}
        if self.source is not None:
            cols['source'] = self.source[sl]
        return LC(cols)

    def integrate_sed(self):
        """``integrate_sed`` (bolometric.py:537-557) of every epoch: trapezoid over the SED sorted by frequency, zero at
        ``freq -+ dfreq`` beyond the ends."""
        o = np.lexsort((self.freq, self.epoch))
        e, f, df, l = self.epoch[o], self.freq[o], self.dfreq[o], self.lum[o]
        same = e[1:] == e[:-1]
        inner = np.where(same, 0.5 * (l[1:] + l[:-1]) * (f[1:] - f[:-1]), 0.)
        out = np.bincount(e[:-1], weights=inner, minlength=self.n_epochs) if len(e) > 1 else np.zeros(self.n_epochs)
        firsts, lasts = self.offsets[:-1], self.offsets[1:] - 1
        ok = lasts >= firsts
        out[ok] += 0.5 * l[firsts[ok]] * df[firsts[ok]] + 0.5 * l[lasts[ok]] * df[lasts[ok]]
        return out * 1e12

    def colors(self, colors):
        """``calc_colors`` (bolometric.py:560-607) of every epoch: dict color -> (mag, dmag, lolim, uplim) arrays."""
        out = {}
        idx = {f: k for k, f in enumerate(self.filters)}
        ne = self.n_epochs

        def first_row(f):
            rows = np.full(ne, -1, np.int64)
            if f in idx:
                r = np.flatnonzero(self.fidx == idx[f])
                ep, at = np.unique(self.epoch[r], return_index=True)
                rows[ep] = r[at]
            return rows

        for color in colors:
            f0, f1 = [filtdict[f] for f in color.split('-')]
            r0, r1 = first_row(f0), first_row(f1)
            both = (r0 >= 0) & (r1 >= 0)
            a, b = np.where(both, r0, 0), np.where(both, r1, 0)
            n0, n1 = self.nondet[a], self.nondet[b]
            with np.errstate(invalid='ignore'):
                mag = np.where(both, np.where(n0 & n1, np.nan, self.absmag[a] - self.absmag[b]), np.nan)
                dmag = np.where(both, (self.dmag[a] ** 2. + self.dmag[b] ** 2.) ** 0.5, np.nan)
            out[color] = (mag, dmag, np.where(both, n0, True), np.where(both, n1, True))
        return out


def Filter_order(f):
    from .filters import Filter
    return Filter.order.index(f.name)


def _write_fixed_width_two_line(table, path):
    """astropy's ``ascii.fixed_width_two_line`` layout (bolometric.py:829-830): names, a row of dashes, then the rows, every
    column padded to its widest entry; NaN entries (masked in the reference's table) are written as ``--``."""
    cols = table.colnames
    cells = []
    for c in cols:
        v = table[c].data
        if v.dtype.kind == 'f':
            cells.append(['--' if np.isnan(x) else repr(float(x)) for x in v])
        else:
            cells.append([str(x) if str(x) else '--' for x in v])
    widths = [max([len(c)] + [len(x) for x in col]) for c, col in zip(cols, cells)]
    with open(path, 'w') as fh:
        fh.write(' '.join(c.rjust(w) for c, w in zip(cols, widths)).rstrip() + '\n')
        fh.write(' '.join('-' * w for w in widths) + '\n')
        for i in range(len(table)):
            fh.write(' '.join(col[i].rjust(w) for col, w in zip(cells, widths)) + '\n')


def calculate_bolometric(lc, z=0., outpath='.', res=1., nwalkers=10, burnin_steps=200, steps=100, priors=None,
                         save_table_as=None, min_nfilt=3, cutoff_freq=np.inf, show=False, colors=None, do_mcmc=True,
                         save_chains=False, use_sigma=False, sigma_type='relative', also_group_by=(), precision='fp64',
                         seed=None, return_sampler=False, return_timing=False):
    """Bolometric light curve from a table of broadband photometry (bolometric.py:648-832).

    Same inputs and output columns as the reference.  B200 design: the per-epoch preparation is done once for the whole
    table (:class:`EpochTable`), the least-squares blackbodies of all epochs are one launch, the MCMC fits of all epochs are
    one launch (one CTA per epoch), and the posterior summaries -- Stefan-Boltzmann and pseudo-bolometric luminosity of every
    sample, medians and 68 % intervals -- are taken on the HBM-resident chains; only ``[epochs, 4, 3]`` numbers come back.
    No corner-plot PDFs are produced (presentation).  Epochs with a single detected filter (reachable with ``min_nfilt <= 1``)
    need the sequential KDE prior of bolometric.py:753-759 and are refused.
    """
    import time
    tm = {}
    t_start = time.perf_counter()
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
    tab = EpochTable(lc, res, also_group_by)
    if np.any((tab.nfilt >= min_nfilt) & (tab.nfilt <= 1)):
        raise NotImplementedError('epochs with a single detected filter are fitted by the reference with a gaussian_kde prior taken '
                                  'from the previous epoch\'s chain (bolometric.py:753-759): sequential and not a built-in prior, so '
                                  'it cannot run batched on the device; use min_nfilt >= 2')
    keep = tab.nfilt >= max(min_nfilt, 2)
    rows, offsets = tab.take(keep)
    ne = int(keep.sum())
    L_int = tab.integrate_sed()[keep]
    cols = tab.colors(colors)
    tm['prepare_s'] = time.perf_counter() - t_start

    # least-squares blackbody of every epoch in ONE launch (the reference: one curve_fit per epoch, bolometric.py:768)
    t1 = time.perf_counter()
    T_range = (priors[0].p_min, priors[0].p_max)
    R_range = (priors[1].p_min, priors[1].p_max)
    temp, radius, dtemp, drad, L_bol, dL_bol, L, fit_status = _lstsq_flat(offsets, tab.freq[rows] * (1. + z), tab.lum[rows], z, [10., 10.],
                                                                          T_range, R_range, cutoff_freq)
    bad = fit_status != 0                                      # bolometric.py:769-770: RuntimeError -> NaN, default start
    for a in (temp, radius, dtemp, drad, L_bol, dL_bol, L):
        a[bad] = np.nan
    p0 = np.where(bad[:, None], 10., np.stack([temp, radius], axis=1)) if ne else np.zeros((0, 2))
    tm['lstsq_s'] = time.perf_counter() - t1

    rng = np.random.default_rng(seed)
    guesses = rng.normal(size=(ne, nwalkers, 2)) + p0[:, None, :]
    guesses[guesses <= 0.] = 1.
    if use_sigma:
        guesses = np.append(guesses, np.abs(rng.normal(size=(ne, nwalkers, 1))), axis=2)

    mc_cols = ['temp_mcmc', 'radius_mcmc', 'dtemp_mcmc0', 'dtemp_mcmc1', 'dradius_mcmc0', 'dradius_mcmc1',
               'L_bol_mcmc', 'dL_bol_mcmc0', 'dL_bol_mcmc1', 'L_mcmc', 'dL_mcmc0', 'dL_mcmc1']
    mc = {c: np.full(ne, np.nan) for c in mc_cols}
    batch = None
    tm.update(problem_build_s=0., sampling_ms=0., summary_s=0.)
    if do_mcmc and ne:
        t2 = time.perf_counter()
        bank = pack_bank(tab.filters, z=z, cutoff_freq=cutoff_freq)
        batch = BatchSampler.from_sed_table(offsets, tab.fidx[rows], tab.lum[rows], tab.dlum[rows], bank, priors, nwalkers,
                                            use_sigma=use_sigma, sigma_type=sigma_type, precision=precision,
                                            seed=None if seed is None else seed + 1)
        tm['problem_build_s'] = time.perf_counter() - t2
        t3 = time.perf_counter()
        batch.run(guesses, burnin_steps, steps)
        tm['sampling_ms'] = batch.last_ms
        tm['sampling_call_s'] = time.perf_counter() - t3
        t4 = time.perf_counter()
        summ = batch.summary(z, cutoff_freq=cutoff_freq)
        status = batch.status
        tm['summary_s'] = time.perf_counter() - t4
        good = status == 0
        for _ in range(int((~good).sum())):
            print('Probability function returned NaN')           # bolometric.py:800-803
        for name, key in (('temp', 'temp'), ('radius', 'radius'), ('L_bol', 'L_bol'), ('L', 'L')):
            v = np.where(good[:, None], summ[key], np.nan)
            pre = {'temp': 'temp_mcmc', 'radius': 'radius_mcmc', 'L_bol': 'L_bol_mcmc', 'L': 'L_mcmc'}[name]
            dpre = {'temp': 'dtemp_mcmc', 'radius': 'dradius_mcmc', 'L_bol': 'dL_bol_mcmc', 'L': 'dL_mcmc'}[name]
            mc[pre], mc[dpre + '0'], mc[dpre + '1'] = v[:, 0], v[:, 1], v[:, 2]
        if save_chains:
            os.makedirs(outpath, exist_ok=True)
            chain = batch.get_chain()
            for i in np.flatnonzero(good):
                np.save(os.path.join(outpath, f'{tab.mjd_med[keep][i]:.3f}.npy'), chain[i].reshape(-1, chain.shape[-1]))

    t5 = time.perf_counter()
    t0 = LC()
    mjd_med = tab.mjd_med[keep]
    base = {'MJD': mjd_med, 'dMJD0': mjd_med - tab.mjd_min[keep], 'dMJD1': tab.mjd_max[keep] - mjd_med, 'temp': temp, 'radius': radius,
            'dtemp': dtemp, 'dradius': drad, 'L_bol': L_bol, 'dL_bol': dL_bol, 'L': L}
    for nm, v in base.items():
        t0[nm] = np.asarray(v, float)
    for nm in mc_cols:
        t0[nm] = mc[nm]
    t0['L_int'] = L_int
    t0['npoints'] = tab.nfilt[keep].astype(int)
    for c in colors:
        m, dm_, lo, up = cols[c]
        t0[c], t0['d({})'.format(c)] = m[keep], dm_[keep]
        t0['lolims({})'.format(c)], t0['uplims({})'.format(c)] = lo[keep].astype(bool), up[keep].astype(bool)
    t0['filts'] = np.array([fs for fs, k in zip(tab.filtstr, keep) if k], dtype='U16')
    if use_src:
        t0['source'] = tab.source[tab.offsets[:-1][keep]] if ne else np.zeros(0, tab.source.dtype)
    for old, new in DEPRECATED_BOLOMETRIC_COLNAMES:
        t0[old] = t0[new]
    warnings.warn('Some column names in the output table have changed (see documentation). Please update your code!')
    if save_table_as is not None and len(t0):
        _write_fixed_width_two_line(t0, save_table_as)
    tm['table_s'] = time.perf_counter() - t5
    tm['total_s'] = time.perf_counter() - t_start
    tm['device_s'] = tm['lstsq_s'] + tm.get('sampling_call_s', 0.) + tm['summary_s']
    out = (t0,)
    if return_sampler:
        out += (batch,)
    if return_timing:
        out += (tm,)
    return out[0] if len(out) == 1 else out
