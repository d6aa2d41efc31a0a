"""Analytic light-curve models, priors and the Planck/filter front end -- device-backed drop-ins.

Same names, call signatures and error behaviour as the reference's ``models.py``; every evaluation
runs in the fused CUDA kernels of ``csrc/lcf_device.cuh`` through the C ABI (``include/lcf.h``).
There is no numpy implementation of the model arithmetic in this package.

Reference map: ``Model`` models.py:51-136, ``BaseShockCooling`` :139-298, ``ShockCooling`` :301,
``ShockCooling2`` :356, ``ShockCooling3`` :433, ``ShockCooling4`` :507, ``BaseCompanionShocking``
:665, ``CompanionShocking{,2,3}`` :848/:921/:983, priors :1048-1098, ``planck_fast`` :1105,
``blackbody_to_filters`` :1131.
"""
import os
import hashlib
import numpy as np

from . import constants as K
from ._capi import MODEL_IDS
from .filters import filtdict, Filter
from .problem import DeviceProblem

k_B, c1, c2, c3, c4 = K.k_B, K.c1, K.c2, K.c3, K.c4


def power(base, exp):
    """Power function that returns zero for any nonpositive base (models.py:42-48); host utility."""
    broadcast = np.broadcast(base, exp)
    zeros = np.zeros(broadcast.shape, float)
    positive = np.asarray(base) > 0.
    with np.errstate(all='ignore'):
        return np.power(base, exp, out=zeros, where=positive)


# -------------------------------------------------------------------------------------------
# Planck + filter integration front end
# -------------------------------------------------------------------------------------------
def _sed_eval(filters, T, R, z, cutoff_freq, ebv, precision='fp64'):
    """R^2 sum_k w_k/(exp(alpha_k/T)-1) for every (T, R) pair and every filter -> [npairs, nfilters]."""
    T = np.asarray(T, float).ravel()
    R = np.asarray(R, float).ravel()
    nf = len(filters)
    prob = DeviceProblem(MODEL_IDS['BlackbodySED'], np.zeros(nf), list(filters), np.ones(nf), np.ones(nf), ndim=2,
                         z=z, cutoff_freq=cutoff_freq, ebv=ebv, precision=precision)
    return prob.model_eval(np.stack([T, R], axis=1))


def planck_fast(nu, T, R, cutoff_freq=np.inf):
    """The Planck spectrum L_nu [W/Hz] at frequencies ``nu`` [THz] (models.py:1105-1128).

    Evaluated on the device as a degenerate filter bank: one "filter" whose samples are the requested
    frequencies with unit weights.  Output shape follows the reference: ``squeeze(outer(T, nu))``.
    """
    nu_a = np.atleast_1d(np.asarray(nu, float))
    T_a = np.asarray(T, float)
    R_a = np.broadcast_to(np.asarray(R, float), T_a.shape)
    n = nu_a.size
    # each frequency is its own 2-sample "filter": (nu, w) and a zero-weight pad (the C ABI wants >= 2 samples)
    alpha = np.repeat(K.c1 * nu_a.ravel(), 2)
    w = np.zeros(2 * n)
    w[0::2] = K.c2 * nu_a.ravel() ** 3 * np.minimum(1., cutoff_freq / nu_a.ravel())
    bank = (np.arange(0, 2 * n + 1, 2, dtype=np.int32), alpha, w, np.zeros(2 * n))
    prob = DeviceProblem(MODEL_IDS['BlackbodySED'], np.zeros(n), np.arange(n), np.ones(n), np.ones(n), ndim=2, bank=bank)
    out = prob.model_eval(np.stack([T_a.ravel(), R_a.ravel()], axis=1))      # [nTR, n]
    return np.squeeze(out.reshape(T_a.shape + nu_a.shape))


def blackbody_to_filters(filters, T, R, z=0., cutoff_freq=np.inf, ebv=0.):
    """Average L_nu of blackbodies through filters (models.py:1131-1165), same dispatch:

    *pointwise* (one (T, R) per filter) iff ``T.ndim == 1 and len(T) == len(filters)``, else *grid*
    (every (T, R) through every filter, output ``[nfilters, *T.shape]``).
    """
    T = np.array(T, float)
    R = np.array(R, float)
    if T.shape != R.shape:
        raise Exception('T & R must have the same shape')
    np.broadcast(T, ebv)  # raises ValueError if not broadcastable, like the reference
    filters = list(np.atleast_1d(filters))
    if np.ndim(ebv) > 0 and np.size(ebv) > 1:
        # per-(T,R) reddening: run each distinct E(B-V) as its own folded bank
        ebv_b = np.broadcast_to(np.asarray(ebv, float), T.shape).ravel()
        out = np.empty((len(filters), T.size))
        for e in np.unique(ebv_b):
            sel = ebv_b == e
            out[:, sel] = _sed_eval(filters, T.ravel()[sel], R.ravel()[sel], z, cutoff_freq, float(e)).T
        if T.ndim == 1 and len(T) == len(filters):
            return np.diagonal(out).copy()
        return out.reshape((len(filters),) + T.shape)
    ebv = float(np.asarray(ebv).ravel()[0]) if np.ndim(ebv) else float(ebv)
    if T.ndim == 1 and len(T) == len(filters):  # pointwise
        uniq = list(dict.fromkeys(filters))
        res = _sed_eval(uniq, T, R, z, cutoff_freq, ebv)                   # [N, nuniq]
        col = np.array([uniq.index(f) for f in filters])
        return res[np.arange(len(filters)), col]
    res = _sed_eval(filters, T, R, z, cutoff_freq, ebv)                    # [T.size, F]
    return res.T.reshape((len(filters),) + T.shape)


# -------------------------------------------------------------------------------------------
# models
# -------------------------------------------------------------------------------------------
class Model:
    """An analytical model, defined by a function and its parameters (models.py:51-136)."""

    input_names = []
    units = []
    output_quantity = 'lum'
    _model_name = None          # key into MODEL_IDS
    precision = 'fp64'          # 'fp64' (reference arithmetic, rtol 1e-9) or 'fp32' (throughput mode, rtol 1e-4)

    def __init__(self, lc=None, redshift=0.):
        if redshift:
            self.z = redshift
        elif lc is not None and 'redshift' in lc.meta:
            self.z = lc.meta['redshift']
        else:
            self.z = 0.
        # instance copies: use_sigma appends '\\sigma' (fitting.py:74-76) without touching the class
        self.input_names = list(type(self).input_names)
        self.units = list(type(self).units)
        self._nmodel = len(type(self).input_names)
        self._ll_cache = {}

    @property
    def nparams(self):
        return len(self.input_names)

    @property
    def axis_labels(self):
        return ['${}$ ({})'.format(var, unit) if unit else '${}$'.format(var)
                for var, unit in zip(self.input_names, self.units)]

    def __repr__(self):
        return f'<{self.__class__.__name__}: z={self.z:.3f}>'

    # -- device plumbing ------------------------------------------------------------------
    def _model_consts(self):
        return ()

    def _extra_problem_kwargs(self):
        return {}

    def _device_problem(self, t, f, y, dy, ndim, **kw):
        return DeviceProblem(MODEL_IDS[self._model_name], t, f, y, dy, ndim=ndim, z=self.z,
                             model_consts=self._model_consts(), precision=kw.pop('precision', self.precision),
                             **self._extra_problem_kwargs(), **kw)

    def __call__(self, *args, **kwargs):
        return self.evaluate(*args, **kwargs)

    def evaluate(self, t_in, f, *params, **kwargs):
        """Evaluate the model at times ``t_in`` and filters ``f`` for one or many parameter sets.

        Same broadcasting contract as the reference (models.py:260, 1161-1164): with scalar parameters and
        ``len(t_in) == len(f)`` the evaluation is pointwise; otherwise every filter is evaluated at every time
        (and every parameter set): output ``[len(f), len(t_in), nsets]`` squeezed.
        """
        if kwargs.get('kappa', 1.) != 1.:
            raise NotImplementedError('kappa != 1 is not part of the device hot path')
        if len(params) < self._nmodel:
            params = tuple(params) + tuple(self._defaults[len(params):])
        params = params[:self._nmodel]
        t_in = np.atleast_1d(np.asarray(t_in, float))
        f = list(np.atleast_1d(f))
        pb = np.broadcast_arrays(*[np.asarray(p, float) for p in params])
        scalar = pb[0].ndim == 0
        P = np.stack([p.ravel() for p in pb], axis=1)          # [nsets, nmodel]
        if scalar and len(t_in) == len(f) and len(t_in) > 1:
            prob = self._device_problem(t_in, f, np.ones(len(f)), np.ones(len(f)), self._nmodel)
            return prob.model_eval(P)[0]
        nt, nf = len(t_in), len(f)
        tt = np.tile(t_in, nf)
        ff = [flt for flt in f for _ in range(nt)]
        prob = self._device_problem(tt, ff, np.ones(nt * nf), np.ones(nt * nf), self._nmodel)
        out = prob.model_eval(P)                               # [nsets, nf*nt]
        out = out.reshape(P.shape[0], nf, nt).transpose(1, 2, 0)
        return np.squeeze(out)

    def log_likelihood(self, lc, p, use_sigma=False, sigma_type='relative'):
        """The log-likelihood of the model given the data in ``lc`` and parameters ``p`` (models.py:93-136)."""
        if sigma_type not in ('relative', 'absolute'):
            raise Exception('sigma_type must either be "relative" or "absolute"')
        # The device problem is cached on the CONTENT of the columns it was packed from (the reference re-reads lc on
        # every call, so an edited column, a recomputed lum or another table of the same length must all be seen).
        t, f = lc['MJD'].data, lc['filter'].data
        y, dy = lc[self.output_quantity].data, lc['d' + self.output_quantity].data
        h = hashlib.blake2b(digest_size=16)
        for a in (t, y, dy):
            h.update(np.ascontiguousarray(a, float).tobytes())
        h.update('\0'.join(getattr(flt, 'name', str(flt)) for flt in f).encode())
        key = (h.digest(), bool(use_sigma), sigma_type, self.precision, self.z)
        prob = self._ll_cache.get(key)
        if prob is None:
            prob = self._device_problem(t, f, y, dy, self._nmodel + (1 if use_sigma else 0),
                                        use_sigma=use_sigma, sigma_type=sigma_type)
            self._ll_cache = {key: prob}
        p = np.asarray(p, float)
        if p.ndim == 1:
            return float(prob.log_likelihood(p[None, :])[0])
        return prob.log_likelihood(p.reshape(p.shape[0], -1).T).reshape(p.shape[1:])


class BaseShockCooling(Model):
    """Sapir & Waxman (2017) shock cooling (models.py:139-298)."""

    def __init__(self, lc=None, redshift=0., n=1.5, RW=False):
        super().__init__(lc, redshift=redshift)
        if n == 1.5:
            self.n, self.A, self.a, self.alpha = 1.5, 0.94, 1.67, 0.8
            self.epsilon_1, self.epsilon_2, self.L_0, self.T_0, self.Tph_to_Tcol = 0.027, 0.086, 2.0e42, 1.61, 1.1
        elif n == 3.:
            self.n, self.A, self.a, self.alpha = 3., 0.79, 4.57, 0.73
            self.epsilon_1, self.epsilon_2, self.L_0, self.T_0, self.Tph_to_Tcol = 0.016, 0.175, 2.1e42, 1.69, 1.0
        else:
            raise ValueError('n can only be 1.5 or 3')
        self.epsilon_T = 2 * self.epsilon_1 - 0.5
        self.epsilon_L = -2 * self.epsilon_2
        if RW:
            self.RW = True
            self.a = 0.
            self.Tph_to_Tcol = 1.2
        else:
            self.RW = False

    def __repr__(self):
        return f'<{self.__class__.__name__}: z={self.z:.3f}, n={self.n:.1f}, RW={self.RW}>'

    def _model_consts(self):
        return (self.A, self.a, self.alpha, self.epsilon_1, self.epsilon_2, self.L_0, self.T_0, self.Tph_to_Tcol)

    @staticmethod
    def t_min(p, kappa=1.):
        """models.py:276-287 (Eq. 17)"""
        v_s, f_rho_M, R = p[0], p[2], p[3]
        t_exp = p[4] if len(p) > 4 else 0.
        return 0.2 * R / v_s * np.maximum(0.5, R ** 0.4 * (f_rho_M * kappa) ** -0.2 * v_s ** -0.7) + t_exp

    @staticmethod
    def t_max(p, kappa=1.):
        """models.py:290-298 (Eq. 24)"""
        R = p[3]
        t_exp = p[4] if len(p) > 4 else 0.
        return 7.4 * (R / kappa) ** 0.55 + t_exp


class ShockCooling(BaseShockCooling):
    input_names = ['v_\\mathrm{s*}', 'M_\\mathrm{env}', 'f_\\rho M', 'R', 't_0']
    units = ['$10^{8.5}$ cm s$^{-1}$', 'M$_\\odot$', 'M$_\\odot$', '$10^{13}$ cm', 'd']
    _model_name = 'ShockCooling'
    _defaults = (None, None, None, None, 0.)


class ShockCooling2(BaseShockCooling):
    input_names = ['T_1', 'L_1', 't_\\mathrm{tr}', 't_0']
    units = ['kK', '$10^{42}$ erg s$^{-1}$', 'd', 'd']
    _model_name = 'ShockCooling2'
    _defaults = (None, None, None, 0.)

    @staticmethod
    def t_min(p, kappa=1.):
        return NotImplemented

    def t_max(self, p, kappa=1.):
        """models.py:422-430"""
        T_1 = p[0]
        t_exp = p[3] if len(p) > 3 else 0.
        return (8.12 / T_1) ** (self.epsilon_T ** -1) + t_exp


class ShockCooling3(BaseShockCooling):
    input_names = ['v_\\mathrm{s*}', 'M_\\mathrm{env}', 'f_\\rho M', 'R', 'd_L', 'E(B-V)', 't_0']
    units = ['$10^{8.5}$ cm s$^{-1}$', 'M$_\\odot$', 'M$_\\odot$', '$10^{13}$ cm', 'Mpc', 'mag', 'd']
    output_quantity = 'flux'
    _model_name = 'ShockCooling3'
    _defaults = (None, None, None, None, None, 0., 0.)

    @staticmethod
    def t_min(p, kappa=1.):
        return BaseShockCooling.t_min([p[0], p[1], p[2], p[3], p[6] if len(p) > 6 else 0.], kappa=kappa)

    @staticmethod
    def t_max(p, kappa=1.):
        return BaseShockCooling.t_max([p[0], p[1], p[2], p[3], p[6] if len(p) > 6 else 0.], kappa=kappa)


class ShockCooling4(Model):
    """Morag, Sapir & Waxman (2023) shock cooling (models.py:507-657), including the operator-precedence
    behaviour of models.py:586 that the reference actually computes."""
    input_names = ['v_\\mathrm{s*}', 'M_\\mathrm{env}', 'f_\\rho M', 'R', 't_0']
    units = ['$10^{8.5}$ cm s$^{-1}$', 'M$_\\odot$', 'M$_\\odot$', '$10^{13}$ cm', 'd']
    _model_name = 'ShockCooling4'
    _defaults = (None, None, None, None, 0.)

    def __init__(self, lc=None, redshift=0.):
        super().__init__(lc, redshift=redshift)
        self.A = 0.9
        self.a = 2.
        self.alpha = 0.5
        self.L_br_0 = 3.69e42
        self.T_col_br_0 = 8.19
        self.t_min_0 = 0.012
        self.t_br_0 = 0.036
        self.t_07eV_0 = 6.86
        self.t_tr_0 = 19.5

    def _model_consts(self):
        return (self.A, self.a, self.alpha, self.L_br_0, self.T_col_br_0, self.t_br_0, self.t_tr_0)

    def t_min(self, p, kappa=1.):
        R = p[3]
        t_exp = p[4] if len(p) > 4 else 0.
        return self.t_min_0 * R + t_exp

    def t_max(self, p, kappa=1.):
        """models.py:644-657, as written there (``t_tr_0 ** sqrt(...)``)."""
        v_s, M_env, f_rho_M, R, t_exp, *_ = p
        t_07eV = self.t_07eV_0 * R ** 0.56 * v_s ** 0.16 * kappa ** -0.61 * f_rho_M ** -0.06
        t_tr = self.t_tr_0 ** np.sqrt(kappa * M_env / v_s)
        return np.minimum(t_07eV, t_tr / self.a) + t_exp


def _load_sifto():
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'sifto.npz'))
    cols = [str(c) for c in d['columns']]
    tab = d['table'][3:]                      # the first three points are ~0 (models.py:661)
    return cols, tab


class BaseCompanionShocking(Model):
    """Kasen (2010) companion shocking + the SiFTO SN Ia template (models.py:665-845)."""

    def __init__(self, lc, redshift=0.):
        from scipy.interpolate import CubicSpline
        super().__init__(lc, redshift=redshift)
        if 'lum' not in lc.colnames:
            if 'absmag' not in lc.colnames:
                lc.calcAbsMag()
            lc.calcLum()
        cols, tab = _load_sifto()
        self._sifto_epoch = tab[:, 0]
        filt_col = lc['filter'].data
        lum = np.asarray(lc['lum'].data, float)
        self.sifto = {}
        have_dlt40 = any(f == filtdict['DLT40'] for f in filt_col)
        for filt in set(filt_col):
            if filt.name == 'unfilt.' and have_dlt40:
                sifto_filt, scale_filt = 'r', filtdict['DLT40']
            elif filt.name == 'DLT40':
                sifto_filt, scale_filt = 'r', filt
            elif filt.char in cols[1:]:
                sifto_filt, scale_filt = filt.char, filt
            else:
                raise Exception('No SiFTO template for filter ' + filt.name)
            col = tab[:, cols.index(sifto_filt)]
            sel = np.array([f == scale_filt for f in filt_col])
            scaled = col * np.max(lum[sel]) / np.max(col)
            self.sifto[filt] = CubicSpline(tab[:, 0], scaled, extrapolate=False)

    def _extra_problem_kwargs(self):
        return {'sifto': self.sifto}

    def t_min(self, p):
        return p[3] + p[4] * self._sifto_epoch.min()

    def t_max(self, p):
        return p[3] + p[4] * self._sifto_epoch.max()


class CompanionShocking(BaseCompanionShocking):
    input_names = ['t_0', 'a', 'M v^7', 't_\\mathrm{max}', 's', 'r_r', 'r_i', 'r_U']
    units = ['d', '$10^{13}$ cm', 'M$_\\mathrm{Ch}$ ($10^9$ cm s$^{-1}$)$^7$', 'd', '', '', '', '']
    _model_name = 'CompanionShocking'
    _defaults = (None, None, None, None, None, 1., 1., 1.)


class CompanionShocking2(BaseCompanionShocking):
    input_names = ['t_0', 'a', 'M v^7', 't_\\mathrm{max}', 's', '\\Delta t_U', '\\Delta t_i']
    units = ['d', '$10^{13}$ cm', 'M$_\\mathrm{Ch}$ ($10^9$ cm s$^{-1}$)$^7$', 'd', '', 'd', 'd']
    _model_name = 'CompanionShocking2'
    _defaults = (None, None, None, None, None, 0., 0.)


class CompanionShocking3(BaseCompanionShocking):
    input_names = ['t_0', 'a', '\\theta', 't_\\mathrm{max}', 's', '\\Delta t_U', '\\Delta t_i']
    units = ['d', '$10^{13}$ cm', 'deg', 'd', '', 'd', 'd']
    _model_name = 'CompanionShocking3'
    _defaults = (None, None, None, None, None, None, None)


class BlackbodySED(Model):
    """The per-epoch SED 'model' behind ``spectrum_mcmc(planck_fast, ...)`` (bolometric.py:154-164): parameters
    (T [kK], R [1000 Rsun]); one filter synthesis per observed point; time is ignored."""
    input_names = ['T', 'R']
    units = ['kK', '1000 R$_\\odot$']
    _model_name = 'BlackbodySED'
    _defaults = (None, None)

    def __init__(self, lc=None, redshift=0., cutoff_freq=np.inf, ebv=0.):
        super().__init__(lc, redshift=redshift)
        self.cutoff_freq = cutoff_freq
        self.ebv = ebv

    def _extra_problem_kwargs(self):
        return {'cutoff_freq': self.cutoff_freq, 'ebv': self.ebv}


# -------------------------------------------------------------------------------------------
# priors (models.py:1048-1098): host objects; the device evaluates the same three families
# -------------------------------------------------------------------------------------------
class Prior:
    def __init__(self, p_min=-np.inf, p_max=np.inf):
        self.p_min = p_min
        self.p_max = p_max

    def __call__(self, p):
        if self.p_min < p < self.p_max:
            return self.logp(p)
        else:
            return -np.inf

    def logp(self, p):
        raise NotImplementedError


class UniformPrior(Prior):
    """dP/dp ∝ 1"""
    def logp(self, p):
        return np.zeros_like(p)


class LogUniformPrior(Prior):
    """dP/dp ∝ 1/p"""
    def __init__(self, p_min=0., p_max=np.inf):
        if p_min < 0.:
            raise ValueError('a log-uniform prior cannot have negative limits')
        super().__init__(p_min, p_max)

    def logp(self, p):
        return -np.log(p)


class GaussianPrior(Prior):
    """dP/dp ∝ exp(-(p - mean)^2 / (2 stddev^2))"""
    def __init__(self, p_min=-np.inf, p_max=np.inf, mean=0., stddev=1.):
        super().__init__(p_min, p_max)
        self.mean = mean
        self.stddev = stddev

    def logp(self, p):
        return -0.5 * ((p - self.mean) / self.stddev) ** 2.
