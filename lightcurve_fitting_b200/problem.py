"""Host-side packing of one fitting problem (light curve or SED epoch + model + priors) for the device.

This is the glue between the reference-shaped Python objects (``Model``, ``Prior``, ``Filter``, ``LC``
columns) and the plain-C problem description of ``include/lcf.h``.  Nothing is computed here except
the one-time FP64 folding of constants into the packed filter bank (``filters.pack_bank``).
"""
import ctypes as C
import numpy as np

from . import _capi
from ._capi import ProblemDesc, dptr, iptr, f64, i32, check, lib
from .filters import pack_bank, filtdict

_PRIOR_KIND = {'UniformPrior': 0, 'LogUniformPrior': 1, 'GaussianPrior': 2}


def _prior_arrays(priors, ndim):
    """Translate Prior objects (models.py:1048-1098) into the flat arrays of the C ABI."""
    if priors is None:
        kind = np.zeros(ndim, np.int32)
        return kind, np.full(ndim, -np.inf), np.full(ndim, np.inf), np.zeros(ndim), np.ones(ndim)
    if len(priors) != ndim:
        raise Exception('priors must have length {:d}'.format(ndim))
    kind, pmin, pmax, mean, std = [], [], [], [], []
    for pr in priors:
        name = type(pr).__name__
        if name not in _PRIOR_KIND:
            # arbitrary callables (e.g. a gaussian_kde.logpdf) cannot run on the device; no CPU fallback
            raise NotImplementedError('prior %r cannot be evaluated on the device; use UniformPrior, '
                                      'LogUniformPrior or GaussianPrior' % (pr,))
        kind.append(_PRIOR_KIND[name])
        pmin.append(pr.p_min)
        pmax.append(pr.p_max)
        mean.append(getattr(pr, 'mean', 0.))
        std.append(getattr(pr, 'stddev', 1.))
    return i32(kind), f64(pmin), f64(pmax), f64(mean), f64(std)


class DeviceProblem:
    """Owns one ``lcf_problem`` handle.

    Parameters
    ----------
    model_id : int
        ``_capi.MODEL_IDS`` value.
    t, filters, y, dy : array-like
        Photometry points (any order; they are grouped by filter internally and results are returned in the
        caller's order).  ``filters`` holds ``Filter`` objects, or integer indices when ``bank`` is given.
    bank : tuple, optional
        Pre-packed ``(offsets, alpha, w, kappa)``; default: packed from the unique ``filters``.
    """

    def __init__(self, model_id, t, filters, y, dy, *, ndim, use_sigma=False, sigma_type='relative', priors=None,
                 z=0., cutoff_freq=np.inf, ebv=0., model_consts=(), precision='fp64', bank=None, sifto=None,
                 dt_filters=None):
        if sigma_type not in ('relative', 'absolute'):
            raise Exception('sigma_type must either be "relative" or "absolute"')   # models.py:126
        if precision not in _capi.PRECISIONS:
            raise ValueError("precision must be 'fp64' or 'fp32'")
        t = f64(np.atleast_1d(t))
        y = f64(np.atleast_1d(y))
        dy = f64(np.atleast_1d(dy))
        n = len(t)
        if bank is None:
            uniq = []
            index = {}
            fidx = np.empty(n, np.int32)
            for i, f in enumerate(filters):
                if f not in index:
                    index[f] = len(uniq)
                    uniq.append(f)
                fidx[i] = index[f]
            bank = pack_bank(uniq, z=z, cutoff_freq=cutoff_freq, ebv=ebv)
            self.filters = uniq
        else:
            fidx = i32(filters)
            self.filters = None
        offsets, alpha, w, kappa = bank
        nfilt = len(offsets) - 1
        self.order = np.argsort(fidx, kind='stable')          # device order -> caller order
        self.inverse = np.empty(n, np.int64)
        self.inverse[self.order] = np.arange(n)
        self.npoints, self.ndim, self.precision = n, ndim, precision
        self.nsamples_per_point = (offsets[1:] - offsets[:-1])[fidx]

        keep = []                                              # arrays that must outlive lcf_problem_create

        def hold(a):
            keep.append(a)
            return a

        d = ProblemDesc()
        d.model_id = model_id
        d.precision = _capi.PRECISIONS[precision]
        d.ndim = ndim
        d.use_sigma = 1 if use_sigma else 0
        d.sigma_type = 0 if sigma_type == 'relative' else 1
        d.npoints = n
        d.nfilters = nfilt
        d.sigma_unit_abs = float(np.median(dy)) if n else 0.    # models.py:124
        for i, v in enumerate(model_consts):
            d.model_consts[i] = float(v)
        d.bank_offsets = iptr(hold(i32(offsets)))
        d.bank_alpha = dptr(hold(f64(alpha)))
        d.bank_w = dptr(hold(f64(w)))
        d.bank_kappa = dptr(hold(f64(kappa)))
        role = np.zeros(nfilt, np.int32)
        if self.filters is not None:
            for k, f in enumerate(self.filters):
                r = 0
                if f.char == 'U':
                    r |= _capi.ROLE_KASEN_RU
                if f.char == 'r':
                    r |= _capi.ROLE_SIFTO_RR
                if f.char == 'i':
                    r |= _capi.ROLE_SIFTO_RI
                if f == filtdict['U']:
                    r |= _capi.ROLE_DT_U
                if f == filtdict['i']:
                    r |= _capi.ROLE_DT_I
                role[k] = r
        d.filter_role = iptr(hold(role))
        if sifto is not None:
            # sifto: dict Filter -> scipy CubicSpline on uniform knots (models.py:717)
            x = None
            coef = []
            for f in self.filters:
                if f not in sifto:
                    raise KeyError(f)
                spl = sifto[f]
                x = spl.x
                coef.append(np.ascontiguousarray(spl.c.T))         # [interval][4], highest power first
            d.sifto_nknots = len(x)
            d.sifto_x0 = float(x[0])
            d.sifto_dx = float(x[1] - x[0])
            if not np.allclose(np.diff(x), x[1] - x[0]):
                raise ValueError('SiFTO knots must be uniform')
            d.sifto_coef = dptr(hold(f64(np.stack(coef))))
        d.t = dptr(hold(f64(t[self.order])))
        d.point_filter = iptr(hold(i32(fidx[self.order])))
        d.y = dptr(hold(f64(y[self.order])))
        d.dy = dptr(hold(f64(dy[self.order])))
        kind, pmin, pmax, mean, std = _prior_arrays(priors, ndim)
        d.prior_kind = iptr(hold(kind))
        d.prior_min = dptr(hold(pmin))
        d.prior_max = dptr(hold(pmax))
        d.prior_mean = dptr(hold(mean))
        d.prior_std = dptr(hold(std))
        self.nmodel = ndim - (1 if use_sigma else 0)
        h = C.c_void_p()
        check(lib().lcf_problem_create(C.byref(d), C.byref(h)))
        self.handle = h

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h is not None and _capi._lib is not None:
            _capi._lib.lcf_problem_destroy(h)
            self.handle = None

    def last_launch(self):
        """Shape and kernel instantiation of the last launch for this problem (``lcf_problem_last_launch``)."""
        a, b, c, v = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        g = C.c_int64()
        check(lib().lcf_problem_last_launch(self.handle, C.byref(a), C.byref(b), C.byref(c), C.byref(g), C.byref(v)))
        ng, nq, ppl = C.c_int64(), C.c_int(), C.c_int()
        check(lib().lcf_problem_last_launch_ex(self.handle, C.byref(ng), C.byref(nq), C.byref(ppl)))
        return {'walkers_per_cta': a.value, 'warps_per_cta': b.value, 'cluster': c.value, 'grid': g.value,
                'groups': ng.value, 'sum_units': nq.value, 'points_per_lane': ppl.value, 'flat': v.value != 3 and g.value != ng.value * c.value,
                'kernel': {0: 'k_pass<generic>', 1: 'k_pass<32 walkers>', 2: 'k_pass<32 walkers, plain>', 3: 'k_ring', 4: 'k_ring<look-ahead>', 5: 'k_pass_seg'}[v.value]}

    # -- evaluation entry points ----------------------------------------------------------
    def model_eval(self, params):
        """params [nsets, n_model_params] -> model values [nsets, npoints] in the caller's point order."""
        p = f64(np.atleast_2d(params))
        out = np.empty((p.shape[0], self.npoints))
        check(lib().lcf_model_eval(self.handle, p.shape[0], dptr(p), dptr(out)))
        return out[:, self.inverse]

    def log_likelihood(self, params):
        p = f64(np.atleast_2d(params))
        if p.shape[1] != self.ndim:
            raise ValueError('expected %d parameters' % self.ndim)
        out = np.empty(p.shape[0])
        check(lib().lcf_log_likelihood(self.handle, p.shape[0], dptr(p), dptr(out)))
        return out

    def log_posterior(self, params, raise_nan=False):
        p = f64(np.atleast_2d(params))
        if p.shape[1] != self.ndim:
            raise ValueError('expected %d parameters' % self.ndim)
        out = np.empty(p.shape[0])
        nan = C.c_int64(0)
        check(lib().lcf_log_posterior(self.handle, p.shape[0], dptr(p), dptr(out), C.byref(nan)))
        if raise_nan and nan.value:
            raise ValueError('Probability function returned NaN')
        return out
