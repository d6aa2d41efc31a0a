"""Broadband filters and the packed filter bank consumed by the CUDA kernels.

Mirrors the parts of the reference's ``filters.py`` that the MCMC hot path touches
(``Filter`` attributes, ``read_curve`` normalisation filters.py:181-214, ``synthesize``
filters.py:288-310, ``extinction_law`` filters.py:14-33, the registry filters.py:369-445).

B200 design: the per-filter transmission curves are *packed* once on the host (FP64) into
three flat vectors -- ``alpha_k = c1 nu'_k``, ``w_k`` (everything that multiplies the Planck
denominator: c2 nu'^3, the UV cutoff, T_norm_per_freq and the trapezoid weights) and the
Fitzpatrick-99 curve ``kappa_k`` per unit E(B-V) -- so that on the device

    synthesize(planck_fast, T, R) = R^2 * sum_k w_k E_k / (exp(alpha_k / T) - 1)

is a single fused multiply-accumulate stream out of shared memory.
"""
import os
from functools import total_ordering

import numpy as np

from . import constants as K
from .filter_registry import REGISTRY

_trapz = getattr(np, 'trapezoid', None) or np.trapz
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')
_curves = None


def _curve(key):
    global _curves
    if _curves is None:
        _curves = np.load(os.path.join(_DATA, 'filter_curves.npz'))
    return _curves[key]


# ---------------------------------------------------------------------------
# Fitzpatrick (1999) extinction curve as implemented by the third-party ``extinction`` package
# (host only: it is evaluated once per filter sample to build kappa_k).
# ---------------------------------------------------------------------------
def _f99_uv(x, c1_, c2_):
    x2 = x * x
    y = x2 - 4.596 ** 2
    d = x2 / (y * y + x2 * 0.99 ** 2)
    k = c1_ + c2_ * x + 3.23 * d
    y5 = np.where(x >= 5.9, x - 5.9, 0.)
    return k + 0.41 * (0.5392 * y5 ** 2 + 0.05644 * y5 ** 3)


_f99_cache = {}


def fitzpatrick99(wave, a_v, r_v=3.1):
    """Extinction A(lambda) in magnitudes for wavelengths in angstroms (``extinction.fitzpatrick99``)."""
    from scipy.interpolate import splrep, splev
    wave = np.atleast_1d(np.asarray(wave, float))
    c2_ = -0.824 + 4.717 / r_v
    c1_ = 2.030 - 3.007 * c2_
    if r_v not in _f99_cache:
        rv2 = r_v * r_v
        with np.errstate(divide='ignore'):
            xk = 1e4 / np.array([np.inf, 26500., 12200., 6000., 5470., 4670., 4110., 2700., 2600.])
        kk = np.array([
            -r_v,
            0.26469 * r_v / 3.1 - r_v,
            0.82925 * r_v / 3.1 - r_v,
            -0.422809 + 1.00270 * r_v + 2.13572e-04 * rv2 - r_v,
            -5.13540e-02 + 1.00216 * r_v - 7.35778e-05 * rv2 - r_v,
            0.700127 + 1.00184 * r_v - 3.32598e-05 * rv2 - r_v,
            1.19456 + 1.01707 * r_v - 5.46959e-03 * rv2 + 7.97809e-04 * rv2 * r_v - 4.45636e-05 * rv2 * rv2 - r_v,
            0., 0.])
        kk[7:] = _f99_uv(xk[7:], c1_, c2_)
        _f99_cache[r_v] = splrep(xk, kk)
    x = 1e4 / wave
    uv = x >= 1e4 / 2700.
    k = np.empty_like(x)
    k[uv] = _f99_uv(x[uv], c1_, c2_)
    if not uv.all():                       # fitpack rejects an empty array (a curve entirely below 2700 A: GALEX FUV)
        k[~uv] = splev(x[~uv], _f99_cache[r_v])
    return a_v / r_v * (k + r_v)


def extinction_law(freq, ebv, rv=3.1):
    """Extinction factor 10^(A/-2.5) at frequencies in THz (reference filters.py:14-33)."""
    A = np.squeeze([fitzpatrick99(K.c_AA_THz / np.asarray(freq, float), rv * e, rv) for e in np.atleast_1d(ebv)])
    return 10. ** (A / -2.5)


class _Trans(dict):
    """Minimal stand-in for the astropy Table the reference keeps in ``Filter.trans``."""

    @property
    def colnames(self):
        return list(self.keys())


@total_ordering
class Filter:
    """A broadband photometric filter (reference filters.py:37-355, hot-path subset)."""

    order = None

    def __init__(self, names, system=None, fnu=3.631e-23, filename='', angstrom=False, offset=0):
        if isinstance(names, (list, tuple)):
            self.name = names[0]
            self.names = list(names)
        else:
            self.name = names
            self.names = [names]
        if len(self.name) == 1:
            self.char = self.name
        else:
            shortest = sorted(self.names, key=len)[0]
            self.char = shortest if len(shortest) == 1 else 'x'
        self.system = system
        self.offset = offset
        self.fnu = fnu
        if fnu is None:
            self.m0 = self.M0 = np.nan
        else:
            self.m0 = 2.5 * np.log10(fnu)
            self.M0 = self.m0 + 90.19
        self.filename = filename
        self.angstrom = angstrom
        self._trans = None
        self._packed = None

    # -- transmission curve -------------------------------------------------------------
    def read_curve(self, force=False):
        """Load and normalise the transmission curve (reference filters.py:181-230)."""
        if (self._trans is None or force) and self.filename:
            raw = _curve(self.filename)
            wl = raw[:, 0] / 10. if self.angstrom else raw[:, 0].astype(float)
            T = raw[:, 1].astype(float)
            idx = np.argsort(wl, kind='stable')
            wl, T = wl[idx], T[idx]
            T = T / np.max(T)
            freq = K._c / (wl * 1e-9) / 1e12
            dwl = _trapz(T, wl)
            self._wl_eff = _trapz(T * wl, wl) / dwl
            self._dwl = dwl
            dfreq = _trapz(T, freq)
            self._freq_eff = _trapz(T * freq, freq) / dfreq
            self._dfreq = -dfreq
            hi = T > 0.5
            left = (wl <= wl[hi].min()) & (T >= 0.1)
            right = (wl >= wl[hi].max()) & (T >= 0.1)
            wl0 = np.interp(0.5, T[left], wl[left])
            wl1 = np.interp(0.5, T[right][::-1], wl[right][::-1])
            freq0 = np.interp(0.5, T[right][::-1], freq[right][::-1])
            freq1 = np.interp(0.5, T[left], freq[left])
            self._wl_range = (self._wl_eff - wl0, wl1 - self._wl_eff)
            self._freq_range = (self._freq_eff - freq0, freq1 - self._freq_eff)
            Tpf = T / freq
            self._trans = _Trans(wl=wl, T=T, freq=freq, T_norm_per_freq=Tpf / _trapz(Tpf, freq))

    @property
    def trans(self):
        self.read_curve()
        return self._trans

    def _prop(self, attr):
        self.read_curve()
        return getattr(self, attr, None)

    wl_eff = property(lambda self: self._prop('_wl_eff'))
    dwl = property(lambda self: self._prop('_dwl'))
    wl_range = property(lambda self: self._prop('_wl_range'))
    freq_eff = property(lambda self: self._prop('_freq_eff'))
    dfreq = property(lambda self: self._prop('_dfreq'))
    freq_range = property(lambda self: self._prop('_freq_range'))

    def extinction(self, ebv, rv=3.1, z=0.):
        """A_lambda at the effective wavelength (reference filters.py:267-286)."""
        if self.wl_eff is not None:
            return fitzpatrick99(np.array([self.wl_eff * 10. / (1. + z)]), ebv * rv, rv)[0]

    # -- packed form for the device -----------------------------------------------------
    def packed(self):
        """(nu [THz], Tn*trapz_weight) of the observed-frame curve; z/cutoff are applied by ``pack_bank``."""
        if self._packed is None:
            tr = self.trans
            if tr is None:
                raise ValueError('filter %s has no transmission curve and cannot be synthesised' % self.name)
            nu = tr['freq']
            tw = np.empty_like(nu)                     # np.trapz(y, nu) == sum(y * tw)
            tw[1:-1] = (nu[2:] - nu[:-2]) / 2.
            tw[0] = (nu[1] - nu[0]) / 2.
            tw[-1] = (nu[-1] - nu[-2]) / 2.
            self._packed = (nu, tr['T_norm_per_freq'] * tw)
        return self._packed

    def synthesize(self, spectrum, *args, z=0., ebv=0., **kwargs):
        """Average L_nu of ``spectrum`` in this filter (reference filters.py:288-310).

        Only the built-in ``planck_fast`` spectrum runs on the device; an arbitrary Python callable cannot, and
        there is no CPU fallback in this package.
        """
        from .models import planck_fast, _sed_eval
        if spectrum is not planck_fast:
            raise NotImplementedError('only spectrum=planck_fast can be synthesised on the device')
        T, R = args[0], args[1]
        cutoff = args[2] if len(args) > 2 else kwargs.get('cutoff_freq', np.inf)
        T = np.asarray(T, float)
        R = np.broadcast_to(np.asarray(R, float), T.shape)
        out = _sed_eval([self], T.ravel(), R.ravel(), z, cutoff, float(ebv))[:, 0]
        return out.reshape(T.shape) if T.ndim else float(out[0])

    def __str__(self):
        return self.name

    def __repr__(self):
        return '<filter ' + self.name + '>'

    def __eq__(self, other):
        return isinstance(other, Filter) and self.name == other.name

    def __lt__(self, other):
        return isinstance(other, Filter) and Filter.order.index(self.name) < Filter.order.index(other.name)

    def __hash__(self):
        return hash(self.name)


all_filters = [Filter(list(names), system, fnu, filename, angstrom) for names, system, fnu, filename, angstrom in REGISTRY]
Filter.order = [f.name for f in all_filters]
filtdict = {}
for _f in all_filters:
    for _n in _f.names:
        filtdict[_n] = _f


_PACK_CACHE = {}


def pack_bank(filters, z=0., cutoff_freq=np.inf, ebv=0.):
    """Pack unique ``filters`` into the flat device bank.

    Returns ``offsets[int32, F+1], alpha, w, kappa`` (float64) with, for sample k of a filter,
      alpha_k = c1 nu_k (1+z);  w_k = c2 nu'_k^3 min(1, nu_c/nu'_k) Tn_k trapz_k 10^(-0.4 ebv kappa_k);
      kappa_k = A_F99(c/nu'_k; a_v = 3.1) (so that A = ebv * kappa_k, exact because F99 is linear in a_v).
    Follows reference filters.py:308-310 + models.py:1127-1128.
    """
    offs, al, ws, ks = [0], [], [], []
    for f in filters:
        key = (f, float(z), float(cutoff_freq))
        hit = _PACK_CACHE.get(key)                        # survey batches pack the same few filters 10^4 times
        if hit is None:
            nu, tnw = f.packed()
            nup = nu * (1. + z)
            kap = fitzpatrick99(K.c_AA_THz / nup, 3.1, 3.1)
            w0 = K.c2 * nup ** 3 * np.minimum(1., cutoff_freq / nup) * tnw
            hit = (K.c1 * nup, w0, kap)
            if len(_PACK_CACHE) < 4096:
                _PACK_CACHE[key] = hit
        alpha, w, kap = hit
        if np.any(np.asarray(ebv) != 0.):
            w = w * 10. ** (-0.4 * float(ebv) * kap)
        al.append(alpha)
        ws.append(w)
        ks.append(kap)
        offs.append(offs[-1] + len(alpha))
    return (np.asarray(offs, np.int32), np.concatenate(al), np.concatenate(ws), np.concatenate(ks))
