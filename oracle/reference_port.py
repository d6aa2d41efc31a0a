"""CPU oracle: a numpy restatement of the reference's MCMC hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lightcurve_fitting_b200/`` imports
this module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker / the
CPU baseline, never as the thing shipped or measured as "ours".

Pinning status
--------------
* Model / Planck / filter-synthesis / likelihood / prior arithmetic: PINNED.
  ``tests/golden/make_golden.py`` imports the reference's *own* ``models.py`` and
  ``filters.py`` from /root/reference (with a tiny stand-in for the astropy
  units/constants/table API those files touch, see ``oracle/refshim``) and stores
  their outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
  this port against them.
* emcee stretch move (draw order) and ``extinction.fitzpatrick99``: third-party
  packages that are neither vendored in /root/reference nor installed here
  (``emcee>=3.1.1`` requirements.txt:4, ``extinction`` requirements.txt:8, both
  unpinned).  Restated from their published algorithms: **parity unpinned** for
  those two pieces.

Every function cites the reference file:line it follows (paths relative to
/root/reference/lightcurve_fitting/).
"""
import os
import numpy as np

_trapz = getattr(np, 'trapezoid', None) or np.trapz

# --------------------------------------------------------------------------
# constants (models.py:10-12, 1101-1102; filters.py:11; bolometric.py:419)
# astropy >= 4 == CODATA 2018 / IAU 2015
# --------------------------------------------------------------------------
_h = 6.62607015e-34          # J s
_kB = 1.380649e-23           # J / K
_c = 299792458.0             # m / s
_e = 1.602176634e-19         # J / eV
_sigma_sb = 2. * np.pi ** 5 * _kB ** 4 / (15. * _h ** 3 * _c ** 2)  # W m-2 K-4
_Rsun = 6.957e8              # m (IAU 2015 nominal)
_au = 1.495978707e11         # m
_pc = _au * 648000. / np.pi   # IAU 2015 B2 (astropy)
_Mpc = 1e6 * _pc

k_B = _kB / _e * 1e3                                              # eV / kK        models.py:10
c3 = (4. * np.pi * (_sigma_sb * 1e7 * _Rsun ** 2 * 1e12)) ** -0.5 / 1000.  # models.py:11
c4 = 1. / (4. * np.pi * _Mpc ** 2.)                               # models.py:12
c1 = _h / _kB * 1e12 / 1e3                                        # kK / THz       models.py:1101
c2 = 8 * np.pi ** 2 * (_h / _c ** 2) * (1000. * _Rsun) ** 2 * 1e36  # models.py:1102
c_AA_THz = _c * 1e10 / 1e12                                       # angstrom THz   filters.py:11
sigma_sb = _sigma_sb * (1000. * _Rsun) ** 2 * 1e12                # bolometric.py:419

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'lightcurve_fitting_b200', 'data')


# --------------------------------------------------------------------------
# extinction.fitzpatrick99 (third party, restated; parity unpinned)
# --------------------------------------------------------------------------
def _f99_uv(x, c1_, c2_):
    x2 = x * x
    y = x2 - 4.596 ** 2
    d = x2 / (y * y + x2 * 0.99 ** 2)
    k = c1_ + c2_ * x + 3.23 * d
    y5 = np.where(x >= 5.9, x - 5.9, 0.)
    return k + 0.41 * (0.5392 * y5 ** 2 + 0.05644 * y5 ** 3)


_F99_SPLINE = {}


def _f99_spline(r_v):
    """Knots of the optical/IR cubic spline for one R_V; built once (the compiled `extinction` package does not refit a
    spline per call either: without the cache half of a `synthesize` call was `splrep`)."""
    if r_v not in _F99_SPLINE:
        from scipy.interpolate import splrep
        c2_ = -0.824 + 4.717 / r_v
        c1_ = 2.030 - 3.007 * c2_
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
        _F99_SPLINE[r_v] = (splrep(xk, kk), c1_, c2_)
    return _F99_SPLINE[r_v]


def fitzpatrick99(wave, a_v, r_v=3.1):
    """A(lambda) in magnitudes for wavelengths in angstrom (extinction package, F99)."""
    from scipy.interpolate import splev
    wave = np.atleast_1d(np.asarray(wave, float))
    tck, c1_, c2_ = _f99_spline(r_v)
    x = 1e4 / wave
    uv = x >= 1e4 / 2700.
    k = np.empty_like(x)
    k[uv] = _f99_uv(x[uv], c1_, c2_)
    if not uv.all():                       # (fitpack rejects an empty array: a curve entirely below 2700 A, e.g. GALEX FUV)
        k[~uv] = splev(x[~uv], tck)
    return a_v / r_v * (k + r_v)


def extinction_law(freq, ebv, rv=3.1):
    """filters.py:14-33"""
    A = np.squeeze([fitzpatrick99(c_AA_THz / freq, rv * e, rv) for e in np.atleast_1d(ebv)])
    return 10. ** (A / -2.5)


# --------------------------------------------------------------------------
# filters (filters.py:37-230, 288-310, 369-445)
# --------------------------------------------------------------------------
_curves = None


def _load_curves():
    global _curves
    if _curves is None:
        _curves = dict(np.load(os.path.join(DATA_DIR, 'filter_curves.npz')))
    return _curves


class OFilter:
    """The subset of ``Filter`` the hot path touches (filters.py:117-230, 288-310)."""

    def __init__(self, name, names, filename, angstrom, fnu, system):
        self.name = name
        self.names = names
        if len(name) == 1:                       # filters.py:125-132
            self.char = name
        else:
            shortest = sorted(names, key=len)[0]
            self.char = shortest if len(shortest) == 1 else 'x'
        self.filename = filename
        self.angstrom = angstrom
        self.fnu = fnu
        self.system = system
        if fnu is None:
            self.m0 = self.M0 = np.nan
        else:
            self.m0 = 2.5 * np.log10(fnu)       # filters.py:155-156
            self.M0 = self.m0 + 90.19
        self._trans = None

    def read_curve(self):                        # filters.py:181-214
        if self._trans is None and self.filename:
            raw = _load_curves()[self.filename]
            wl = raw[:, 0] / 10. if self.angstrom else raw[:, 0].copy()   # nm
            T = raw[:, 1].copy()
            order = np.argsort(wl, kind='stable')
            wl, T = wl[order], T[order]
            T = T / np.max(T)
            freq = _c / (wl * 1e-9) / 1e12       # THz, descending
            dfreq = _trapz(T, freq)
            freq_eff = _trapz(T * freq, freq) / dfreq
            T_per_freq = T / freq
            Tn = T_per_freq / _trapz(T_per_freq, freq)
            self._trans = {'wl': wl, 'T': T, 'freq': freq, 'T_norm_per_freq': Tn}
            self.freq_eff = freq_eff
            self.dfreq = -dfreq

    @property
    def trans(self):
        self.read_curve()
        return self._trans

    def synthesize(self, spectrum, *args, z=0., ebv=0., **kwargs):   # filters.py:288-310
        freq = self.trans['freq'] * (1. + z)
        return _trapz(spectrum(freq, *args, **kwargs) * extinction_law(freq, ebv)
                      * self.trans['T_norm_per_freq'], self.trans['freq'])

    def __repr__(self):
        return '<ofilter ' + self.name + '>'

    def __eq__(self, other):
        return isinstance(other, OFilter) and self.name == other.name

    def __hash__(self):
        return hash(self.name)


def _build_registry():
    # the registry table (names -> file, unit, zero point) is pure data and is
    # shared with the product package; it restates filters.py:369-440
    from lightcurve_fitting_b200.filter_registry import REGISTRY
    fd, allf = {}, []
    for names, system, fnu, filename, angstrom in REGISTRY:
        f = OFilter(names[0], list(names), filename, angstrom, fnu, system)
        allf.append(f)
        for n in names:
            fd[n] = f
    return fd, allf


filtdict, all_filters = _build_registry()


# --------------------------------------------------------------------------
# models.py:42-48
# --------------------------------------------------------------------------
def power(base, exp):
    broadcast = np.broadcast(base, exp)
    zeros = np.zeros(broadcast.shape, float)
    positive = np.asarray(base) > 0.
    with np.errstate(all='ignore'):
        return np.power(base, exp, out=zeros, where=positive)


def planck_fast(nu, T, R, cutoff_freq=np.inf):   # models.py:1105-1128
    with np.errstate(all='ignore'):
        return c2 * np.squeeze(np.multiply.outer(R ** 2, nu ** 3 * np.minimum(1., cutoff_freq / nu))
                               * power(np.exp(c1 * np.multiply.outer(power(T, -1.), nu)) - 1., -1.))


def blackbody_to_filters(filters, T, R, z=0., cutoff_freq=np.inf, ebv=0.):   # models.py:1131-1165
    T = np.array(T)
    R = np.array(R)
    if T.shape != R.shape:
        raise Exception('T & R must have the same shape')
    np.broadcast(T, ebv)
    if T.ndim == 1 and len(T) == len(filters):   # pointwise
        y_fit = np.array([f.synthesize(planck_fast, t, r, cutoff_freq, z=z, ebv=ebv)
                          for f, t, r in zip(filters, T, R)])
    else:
        y_fit = np.array([f.synthesize(planck_fast, T, R, cutoff_freq, z=z, ebv=ebv) for f in filters])
    return y_fit


class Model:                                     # models.py:51-136
    input_names = []
    output_quantity = 'lum'

    def __init__(self, redshift=0.):
        self.z = redshift

    @property
    def nparams(self):
        return len(self.input_names)

    def __call__(self, *a, **k):
        return self.evaluate(*a, **k)

    def log_likelihood(self, t, f, y, dy, p, use_sigma=False, sigma_type='relative'):   # models.py:93-136
        if sigma_type == 'relative':
            sigma_units = dy
        elif sigma_type == 'absolute':
            sigma_units = np.median(dy)
        else:
            raise Exception('sigma_type must either be "relative" or "absolute"')
        if use_sigma:
            y_fit = self(t, f, *p[:-1])
            sigma = np.sqrt(dy ** 2. + (p[-1] * sigma_units) ** 2.)
        else:
            y_fit = self(t, f, *p)
            sigma = dy
        with np.errstate(all='ignore'):
            return -0.5 * np.sum(np.log(2 * np.pi * sigma ** 2.) + ((y - y_fit) / sigma) ** 2.)


class BaseShockCooling(Model):                   # models.py:139-269
    def __init__(self, redshift=0., n=1.5, RW=False):
        super().__init__(redshift)
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
        self.RW = bool(RW)
        if RW:
            self.a = 0.
            self.Tph_to_Tcol = 1.2

    def temperature_radius(self, t_in, v_s, M_env, f_rho_M, R, t_exp=0., kappa=1.):   # models.py:231-269
        with np.errstate(all='ignore'):
            t = np.reshape(t_in, (-1, 1)) - t_exp
            L_RW = self.L_0 * power(t ** 2 * v_s / (f_rho_M * kappa), -self.epsilon_2) * v_s ** 2 * R / kappa
            t_tr = 19.5 * (kappa * M_env / v_s) ** 0.5
            L = L_RW * self.A * np.exp(-power(self.a * t / t_tr, self.alpha))
            T_ph = self.T_0 * power(t ** 2 * v_s ** 2 / (f_rho_M * kappa), self.epsilon_1) \
                * kappa ** -0.25 * power(t, -0.5) * R ** 0.25
            T_col = T_ph * self.Tph_to_Tcol
            T_K = np.squeeze(T_col) / k_B
            R_bb = c3 * np.squeeze(L) ** 0.5 * power(T_K, -2.)
        return T_K, R_bb


class ShockCooling(BaseShockCooling):            # models.py:301-353
    input_names = ['v_s', 'M_env', 'f_rho_M', 'R', 't_0']

    def evaluate(self, t_in, f, v_s, M_env, f_rho_M, R, t_exp=0., kappa=1.):
        T_K, R_bb = self.temperature_radius(t_in, v_s, M_env, f_rho_M, R, t_exp, kappa)
        return blackbody_to_filters(f, T_K, R_bb, self.z)


class ShockCooling2(BaseShockCooling):           # models.py:356-411
    input_names = ['T_1', 'L_1', 't_tr', 't_0']

    def evaluate(self, t_in, f, T_1, L_1, t_tr, t_exp=0.):
        with np.errstate(all='ignore'):
            t = np.reshape(t_in, (-1, 1)) - t_exp
            T_K = np.squeeze(T_1 * power(t, self.epsilon_T))
            L = np.squeeze(L_1 * np.exp(-power(self.a * t / t_tr, self.alpha)) * power(t, self.epsilon_L)) * 1e42
            R_bb = c3 * L ** 0.5 * power(T_K, -2.)
        return blackbody_to_filters(f, T_K, R_bb, self.z)


class ShockCooling3(BaseShockCooling):           # models.py:433-496
    input_names = ['v_s', 'M_env', 'f_rho_M', 'R', 'd_L', 'E(B-V)', 't_0']
    output_quantity = 'flux'

    def evaluate(self, t_in, f, v_s, M_env, f_rho_M, R, dist, ebv=0., t_exp=0., kappa=1.):
        T_K, R_bb = self.temperature_radius(t_in, v_s, M_env, f_rho_M, R, t_exp, kappa)
        lum = blackbody_to_filters(f, T_K, R_bb, self.z, ebv=ebv)
        return c4 * lum / dist ** 2.


class ShockCooling4(Model):                      # models.py:507-632
    input_names = ['v_s', 'M_env', 'f_rho_M', 'R', 't_0']

    def __init__(self, redshift=0.):
        super().__init__(redshift)
        self.A, self.a, self.alpha = 0.9, 2., 0.5
        self.L_br_0, self.T_col_br_0, self.t_br_0, self.t_tr_0 = 3.69e42, 8.19, 0.036, 19.5

    def temperature_radius(self, t_in, v_s, M_env, f_rho_M, R, t_exp=0., kappa=1.):   # models.py:583-597
        with np.errstate(all='ignore'):
            t_br = self.t_br_0 * R ** 1.26 * v_s ** -1.13 * f_rho_M ** -0.13
            L_br = self.L_br_0 * R ** 0.78 * v_s ** 2.11 * f_rho_M ** 0.11 * kappa ** -0.89
            # NB right-associative ** chain, exactly as written at models.py:586
            T_col_br = self.T_col_br_0 * R ** -0.32 * v_s ** 0.58 ** f_rho_M ** 0.03 * kappa ** -0.22
            t_tr = self.t_tr_0 * np.sqrt(kappa * M_env / v_s)
            t = np.reshape(t_in, (-1, 1)) - t_exp
            ttilde = t / t_br
            L = L_br * (power(ttilde, -4. / 3.)
                        + self.A * np.exp(-power(self.a * t / t_tr, self.alpha)) * power(ttilde, -0.17))
            T_col = T_col_br * np.minimum(0.97 * power(ttilde, -1. / 3.), power(ttilde, -0.45))
            T_K = np.squeeze(T_col) / k_B
            R_bb = c3 * np.squeeze(L) ** 0.5 * power(T_K, -2.)
        return T_K, R_bb

    def evaluate(self, t_in, f, v_s, M_env, f_rho_M, R, t_exp=0., kappa=1.):   # models.py:599-632
        T_K, R_bb = self.temperature_radius(t_in, v_s, M_env, f_rho_M, R, t_exp, kappa)
        lum_blackbody = blackbody_to_filters(f, T_K, R_bb, self.z)
        lum_suppressed = blackbody_to_filters(f, 0.74 * T_K, 0.74 ** -2. * R_bb, self.z)
        return np.minimum(lum_blackbody, lum_suppressed)


def load_sifto():
    """models.py:660-661: the first three rows are dropped."""
    d = np.load(os.path.join(DATA_DIR, 'sifto.npz'))
    return [str(c) for c in d['columns']], d['table'][3:]


class BaseCompanionShocking(Model):              # models.py:665-827
    def __init__(self, filters, lum, redshift=0.):
        """`filters`, `lum`: the light curve's per-point filter objects and luminosities
        (the reference takes them from ``lc`` at models.py:696-717)."""
        from scipy.interpolate import CubicSpline
        super().__init__(redshift)
        cols, tab = load_sifto()
        epoch = tab[:, 0]
        filters = np.asarray(filters, dtype=object)
        lum = np.asarray(lum, float)
        self.sifto = {}
        have_dlt40 = any(f.name == 'DLT40' for f in filters)
        for filt in set(filters):
            if filt.name == 'unfilt.' and have_dlt40:
                sifto_filt, scale_filt = 'r', filtdict['DLT40']
            elif filt.name == 'DLT40':
                sifto_filt, scale_filt = 'r', filt
            elif filt.char in cols[1:]:
                sifto_filt, scale_filt = filt.char, filt
            else:
                raise Exception('No SiFTO template for filter ' + filt.name)
            col = tab[:, cols.index(sifto_filt)]
            sel = np.array([f == scale_filt for f in filters])
            scaled = col * np.max(lum[sel]) / np.max(col)
            self.sifto[filt] = CubicSpline(epoch, scaled, extrapolate=False)

    @staticmethod
    def temperature_radius(t_in, t_exp, a13, Mc_v9_7, kappa=1.):   # models.py:727-755
        with np.errstate(all='ignore'):
            t = np.reshape(t_in, (-1, 1)) - t_exp
            T_kasen = np.squeeze(25. * power(a13 ** 36. * Mc_v9_7 * kappa ** -35. * power(t, -74.), 1. / 144.))
            R_kasen = np.squeeze(2.7 * power(kappa * Mc_v9_7 * t ** 7., 1. / 9.))
        return T_kasen, R_kasen

    def companion_shocking(self, t_in, f, t_exp, a13, Mc_v9_7, kappa=1.):   # models.py:757-784
        T_kasen, R_kasen = self.temperature_radius(t_in, t_exp, a13, Mc_v9_7, kappa)
        return blackbody_to_filters(f, T_kasen, R_kasen, self.z)

    def stretched_sifto(self, t_in, f, t_peak, stretch, dtU=None, dti=None):   # models.py:786-827
        dt_peak = {}
        if dtU is not None:
            dt_peak[filtdict['U']] = dtU
        if dti is not None:
            dt_peak[filtdict['i']] = dti
        t_wrt_peak = np.squeeze(np.reshape(t_in, (-1, 1)) - t_peak)
        if t_wrt_peak.ndim <= 1 and len(t_wrt_peak) == len(f):
            Lnu = np.array([self.sifto[filt]((t - dt_peak.get(filt, 0.)) / stretch)
                            for t, filt in zip(t_wrt_peak, f)])
        elif t_wrt_peak.ndim <= 1:
            Lnu = np.array([self.sifto[filt]((t_wrt_peak - dt_peak.get(filt, 0.)) / stretch) for filt in f])
        else:
            Lnu = np.array([np.transpose([self.sifto[filt]((t - dt) / s) for t, dt, s in
                                          zip(t_wrt_peak.T, dt_peak.get(filt, np.zeros_like(stretch)), stretch)])
                            for filt in f])
        Lnu[np.isnan(Lnu)] = 0.
        return Lnu


class CompanionShocking(BaseCompanionShocking):  # models.py:848-918
    input_names = ['t_0', 'a', 'M v^7', 't_max', 's', 'r_r', 'r_i', 'r_U']

    def evaluate(self, t_in, f, t_exp, a13, Mc_v9_7, t_peak, stretch, rr=1., ri=1., rU=1., kappa=1.):
        Lk = self.companion_shocking(t_in, f, t_exp, a13, Mc_v9_7, kappa)
        Ls = self.stretched_sifto(t_in, f, t_peak, stretch)
        sf = {'r': rr, 'i': ri}
        kf = {'U': rU}
        return np.array([L1 * kf.get(filt.char, 1.) + L2 * sf.get(filt.char, 1.)
                         for L1, L2, filt in zip(Lk, Ls, f)])


class CompanionShocking2(BaseCompanionShocking):  # models.py:921-980
    input_names = ['t_0', 'a', 'M v^7', 't_max', 's', 'dt_U', 'dt_i']

    def evaluate(self, t_in, f, t_exp, a13, Mc_v9_7, t_peak, stretch, dtU=0., dti=0., kappa=1.):
        Lk = self.companion_shocking(t_in, f, t_exp, a13, Mc_v9_7, kappa)
        Ls = self.stretched_sifto(t_in, f, t_peak, stretch, dtU, dti)
        return Lk + Ls


class CompanionShocking3(BaseCompanionShocking):  # models.py:983-1045
    input_names = ['t_0', 'a', 'theta', 't_max', 's', 'dt_U', 'dt_i']

    def evaluate(self, t_in, f, t_exp, a13, theta, t_peak, stretch, dtU, dti, kappa=1.):
        Lk = self.companion_shocking(t_in, f, t_exp, a13, 1., kappa)
        Ls = self.stretched_sifto(t_in, f, t_peak, stretch, dtU, dti)
        theta_rad = np.deg2rad(theta)
        frac = (0.5 * np.cos(theta_rad) + 0.5) * (0.14 * theta_rad ** 2. - 0.4 * theta_rad + 1.)
        return Lk * frac + Ls


class BlackbodySED(Model):
    """The per-epoch SED 'model' of bolometric.py:154-164: one ``synthesize(planck_fast, T, R,
    cutoff_freq)`` per observed filter; time is ignored."""
    input_names = ['T', 'R']

    def __init__(self, redshift=0., cutoff_freq=np.inf, ebv=0.):
        super().__init__(redshift)
        self.cutoff_freq = cutoff_freq
        self.ebv = ebv

    def evaluate(self, t_in, f, T, R):
        return np.array([filt.synthesize(planck_fast, T, R, cutoff_freq=self.cutoff_freq, z=self.z, ebv=self.ebv)
                         for filt in f])


# --------------------------------------------------------------------------
# priors (models.py:1048-1098)
# --------------------------------------------------------------------------
class Prior:
    def __init__(self, p_min=-np.inf, p_max=np.inf):
        self.p_min, self.p_max = p_min, p_max

    def __call__(self, p):
        if self.p_min < p < self.p_max:
            return self.logp(p)
        return -np.inf


class UniformPrior(Prior):
    def logp(self, p):
        return np.zeros_like(p)


class LogUniformPrior(Prior):
    def __init__(self, p_min=0., p_max=np.inf):
        if p_min < 0.:
            raise ValueError('a log-uniform prior cannot have negative limits')
        super().__init__(p_min, p_max)

    def logp(self, p):
        return -np.log(p)


class GaussianPrior(Prior):
    def __init__(self, p_min=-np.inf, p_max=np.inf, mean=0., stddev=1.):
        super().__init__(p_min, p_max)
        self.mean, self.stddev = mean, stddev

    def logp(self, p):
        return -0.5 * ((p - self.mean) / self.stddev) ** 2.


def make_log_posterior(model, priors, t, f, y, dy, use_sigma=False, sigma_type='relative'):
    """fitting.py:121-128 (and bolometric.py:154-164 when ``model`` is a BlackbodySED)."""
    def log_posterior(p):
        log_prior = 0.
        for prior, p_i in zip(priors, p):
            log_prior += prior(p_i)
        if np.isinf(log_prior):
            return log_prior
        return log_prior + model.log_likelihood(t, f, y, dy, p, use_sigma=use_sigma, sigma_type=sigma_type)
    return log_posterior


# --------------------------------------------------------------------------
# emcee 3.1.x EnsembleSampler + StretchMove(a=2), restated (parity unpinned).
# Draw order: SURVEY.md appendix B.
# --------------------------------------------------------------------------
class StretchReplay:
    """Serial affine-invariant stretch move in emcee's draw order on a legacy
    ``np.random.RandomState``.  ``record=True`` stores every draw so that the CUDA
    path can be driven with identical (split, z, partner, log u) values."""

    def __init__(self, nwalkers, ndim, log_prob_fn, random_state=None, a=2.0, randomize_split=True):
        self.W, self.D, self.f, self.a = nwalkers, ndim, log_prob_fn, a
        self.random = random_state if random_state is not None else np.random.RandomState()
        self.randomize_split = randomize_split
        self.reset()

    def reset(self):
        self._chain, self._lnp = [], []
        self.accepted = np.zeros(self.W)
        self.iteration = 0
        self.draws = []

    def _lnprob(self, q):
        if np.any(np.isinf(q)):
            raise ValueError('At least one parameter value was infinite')
        if np.any(np.isnan(q)):
            raise ValueError('At least one parameter value was NaN')
        lp = np.array([float(self.f(p)) for p in q])
        if np.any(np.isnan(lp)):
            raise ValueError('Probability function returned NaN')
        return lp

    def run_mcmc(self, initial, nsteps, log_prob0=None, record=False):
        coords = np.array(initial, float)
        if coords.shape != (self.W, self.D):
            raise ValueError('incompatible input dimensions')
        if self.W < 2 * self.D:
            raise RuntimeError('It is unadvisable to use a red-blue move with fewer walkers than twice the '
                               'number of dimensions.')
        lnp = self._lnprob(coords) if log_prob0 is None else np.array(log_prob0, float)
        for _ in range(nsteps):
            self.random.choice(1, p=[1.0])                         # move selection: one uniform
            inds = np.arange(self.W) % 2
            if self.randomize_split:
                self.random.shuffle(inds)
            step_draws = {'inds': inds.copy(), 'halves': []}
            for split in (0, 1):
                S1 = inds == split
                s, c = coords[S1], coords[~S1]
                Ns, Nc = len(s), len(c)
                zz = ((self.a - 1.) * self.random.rand(Ns) + 1) ** 2. / self.a
                factors = (self.D - 1.) * np.log(zz)
                rint = self.random.randint(Nc, size=(Ns,))
                q = c[rint] - (c[rint] - s) * zz[:, None]
                nlp = self._lnprob(q)
                logu = np.empty(Ns)
                acc = np.zeros(Ns, bool)
                margin = np.empty(Ns)                               # lnpdiff - ln u of every decision (how close a call it was)
                for i, j in enumerate(np.flatnonzero(S1)):
                    lnpdiff = factors[i] + nlp[i] - lnp[j]
                    logu[i] = np.log(self.random.rand())
                    margin[i] = lnpdiff - logu[i]
                    if lnpdiff > logu[i]:
                        acc[i] = True
                idx = np.flatnonzero(S1)[acc]
                coords[idx] = q[acc]
                lnp[idx] = nlp[acc]
                self.accepted[idx] += 1
                step_draws['halves'].append({'z': zz, 'rint': rint, 'logu': logu, 'margin': margin, 'nlp': nlp.copy(),
                                             'walkers': np.flatnonzero(S1)})
            if record:
                self.draws.append(step_draws)
            self._chain.append(coords.copy())
            self._lnp.append(lnp.copy())
            self.iteration += 1
        return coords, lnp, self.random.get_state()

    def get_chain(self, flat=False):
        ch = np.array(self._chain).reshape(-1, self.W, self.D)
        return ch.reshape(-1, self.D) if flat else ch

    def get_log_prob(self, flat=False):
        lp = np.array(self._lnp).reshape(-1, self.W)
        return lp.reshape(-1) if flat else lp

    @property
    def chain(self):
        return np.swapaxes(self.get_chain(), 0, 1)

    @property
    def flatchain(self):
        return self.get_chain(flat=True)

    @property
    def acceptance_fraction(self):
        return self.accepted / max(self.iteration, 1)


# --------------------------------------------------------------------------
# bolometric post-processing (bolometric.py:32-59, 422-480)
# --------------------------------------------------------------------------
def pseudo(temp, radius, z, filter0=None, filter1=None, cutoff_freq=np.inf):
    filter0 = filter0 or filtdict['I']
    filter1 = filter1 or filtdict['U']
    filter0.read_curve()
    filter1.read_curve()
    freq0 = filter0.freq_eff - filter0.dfreq / 2.
    freq1 = filter1.freq_eff + filter1.dfreq / 2.
    x_optical = np.arange(freq0, freq1)
    y_optical = planck_fast(x_optical * (1. + z), temp, radius, cutoff_freq)
    return _trapz(y_optical) * 1e12


def stefan_boltzmann(temp, radius, dtemp=None, drad=None, covTR=None):   # bolometric.py:422-453
    lum = 4 * np.pi * radius ** 2 * sigma_sb * temp ** 4
    if dtemp is None or drad is None or covTR is None:
        return lum
    dlum = 8 * np.pi * sigma_sb * (radius ** 2 * temp ** 8 * drad ** 2
                                   + 4 * radius ** 4 * temp ** 6 * dtemp ** 2
                                   + 4 * radius ** 3 * temp ** 7 * covTR) ** 0.5
    return lum, dlum


def median_and_unc(x, perc_contained=68.):
    q = 50. + np.array([-perc_contained / 2., 0., perc_contained / 2.])
    percentiles = np.percentile(x, q, axis=0)
    lower, upper = np.diff(percentiles, axis=0)
    return percentiles[1], lower, upper


# ---------------------------------------------------------------------------------------------------
# Convergence diagnostics.  Not called by the reference; definitions restated from memory of
# emcee 3.1.x ``emcee/autocorr.py`` (function_1d, auto_window, integrated_time) -- parity unpinned -- and the usual
# split Gelman-Rubin statistic.  Test infrastructure like the rest of this file.
# ---------------------------------------------------------------------------------------------------
def _next_pow_two(n):
    i = 1
    while i < n:
        i = i << 1
    return i


def autocorr_function_1d(x):
    """Normalised autocorrelation function of a 1-D series via zero-padded FFT (emcee.autocorr.function_1d)."""
    x = np.atleast_1d(x)
    n = _next_pow_two(len(x))
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[:len(x)].real
    acf /= acf[0]
    return acf


def autocorr_auto_window(taus, c):
    m = np.arange(len(taus)) < c * taus
    if np.any(m):
        return int(np.argmin(m))
    return len(taus) - 1


def integrated_time(x, c=5):
    """Integrated autocorrelation time per dimension of a chain [n_t, n_w, n_d] (emcee.autocorr.integrated_time without
    the tolerance check).  Returns (tau [n_d], window [n_d])."""
    x = np.asarray(x, float)
    n_t, n_w, n_d = x.shape
    tau, win = np.empty(n_d), np.empty(n_d, int)
    for d in range(n_d):
        f = np.zeros(n_t)
        for k in range(n_w):
            f += autocorr_function_1d(x[:, k, d])
        f /= n_w
        taus = 2.0 * np.cumsum(f) - 1.0
        win[d] = autocorr_auto_window(taus, c)
        tau[d] = taus[win[d]]
    return tau, win


def split_rhat(x):
    """Split Gelman-Rubin statistic per dimension: every walker's chain is cut into two halves of floor(n_t/2) steps."""
    x = np.asarray(x, float)
    h = x.shape[0] // 2
    halves = np.concatenate([x[:h], x[h:2 * h]], axis=1)          # [h, 2 n_w, n_d]
    means = halves.mean(axis=0)
    wv = halves.var(axis=0, ddof=1).mean(axis=0)
    b_over_h = means.var(axis=0, ddof=1)
    return np.sqrt(((h - 1) / h * wv + b_over_h) / wv)


def blackbody_lstsq(freq, lum, z, p0=None, T_range=(1., 100.), R_range=(0.01, 1000.), cutoff_freq=np.inf):
    """Chi-square blackbody fit of one SED with scipy's curve_fit, as bolometric.py:483-531 does it (``freq``/``lum`` are
    the epoch's ``freq`` and ``lum`` columns).  Returns (temp, radius, dtemp, drad, lum, dlum, L_opt)."""
    import warnings
    from scipy.optimize import curve_fit, OptimizeWarning
    if p0 is None:
        p0 = [10., 10.]

    def planck_cutoff(nu, T, R):
        return planck_fast(nu, T, R, cutoff_freq)

    with warnings.catch_warnings():
        if len(freq) <= 2:
            warnings.simplefilter('ignore', OptimizeWarning)
        p0, cov = curve_fit(planck_cutoff, np.asarray(freq, float) * (1. + z), np.asarray(lum, float), p0=p0,
                            bounds=([T_range[0], R_range[0]], [T_range[1], R_range[1]]))
    temp, radius = p0
    dtemp, drad = np.sqrt(np.diag(cov))
    lum_bb, dlum = stefan_boltzmann(temp, radius, dtemp, drad, cov[0, 1])
    L_opt = pseudo(temp, radius, z, cutoff_freq=cutoff_freq)
    return temp, radius, dtemp, drad, lum_bb, dlum, L_opt
