"""Astropy-free stand-in for the reference's ``LC`` table (lightcurve.py:62-359, 878-1000).

Out of the MCMC hot path: it only prepares the ``y``/``dy`` vectors the
kernels consume.  The real ``LC`` is an ``astropy.table.Table`` subclass; astropy is not available
in this image, so this class offers the same surface the hot-path drivers use -- ``lc['col']``
(with ``.data``), ``lc.meta``, ``colnames``, ``where``, ``calcFlux``, ``calcMag``, ``calcAbsMag``,
``calcLum``, ``bin`` -- on top of a dict of numpy arrays.  Any object that quacks the same way
(including a real astropy ``LC``) is accepted by ``lightcurve_mcmc`` / ``calculate_bolometric``.
"""
import os
import numpy as np

from .filters import filtdict, Filter


class Column(np.ndarray):
    """ndarray with the ``.data`` / ``.value`` accessors the reference uses on table columns."""

    def __new__(cls, arr):
        return np.asarray(arr).view(cls)

    @property
    def data(self):
        return np.asarray(self)

    value = data


def mag2flux(mag, dmag=np.nan, zp=0., nondet=None, nondetSigmas=3.):
    """lightcurve.py:912-941"""
    mag = np.asarray(mag, float)
    flux = 10 ** ((zp - mag) / 2.5)
    dflux = np.log(10) / 2.5 * flux * np.asarray(dmag, float)
    if nondet is not None and np.any(nondet):
        nondet = np.asarray(nondet, bool)
        dflux[nondet] = flux[nondet] / nondetSigmas
        flux[nondet] = 0
    return flux, dflux


def flux2mag(flux, dflux=np.array(np.nan), zp=0., nondet=None, nondetSigmas=3.):
    """lightcurve.py:878-909"""
    flux = np.array(flux, float)
    dflux = np.array(dflux, float)
    if nondet is not None and np.any(nondet):
        nondet = np.asarray(nondet, bool)
        flux[nondet] = nondetSigmas * dflux[nondet]
        dflux[nondet] = np.nan
    with np.errstate(all='ignore'):
        mag = -2.5 * np.log10(flux, out=np.full_like(flux, -np.inf), where=flux > 0.) + zp
        dmag = 2.5 * dflux / (flux * np.log(10))
    return mag, dmag


def binflux(time, flux, dflux, delta=0.2, include_zero=True):
    """lightcurve.py:944-1000"""
    time, flux, dflux = (np.asarray(a, float) for a in (time, flux, dflux))
    bt, bf, bd = [], [], []
    while len(flux) > 0:
        grp = np.abs(time - time[0]) <= delta
        tg, fg, dg = time[grp], flux[grp], dflux[grp]
        zeros = (dg == 0) | (dg == 999) | (dg == 9999) | (dg == -1) | np.isnan(dg)
        if zeros.any() and include_zero:
            x, y, z = np.mean(tg), np.mean(fg), 0.
        else:
            tg, fg, dg = tg[~zeros], fg[~zeros], dg[~zeros]
            x = np.mean(tg)
            y = np.sum(fg * dg ** -2) / np.sum(dg ** -2)
            z = np.sum(dg ** -2) ** -0.5
        bt.append(x)
        bf.append(y)
        bd.append(z)
        time, flux, dflux = time[~grp], flux[~grp], dflux[~grp]
    return np.array(bt), np.array(bf), np.array(bd)


class LC:
    """Dict-of-columns light curve with the methods of the reference ``LC`` that the fitters call."""

    def __init__(self, data=None, meta=None, **columns):
        self._cols = {}
        self.meta = dict(meta or {})
        self.nondetSigmas = 3.
        self.groupby = {'filter', 'source'}
        src = {}
        if isinstance(data, LC):
            src = {k: np.array(v) for k, v in data._cols.items()}
            self.meta = dict(data.meta) if meta is None else self.meta
            self.nondetSigmas = data.nondetSigmas
            self.groupby = set(data.groupby)
        elif isinstance(data, dict):
            src = data
        src = dict(src, **columns)
        for k, v in src.items():
            self[k] = v
        if 'filter' in self._cols and not all(isinstance(f, Filter) for f in self._cols['filter']):
            self.filters_to_objects()

    # -- table protocol -------------------------------------------------------------------
    @property
    def colnames(self):
        return list(self._cols.keys())

    def __len__(self):
        return len(next(iter(self._cols.values()))) if self._cols else 0

    def __contains__(self, key):
        return key in self._cols

    def __getitem__(self, key):
        if isinstance(key, str):
            return Column(self._cols[key])
        out = LC(meta=self.meta)
        for k, v in self._cols.items():
            out._cols[k] = np.asarray(v)[key]
        out.nondetSigmas = self.nondetSigmas
        out.groupby = set(self.groupby)
        return out

    def __setitem__(self, key, value):
        arr = np.asarray(value)
        if arr.ndim == 0 and self._cols:
            arr = np.tile(arr, len(self))
        self._cols[key] = np.array(arr)

    def copy(self):
        return LC(self)

    def get(self, key, default=None):
        if key in self._cols:
            return Column(self._cols[key])
        return np.tile(default, len(self))

    def filters_to_objects(self):
        """lightcurve.py:163-181 (without the Swift telescope special case columns when absent)"""
        names = self._cols['filter']
        filters = np.array([f if isinstance(f, Filter) else filtdict.get(str(f), filtdict['?']) for f in names], dtype=object)
        is_swift = np.zeros(len(filters), bool)
        if 'telescope' in self._cols:
            tel = self._cols['telescope']
            for nm in ('Swift', 'UVOT', 'Swift/UVOT', 'Swift+UVOT'):
                is_swift |= tel == nm
        if 'source' in self._cols:
            is_swift |= self._cols['source'] == 'SOUSA'
        if is_swift.any():
            for filt, swiftfilt in zip('UBV', 'sbv'):
                sel = is_swift & np.array([str(n) == filt for n in names])
                filters[sel] = filtdict[swiftfilt]
        self._cols['filter'] = filters

    def where(self, **kwargs):
        """lightcurve.py:87-134"""
        use = np.ones(len(self), bool)
        for col, val in kwargs.items():
            if col.startswith('filter'):
                if isinstance(val, str):
                    val = filtdict[val]
                elif isinstance(val, list):
                    val = [filtdict[v] if isinstance(v, str) else v for v in val]
            base = col.replace('_not', '').replace('_min', '').replace('_max', '')
            arr = self._cols[base]
            if isinstance(val, list):
                if '_not' in col:
                    use1 = np.ones(len(self), bool)
                    for v in val:
                        use1 &= np.array([a != v for a in arr])
                else:
                    use1 = np.zeros(len(self), bool)
                    for v in val:
                        use1 |= np.array([a == v for a in arr])
            elif '_min' in col:
                use1 = arr >= val
            elif '_max' in col:
                use1 = arr <= val
            elif '_not' in col:
                use1 = np.array([a != val for a in arr]) if val is not None else np.array([a is not None for a in arr])
            else:
                use1 = np.array([a == val for a in arr]) if val is not None else np.array([a is None for a in arr])
            use &= use1
        return self[use]

    @property
    def zp(self):
        return np.array([f.m0 for f in self._cols['filter']])

    def _nondet(self):
        return np.asarray(self._cols['nondet'], bool) if 'nondet' in self._cols else np.zeros(len(self), bool)

    def calcFlux(self, nondetSigmas=None, zp=None):
        """lightcurve.py:189-204"""
        if nondetSigmas is not None:
            self.nondetSigmas = nondetSigmas
        if zp is None:
            zp = self.zp
        self['flux'], self['dflux'] = mag2flux(self._cols['mag'], self._cols['dmag'], zp, self._nondet(), self.nondetSigmas)

    def findNondet(self, nondetSigmas=None):
        if nondetSigmas is not None:
            self.nondetSigmas = nondetSigmas
        self['nondet'] = self._cols['flux'] < self.nondetSigmas * self._cols['dflux']

    def calcMag(self, nondetSigmas=None, zp=None):
        """lightcurve.py:253-269"""
        if nondetSigmas is not None:
            self.nondetSigmas = nondetSigmas
        self.findNondet()
        if zp is None:
            zp = self.zp
        self['mag'], self['dmag'] = flux2mag(self._cols['flux'], self._cols['dflux'], zp, self._nondet(), self.nondetSigmas)

    def calcAbsMag(self, dm=None, extinction=None, hostext=None, ebv=None, rv=None, host_ebv=None, host_rv=None,
                   redshift=None):
        """lightcurve.py:271-345 (a redshift-only distance needs astropy's Planck18: give ``dm``)"""
        if redshift is not None:
            self.meta['redshift'] = redshift
        elif 'redshift' not in self.meta:
            self.meta['redshift'] = 0.
        if dm is not None:
            self.meta['dm'] = dm
        elif 'dm' not in self.meta and self.meta.get('redshift'):
            raise ValueError("meta['dm'] is required: the Planck18 redshift-distance relation needs astropy")
        elif 'dm' not in self.meta:
            self.meta['dm'] = 0.
        if ebv is None:
            ebv = self.meta.get('ebv')
        if host_ebv is None:
            host_ebv = self.meta.get('host_ebv')
        if rv is None:
            rv = self.meta.get('rv', 3.1)
        if host_rv is None:
            host_rv = self.meta.get('host_rv', 3.1)
        filts = set(self._cols['filter'])
        if extinction is not None:
            self.meta['extinction'] = extinction
        elif 'extinction' not in self.meta:
            self.meta['extinction'] = {f.name: f.extinction(ebv, rv) for f in filts
                                       if f.filename and ebv is not None}
        if hostext is not None:
            self.meta['hostext'] = hostext
        elif 'hostext' not in self.meta:
            self.meta['hostext'] = {f.name: f.extinction(host_ebv, host_rv, self.meta.get('z', 0.)) for f in filts
                                    if f.filename and host_ebv is not None}
        absmag = np.asarray(self._cols['mag'], float) - self.meta['dm']
        farr = self._cols['filter']
        for filtobj in filts:
            sel = np.array([f == filtobj for f in farr])
            for key in ('extinction', 'hostext'):
                for nm in filtobj.names:
                    if nm in self.meta[key]:
                        absmag[sel] -= self.meta[key][nm]
                        break
        self['absmag'] = absmag

    def calcLum(self, nondetSigmas=None):
        """lightcurve.py:347-359"""
        if nondetSigmas is not None:
            self.nondetSigmas = nondetSigmas
        self['lum'], self['dlum'] = mag2flux(self._cols['absmag'], self._cols['dmag'], self.zp + 90.19, self._nondet(),
                                             self.nondetSigmas)

    def bin(self, delta=0.3, groupby=None):
        """lightcurve.py:206-238"""
        if groupby is not None:
            self.groupby = groupby
        keys = [k for k in self.groupby if k in self._cols]
        groups = {}
        for i in range(len(self)):
            groups.setdefault(tuple(self._cols[k][i] for k in keys), []).append(i)
        cols = {'MJD': [], 'flux': [], 'dflux': []}
        for k in keys:
            cols[k] = []
        for gk in sorted(groups, key=lambda g: tuple(str(x) for x in g)):
            idx = np.array(groups[gk])
            mjd, flux, dflux = binflux(self._cols['MJD'][idx], self._cols['flux'][idx], self._cols['dflux'][idx], delta)
            cols['MJD'].extend(mjd)
            cols['flux'].extend(flux)
            cols['dflux'].extend(dflux)
            for k, v in zip(keys, gk):
                cols[k].extend([v] * len(mjd))
        out = LC(meta=self.meta)
        for k, v in cols.items():
            out._cols[k] = np.array(v, dtype=object if k == 'filter' else None)
        out.nondetSigmas = self.nondetSigmas
        return out

    @classmethod
    def example(cls):
        """The bundled SN 2016bkv light curve with the metadata of docs/source/usage.rst:46-49."""
        d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'SN2016bkv.npz'))
        lc = cls({k: d[k] for k in ('MJD', 'mag', 'dmag', 'filter', 'source', 'nondet')})
        lc.meta.update(dm=30.79, ebv=0.016, host_ebv=0., redshift=0.002)
        return lc
