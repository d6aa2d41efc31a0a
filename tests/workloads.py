"""Test-side view of the synthetic workloads: the product objects come from
``lightcurve_fitting_b200.synthetic``; the matching CPU-oracle objects are built here."""
import numpy as np

from lightcurve_fitting_b200 import synthetic
from oracle import reference_port as rp


def oracle_filters(names):
    return np.array([rp.filtdict[n] for n in names], dtype=object)


def oracle_model(model_name, z, filter_names=None, y=None, model_kwargs=None):
    cls = getattr(rp, model_name)
    if model_name.startswith('CompanionShocking'):
        return cls(oracle_filters(filter_names), y, redshift=z)
    return cls(redshift=z, **(model_kwargs or {}))


def oracle_truth(model_name, t, filter_names, params, z):
    m = oracle_model(model_name, z)
    return np.asarray(m(np.asarray(t, float), oracle_filters(filter_names), *params), float)


def oracle_for(wl):
    """(model, priors, log_posterior) of the oracle for a Workload."""
    model = oracle_model(wl.model_name, wl.z, wl.filter_names, wl.y, wl.model_kwargs)
    priors = wl.priors(ns=rp)
    f = oracle_filters(wl.filter_names)
    lp = rp.make_log_posterior(model, priors, wl.t, f, wl.y, wl.dy, use_sigma=wl.use_sigma, sigma_type=wl.sigma_type)
    return model, priors, lp


def oracle_log_posterior(wl):
    return oracle_for(wl)[2]


def kasen_sifto_truth(t, filter_names, z):
    """A smooth positive SN-Ia-like light curve (SiFTO template per filter, scaled to 1e21 W/Hz at peak)."""
    cols, tab = rp.load_sifto()
    t = np.asarray(t, float)
    tpk = 58000.
    out = np.empty(len(t))
    for i, (ti, fn) in enumerate(zip(t, filter_names)):
        col = tab[:, cols.index(fn)]
        out[i] = 1e21 * np.interp(ti - tpk, tab[:, 0], col / col.max(), left=0., right=0.) + 2e19
    return out


def example_sc4(**kw):
    return synthetic.example_sc4(**kw)


def synthetic_sc3(**kw):
    return synthetic.synthetic_sc3(oracle_truth, **kw)


def synthetic_sc4(**kw):
    return synthetic.synthetic_sc4(oracle_truth, **kw)


def synthetic_cs3(**kw):
    return synthetic.synthetic_cs3(kasen_sifto_truth, **kw)


def sed_epoch(rng, **kw):
    return synthetic.sed_epoch(oracle_truth, rng, **kw)
