"""``lightcurve_mcmc`` -- drop-in for the reference driver (fitting.py:16-168).

Same signature, same error messages, same sequence of global-RNG draws (one ``np.random.rand(nwalkers, ndim)`` for the
starting positions, fitting.py:132); additive keywords ``precision`` ('fp64' default = reference arithmetic, 'fp32' =
throughput mode) and ``seed``.  The reference's ``log_posterior`` closure and emcee sampler (fitting.py:121-145) are a
:class:`DeviceProblem` and the device-resident :class:`EnsembleSampler` here; the optional burn-in / chain figure is
drawn by :mod:`lightcurve_fitting_b200.plotting` (needs matplotlib; presentation is outside the hot path).
"""
import warnings
import numpy as np

from .models import UniformPrior
from .sampler import EnsembleSampler

PRIOR_WARNING = 'The p_max/p_min keywords are deprecated. Use the priors keyword instead.'
MODEL_KWARGS_WARNING = 'The model_kwargs keyword is deprecated. These are now included in the model intialization.'

# what each output quantity needs from the light curve before it can be fitted (fitting.py:68-72)
_PREPARE = {'flux': ('calcFlux',), 'lum': ('calcAbsMag', 'calcLum')}


def build_problem(lc, model, priors, use_sigma=False, sigma_type='relative', precision=None):
    """The device counterpart of the ``log_posterior`` closure (fitting.py:121-128)."""
    q = model.output_quantity
    return model._device_problem(lc['MJD'].data, lc['filter'].data, lc[q].data, lc['d' + q].data,
                                 model._nmodel + (1 if use_sigma else 0), use_sigma=use_sigma, sigma_type=sigma_type,
                                 priors=priors, precision=precision or model.precision)


def _vector(value, ndim, wrong_length):
    """``value`` as a float vector of length ``ndim``; ``wrong_length`` is the message of the reference's exception."""
    if len(value) != ndim:             # len(None) raises TypeError, as the reference does for a missing p_up
        raise Exception(wrong_length)
    return np.array(value, float)


def _start_box(ndim, p_min, p_max, p_lo, p_up):
    """Prior box (deprecated ``p_min``/``p_max`` keywords) and start box of fitting.py:81-106, as four vectors."""
    box = {}
    for key, value, fill in (('p_min', p_min, -np.inf), ('p_max', p_max, np.inf)):
        if value is None:
            box[key] = np.full(ndim, fill)
        else:
            box[key] = _vector(value, ndim, PRIOR_WARNING)
            warnings.warn(PRIOR_WARNING)
    box['p_lo'] = box['p_min'] if p_lo is None else _vector(p_lo, ndim, 'p_lo must have length {:d}'.format(ndim))
    box['p_up'] = _vector(p_up, ndim, 'p_up must have length {:d}'.format(ndim))
    return box


def _check_start_inside_priors(names, priors, p_lo, p_up):
    """fitting.py:115-119: the start box must lie inside every prior's support."""
    for param, prior, lo, up in zip(names, priors, p_lo, p_up):
        for edge, value, limit, outside in (('p_lo', lo, prior.p_min, lo < prior.p_min), ('p_up', up, prior.p_max, up > prior.p_max)):
            if outside:
                raise Exception('starting guess for {} ({} = {}) is outside prior ({} = {})'.format(
                    param, edge, value, 'p_min' if edge == 'p_lo' else 'p_max', limit))


def lightcurve_mcmc(lc, model, priors=None, p_min=None, p_max=None, p_lo=None, p_up=None,
                    nwalkers=100, nsteps=1000, nsteps_burnin=1000, model_kwargs=None,
                    show=False, save_plot_as='', save_sampler_as='', use_sigma=False, sigma_type='relative',
                    precision=None, seed=None):
    """Fit an analytical model to observed photometry with an MCMC routine running on the GPU.

    Parameters and return value follow the reference (fitting.py:19-63); the returned sampler exposes
    ``chain``, ``flatchain``, ``get_chain()``, ``get_log_prob()``, ``acceptance_fraction``.
    """
    if model_kwargs is not None:
        raise Exception(MODEL_KWARGS_WARNING)
    for method in _PREPARE.get(model.output_quantity, ()):
        getattr(lc, method)()
    if use_sigma and model.input_names[-1] != '\\sigma':       # fitting.py:74-76 (class-level lists)
        model.input_names.append('\\sigma')
        model.units.append('')
    ndim = model.nparams

    box = _start_box(ndim, p_min, p_max, p_lo, p_up)
    if priors is None:
        priors = [UniformPrior(lo, hi) for lo, hi in zip(box['p_min'], box['p_max'])]
    elif len(priors) != ndim:
        raise Exception('priors must have length {:d}'.format(ndim))
    _check_start_inside_priors(model.input_names, priors, box['p_lo'], box['p_up'])

    problem = build_problem(lc, model, priors, use_sigma=use_sigma, sigma_type=sigma_type, precision=precision)
    sampler = EnsembleSampler(nwalkers, ndim, problem, seed=seed)
    starting_guesses = np.random.rand(nwalkers, ndim) * (box['p_up'] - box['p_lo']) + box['p_lo']
    pos, _, _ = sampler.run_mcmc(starting_guesses, nsteps_burnin)

    figure = None
    if show or save_plot_as:
        from . import plotting
        figure = plotting.ChainFigure(model.axis_labels[:ndim])
        figure.draw(0, sampler.chain, 'During Burn In')

    sampler.reset()
    sampler.run_mcmc(None, nsteps, skip_initial_state_check=True)
    if save_sampler_as:
        np.save(save_sampler_as, sampler.flatchain)
        print('saving sampler.flatchain as ' + save_sampler_as)
    if figure is not None:
        figure.draw(1, sampler.chain, 'After Burn In')
        figure.finish(save_plot_as, show)
    return sampler
