"""``lightcurve_mcmc`` -- drop-in for the reference driver (fitting.py:16-168).

Same positional/keyword signature and validation; additive keywords: ``precision`` ('fp64' default =
reference arithmetic, 'fp32' = throughput mode) and ``seed``.  The ``log_posterior`` closure and the emcee
sampler of the reference (fitting.py:121-145) are replaced by a :class:`DeviceProblem` and the
device-resident :class:`EnsembleSampler`.
"""
import warnings
import numpy as np

from .models import UniformPrior
from .sampler import EnsembleSampler

PRIOR_WARNING = 'The p_max/p_min keywords are deprecated. Use the priors keyword instead.'
MODEL_KWARGS_WARNING = 'The model_kwargs keyword is deprecated. These are now included in the model intialization.'


def build_problem(lc, model, priors, use_sigma=False, sigma_type='relative', precision=None):
    """The device counterpart of the ``log_posterior`` closure (fitting.py:121-128)."""
    q = model.output_quantity
    return model._device_problem(lc['MJD'].data, lc['filter'].data, lc[q].data, lc['d' + q].data,
                                 model._nmodel + (1 if use_sigma else 0), use_sigma=use_sigma, sigma_type=sigma_type,
                                 priors=priors, precision=precision or model.precision)


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

    if model.output_quantity == 'flux':
        lc.calcFlux()
    elif model.output_quantity == 'lum':
        lc.calcAbsMag()
        lc.calcLum()

    if use_sigma and model.input_names[-1] != '\\sigma':
        model.input_names.append('\\sigma')
        model.units.append('')

    ndim = model.nparams

    # DEPRECATED
    if p_min is None:
        p_min = np.tile(-np.inf, ndim)
    elif len(p_min) == ndim:
        p_min = np.array(p_min, float)
        warnings.warn(PRIOR_WARNING)
    else:
        raise Exception(PRIOR_WARNING)

    # DEPRECATED
    if p_max is None:
        p_max = np.tile(np.inf, ndim)
    elif len(p_max) == ndim:
        p_max = np.array(p_max, float)
        warnings.warn(PRIOR_WARNING)
    else:
        raise Exception(PRIOR_WARNING)

    if p_lo is None:
        p_lo = p_min
    elif len(p_lo) == ndim:
        p_lo = np.array(p_lo, float)
    else:
        raise Exception('p_lo must have length {:d}'.format(ndim))

    if len(p_up) == ndim:
        p_up = np.array(p_up, float)
    else:
        raise Exception('p_up must have length {:d}'.format(ndim))

    if priors is None:
        priors = [UniformPrior(p0, p1) for p0, p1 in zip(p_min, p_max)]
    elif len(priors) != ndim:
        raise Exception('priors must have length {:d}'.format(ndim))

    for param, prior, p0, p1 in zip(model.input_names, priors, p_lo, p_up):
        if p0 < prior.p_min:
            raise Exception(f'starting guess for {param} (p_lo = {p0}) is outside prior (p_min = {prior.p_min})')
        if p1 > prior.p_max:
            raise Exception(f'starting guess for {param} (p_up = {p1}) is outside prior (p_max = {prior.p_max})')

    problem = build_problem(lc, model, priors, use_sigma=use_sigma, sigma_type=sigma_type, precision=precision)
    sampler = EnsembleSampler(nwalkers, ndim, problem, seed=seed)

    starting_guesses = np.random.rand(nwalkers, ndim) * (p_up - p_lo) + p_lo
    pos, _, _ = sampler.run_mcmc(starting_guesses, nsteps_burnin)

    if show or save_plot_as:
        import matplotlib.pyplot as plt
        fig, ax = plt.subplots(ndim, 2, figsize=(12., 2. * ndim))
        ax1 = ax[:, 0]
        for i in range(ndim):
            ax1[i].plot(sampler.chain[:, :, i].T, 'k', alpha=0.2)
            ax1[i].set_ylabel(model.axis_labels[i])
        ax1[0].set_title('During Burn In')
        ax1[-1].set_xlabel('Step Number')

    sampler.reset()
    sampler.run_mcmc(None, nsteps, skip_initial_state_check=True)
    if save_sampler_as:
        np.save(save_sampler_as, sampler.flatchain)
        print('saving sampler.flatchain as ' + save_sampler_as)

    if show or save_plot_as:
        ax2 = ax[:, 1]
        for i in range(ndim):
            ax2[i].plot(sampler.chain[:, :, i].T, 'k', alpha=0.2)
            ax2[i].set_ylabel(model.axis_labels[i])
            ax2[i].yaxis.set_label_position('right')
            ax2[i].yaxis.tick_right()
        ax2[0].set_title('After Burn In')
        ax2[-1].set_xlabel('Step Number')
        fig.tight_layout()
        if save_plot_as:
            print('saving chain plot as ' + save_plot_as)
            fig.savefig(save_plot_as)
        if show:
            plt.show()

    return sampler
