"""lightcurve_fitting_b200 -- B200-native (sm_100a) MCMC hot path of griffin-h/lightcurve_fitting.

Drop-in names: ``lightcurve_mcmc``, the ``Model`` classes, the ``Prior`` classes, ``blackbody_to_filters``,
``planck_fast``, ``spectrum_mcmc`` / ``blackbody_mcmc``, ``calculate_bolometric``.  All numerical work runs in
hand-written CUDA kernels behind the C ABI of ``include/lcf.h`` (``liblcf_b200.so``); there is no CPU fallback.
"""
__version__ = '0.1.0'

from . import filters, lightcurve, models, fitting, bolometric  # noqa: F401
from .fitting import lightcurve_mcmc  # noqa: F401
from .bolometric import spectrum_mcmc, blackbody_mcmc, calculate_bolometric  # noqa: F401
from .lightcurve import LC  # noqa: F401
