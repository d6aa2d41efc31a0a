"""emcee stand-in: EnsembleSampler backed by the oracle's restatement of the stretch move (parity unpinned)."""
import numpy as np
from oracle.reference_port import StretchReplay


class EnsembleSampler(StretchReplay):
    def __init__(self, nwalkers, ndim, log_prob_fn, **kw):
        rs = np.random.RandomState()
        rs.set_state(np.random.get_state())      # emcee copies numpy's global legacy RNG state
        super().__init__(nwalkers, ndim, log_prob_fn, random_state=rs)

    def run_mcmc(self, initial, nsteps, progress=False, progress_kwargs=None, skip_initial_state_check=False, **kw):
        return super().run_mcmc(initial, nsteps)
