"""Device-resident affine-invariant ensemble sampler with the ``emcee.EnsembleSampler`` surface the
reference relies on (fitting.py:130-153, bolometric.py:167-183): ``run_mcmc`` returning a 3-iterable
state, ``reset``, ``chain``, ``flatchain``, ``get_chain``, ``get_log_prob``, ``acceptance_fraction``.

The walker ensemble, its log-probabilities and the stored chain live in HBM for the whole run; each
stretch-move half-step is one fused kernel launch (proposal + prior + model + filter integration +
chi-square + accept + chain write-back).
"""
import ctypes as C
import numpy as np

from ._capi import lib, check, dptr, iptr, f64, i32


class AutocorrError(Exception):
    """Raised when the chain is too short for a reliable autocorrelation time (emcee.autocorr.AutocorrError)."""

    def __init__(self, tau, *args, **kwargs):
        self.tau = tau
        super().__init__(*args, **kwargs)


class State:
    """Minimal ``emcee.State``: unpacks to ``(coords, log_prob, random_state)``.

    The state ``run_mcmc`` returns is fetched from the device on first use (``fetch``: a callable returning
    ``(coords, log_prob)``): a caller that streams the chain to host buffers already holds the final positions and does not
    pay a second device-to-host copy.  The sampler materialises a pending state before it moves the ensemble again."""

    def __init__(self, coords=None, log_prob=None, random_state=None, fetch=None):
        self._coords, self._log_prob, self.random_state, self._fetch = coords, log_prob, random_state, fetch

    def materialize(self):
        if self._fetch is not None:
            self._coords, self._log_prob = self._fetch()
            self._fetch = None
        return self

    @property
    def coords(self):
        return self.materialize()._coords

    @property
    def log_prob(self):
        return self.materialize()._log_prob

    def __iter__(self):
        return iter((self.coords, self.log_prob, self.random_state))

    def __len__(self):
        return 3

    def __getitem__(self, i):
        return (self.coords, self.log_prob, self.random_state)[i]


def seed_from_global_rng():
    """Key for the device RNG derived from numpy's global legacy generator WITHOUT advancing it: emcee copies the global
    state into a private RandomState, so ``np.random.seed(s)`` makes a fit reproducible and the
    starting positions drawn afterwards (fitting.py:132) see the same stream as in the reference."""
    import hashlib
    kind, keys, pos, has_gauss, cached = np.random.get_state()
    h = hashlib.blake2b(digest_size=8)
    h.update(np.ascontiguousarray(keys).tobytes())
    h.update(repr((kind, int(pos), int(has_gauss), float(cached))).encode())
    return int.from_bytes(h.digest(), 'little') >> 1


def walkers_independent(coords):
    """emcee's initial-state check: finite, non-degenerate, condition number <= 1e8."""
    if not np.all(np.isfinite(coords)):
        return False
    Cm = coords - np.mean(coords, axis=0)[None, :]
    colmax = np.amax(np.abs(Cm), axis=0)
    if np.any(colmax == 0):
        return False
    Cm = Cm / colmax
    Cm = Cm / np.sqrt(np.sum(Cm ** 2, axis=0))
    return np.linalg.cond(Cm.astype(float)) <= 1e8


class EnsembleSampler:
    """Stretch-move ensemble sampler over a :class:`~lightcurve_fitting_b200.problem.DeviceProblem`.

    Parameters
    ----------
    nwalkers, ndim : int
    problem : DeviceProblem
        Replaces emcee's ``log_prob_fn``: the log-posterior is evaluated inside the kernels.
    seed : int, optional
        Key of the counter-based device RNG.  Default: a hash of numpy's global legacy RNG state (not advanced), so
        that ``np.random.seed(s)`` before a fit makes it reproducible, as with the reference.
    rank, world : int
        Shard one ensemble over ``world`` GPUs (see ``parallel.py``).
    """

    def __init__(self, nwalkers, ndim, problem, seed=None, rank=0, world=1):
        if ndim != problem.ndim:
            raise ValueError('ndim does not match the problem')
        self.nwalkers, self.ndim, self.problem = int(nwalkers), int(ndim), problem
        if seed is None:
            seed = seed_from_global_rng()
        self.seed = int(seed)
        h = C.c_void_p()
        check(lib().lcf_ensemble_create(problem.handle, self.nwalkers, C.c_uint64(self.seed), rank, world, C.byref(h)))
        self.handle = h
        self.iteration = 0
        self.last_ms = 0.
        self.last_launches = 0

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h is not None:
            try:
                lib().lcf_ensemble_destroy(h)
            except Exception:
                pass
            self.handle = None

    # -- emcee surface ----------------------------------------------------------------------
    def _set_initial(self, initial, skip_initial_state_check):
        if isinstance(initial, State):
            coords, log_prob = initial.coords, initial.log_prob
        else:
            coords, log_prob = initial, None
        coords = f64(np.atleast_2d(coords))
        if coords.shape != (self.nwalkers, self.ndim):
            raise ValueError('incompatible input dimensions')
        if not skip_initial_state_check and not walkers_independent(coords):
            raise ValueError('Initial state has a large condition number. Make sure that your walkers are linearly '
                             'independent for the best performance')
        lp = None if log_prob is None else f64(log_prob)
        check(lib().lcf_ensemble_set_state(self.handle, dptr(coords), dptr(lp) if lp is not None else None))

    def _fetch_state(self):
        coords = np.empty((self.nwalkers, self.ndim))
        lp = np.empty(self.nwalkers)
        check(lib().lcf_ensemble_get_state(self.handle, dptr(coords), dptr(lp)))
        return coords, lp

    def _state(self, lazy=False):
        if not lazy:
            return State(*self._fetch_state())
        self._pending = State(fetch=self._fetch_state)
        return self._pending

    def _settle(self):
        """Materialise the state handed out by the previous run before the ensemble moves on."""
        pending = getattr(self, '_pending', None)
        if pending is not None:
            pending.materialize()
            self._pending = None

    def _timing(self):
        ms = C.c_double(0.)
        n = C.c_int64(0)
        check(lib().lcf_ensemble_last_timing(self.handle, C.byref(ms), C.byref(n)))
        self.last_ms, self.last_launches = ms.value, n.value

    def run_mcmc(self, initial_state, nsteps, progress=False, progress_kwargs=None, skip_initial_state_check=False,
                 store=True, chain_out=None, log_prob_out=None, **kwargs):
        """Iterate the stretch move ``nsteps`` times from ``initial_state`` ([nwalkers, ndim] or a State; ``None``
        continues from the current position).  Returns a State that unpacks to (coords, log_prob, random_state).

        ``chain_out`` [nsteps, nwalkers, ndim] and ``log_prob_out`` [nsteps, nwalkers] (C-contiguous float64, ideally
        page-locked) make the run stream every finished step to the host while the next ones are sampled."""
        self._settle()
        if initial_state is not None:
            self._set_initial(initial_state, skip_initial_state_check)
        if chain_out is not None or log_prob_out is not None:
            if chain_out is None or log_prob_out is None or not store:
                raise ValueError('chain_out and log_prob_out go together and imply store=True')
            if (chain_out.shape != (int(nsteps), self.nwalkers, self.ndim) or log_prob_out.shape != (int(nsteps), self.nwalkers)
                    or chain_out.dtype != np.float64 or log_prob_out.dtype != np.float64
                    or not chain_out.flags.c_contiguous or not log_prob_out.flags.c_contiguous):
                raise ValueError('chain_out / log_prob_out must be C-contiguous float64 [nsteps, nwalkers(, ndim)]')
            check(lib().lcf_ensemble_run_to_host(self.handle, int(nsteps), dptr(chain_out), dptr(log_prob_out)))
        else:
            check(lib().lcf_ensemble_run(self.handle, int(nsteps), 1 if store else 0))
        self._timing()
        self.iteration += int(nsteps) if store else 0
        return self._state(lazy=True)

    def run_replay(self, initial_state, draws, store=True):
        """Drive the move with injected draws in emcee's order (see ``lcf_ensemble_run_replay``).

        ``draws`` is the list recorded by the oracle's ``StretchReplay``: one dict per step with ``inds`` (split
        label per walker) and ``halves`` = [{'z', 'rint', 'logu'}, ...] for split 0 and split 1.
        """
        self._settle()
        if initial_state is not None:
            self._set_initial(initial_state, True)
        S, W = len(draws), self.nwalkers
        split = np.empty((S, W), np.int32)
        z = np.empty((S, W))
        rint = np.empty((S, W), np.int32)
        logu = np.empty((S, W))
        for s, d in enumerate(draws):
            split[s] = d['inds']
            z[s] = np.concatenate([h['z'] for h in d['halves']])
            rint[s] = np.concatenate([h['rint'] for h in d['halves']])
            logu[s] = np.concatenate([h['logu'] for h in d['halves']])
        check(lib().lcf_ensemble_run_replay(self.handle, S, 1 if store else 0, iptr(split), dptr(z), iptr(rint), dptr(logu)))
        self._timing()
        self.iteration += S if store else 0
        return self._state()

    def reserve(self, nsteps):
        """Make room in HBM for ``nsteps`` more stored steps now (otherwise the chain buffer grows inside ``run_mcmc``, like
        emcee's ``Backend.grow``); additive, not part of the emcee surface."""
        check(lib().lcf_ensemble_reserve(self.handle, int(nsteps)))

    def reset(self):
        check(lib().lcf_ensemble_reset(self.handle))
        self.iteration = 0

    def get_chain(self, flat=False, thin=1, discard=0, out=None):
        n = lib().lcf_ensemble_nstored(self.handle)
        if out is None:
            out = np.empty((n, self.nwalkers, self.ndim))
        else:                                    # caller-provided (e.g. pinned) buffer
            out = out.reshape(-1)[:n * self.nwalkers * self.ndim].reshape(n, self.nwalkers, self.ndim)
        check(lib().lcf_ensemble_get_chain(self.handle, dptr(out)))
        out = out[discard + thin - 1::thin]
        return out.reshape(-1, self.ndim) if flat else out

    def get_log_prob(self, flat=False, thin=1, discard=0, out=None):
        n = lib().lcf_ensemble_nstored(self.handle)
        if out is None:
            out = np.empty((n, self.nwalkers))
        else:
            out = out.reshape(-1)[:n * self.nwalkers].reshape(n, self.nwalkers)
        check(lib().lcf_ensemble_get_log_prob(self.handle, dptr(out)))
        out = out[discard + thin - 1::thin]
        return out.reshape(-1) if flat else out

    @property
    def chain(self):
        """[nwalkers, nsteps, ndim] (emcee's deprecated but still-used layout, fitting.py:139)."""
        return np.swapaxes(self.get_chain(), 0, 1)

    @property
    def flatchain(self):
        """[nsteps * nwalkers, ndim], step-major (fitting.py:147)."""
        return self.get_chain(flat=True)

    @property
    def lnprobability(self):
        return np.swapaxes(self.get_log_prob(), 0, 1)

    @property
    def flatlnprobability(self):
        return self.get_log_prob(flat=True)

    def get_autocorr_time(self, discard=0, c=5, tol=50, quiet=False, max_lag=0, return_window=False):
        """Integrated autocorrelation time per dimension, ``emcee.EnsembleSampler.get_autocorr_time`` semantics (Sokal's
        automatic window with constant ``c``; ``AutocorrError`` unless the chain is longer than ``tol`` times the estimate
        or ``quiet``), computed on the device-resident chain."""
        tau = np.empty(self.ndim)
        win = np.empty(self.ndim, np.int64)
        check(lib().lcf_ensemble_diagnostics(self.handle, int(discard), float(c), int(max_lag), dptr(tau),
                                             win.ctypes.data_as(C.POINTER(C.c_int64)), None))
        n = lib().lcf_ensemble_nstored(self.handle) - int(discard)
        if tol > 0 and np.any(tol * tau > n):
            msg = ('The chain is shorter than {0} times the integrated autocorrelation time for {1} parameter(s). Use this '
                   'estimate with caution and run a longer chain!\nN/{0} = {2:.0f};\ntau: {3}').format(
                       tol, int(np.sum(tol * tau > n)), n / tol, tau)
            if not quiet:
                raise AutocorrError(tau, msg)
            import warnings
            warnings.warn(msg)
        return (tau, win) if return_window else tau

    def get_split_rhat(self, discard=0):
        """Split Gelman-Rubin statistic per dimension over the 2 x nwalkers half-chains (device-side moments)."""
        rhat = np.empty(self.ndim)
        check(lib().lcf_ensemble_diagnostics(self.handle, int(discard), 5., 0, None, None, dptr(rhat)))
        return rhat

    @property
    def acceptance_fraction(self):
        acc = np.empty(self.nwalkers, np.int64)
        check(lib().lcf_ensemble_get_accepted(self.handle, acc.ctypes.data_as(C.POINTER(C.c_int64))))
        return acc / max(self.iteration, 1)
