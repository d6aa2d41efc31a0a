"""Multi-GPU execution of the hot path: one process per GPU (``torch.distributed``, NCCL over NVLink 5).

Two modes:

* **independent problems** (light curves / SED epochs): ``shard_items`` hands each rank a disjoint subset; no
  data-path collective.
* **one ensemble split across GPUs**: every rank keeps a full replica of the walker positions, stored
  colour-major ``[even walkers | odd walkers]``.  In half-step ``h`` a rank proposes / evaluates / accepts only
  its contiguous slice of colour ``h`` (fused kernel), then the updated slices are exchanged with ONE in-place
  all-gather of that colour block, stream-ordered behind the kernel (the library launches on torch's current
  stream).  Colour ``1-h`` is never written during the half-step, so no other synchronisation is needed.
  The device RNG is keyed by the *global* walker index: chains are independent of the number of GPUs.

torch is plumbing here (process group, streams, a zero-copy view of the library's device buffers).
"""
import ctypes as C
import numpy as np

from ._capi import lib, check


def shard_items(n_items, rank, world):
    """Indices of the items (light curves / epochs) owned by ``rank``: round-robin, no communication."""
    return list(range(rank, n_items, world))


def half_slices(nwalkers, rank, world):
    """(begin, count) of this rank's slice inside each colour block -- mirrors ``lcf_ensemble_create``."""
    n0 = (nwalkers + 1) // 2
    out = []
    for n in (n0, nwalkers - n0):
        per = (n + world - 1) // world
        b, e = min(n, per * rank), min(n, per * (rank + 1))
        out.append((b, e - b))
    return n0, out


def exchange_half(block, rank, world, group=None):
    """In-place all-gather of one colour block ``[n, D]`` whose ``rank``-th equal slice is up to date."""
    import torch.distributed as dist
    n = block.shape[0]
    if n % world:
        raise ValueError('sharded ensembles need the half-ensemble size to be a multiple of the number of GPUs')
    per = n // world
    mine = block[rank * per:(rank + 1) * per]
    if block.is_cuda:
        dist.all_gather_into_tensor(block, mine, group=group)
    else:  # gloo (CPU tests): gather into a list of views
        import torch
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine.contiguous(), group=group)
        for r, p in enumerate(parts):
            block[r * per:(r + 1) * per] = p
    return block


class _DevArray:
    """Zero-copy torch view of a raw device pointer owned by the C library."""

    def __init__(self, ptr, shape, typestr='<f8'):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False),
                                         'version': 2, 'strides': None}


class ShardedEnsemble:
    """One ensemble of ``nwalkers`` walkers split across the ranks of a torch.distributed group."""

    def __init__(self, problem, nwalkers, seed, rank, world, group=None):
        import torch
        from .sampler import EnsembleSampler
        self.torch = torch
        self.rank, self.world, self.group = rank, world, group
        self.nwalkers, self.ndim = int(nwalkers), problem.ndim
        n0 = (self.nwalkers + 1) // 2
        if world > 1 and (n0 % world or (self.nwalkers - n0) % world):
            raise ValueError('nwalkers/2 must be a multiple of the number of GPUs')
        self.sampler = EnsembleSampler(nwalkers, self.ndim, problem, seed=seed, rank=rank, world=world)
        h = self.sampler.handle
        # kernels and collectives share torch's current stream
        check(lib().lcf_ensemble_set_stream(h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        dc, dl, st = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n0c = C.c_int64()
        ob, oc = (C.c_int64 * 2)(), (C.c_int64 * 2)()
        check(lib().lcf_ensemble_device_view(h, C.byref(dc), C.byref(dl), C.byref(st), C.byref(n0c), ob, oc))
        self.n0 = n0c.value
        self.coords = torch.as_tensor(_DevArray(dc.value, (self.nwalkers, self.ndim)), device='cuda')
        self.logp = torch.as_tensor(_DevArray(dl.value, (self.nwalkers,)), device='cuda')
        self.own = [(ob[0], oc[0]), (ob[1], oc[1])]

    def set_state(self, coords):
        self.sampler._set_initial(coords, True)

    def reserve(self, nsteps):
        check(lib().lcf_ensemble_reserve(self.sampler.handle, int(nsteps)))

    def run(self, nsteps, store=False):
        """``nsteps`` stretch-move iterations; one fused kernel + one all-gather per half-step."""
        L, h = lib(), self.sampler.handle
        if store:
            self.reserve(nsteps)
        blocks = (self.coords[:self.n0], self.coords[self.n0:])
        for _ in range(int(nsteps)):
            for half in (0, 1):
                check(L.lcf_ensemble_half_step(h, half, 1 if store else 0))
                if self.world > 1:
                    exchange_half(blocks[half], self.rank, self.world, self.group)
            check(L.lcf_ensemble_end_step(h, 1 if store else 0))
        if store:
            self.sampler.iteration += int(nsteps)

    def finish(self):
        """Make log-probabilities consistent on every rank and surface NaN errors (emcee: ValueError)."""
        if self.world > 1:
            exchange_half(self.logp[:self.n0].unsqueeze(1), self.rank, self.world, self.group)
            exchange_half(self.logp[self.n0:].unsqueeze(1), self.rank, self.world, self.group)
        self.torch.cuda.current_stream().synchronize()
        check(lib().lcf_ensemble_sync(self.sampler.handle))

    def own_walkers(self):
        """(first, count) of the logical walkers this rank owns (both colours of its slice are contiguous)."""
        (b0, c0), (b1, c1) = self.own
        if (b0, c0) != (b1, c1):
            raise ValueError('uneven colour slices: use gather_chain()')
        return 2 * b0, 2 * c0

    def get_own_chain(self, out_chain=None, out_log_prob=None):
        """This rank's part of the stored chain: ([nsteps, count, ndim], [nsteps, count]); D2H of own rows only."""
        from ._capi import dptr
        first, count = self.own_walkers()
        n = lib().lcf_ensemble_nstored(self.sampler.handle)
        if out_chain is None:
            out_chain = np.empty((n, count, self.ndim))
        if out_log_prob is None:
            out_log_prob = np.empty((n, count))
        ch = out_chain.reshape(-1)[:n * count * self.ndim].reshape(n, count, self.ndim)
        lp = out_log_prob.reshape(-1)[:n * count].reshape(n, count)
        check(lib().lcf_ensemble_get_chain_slice(self.sampler.handle, first, count, dptr(ch), dptr(lp)))
        return ch, lp

    def gather_chain(self):
        """Full stored chain ``[nsteps, nwalkers, ndim]`` on every rank (each rank stored only its walkers)."""
        import torch.distributed as dist
        chain = self.sampler.get_chain()
        if self.world == 1:
            return chain
        own = np.zeros(self.nwalkers, bool)
        for colour, (b, c) in enumerate(self.own):
            own[2 * np.arange(b, b + c) + colour] = True
        t = self.torch.from_numpy(np.where(own[None, :, None], chain, 0.)).cuda()
        dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()
