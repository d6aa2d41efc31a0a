"""Multi-GPU execution of the hot path: one process per GPU (``torch.distributed``, NCCL over NVLink 5).

Two modes:

* **independent problems** (light curves / SED epochs): ``shard_items`` hands each rank a disjoint subset; no
  data-path collective.
* **one ensemble split across GPUs**: every rank keeps a full replica of the walker positions, stored
  colour-major ``[even walkers | odd walkers]``.  In half-step ``h`` a rank proposes / evaluates / accepts only
  its contiguous slice of colour ``h`` (fused kernel).  Colour ``1-h`` is never written during the half-step.
  The device RNG is keyed by the *global* walker index: chains are independent of the number of GPUs.
  Two ways to publish the updated slice:

  - ``exchange='p2p'`` (default on CUDA): the peers' replicas are mapped into every process once (cudaIpc handles
    exchanged through torch.distributed) and the kernel's accept epilogue stores each accepted walker straight into
    them over NVLink; the last CTA of a launch publishes a half-step counter into the peers' flag arrays and the
    next launch spins on its own flags.  Compute and exchange are ONE kernel, whole chains run from one C call with
    no NCCL collective and no host round trip per half-step.
  - ``exchange='nccl'``: ONE in-place ``all_gather_into_tensor`` of the colour block per half-step, stream-ordered
    behind the kernel (the baseline; also what the gloo CPU tests exercise).

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

    def __init__(self, problem, nwalkers, seed, rank, world, group=None, exchange='p2p'):
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
        self.fused = False
        self._dirty = False                  # sampling kernels may still be in flight on some rank
        if world > 1 and exchange == 'p2p':
            self._attach_peers()

    def _attach_peers(self):
        """Exchange cudaIpc handles of (coords, log_prob, flags) and map every peer's replica (same node)."""
        import torch.distributed as dist
        torch, L, h = self.torch, lib(), self.sampler.handle
        buf = C.create_string_buffer(192)
        check(L.lcf_ensemble_ipc_export(h, buf))
        mine = torch.tensor(list(buf.raw), dtype=torch.uint8, device='cuda')
        allh = torch.empty(self.world * 192, dtype=torch.uint8, device='cuda')
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        check(L.lcf_ensemble_peers_attach_ipc(h, bytes(allh.cpu().numpy().tobytes())))
        dist.barrier(group=self.group)
        self.fused = True

    def _quiesce(self):
        """Every rank's kernels (which store into the other replicas) have finished."""
        if self.world > 1:
            import torch.distributed as dist
            self.torch.cuda.current_stream().synchronize()
            check(lib().lcf_ensemble_sync(self.sampler.handle))
            dist.barrier(group=self.group)
            self._dirty = False

    def close(self):
        """Unmap the peers' replicas on every rank before any rank frees its own (collective when fused)."""
        if self.fused:
            import torch.distributed as dist
            self._quiesce()
            check(lib().lcf_ensemble_peers_detach(self.sampler.handle))
            self.fused = False
            dist.barrier(group=self.group)

    def set_state(self, coords):
        """Start positions ``[nwalkers, ndim]`` in emcee's walker order (the same array on every rank, as in the reference's
        ``run_mcmc(initial)``).  Each rank uploads and evaluates ONLY the walkers it owns; their rows and log-probabilities
        reach the other replicas by posted NVLink stores from a device kernel (fused exchange) or one all-gather per colour
        block (NCCL exchange).  One barrier before (only when a run is still in flight) and one after."""
        if self.world > 1 and self.coords.is_cuda and self._even():
            import torch.distributed as dist
            from ._capi import dptr
            if self._dirty:
                self._quiesce()              # the peers' kernels read the replicas about to be overwritten
            coords = np.asarray(coords, float)
            if coords.shape != (self.nwalkers, self.ndim):
                raise ValueError('incompatible input dimensions')
            first, count = self.own_walkers()
            mine = np.ascontiguousarray(coords[first:first + count])
            check(lib().lcf_ensemble_set_state_slice(self.sampler.handle, first, count, dptr(mine)))
            if not self.fused:
                for block in (self.coords[:self.n0], self.coords[self.n0:], self.logp[:self.n0].unsqueeze(1),
                              self.logp[self.n0:].unsqueeze(1)):
                    exchange_half(block, self.rank, self.world, self.group)
                self.torch.cuda.current_stream().synchronize()
            dist.barrier(group=self.group)   # every replica is complete before anyone steps
            self._dirty = False
        else:
            if self.fused:
                self._quiesce()
            self.sampler._set_initial(coords, True)
            if self.fused:
                self._quiesce()

    def _even(self):
        (b0, c0), (b1, c1) = self.own
        return (b0, c0) == (b1, c1) and c0 * 2 * self.world == self.nwalkers

    def reserve(self, nsteps):
        check(lib().lcf_ensemble_reserve(self.sampler.handle, int(nsteps)))

    def run(self, nsteps, store=False, chain_out=None, log_prob_out=None):
        """``nsteps`` stretch-move iterations.  Fused exchange: one C call for the whole run; ``chain_out``
        [nsteps, own count, ndim] / ``log_prob_out`` [nsteps, own count] (page-locked) then receive this rank's walkers
        step by step while the next steps are sampled.  NCCL exchange: one kernel + one all-gather per half-step."""
        L, h = lib(), self.sampler.handle
        self._dirty = True
        if store:
            self.reserve(nsteps)
        if self.fused and chain_out is not None:
            from ._capi import dptr
            first, count = self.own_walkers()
            if chain_out.shape != (int(nsteps), count, self.ndim) or log_prob_out.shape != (int(nsteps), count):
                raise ValueError('chain_out / log_prob_out must be [nsteps, own walkers(, ndim)]')
            check(L.lcf_ensemble_run_to_host_slice(h, int(nsteps), first, count, dptr(chain_out), dptr(log_prob_out)))
            self.sampler.iteration += int(nsteps)
            return
        if self.fused:                       # compute + exchange are one kernel: the whole run is one C call
            check(L.lcf_ensemble_run(h, int(nsteps), 1 if store else 0))
            if store:
                self.sampler.iteration += int(nsteps)
            return
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
        if self.fused:
            self._quiesce()                  # log-probabilities travelled with the positions
        elif self.world > 1:
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
