#!/usr/bin/env python
"""Multi-GPU correctness check of the shared ensemble (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multigpu.py

The chain of a world-size-N ensemble must be bit-identical to the single-GPU chain of the same seed (the device RNG is
keyed by the global walker index; the launch shape is pinned so that the FP32 partial sums are taken in the same order),
both with the fused peer-memory exchange ('p2p') and with the NCCL all-gather ('nccl').
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import bench
    from lightcurve_fitting_b200 import _capi
    from lightcurve_fitting_b200.parallel import ShardedEnsemble
    from lightcurve_fitting_b200.sampler import EnsembleSampler
    _capi.check(_capi.lib().lcf_set_device(local))
    _capi.check(_capi.lib().lcf_set_tuning_ex(32, 8, 1))
    wl = bench.workload(bench.device_truth, 200)
    prob = wl.device_problem('fp32')
    W, nsteps = 64 * world * 4, 12
    p0 = wl.start(W, np.random.default_rng(3))
    single = EnsembleSampler(W, wl.ndim, prob, seed=77)
    single.run_mcmc(p0, nsteps, skip_initial_state_check=True)
    ref, ref_lp = single.get_chain(), single.get_log_prob()
    ok = True
    for exchange in ('p2p', 'nccl'):
        ens = ShardedEnsemble(prob, W, seed=77, rank=rank, world=world, exchange=exchange)
        ens.set_state(p0)
        ens.run(nsteps, store=True)
        ens.finish()
        ch, lp = ens.get_own_chain()
        first, count = ens.own_walkers()
        same = np.array_equal(ch, ref[:, first:first + count]) and np.array_equal(lp, ref_lp[:, first:first + count])
        st = ens.sampler._state()
        same_state = np.array_equal(st.coords, ref[-1]) and np.array_equal(st.log_prob, ref_lp[-1])
        print('[rank %d] %s: own chain identical to single-GPU: %s; full replica identical: %s (fused=%s)'
              % (rank, exchange, same, same_state, ens.fused), flush=True)
        ok = ok and same and same_state
        ens.close()
        del ens
        dist.barrier()
    t = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('MULTIGPU CHECK', 'OK' if int(t.item()) else 'FAILED', flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) else 1)


if __name__ == '__main__':
    main()
