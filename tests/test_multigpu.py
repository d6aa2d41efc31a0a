"""Multi-GPU test (needs >= 2 GPUs on one node; skipped otherwise): one process per GPU under torchrun, the shared ensemble's
chains must be bit-identical to the single-GPU chain with the fused peer-memory exchange and with the NCCL all-gather
(tools/check_multigpu.py).  The CPU suite covers the exchange logic with a 2-rank gloo group (test_host_cpu.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpus() < 2, reason='needs at least 2 GPUs')
def test_shared_ensemble_bit_identical_across_gpu_counts():
    n = min(_ngpus(), 8)
    n = 1 << (n.bit_length() - 1)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n), '--master-addr', '127.0.0.1',
           '--master-port', '29541', os.path.join(ROOT, 'tools', 'check_multigpu.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and 'MULTIGPU CHECK OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
