"""-m gpu, needs >= 2 GPUs on the box (gpurun --gpus 2): numerics of the data-parallel search step on hardware --
the peer-memory fused optimiser step vs NCCL all-reduce + FusedAdam vs the chunked CPU oracle (tests/dist_gpu_worker.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.timeout(600)
def test_peer_memory_step_matches_nccl_and_oracle():
    world = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1',
           '--master-port', '29631', os.path.join(ROOT, 'tests', 'dist_gpu_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=540)
    assert 'DIST_OK' in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
