"""Sharded state vector on real GPUs (needs >= 2 devices on the box; skipped
otherwise): tools/dist_check.py under torchrun compares the sharded result with
the single-GPU simulator for both dtypes and measures the peer-memory swap."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize('nproc', [2, 4, 8])
def test_sharded_matches_single_gpu(nproc):
    import torch

    if torch.cuda.device_count() < nproc:
        pytest.skip(f'needs {nproc} GPUs, have {torch.cuda.device_count()}')
    env = dict(os.environ, B2Q_SWAP_NLOCAL='26')
    port = 29600 + nproc
    proc = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={nproc}',
         '--master-addr', '127.0.0.1', '--master-port', str(port),
         os.path.join(ROOT, 'tools', 'dist_check.py')],
        capture_output=True, text=True, timeout=1200, env=env, cwd=ROOT,
    )
    out = proc.stdout + proc.stderr
    assert 'DIST CHECK PASSED' in out, out[-4000:]
