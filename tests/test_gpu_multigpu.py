"""Multi-GPU equivalence as a collected test: when at least two GPUs are visible this launches
``tests/multigpu_check.py`` under ``torch.distributed.run`` with two ranks (one process per GPU) and requires its
contract -- first forward / first loss within 1e-5 of the one-GPU engine, six Adam steps within 1e-4, embeddings
bit-identical on every rank.  On a one-GPU box it is skipped (the driver's scaling run reports the same numbers
through ``bench.py``'s ``equivalence`` object); the host-side partition logic runs on CPU in test_partition.py."""
import os
import socket
import subprocess
import sys

import pytest

from decagon_b200 import _lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize('env', [{}, {'DGN_DISABLE_STAGED': '1'}, {'DGN_CUDA_GRAPH': '0'}])
def test_two_ranks_match_one_gpu(env):
    """env = {}: the staged (shared-memory) kernels with the step replayed as a CUDA graph; DGN_DISABLE_STAGED: the
    relation partition of a group on the gather path (what config #5's 6 450-drug group takes); DGN_CUDA_GRAPH=0:
    every kernel issued from the host."""
    if _lib.device_count() < 2:
        pytest.skip('needs two GPUs (gpurun --gpus 2)')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'multigpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, **env))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'multigpu_check OK' in out.stdout
