"""Host-side multi-GPU logic on CPU: the relation partition (the library's own rule through the C ABI) and
the all-gather of the exchange handles over a world_size-2 gloo group."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from decagon_b200 import _lib


def test_partition_is_balanced_and_deterministic():
    rng = np.random.RandomState(0)
    w = np.exp(rng.uniform(np.log(500), np.log(28568), size=1928)).astype(np.int64)
    for world in (1, 2, 4, 8):
        owner = _lib.partition_relations(w, world)
        assert owner.min() >= 0 and owner.max() == world - 1
        loads = np.array([w[owner == r].sum() for r in range(world)])
        assert loads.max() <= 1.01 * loads.mean()  # longest-processing-time: within 1 %
        assert np.array_equal(owner, _lib.partition_relations(w, world))


def test_partition_rejects_bad_arguments():
    with pytest.raises(ValueError):
        _lib.partition_relations(np.array([1, 2, 3]), 0)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    # what Engine.connect does with the 64-byte IPC handles; plus every rank derives the same owners
    handle = bytes([rank]) * 64
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    w = np.arange(1, 101, dtype=np.int64) ** 2
    owner = _lib.partition_relations(w, world)
    mine = [int(k) for k in np.nonzero(owner == rank)[0]]
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    out.put((rank, handles, everyone))
    dist.destroy_process_group()


def test_handles_and_owners_agree_over_gloo():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, handles, everyone in results:
        assert handles == [bytes([r]) * 64 for r in range(world)]
        assert sorted(k for part in everyone for k in part) == list(range(100))  # a partition: disjoint, complete
        assert everyone == results[0][2]
