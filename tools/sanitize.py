"""Workload for ``compute-sanitizer --tool memcheck`` (GPU box):

    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize.py

Two small graphs so that every kernel family of the training step runs once or twice under the
tool: the toy graph (gather-path SpMM, all four decoders) and a mini polypharmacy-shape graph
(staged SpMM kernels with TMA bulk copies, the tcgen05 layer-2 kernels, fused Adam), plus the
all-pairs / edge-scoring kernels.  No torch import (the engine is ctypes over the C ABI), kernels are
issued one by one (``DGN_CUDA_GRAPH=0``) so that a report names the launch it belongs to.
"""
import os
import sys
import time

os.environ.setdefault('DGN_CUDA_GRAPH', '0')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402

import common  # noqa: E402
from common import Case  # noqa: E402
from decagon_b200 import datasets  # noqa: E402


def drive(name, case, steps):
    t0 = time.time()
    eng = case.engine()
    eng.reset_optimizer()
    for step, (r, batch) in enumerate(case.batches(steps)):
        loss = eng.train_step(r, batch, negatives=None, dropout=0.1, seed=7, step=step, apply_update=True)
        assert np.isfinite(float(loss)), (name, step, loss)
    eng.forward(0.0, 7, 0)
    rng = np.random.RandomState(0)
    for g in case.graph.groups:
        r = eng.flat_index[(g, 0)]
        n_i, n_j = case.graph.n_nodes[g[0]], case.graph.n_nodes[g[1]]
        eng.predict(r)
        edges = np.stack([rng.randint(0, n_i, 300), rng.randint(0, n_j, 300)], axis=1)
        eng.predict_edges(r, edges)
        eng.rank_edges(r, edges, top=50)
        rel_k = rng.randint(0, case.graph.K[g], 300)
        eng.evaluate_edges(case.graph.groups.index(g), rel_k, edges, labels=rng.randint(0, 2, 300))
    eng.sync()
    eng.close()
    print('%s: %d steps + scoring in %.1fs' % (name, steps, time.time() - t0), flush=True)


if __name__ == '__main__':
    drive('toy (gather path, mixed decoders)', Case(datasets.toy_graph(common.MIXED_DECODERS)), 2)
    drive('mini polypharmacy shape (staged path, tensor-core layer 2)', Case(common.mini_poly(), batch_size=64), 3)
    drive('padded hidden sizes (48, 5)', Case(datasets.toy_graph(), hidden1=48, hidden2=5), 1)
    print('sanitize workload done')
