#!/usr/bin/env python
"""Generates the committed golden fixtures by running the REFERENCE itself in this container.

    python tests/golden/make_golden.py          (needs /root/reference; not needed at test time)

What is pinned here:

* ``iterator_tiny.npz``   -- full arrays of the reference's ``EdgeMinibatchIterator``
  (``decagon/deep/minibatch.py``, imported UNMODIFIED) on a small 3-type graph: every split, every
  normalised adjacency tuple (float64), the flat relation index and one epoch of minibatches.
* ``iterator_digests.json`` -- SHA-256 digests of the same arrays for BASELINE config #1 (toy) and
  config #3 (polypharmacy shape, 1932 relation matrices), too large to commit as arrays.
* ``nppredictor.npz``     -- the reference's own numpy statement of the DEDICOM all-pairs score
  (``main/Predictor/NpPredictor.py:293-313``: ``Z @ D @ R @ D @ Z.T`` -> sigmoid -> np.take) executed
  from its source (the three methods are compiled out of the file because importing package
  ``main`` needs TensorFlow) on the reference's dumped parameters
  (``ndarray-dumpGlobalRelations.npy``, ``ndarray-dumpEmbeddingImportance.npyz.npz``) and a seeded Z.

TensorFlow 1.8 cannot run here, so the TF arithmetic of layers.py / model.py / optimizer.py has no
golden vectors: parity at that boundary is UNPINNED (see oracle/decagon_oracle.py).
"""
import ast
import contextlib
import hashlib
import importlib.util
import io
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from decagon.deep.minibatch import EdgeMinibatchIterator as RefIterator  # noqa: E402  (the reference)
from decagon_b200 import datasets  # noqa: E402

_spec = importlib.util.spec_from_file_location('ref_sparse', os.path.join(REF, 'main/Utils/Sparse.py'))
ref_sparse = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(ref_sparse)

PLACEHOLDERS = {k: k for k in ['batch', 'batch_edge_type_idx', 'batch_row_edge_type', 'batch_col_edge_type', 'dropout']}
SPLITS = ['train_edges', 'val_edges', 'val_edges_false', 'test_edges', 'test_edges_false']


def to_reference_matrices(adj_mats):
    """Same matrices, wrapped in the REFERENCE's RelationCsrMatrix with the same twin links."""
    out, by_id = {}, {}
    for et, mtxs in adj_mats.items():
        out[et] = []
        for m in mtxs:
            r = ref_sparse.RelationCsrMatrix(sp.csr_matrix(m))
            r.isTranspose = m.isTranspose
            by_id[m.id] = r
            out[et].append(r)
    for et, mtxs in adj_mats.items():
        for m, r in zip(mtxs, out[et]):
            if m.transposedMtxLink is not None:
                r.transposedMtxLink = by_id[m.transposedMtxLink.id]
    return out


def run_reference_iterator(inputs, seed, batch_size, val_test_size, epoch_seed):
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        it = RefIterator(to_reference_matrices(inputs.adj_mats), inputs.feat, inputs.edge_types, {},
                         batch_size=batch_size, val_test_size=val_test_size)
    arrays = {}
    for r, (i, j, k) in it.idx2edge_type.items():
        for name in SPLITS:
            arrays['r%d/%s' % (r, name)] = np.asarray(getattr(it, name)[i, j][k])
        coords, values, shape = it.adj_train[i, j][k]
        arrays['r%d/adj_coords' % r] = np.asarray(coords)
        arrays['r%d/adj_values' % r] = np.asarray(values)
        arrays['r%d/adj_shape' % r] = np.asarray(shape)
    arrays['flat'] = np.array([it.idx2edge_type[r] for r in range(len(it.idx2edge_type))])
    np.random.seed(epoch_seed)
    it.shuffle()
    seq, batches = [], []
    while not it.end():
        fd = it.next_minibatch_feed_dict(PLACEHOLDERS)
        seq.append(fd['batch_edge_type_idx'])
        batches.append(np.asarray(fd['batch']))
    arrays['epoch/relation'] = np.array(seq)
    arrays['epoch/batches'] = np.stack(batches) if batches else np.zeros((0, batch_size, 2))
    return arrays


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()


def digest_by_kind(arrays):
    """One digest per array kind: SHA-256 over the per-relation digests in flat relation order
    (a checksum of checksums -- 15 k arrays at the polypharmacy shape)."""
    kinds = {}
    n_rel = len(arrays['flat'])
    for key in ['flat', 'epoch/relation', 'epoch/batches']:
        kinds[key] = digest(arrays[key])
    for kind in SPLITS + ['adj_coords', 'adj_values', 'adj_shape']:
        h = hashlib.sha256()
        for r in range(n_rel):
            h.update(digest(arrays['r%d/%s' % (r, kind)]).encode())
        kinds[kind] = h.hexdigest()
    return kinds


def tiny_graph():
    return datasets.polypharmacy_graph(n_types=3, seed=11, n_proteins=60, n_drugs=40, n_ppi=400, n_targets=150,
                                       n_pairs=500, n_ddi=900, min_size=200, max_size=400)


def nppredictor_vectors():
    src = open(os.path.join(REF, 'main/Predictor/NpPredictor.py')).read()
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'NpPredictor'][0]
    keep = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ('_predictEdges', '_getSampledPredictions', '_sigmoid')]
    mod = ast.Module(body=[ast.ClassDef(name='NpPredictor', bases=[], keywords=[], body=keep, decorator_list=[])], type_ignores=[])
    ast.fix_missing_locations(mod)

    class Holder:
        pass

    holder = Holder()
    ns = {'np': np, 'predsInfoHolder': holder}
    exec(compile(mod, 'NpPredictor.py', 'exec'), ns)
    predictor = ns['NpPredictor']()
    R = np.load(os.path.join(REF, 'ndarray-dumpGlobalRelations.npy'))
    D = np.load(os.path.join(REF, 'ndarray-dumpEmbeddingImportance.npyz.npz'))['arr_0']
    rng = np.random.RandomState(7)
    Z = rng.uniform(-0.5, 0.5, size=(40, 32)).astype(np.float32)
    edges = np.stack([rng.randint(0, 40, 64), rng.randint(0, 40, 64)], axis=1)
    holder.globalInteraction, holder.embeddings = R, Z
    out = {'R': R, 'D': D, 'Z': Z, 'edges': edges}
    for k in range(D.shape[0]):
        out['pred%d' % k] = predictor._predictEdges(D[k], edges, 1)
    return out


def main():
    tiny = run_reference_iterator(tiny_graph(), seed=5, batch_size=32, val_test_size=0.1, epoch_seed=6)
    np.savez_compressed(os.path.join(HERE, 'iterator_tiny.npz'), **tiny)
    digests = {}
    for name, inputs, bs, vf in [('toy', datasets.toy_graph(), 512, 0.05), ('poly', datasets.polypharmacy_graph(), 512, 0.05)]:
        arrays = run_reference_iterator(inputs, seed=0, batch_size=bs, val_test_size=vf, epoch_seed=1)
        if name == 'poly':  # one epoch is 144k batches: keep the first 2000
            arrays['epoch/relation'] = arrays['epoch/relation'][:2000]
            arrays['epoch/batches'] = arrays['epoch/batches'][:2000]
        digests[name] = {k: digest(v) for k, v in arrays.items()} if name == 'toy' else digest_by_kind(arrays)
        print(name, len(arrays), 'arrays,', len(arrays['epoch/relation']), 'steps')
    json.dump(digests, open(os.path.join(HERE, 'iterator_digests.json'), 'w'), indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, 'nppredictor.npz'), **nppredictor_vectors())
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == '__main__':
    main()
