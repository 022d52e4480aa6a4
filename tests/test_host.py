"""CPU tests (run with -m "not gpu"): the host half of the hot path against the reference's own
outputs (committed golden fixtures, tests/golden/make_golden.py), the oracle against those
fixtures and against torch autograd, and the C ABI surface of the shared library."""
import hashlib
import json
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

import common
from decagon_b200 import _lib, datasets
from decagon_b200.deep.minibatch import EdgeMinibatchIterator, normalize_adjacency
from decagon_b200.sparse import RelationCsrMatrix
from oracle import decagon_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
PLACEHOLDERS = {k: k for k in common.PLACEHOLDER_KEYS}
SPLITS = ['train_edges', 'val_edges', 'val_edges_false', 'test_edges', 'test_edges_false']


def iterator_arrays(inputs, seed, batch_size, val_test_size, epoch_seed, max_steps=None):
    np.random.seed(seed)
    it = EdgeMinibatchIterator(inputs.adj_mats, inputs.feat, inputs.edge_types, {}, batch_size=batch_size,
                               val_test_size=val_test_size)
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
    while not it.end() and (max_steps is None or len(seq) < max_steps):
        fd = it.next_minibatch_feed_dict(PLACEHOLDERS)
        seq.append(fd['batch_edge_type_idx'])
        batches.append(np.asarray(fd['batch']))
    arrays['epoch/relation'] = np.array(seq)
    arrays['epoch/batches'] = np.stack(batches)
    return arrays


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()


def tiny_graph():
    return datasets.polypharmacy_graph(n_types=3, seed=11, n_proteins=60, n_drugs=40, n_ppi=400, n_targets=150,
                                       n_pairs=500, n_ddi=900, min_size=200, max_size=400)


# ------------------------------------------------------------------ iterator vs the reference
def test_iterator_matches_reference_arrays():
    """Every split, normalised tuple (float64, bit for bit) and minibatch of one epoch equal what
    the reference's minibatch.py produced under the same np.random seeds."""
    gold = np.load(os.path.join(GOLDEN, 'iterator_tiny.npz'))
    mine = iterator_arrays(tiny_graph(), seed=5, batch_size=32, val_test_size=0.1, epoch_seed=6)
    assert set(gold.files) == set(mine)
    for key in gold.files:
        g, m = gold[key], mine[key]
        assert g.dtype == m.dtype and g.shape == m.shape, key
        assert np.array_equal(g, m), key


def test_iterator_digests_toy():
    gold = json.load(open(os.path.join(GOLDEN, 'iterator_digests.json')))['toy']
    mine = iterator_arrays(datasets.toy_graph(), seed=0, batch_size=512, val_test_size=0.05, epoch_seed=1)
    assert len(mine['epoch/relation']) == 492
    assert set(gold) == set(mine)
    for key, d in gold.items():
        assert digest(mine[key]) == d, key


def test_iterator_digests_polypharmacy_shape():
    """BASELINE config #3 at full size (1932 relation matrices, 21.6 M non-zeros): a checksum of
    checksums per array kind."""
    gold = json.load(open(os.path.join(GOLDEN, 'iterator_digests.json')))['poly']
    mine = iterator_arrays(datasets.polypharmacy_graph(), seed=0, batch_size=512, val_test_size=0.05, epoch_seed=1,
                           max_steps=2000)
    n_rel = len(mine['flat'])
    assert n_rel == 1932
    for key in ['flat', 'epoch/relation', 'epoch/batches']:
        assert digest(mine[key]) == gold[key], key
    for kind in SPLITS + ['adj_coords', 'adj_values', 'adj_shape']:
        h = hashlib.sha256()
        for r in range(n_rel):
            h.update(digest(mine['r%d/%s' % (r, kind)]).encode())
        assert h.hexdigest() == gold[kind], kind


def test_precomputed_drug_drug_edges_and_twins():
    """The drug_drug_test_edges path (minibatch.py:235-253) and the transposed-twin mirroring
    (minibatch.py:137-172)."""
    g = datasets.toy_graph()
    rng = np.random.RandomState(1)
    dd = {k: {'positive': rng.randint(0, 400, (20, 2)), 'negative': rng.randint(0, 400, (20, 2))} for k in range(3)}
    np.random.seed(3)
    it = EdgeMinibatchIterator(g.adj_mats, g.feat, g.edge_types, dd, batch_size=512, val_test_size=0.05)
    for k in range(3):
        assert np.array_equal(it.val_edges[1, 1][k], dd[k]['positive'])
        assert len(it.train_edges[1, 1][k]) == g.adj_mats[1, 1][k].nnz  # nothing masked out
        c, v, s = it.adj_train[1, 1][k]
        ct, vt, st = it.adj_train[1, 1][k + 3]
        assert np.array_equal(np.flip(c, axis=1), ct) and vt is v and st == (s[1], s[0])
        assert np.array_equal(np.flip(it.train_edges[1, 1][k], axis=1), it.train_edges[1, 1][k + 3])
    c01, v01, _ = it.adj_train[0, 1][0]
    c10, v10, _ = it.adj_train[1, 0][0]
    assert np.array_equal(np.flip(c01, axis=1), c10) and v10 is v01


def test_normalize_adjacency_formula():
    """square: D^-1/2 (A+I)^T D^-1/2, rect: Dr^-1/2 A Dc^-1/2 (minibatch.py:80-93)."""
    rng = np.random.RandomState(2)
    a = sp.csr_matrix((rng.rand(30, 30) < 0.2).astype(float))
    coords, values, shape = normalize_adjacency(a)
    dense = sp.csr_matrix((values, (coords[:, 0], coords[:, 1])), shape=shape).toarray()
    ah = a.toarray() + np.eye(30)
    d = ah.sum(1) ** -0.5
    assert np.allclose(dense, (d[:, None] * ah.T * d[None, :]), rtol=1e-14, atol=0)
    b = sp.csr_matrix((rng.rand(20, 35) < 0.2).astype(float))
    b[3, :] = 0
    b.eliminate_zeros()
    coords, values, shape = normalize_adjacency(b)
    dense = sp.csr_matrix((values, (coords[:, 0], coords[:, 1])), shape=shape).toarray()
    bd = b.toarray()
    with np.errstate(divide='ignore'):
        dr, dc = np.nan_to_num(bd.sum(1) ** -0.5), np.nan_to_num(bd.sum(0) ** -0.5)
    assert np.allclose(dense, dr[:, None] * bd * dc[None, :], rtol=1e-14, atol=0)
    assert np.all(np.diff(coords[:, 0]) >= 0)  # rect tuples are row-major


def test_live_reference_when_available():
    """In the build container the reference is importable: compare on a fresh random graph."""
    if not os.path.isdir('/root/reference/decagon'):
        pytest.skip('reference tree not present on this machine')
    import contextlib
    import importlib.util
    import io
    import sys
    sys.path.insert(0, os.path.join(GOLDEN))
    sys.path.insert(0, '/root/reference')
    import make_golden
    inputs = common.mini_poly(n_types=5, seed=9)
    ref = make_golden.run_reference_iterator(inputs, seed=21, batch_size=64, val_test_size=0.07, epoch_seed=22)
    mine = iterator_arrays(inputs, seed=21, batch_size=64, val_test_size=0.07, epoch_seed=22)
    assert set(ref) == set(mine)
    for key in ref:
        assert ref[key].dtype == mine[key].dtype and np.array_equal(ref[key], mine[key]), key


# ------------------------------------------------------------------ oracle
def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = O.philox4x32(*[np.array([c]) for c in ctr], key[0], key[1])
        assert tuple(int(x[0]) for x in got) == want


def test_oracle_dedicom_matches_nppredictor():
    """The oracle's all-pairs DEDICOM score against the reference's numpy statement
    (NpPredictor._predictEdges run from its source on the reference's dumped R and D_k)."""
    gold = np.load(os.path.join(GOLDEN, 'nppredictor.npz'))
    R, D, Z, edges = gold['R'], gold['D'], gold['Z'], gold['edges']
    n = Z.shape[0]
    adj = {(0, 0): [sp.identity(n, format='csr') for _ in range(D.shape[0])]}
    graph = O.Graph({0: n}, [(0, 0)], adj, {0: sp.identity(n, format='csr')}, {(0, 0): 'dedicom'})
    p = {'W1': {}, 'W2': {(0, 0): np.zeros((D.shape[0], 64, 32))}, 'R': {(0, 0): R.astype(np.float64)},
         'D': {(0, 0): np.stack([np.diag(D[k]) for k in range(D.shape[0])]).astype(np.float64)}}
    for k in range(D.shape[0]):
        assert np.count_nonzero(D[k] - np.diag(np.diag(D[k]))) == 0  # latent_varies are diagonal (model.py:132)
        pred = O.predict_all_pairs(graph, p, {0: Z.astype(np.float64)}, (0, 0), k)
        want = gold['pred%d' % k]
        assert np.array_equal(want[:, :2], edges)
        assert np.abs(O.sampled_scores(pred, edges) - want[:, 2]).max() <= 2e-6  # reference ran in float32


def test_oracle_backward_matches_autograd():
    import torch
    from oracle.torch_ref import TorchDecagon
    c = common.Case(datasets.toy_graph(common.MIXED_DECODERS))
    for step, (r, batch) in enumerate(c.batches(4)):
        g, k = c.graph.flat[r]
        negs = O.sample_negatives(c.thresholds(r), len(batch), r, step, 7)
        masks = O.masks_for(c.graph, 0.1, step, 7)
        for kind in ('hinge', 'xent'):
            loss, pos, neg, grads, _ = O.train_step_grads(c.graph, c.p64, g, k, batch, negs, 0.1, masks, kind)
            T = TorchDecagon(c.graph, c.p64, torch.float64)
            cost, tpos, tneg, _ = T.grads(g, k, batch, negs, 0.1, masks, kind=kind)
            assert abs(loss - float(cost.detach())) <= 1e-10 * abs(loss)
            assert common.rel_err(pos, tpos.detach().numpy()) <= 1e-12
            tg = T.grad_dict()
            for name in grads:
                for gg in grads[name]:
                    assert common.rel_err(grads[name][gg], tg[name][gg]) <= 1e-10, (name, gg)


def test_oracle_float32_twin_is_close_to_float64():
    c = common.Case(datasets.toy_graph())
    r, batch = c.batches(1)[0]
    g, k = c.graph.flat[r]
    negs = O.sample_negatives(c.thresholds(r), len(batch), r, 0, 7)
    l64, _, _, g64, Z64 = O.train_step_grads(c.graph, c.p64, g, k, batch, negs)
    l32, _, _, g32, Z32 = O.train_step_grads(c.graph, O.cast_params(c.p32, np.float32), g, k, batch, negs)
    assert abs(l64 - l32) <= 1e-4 * abs(l64)
    for t in Z64:
        assert common.rel_err(Z32[t], Z64[t]) <= 1e-5


def test_adam_tf1_first_steps():
    """alpha_t = lr sqrt(1-b2^t)/(1-b1^t); first update of a constant gradient is ~lr * sign(g)."""
    p = {'W2': {(0, 0): np.zeros((1, 2, 2), dtype=np.float32)}}
    g = {'W2': {(0, 0): np.full((1, 2, 2), 0.5, dtype=np.float32)}}
    adam = O.AdamTF1(p, lr=1e-3)
    adam.apply(p, g)
    assert np.allclose(p['W2'][0, 0], -1e-3, rtol=1e-4)
    adam.apply(p, {'W2': {(0, 0): np.zeros((1, 2, 2), dtype=np.float32)}})
    assert np.all(p['W2'][0, 0] < -1e-3)  # zero gradient: the variable still moves (dense Adam)


def test_sampler_distribution():
    deg = np.array([0, 1, 16, 81, 0, 256], dtype=np.float64)
    thr = O.sampler_thresholds(deg)
    draws = O.sample_negatives(thr, 200000, 3, 0, 99)
    freq = np.bincount(draws, minlength=len(deg)) / 200000.0
    w = deg ** 0.75
    assert freq[0] == 0 and freq[4] == 0
    assert np.abs(freq - w / w.sum()).max() < 5e-3


# ------------------------------------------------------------------ C ABI surface
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'decagon_b200.h')).read()
    declared = set(re.findall(r'\b(dgn_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no prototypes found'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.dgn_version() >= 100


def test_csr_from_coo_host():
    rng = np.random.RandomState(0)
    a = sp.random(57, 91, 0.08, random_state=rng, format='coo')
    perm = rng.permutation(a.nnz)
    rowptr, col, val = _lib.csr_from_coo(57, 91, a.row[perm], a.col[perm], a.data[perm])
    ref = a.tocsr()
    ref.sort_indices()
    assert np.array_equal(rowptr, ref.indptr) and np.array_equal(col, ref.indices)
    assert np.array_equal(val, ref.data.astype(np.float32))
    rowptr, col, val = _lib.csr_from_coo(4, 4, [], [], [])
    assert np.array_equal(rowptr, np.zeros(5, dtype=np.int32)) and len(col) == 0
    with pytest.raises(ValueError):
        _lib.csr_from_coo(3, 3, [5], [0], [1.0])


def test_sampler_thresholds_match_oracle():
    g = datasets.toy_graph()
    for t in (0, 1):
        for d in g.degrees[t]:
            assert np.array_equal(_lib.sampler_thresholds(d), O.sampler_thresholds(d))
    with pytest.raises(ValueError):
        _lib.sampler_thresholds(np.zeros(4))


def test_no_device_fails_loudly():
    if _lib.device_count() > 0:
        pytest.skip('a CUDA device is visible')
    from decagon_b200.engine import Engine
    g = datasets.toy_graph()
    with pytest.raises(_lib.DecagonB200Error, match='no CPU fallback'):
        Engine(g.n_nodes, g.num_feat, g.edge_types, g.edge_type2decoder)


def test_hidden_sizes_accepted_at_the_boundary():
    """``model.py:68,80`` take any FLAGS.hidden1 / hidden2; the library accepts hidden1 <= 128 and hidden2 <= 32 (padded
    to its panel widths inside) and refuses anything larger with DGN_ERR_UNSUPPORTED -- checked before a device is
    looked for, so the refusal is observable here too."""
    from decagon_b200.engine import Engine
    g = datasets.toy_graph()
    for h1, h2 in ((129, 32), (64, 33), (0, 32), (64, 0)):
        with pytest.raises(NotImplementedError, match='not supported'):
            Engine(g.n_nodes, g.num_feat, g.edge_types, g.edge_type2decoder, hidden1=h1, hidden2=h2)
    if _lib.device_count() == 0:
        for h1, h2 in ((100, 20), (1, 1), (128, 32)):
            with pytest.raises(_lib.DecagonB200Error, match='no CPU fallback'):  # past the size check
                Engine(g.n_nodes, g.num_feat, g.edge_types, g.edge_type2decoder, hidden1=h1, hidden2=h2)


def test_relation_matrix_types():
    m = RelationCsrMatrix(sp.identity(4, format='csr'))
    t = m.transpose(copy=True, setId=True)
    assert m.isTranspose and t.isTranspose and m.transposedMtxLink is t and t.transposedMtxLink is m
    assert m.id != t.id and m.tocoo().id == m.id and m.tocoo().tocsr().transposedMtxLink is t


def test_ndarray_io_predict_matches_nppredictor_golden():
    """ndarray_io.np_predict_edges restates NpPredictor._predictEdges (NpPredictor.py:304-319): same numbers as the
    reference's own source run on its dumped R / D_k (tests/golden/nppredictor.npz)."""
    from decagon_b200 import ndarray_io
    gold = np.load(os.path.join(GOLDEN, 'nppredictor.npz'))
    R, D, Z, edges = gold['R'], gold['D'], gold['Z'], gold['edges']
    for k in range(D.shape[0]):
        want = gold['pred%d' % k]
        got = ndarray_io.np_predict_edges(Z, D[k], R, edges)
        assert np.abs(got - want[:, 2]).max() <= 2e-6


def test_multi_hot_features_and_unmask_schedule():
    """Host helpers around the widened path: the multi-hot feature generator (shape of the reference's drug features,
    DecagonPublicDataNodeFeaturesBuilder.py:34-51) and RandomMaskingActiveLearner._updateMask's schedule (:166-171)."""
    from decagon_b200.active_learning import num_to_unmask
    x = datasets.multi_hot_features(97, 213, per_row=9, seed=4)
    assert x.shape == (97, 213) and x.nnz > 97 and set(np.unique(x.data)) == {1.0}
    assert np.all(np.diff(x.indptr) >= 1)                      # every node has at least one feature
    assert np.array_equal(x.toarray(), datasets.multi_hot_features(97, 213, per_row=9, seed=4).toarray())
    inputs = datasets.toy_graph(features={1: datasets.multi_hot_features(400, 150, per_row=6, seed=3)})
    assert inputs.num_feat == {0: 500, 1: 150} and inputs.nonzero_feat[1] == inputs.feat[1][1].sum()
    # 1 %, 1 %, 2 %, 4 %, ... of the data set, capped at 100 % in total
    sizes = [num_to_unmask(10000, i) for i in range(8)]
    assert sizes == [100, 100, 200, 400, 800, 1600, 3200, 3600] and sum(sizes) == 10000


def test_sparse_relation_masks_equal_the_dense_reference_formula():
    """RandomMaskingActiveLearner._updateMask / _applyMask (RandomMaskingActiveLearner.py:166-200) keep dense n x n
    masks and multiply them into mtx.toarray(); SparseRelationMasks keeps one bit per non-zero.  Same matrices."""
    import scipy.sparse as sp
    from decagon_b200.active_learning import SparseRelationMasks
    rng = np.random.RandomState(0)
    n = 60
    mats = {}
    for rel in (3, 7, 11):
        a = (rng.rand(n, n) < 0.08).astype(np.float64)
        a[5] = 0  # an empty row
        mats[rel] = sp.csr_matrix(np.maximum(a, a.T))
    grid = np.array([(rel, r, c) for rel in mats for r in range(n) for c in range(n)])
    dense_masks = {rel: np.zeros((n, n)) for rel in mats}
    sparse = SparseRelationMasks(mats)
    remaining = grid
    for round_ in range(4):
        pick = rng.choice(len(remaining), size=len(remaining) // 5, replace=False)
        for rel, r, c in remaining[pick]:               # the reference's loop (:173-174)
            dense_masks[rel][r, c] = 1
        sparse.unmask(remaining[pick])
        remaining = np.delete(remaining, pick, axis=0)
        got = sparse.apply()
        for rel in mats:
            want = sp.csr_matrix(np.multiply(dense_masks[rel], mats[rel].toarray()))   # _applyMask (:187-190)
            assert (got[rel] != want).nnz == 0 and got[rel].nnz == want.nnz
            assert np.array_equal(got[rel].indptr, want.indptr) and np.array_equal(got[rel].indices, want.indices)


def test_feed_dict_token_semantics():
    """minibatch.FeedDict: update_feed_dict stamps the token that lets Session skip the walk over every adjacency
    tuple; only replacing a sparse (adjacency / feature) entry clears it; plain dicts never carry one."""
    from decagon_b200 import tf_compat as tf
    from decagon_b200.deep.minibatch import FeedDict
    inputs = datasets.toy_graph()
    ph = {'batch': tf.placeholder(tf.int32), 'batch_edge_type_idx': tf.placeholder(tf.int32),
          'batch_row_edge_type': tf.placeholder(tf.int32), 'batch_col_edge_type': tf.placeholder(tf.int32),
          'dropout': tf.placeholder_with_default(0., shape=())}
    ph.update({'adj_mats_%d,%d,%d' % (i, j, k): tf.sparse_placeholder(tf.float32)
               for i, j in inputs.edge_types for k in range(inputs.edge_types[i, j])})
    ph.update({'feat_%d' % i: tf.sparse_placeholder(tf.float32) for i, _ in inputs.edge_types})
    np.random.seed(0)
    it = EdgeMinibatchIterator(inputs.adj_mats, inputs.feat, inputs.edge_types, {}, batch_size=512, val_test_size=0.05)
    fd = it.next_minibatch_feed_dict(ph)
    assert isinstance(fd, FeedDict) and isinstance(fd, dict) and fd.graph_token is None and len(fd) == 4
    fd = it.update_feed_dict(fd, 0.1, ph)
    assert fd.graph_token == (id(it), id(ph)) and len(fd) == 4 + 10 + 2 + 1
    assert fd[ph['adj_mats_1,1,3']] is it.adj_train[1, 1][3] and fd[ph['dropout']] == 0.1
    fd[ph['dropout']] = 0.0
    fd[ph['batch_edge_type_idx']] = 3
    assert fd.graph_token is not None
    fd[ph['feat_0']] = it.feat[0]
    assert fd.graph_token is None
    plain = it.update_feed_dict(dict(it.next_minibatch_feed_dict(ph)), 0.1, ph)
    assert type(plain) is dict and len(plain) == 17
    other = it.update_feed_dict(it.next_minibatch_feed_dict(ph), 0.1, ph)
    other.update({ph['dropout']: 0.5})
    assert other.graph_token is None


def test_choice_is_randint():
    """The schedule draws relation r = free[randint(0, len(free))]; the reference writes np.random.choice(free)
    (minibatch.py:292).  Same values, same generator state afterwards."""
    for n in (1, 2, 7, 1929):
        free = list(range(5, 5 + n))
        np.random.seed(9)
        a = [np.random.choice(free) for _ in range(50)]
        sa = np.random.get_state()
        np.random.seed(9)
        b = [free[np.random.randint(0, len(free))] for _ in range(50)]
        sb = np.random.get_state()
        assert a == b and sa[2] == sb[2] and np.array_equal(sa[1], sb[1])


def test_lazy_feed_dict_behaves_like_a_dict():
    from decagon_b200.deep.minibatch import FeedDict
    base = {('adj', k): (k, k) for k in range(50)}
    fd = FeedDict({'batch': 1, 'idx': 2})
    fd.attach(base, 'token')
    dict.__setitem__(fd, 'dropout', 0.1)
    assert fd['batch'] == 1 and ('adj', 3) in fd and fd.get(('adj', 4)) == (4, 4) and fd.get('nope', 7) == 7
    assert fd._pending is not None                      # nothing forced the merge yet
    assert fd[('adj', 7)] == (7, 7) and fd._pending is None   # a lookup miss merges
    fd2 = FeedDict({'batch': 1})
    fd2.attach(base, 'token')
    assert len(fd2) == 51 and len(dict(fd2)) == 51 and set(fd2.keys()) == set(base) | {'batch'}
    fd3 = FeedDict({'batch': 1, ('adj', 0): 'mine'})
    fd3.attach(base, 'token')
    assert dict(fd3)[('adj', 0)] == 'mine' and fd3.copy().graph_token == 'token'
    with pytest.raises(KeyError):
        fd3['missing']
    assert {**fd2}['batch'] == 1 and sorted(k for k in fd2 if k != 'batch') == sorted(base)
