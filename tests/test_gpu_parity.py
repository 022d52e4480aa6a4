"""Parity of the CUDA path (through the C ABI, via decagon_b200.engine) against the CPU oracle.

Tolerances (BASELINE.json north_star): indices bit-exact; embeddings, logits, losses and
gradients within rel-err 1e-5 of the float64 oracle, in BOTH metrics: max|a-b| / max|b| and
||a-b||_2 / ||b||_2 per tensor (common.assert_close).
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import common
from common import Case, assert_close, rel_err
from decagon_b200 import _lib, datasets
from oracle import decagon_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5
SEED = 42


@pytest.fixture(scope='module')
def toy():
    c = Case(datasets.toy_graph())
    c.eng = c.engine()
    return c


@pytest.fixture(scope='module')
def toy_mixed():
    c = Case(datasets.toy_graph(common.MIXED_DECODERS))
    c.eng = c.engine()
    return c


@pytest.fixture(scope='module')
def mini():
    c = Case(common.mini_poly(), batch_size=64)
    c.eng = c.engine()
    return c


def check_forward(c, eng, rate, step):
    masks = O.masks_for(c.graph, rate, step, SEED)
    Z, cache = O.encoder_forward(c.graph, c.p64, rate, masks)
    eng.forward(rate, SEED, step)
    for t in Z:
        assert_close(eng.hidden1_of(t), cache['H'][t], TOL, ('hidden1', t))
        assert_close(eng.embeddings(t), Z[t], TOL, ('embeddings', t))
    for gi, g in enumerate(c.graph.groups):
        assert_close(eng.tensor(_lib.TENSOR_LAYER1_GROUP, gi), cache['Y1'][g], TOL, ('layer1', g))
        assert_close(eng.tensor(_lib.TENSOR_LAYER2_GROUP, gi), cache['Y2'][g], TOL, ('layer2', g))
    return Z


def relu_pattern(c, eng, H_ref):
    """The device's ReLU activation pattern (hidden1 > 0), after checking that it differs from the oracle's only
    where the pre-activation is zero to rounding: ReLU has no derivative there and float32 / float64 may land on
    different sides (seen at the config-#3 shape: 1.2 M pre-activations, a handful within 1e-9 of zero)."""
    pattern = {}
    for t, h in H_ref.items():
        dev = eng.hidden1_of(t)
        pattern[t] = dev > 0
        flipped = pattern[t] != (h > 0)
        if flipped.any():
            tiny = 1e-6 * np.abs(h).max()
            assert np.abs(h[flipped]).max() <= tiny and np.abs(dev[flipped]).max() <= tiny, ('relu pattern', t)
            assert flipped.sum() <= 1e-4 * flipped.size
    return pattern


def check_grads(c, eng, r, batch, rate, step, loss_kind, negs=None):
    g, k = c.graph.flat[r]
    if negs is None:
        negs = O.sample_negatives(c.thresholds(r), len(batch), r, step, SEED)
    masks = O.masks_for(c.graph, rate, step, SEED)
    got = eng.train_step(r, batch, negatives=negs, loss=loss_kind, dropout=rate, seed=SEED, step=step,
                         apply_update=False)
    _, cache = O.encoder_forward(c.graph, c.p64, rate, masks)
    pattern = relu_pattern(c, eng, cache['H'])
    loss, pos, neg, grads, Z = O.train_step_grads(c.graph, c.p64, g, k, batch, negs, rate, masks, loss_kind,
                                                  relu_mask=pattern)
    gpos, gneg, gsamples = eng.last_batch_outputs(len(batch))
    assert np.array_equal(gsamples, negs)
    assert_close(gpos, pos, TOL, 'outputs')
    assert_close(gneg, neg, TOL, 'neg_outputs')
    assert abs(float(got) - loss) <= TOL * max(abs(loss), 1.0), (float(got), loss)
    eg = eng.get_grads()
    for name in grads:
        for gg in grads[name]:
            assert_close(eg[name][gg], grads[name][gg], TOL, ('grad', name, gg))
    return eg


def test_csr_bit_exact(toy):
    """Device CSR == scipy's canonical CSR of the reference tuple (indices bit-exact, values =
    float32 cast of the float64 normalisation)."""
    for r, (g, k) in enumerate(toy.graph.flat):
        coords, values, shape = toy.it.adj_train[g][k]
        ref = sp.csr_matrix((np.asarray(values).astype(np.float32), (coords[:, 0], coords[:, 1])), shape=shape)
        ref.sort_indices()
        rowptr, col, val = toy.eng.get_csr(r)
        assert np.array_equal(rowptr, ref.indptr) and np.array_equal(col, ref.indices)
        assert np.array_equal(val, ref.data)


@pytest.mark.parametrize('rate', [0.0, 0.1])
def test_forward_toy(toy, rate):
    check_forward(toy, toy.eng, rate, step=3)


@pytest.mark.parametrize('rate', [0.0, 0.1])
def test_forward_mini_staged(mini, rate):
    check_forward(mini, mini.eng, rate, step=1)


@pytest.mark.parametrize('loss_kind', ['hinge', 'xent'])
@pytest.mark.parametrize('rate', [0.0, 0.1])
def test_train_step_grads_toy(toy, loss_kind, rate):
    for step, (r, batch) in enumerate(toy.batches(4)):
        check_grads(toy, toy.eng, r, batch, rate, step, loss_kind)


@pytest.mark.parametrize('loss_kind', ['hinge', 'xent'])
def test_train_step_grads_all_decoders(toy_mixed, loss_kind):
    for step, (r, batch) in enumerate(toy_mixed.batches(4)):
        check_grads(toy_mixed, toy_mixed.eng, r, batch, 0.1, step, loss_kind)


@pytest.mark.parametrize('kind', ['innerproduct', 'distmult', 'bilinear', 'dedicom'])
def test_each_decoder_everywhere(kind):
    """Config #2: every group set to the same decoder kind."""
    c = Case(datasets.toy_graph({g: kind for g in datasets.DEFAULT_DECODERS}))
    eng = c.engine()
    for step, (r, batch) in enumerate(c.batches(4)):
        check_grads(c, eng, r, batch, 0.0, step, 'hinge')
    eng.close()


def test_train_step_grads_mini_staged(mini):
    for step, (r, batch) in enumerate(mini.batches(8)):
        check_grads(mini, mini.eng, r, batch, 0.1, step, 'hinge')


def test_staged_matches_gather():
    """The staged (shared-memory) SpMM and the gather SpMM are two code paths for one result."""
    c = Case(common.mini_poly(seed=3), batch_size=64)
    os.environ['DGN_DISABLE_STAGED'] = '1'
    try:
        gather = c.engine()
    finally:
        del os.environ['DGN_DISABLE_STAGED']
    staged = c.engine()
    for eng in (gather, staged):
        eng.forward(0.1, SEED, 0)
    for t in c.graph.n_nodes:
        assert rel_err(staged.embeddings(t), gather.embeddings(t)) <= TOL
    check_forward(c, gather, 0.1, 0)
    r, batch = c.batches(1)[0]
    check_grads(c, gather, r, batch, 0.1, 0, 'hinge')
    gather.close()
    staged.close()


def test_negative_sampler_bit_exact(toy):
    for step, (r, batch) in enumerate(toy.batches(6)):
        toy.eng.train_step(r, batch, negatives=None, seed=SEED, step=step, apply_update=False)
        _, _, samples = toy.eng.last_batch_outputs(len(batch))
        assert np.array_equal(samples, O.sample_negatives(toy.thresholds(r), len(batch), r, step, SEED))


def test_adam_update_tf1():
    """Five optimizer steps: the parameters must follow TF-1.8 ApplyAdam applied to the
    gradients the device itself produced (isolates the update rule and the beta-power
    schedule); every variable moves, also those with zero gradient."""
    c = Case(datasets.toy_graph())
    eng = c.engine()
    eng.reset_optimizer()
    p = O.cast_params(eng.get_params(), np.float32)
    adam = O.AdamTF1(p, lr=1e-3)
    for step, (r, batch) in enumerate(c.batches(5)):
        eng.train_step(r, batch, negatives=None, seed=SEED, step=step, dropout=0.1, apply_update=True)
        grads = eng.get_grads()
        adam.apply(p, grads)
        now = eng.get_params()
        for name in p:
            for g in p[name]:
                assert np.abs(now[name][g] - p[name][g]).max() <= 2e-7, (step, name, g)
    eng.close()


def test_multi_step_training_matches_oracle():
    """Ten full training steps (forward, decode, backward, Adam) on both sides with identical
    batches, negatives and dropout masks: losses agree within 1e-4 and the loss goes down."""
    c = Case(datasets.toy_graph())
    eng = c.engine()
    eng.reset_optimizer()
    p = O.cast_params(c.p32, np.float64)
    adam = O.AdamTF1(p, lr=1e-3)
    losses_dev, losses_ref = [], []
    for step, (r, batch) in enumerate(c.batches(10)):
        g, k = c.graph.flat[r]
        negs = O.sample_negatives(c.thresholds(r), len(batch), r, step, SEED)
        masks = O.masks_for(c.graph, 0.1, step, SEED)
        loss, _, _, grads, _ = O.train_step_grads(c.graph, p, g, k, batch, negs, 0.1, masks, 'hinge')
        adam.apply(p, grads)
        losses_ref.append(loss)
        losses_dev.append(float(eng.train_step(r, batch, negatives=None, seed=SEED, step=step, dropout=0.1)))
    assert rel_err(losses_dev, losses_ref) <= 1e-4, (losses_dev, losses_ref)
    eng.close()


def test_predictions(toy_mixed):
    """optimizer.predictions for one relation of every decoder kind + sampled sigmoid scores."""
    c, eng = toy_mixed, toy_mixed.eng
    Z = check_forward(c, eng, 0.0, 0)
    rng = np.random.RandomState(0)
    for r, (g, k) in enumerate(c.graph.flat):
        ref = O.predict_all_pairs(c.graph, c.p64, Z, g, k)
        assert rel_err(eng.predict(r), ref) <= TOL, (g, k)
        edges = np.stack([rng.randint(0, ref.shape[0], 200), rng.randint(0, ref.shape[1], 200)], axis=1)
        assert rel_err(eng.predict_edges(r, edges), O.sampled_scores(ref, edges)) <= TOL
        glb, loc = O.relation_matrices(c.graph, c.p64, g, k)
        eglb, eloc = eng.relation_matrices(r)
        assert rel_err(eglb, glb) <= 1e-7 and rel_err(eloc, loc) <= 1e-7


def test_predict_all_relations_tensor_cores(mini, toy_mixed):
    """BASELINE config #4 in small: every relation of a group in one call (evaluateAll,
    DecagonAccuracyEvaluator.py:57-91).  The tcgen05 (3 x TF32) kernel against the float64 oracle and against
    the CUDA-core kernel, on ragged tiles (97 and 400 / 500 nodes are not multiples of the 128-wide tiles)."""
    import torch
    for c in (mini, toy_mixed):
        eng = c.eng
        Z = check_forward(c, eng, 0.0, 0)
        for g in c.graph.groups:
            K = c.graph.K[g]
            r0 = eng.flat_index[(g, 0)]
            n_i, n_j = c.graph.n_nodes[g[0]], c.graph.n_nodes[g[1]]
            outs = {}
            for mode in ('0', '1'):
                os.environ['DGN_PREDICT_FFMA'] = mode
                try:
                    buf = torch.full((K, n_i, n_j), float('nan'), dtype=torch.float32, device='cuda')
                    eng.predict_relations_dev(r0, K, buf.data_ptr())
                    eng.sync()
                    outs[mode] = buf.cpu().numpy()
                finally:
                    del os.environ['DGN_PREDICT_FFMA']
            for k in range(K):
                ref = O.predict_all_pairs(c.graph, c.p64, Z, g, k)
                assert rel_err(outs['0'][k], ref) <= TOL, (g, k)
                assert rel_err(outs['1'][k], ref) <= TOL, (g, k)


def test_param_roundtrip_and_errors(toy):
    eng = toy.eng
    back = eng.get_params()
    for name in toy.p32:
        for g in toy.p32[name]:
            assert np.array_equal(back[name][g], toy.p32[name][g])
    with pytest.raises(ValueError):
        eng.train_step(0, np.array([[0, 10 ** 6]], dtype=np.int32))
    with pytest.raises(ValueError):
        eng.train_step(10 ** 4, np.zeros((4, 2), dtype=np.int32))
    with pytest.raises(ValueError):
        eng.set_param(_lib.PARAM_W2, (0, 0), 0, np.zeros((3, 3), dtype=np.float32))


def test_hidden_sizes():
    """hidden1 = 32 and 128 exercise the 1- and 4-panel kernels."""
    for h1 in (32, 128):
        c = Case(common.mini_poly(n_types=10, seed=h1), batch_size=64, hidden1=h1)
        eng = c.engine()
        check_forward(c, eng, 0.1, 0)
        r, batch = c.batches(1)[0]
        check_grads(c, eng, r, batch, 0.1, 0, 'hinge')
        eng.close()


def test_diverged_decoder_reports_nan_loss():
    """dZ is scattered on 2^-40 fixed-point integers (sums independent of the order of the atomics, DESIGN.md
    section 4); a contribution of 2^22 or more does not fit.  The step then reports a NaN loss -- what the
    reference's float arithmetic shows for a diverged model -- instead of a finite loss over wrong gradients."""
    c = Case(datasets.toy_graph())
    eng = c.engine()
    r, batch = c.batches(1)[0]
    g, k = c.graph.flat[r]
    assert np.isfinite(float(eng.train_step(r, batch, negatives=None, seed=SEED, step=0, apply_update=False)))
    eng.set_param(_lib.PARAM_DEC_LOCAL, g, None, c.p32['D'][g] * np.float32(1e8))
    assert np.isnan(float(eng.train_step(r, batch, negatives=None, seed=SEED, step=0, apply_update=False)))
    eng.set_param(_lib.PARAM_DEC_LOCAL, g, None, c.p32['D'][g])
    assert np.isfinite(float(eng.train_step(r, batch, negatives=None, seed=SEED, step=1, apply_update=False)))
    eng.close()


@pytest.mark.parametrize('h1,h2', [(64, 16), (48, 5), (100, 20)])
def test_any_hidden_size(h1, h2):
    """``model.py:68,80`` take any ``FLAGS.hidden1`` / ``FLAGS.hidden2``.  The device works on 32 / 64 / 128 hidden
    and 32 embedding columns; smaller sizes are zero-padded at the C ABI (W1 / W2 columns, W2 rows, decoder rows /
    columns), which is the caller's model term by term: every tensor crosses the boundary in the caller's shape and
    matches the oracle of that shape -- all four decoder kinds, both losses, all-pairs scores, and a multi-step
    Adam run (padding that leaked into the parameters would show up in the later losses)."""
    c = Case(datasets.toy_graph(common.MIXED_DECODERS), hidden1=h1, hidden2=h2)
    eng = c.engine()
    assert eng.n_params() == sum(v.size for name in c.p32 for v in c.p32[name].values())
    back = eng.get_params()
    for name in c.p32:
        for g in c.p32[name]:
            assert back[name][g].shape == c.p32[name][g].shape and np.array_equal(back[name][g], c.p32[name][g])
    Z = check_forward(c, eng, 0.1, 0)
    assert all(eng.embeddings(t).shape == (c.graph.n_nodes[t], h2) for t in Z)
    assert all(eng.hidden1_of(t).shape == (c.graph.n_nodes[t], h1) for t in Z)
    for step, g in enumerate(c.graph.groups):
        batch = np.asarray(c.it.train_edges[g][0][:256], dtype=np.int32)
        check_grads(c, eng, eng.flat_index[(g, 0)], batch, 0.1, step, 'hinge' if step % 2 == 0 else 'xent')
    Z = check_forward(c, eng, 0.0, 0)
    for g in c.graph.groups:
        r = eng.flat_index[(g, 0)]
        assert rel_err(eng.predict(r), O.predict_all_pairs(c.graph, c.p64, Z, g, 0)) <= TOL, g
        glb, loc = O.relation_matrices(c.graph, c.p64, g, 0)
        eglb, eloc = eng.relation_matrices(r)
        assert eglb.shape == (h2, h2) and rel_err(eglb, glb) <= 1e-7 and rel_err(eloc, loc) <= 1e-7
    # saved embeddings loaded back (NpPredictor's use case) are hidden2-wide too
    rng = np.random.RandomState(h2)
    Zs = {t: rng.randn(c.graph.n_nodes[t], h2).astype(np.float32) for t in Z}
    for t in Zs:
        eng.set_embeddings(t, Zs[t])
        assert np.array_equal(eng.embeddings(t), Zs[t])
    g = c.graph.groups[-1]
    want = O.predict_all_pairs(c.graph, c.p64, {t: Zs[t].astype(np.float64) for t in Zs}, g, 0)
    assert rel_err(eng.predict(eng.flat_index[(g, 0)]), want) <= TOL
    eng.reset_optimizer()
    p = O.cast_params(c.p32, np.float64)
    adam = O.AdamTF1(p, lr=1e-3)
    losses_dev, losses_ref = [], []
    for step, (r, batch) in enumerate(c.batches(8)):
        g, k = c.graph.flat[r]
        negs = O.sample_negatives(c.thresholds(r), len(batch), r, step, SEED)
        masks = O.masks_for(c.graph, 0.1, step, SEED)
        loss, _, _, grads, _ = O.train_step_grads(c.graph, p, g, k, batch, negs, 0.1, masks, 'hinge')
        adam.apply(p, grads)
        losses_ref.append(loss)
        losses_dev.append(float(eng.train_step(r, batch, negatives=None, seed=SEED, step=step, dropout=0.1)))
    assert rel_err(losses_dev, losses_ref) <= 1e-4, (losses_dev, losses_ref)
    eng.close()


@pytest.mark.parametrize('h1', [32, 64])
def test_dense_layer2_tensor_cores_match_cuda_cores(h1):
    """project / dw2 / dh on tcgen05 (TF32 split, dense_tc.cu; the default) and on the CUDA cores
    (DGN_DENSE_FFMA=1, dense.cu): both against the float64 oracle, dropout on and off, on a graph whose node
    counts leave ragged 128-row tiles (300 and 97 rows)."""
    c = Case(common.mini_poly(n_types=9, seed=40 + h1), batch_size=64, hidden1=h1)
    os.environ['DGN_DENSE_FFMA'] = '1'
    try:
        ffma = c.engine()
    finally:
        del os.environ['DGN_DENSE_FFMA']
    tcore = c.engine()
    for rate in (0.0, 0.1):
        for eng in (ffma, tcore):
            check_forward(c, eng, rate, 0)
            for step, (r, batch) in enumerate(c.batches(3)):
                check_grads(c, eng, r, batch, rate, step, 'hinge')
        for t in c.graph.n_nodes:
            assert rel_err(tcore.embeddings(t), ffma.embeddings(t)) <= TOL
    ffma.close()
    tcore.close()


@pytest.mark.parametrize('graph_kind', ['toy', 'mini'])
def test_general_sparse_features(graph_kind):
    """Non-identity node features (the reference's public data gives drugs multi-hot side-effect
    features, DecagonPublicDataNodeFeaturesBuilder.py:34-51): layer 1 is X_j W1_k with dropout on the feature
    non-zeros (layers.py:23-31, 89).  Forward, every gradient and an Adam step against the float64 oracle, on
    the gather path (toy) and on the staged path (mini: many small relations)."""
    if graph_kind == 'toy':
        feats = {1: datasets.multi_hot_features(400, 150, per_row=6, seed=3)}
        c = Case(datasets.toy_graph(features=feats))
    else:
        feats = {1: datasets.multi_hot_features(97, 213, per_row=9, seed=4),
                 0: datasets.multi_hot_features(300, 64, per_row=3, seed=5)}
        c = Case(common.mini_poly(n_types=10, seed=9, features=feats), batch_size=64)
    assert c.inputs.num_feat[1] == feats[1].shape[1] and c.inputs.nonzero_feat[1] == feats[1].nnz
    eng = c.engine()
    assert eng.get_param(_lib.PARAM_W1, (1, 1), 0).shape == (feats[1].shape[1], 64)
    for rate in (0.0, 0.1):
        check_forward(c, eng, rate, 2)
        for step, (r, batch) in enumerate(c.batches(4)):
            check_grads(c, eng, r, batch, rate, step, 'hinge')
    # optimizer steps (no fused Adam on this path): TF-1.8 ApplyAdam applied to the gradients the device produced
    eng.reset_optimizer()
    p = O.cast_params(eng.get_params(), np.float32)
    adam = O.AdamTF1(p, lr=1e-3)
    for step, (r, batch) in enumerate(c.batches(3)):
        eng.train_step(r, batch, negatives=None, seed=SEED, step=step, dropout=0.1, apply_update=True)
        adam.apply(p, eng.get_grads())
        now = eng.get_params()
        for name in p:
            for gg in p[name]:
                assert np.abs(now[name][gg] - p[name][gg]).max() <= 2e-7, (step, name, gg)
    eng.close()


@pytest.mark.parametrize('env', [{'DGN_SINGLE_STREAM': '1'}, {'DGN_PROJECT_SS': '1'}, {'DGN_FUSE_ADAM': '0'},
                                 {'DGN_DISABLE_TSTAGED': '1'}, {'DGN_MASK_CTAS': '1'}, {'DGN_CUDA_GRAPH': '0'},
                                 {'DGN_SIDE_LANES': '4'}, {'DGN_GATHER_ROWSUMS': '1'}, {'DGN_MASK_AHEAD': '1'}, {'DGN_SYNC_LOSS': '1'},
                                 {'DGN_BWD_ROW_ORDER': 'length'}, {'DGN_BWD_ROW_ORDER': 'address'}])
def test_alternate_code_paths(env):
    """The switches that select the non-default kernels / schedules (one stream instead of the lanes + mask stream,
    the shared-memory projection instead of the tensor-memory one, unfused Adam, gather-path backward) give the
    same results: forward, gradients and two optimizer steps against the oracle on the staged mini graph."""
    c = Case(common.mini_poly(n_types=9, seed=77), batch_size=64)
    os.environ.update(env)
    try:
        eng = c.engine()
        check_forward(c, eng, 0.1, 1)
        for step, (r, batch) in enumerate(c.batches(2)):
            check_grads(c, eng, r, batch, 0.1, step, 'hinge')
        eng.reset_optimizer()
        p = O.cast_params(eng.get_params(), np.float64)
        adam = O.AdamTF1(p, lr=1e-3)
        for step, (r, batch) in enumerate(c.batches(2)):
            g, k = c.graph.flat[r]
            negs = O.sample_negatives(c.thresholds(r), len(batch), r, step, SEED)
            masks = O.masks_for(c.graph, 0.1, step, SEED)
            _, _, _, grads, _ = O.train_step_grads(c.graph, p, g, k, batch, negs, 0.1, masks, 'hinge')
            adam.apply(p, grads)
            eng.train_step(r, batch, negatives=negs, seed=SEED, step=step, dropout=0.1, apply_update=True)
        now = eng.get_params()
        for name in ('W1', 'W2', 'R', 'D'):
            for gg in now[name]:
                assert rel_err(now[name][gg], p[name][gg]) <= 1e-4, (env, name, gg)
        eng.close()
    finally:
        for k in env:
            del os.environ[k]


# --------------------------------------------------------------------------- fused Adam
def _same(a, b):
    return all(np.array_equal(a[n][g], b[n][g]) for n in a for g in a[n])


def test_fused_adam_updates_w1_like_tf1():
    """The layer-1 backward of the many-relation group applies TF-1.8 ApplyAdam inside spmm_tstaged_kernel (95 % of all
    parameters at the polypharmacy shape; the gradient never reaches HBM).  Three steps on a staged graph: the step
    is run once with apply_update = 0 (gradients materialised and checked against the float64 oracle), then again
    with the update; EVERY variable -- W1 included -- must equal AdamTF1 (oracle) applied to those gradients within
    2e-7.  The unfused path (DGN_FUSE_ADAM=0) and the keep-gradients path (what fetching [opt_op, grads_vars]
    uses, optimizer.py:111-114) must give the same parameters bit for bit."""
    c = Case(common.mini_poly(n_types=11, seed=21), batch_size=64)
    fused = c.engine()
    keep = c.engine()
    keep.keep_gradients(True)
    os.environ['DGN_FUSE_ADAM'] = '0'
    try:
        unfused = c.engine()
    finally:
        del os.environ['DGN_FUSE_ADAM']
    for e in (fused, keep, unfused):
        e.reset_optimizer()
    p = fused.get_params()
    adam = O.AdamTF1(p, lr=1e-3)
    for step, (r, batch) in enumerate(c.batches(3)):
        negs = O.sample_negatives(c.thresholds(r), len(batch), r, step, SEED)
        kw = dict(negatives=negs, seed=SEED, step=step, dropout=0.1)
        c.p64 = O.cast_params(p, np.float64)  # the oracle follows the device parameters
        grads = check_grads(c, fused, r, batch, 0.1, step, 'hinge', negs=negs)
        for e in (fused, keep, unfused):
            e.train_step(r, batch, apply_update=True, **kw)
        with pytest.raises(ValueError):      # the fused step never stored the W1 gradient of the staged group
            fused.get_grad(_lib.PARAM_W1, (1, 1))
        assert _same(keep.get_grads(), grads), 'keep_gradients: gradients differ from the apply_update=0 run'
        adam.apply(p, grads)
        now = fused.get_params()
        for name in p:
            for g in p[name]:
                assert np.abs(now[name][g] - p[name][g]).max() <= 2e-7, (step, name, g)
        assert _same(now, unfused.get_params()), 'fused and unfused Adam differ'
        assert _same(now, keep.get_params()), 'keep-gradients path differs'
        for name in p:                       # continue from the device state: one update per comparison
            for g in p[name]:
                p[name][g][...] = now[name][g]
    assert not np.array_equal(now['W1'][(1, 1)], c.p32['W1'][(1, 1)])
    for e in (fused, keep, unfused):
        e.close()


# --------------------------------------------------------------------------- config #3 shape
@pytest.fixture(scope='module')
def poly():
    c = Case(datasets.polypharmacy_graph())
    c.eng = c.engine()
    yield c
    c.eng.close()


def test_config3_forward_and_grads(poly):
    """BASELINE config #3 as bench.py times it: 19 085 proteins, 645 drugs, 1932 relation matrices, dropout 0.1.
    These are the kernel instantiations of the headline numbers (spmm_staged3_kernel<6>, spmm_tstaged_kernel<1|2>
    at 1928 relations, project_ts / dw2_tc / dh_tc with 5 full row tiles + 1 ragged) against the float64 oracle
    (layers.py:85-118, optimizer.py:63-127): hidden1, embeddings, per-group layer outputs, scores, loss and EVERY
    gradient (all 83 M layer-1 weights included) at 1e-5 in both metrics."""
    c, eng = poly, poly.eng
    check_forward(c, eng, 0.1, step=2)
    # one batch of a drug-drug relation (dedicom) and one of the protein-protein group (bilinear)
    seen = set()
    for step, (r, batch) in enumerate(c.batches(4)):
        g, _ = c.graph.flat[r]
        if g in seen or g not in ((1, 1), (0, 0)):
            continue
        seen.add(g)
        check_grads(c, eng, r, batch, 0.1, step, 'hinge')
    assert (1, 1) in seen


def test_config3_fused_adam(poly):
    """One optimizer step at the config-#3 shape: the fused update of W1 (1928 x 645 x 64) against AdamTF1 applied
    to the gradients of the same step."""
    c, eng = poly, poly.eng
    eng.set_params(c.p32)
    eng.reset_optimizer()
    r, batch = next((r, b) for r, b in c.batches(4) if c.graph.flat[r][0] == (1, 1))
    negs = O.sample_negatives(c.thresholds(r), len(batch), r, 7, SEED)
    kw = dict(negatives=negs, seed=SEED, step=7, dropout=0.1)
    eng.train_step(r, batch, apply_update=False, **kw)
    grads = eng.get_grads()
    p = {name: {g: a.copy() for g, a in d.items()} for name, d in c.p32.items()}  # cast_params would alias c.p32
    O.AdamTF1(p, lr=1e-3).apply(p, grads)
    eng.train_step(r, batch, apply_update=True, **kw)
    now = eng.get_params()
    for name in p:
        for g in p[name]:
            assert np.abs(now[name][g] - p[name][g]).max() <= 2e-7, (name, g)
    assert np.abs(now['W1'][(1, 1)] - c.p32['W1'][(1, 1)]).max() > 1e-4   # the weights did move
    eng.set_params(c.p32)
    eng.reset_optimizer()


def test_config3_all_pairs(poly):
    """predict_tc_kernel at its bench shape (645 x 645, column tiles of 224): 8 relation matrices of the drug-drug
    group against optimizer.predict (optimizer.py:87-106) in float64."""
    import torch
    c, eng = poly, poly.eng
    eng.set_params(c.p32)
    Z, _ = O.encoder_forward(c.graph, c.p64, 0.0, None)
    eng.forward(0.0, SEED, 0)
    g = (1, 1)
    n = c.graph.n_nodes[1]
    K = c.graph.K[g]
    for k0 in (0, K // 2, K - 4):
        buf = torch.full((4, n, n), float('nan'), dtype=torch.float32, device='cuda')
        eng.predict_relations_dev(eng.flat_index[(g, k0)], 4, buf.data_ptr())
        eng.sync()
        out = buf.cpu().numpy()
        for q in range(4):
            assert_close(out[q], O.predict_all_pairs(c.graph, c.p64, Z, g, k0 + q), TOL, (g, k0 + q))


def test_cuda_graph_replay_is_bit_identical():
    """The training step replayed as ONE CUDA graph (default from the second step of a configuration on; per-step
    values -- batch, relation, decoder pointers, dropout stream words, Adam coefficients -- come from the device
    block StepDyn) against the same steps issued kernel by kernel (DGN_CUDA_GRAPH=0): losses, negatives and every
    parameter bit for bit over steps that change relation group, dropout rate and update mode."""
    c = Case(common.mini_poly(n_types=10, seed=31), batch_size=64)
    os.environ['DGN_MASK_AHEAD'] = '1'   # graph replay (default) + layer-2 keep words drawn one step ahead
    try:
        graphed = c.engine()
    finally:
        del os.environ['DGN_MASK_AHEAD']
    os.environ['DGN_CUDA_GRAPH'] = '0'
    try:
        direct = c.engine()
    finally:
        del os.environ['DGN_CUDA_GRAPH']
    for e in (graphed, direct):
        e.reset_optimizer()
    plan = [(0.1, True)] * 6 + [(0.0, True)] * 3 + [(0.1, False)] * 3 + [(0.1, True)] * 4
    for step, ((r, batch), (rate, update)) in enumerate(zip(c.batches(len(plan)), plan)):
        out = []
        for e in (graphed, direct):
            loss = e.train_step(r, batch, negatives=None, seed=SEED, step=step, dropout=rate, apply_update=update)
            out.append((np.float32(loss), e.last_batch_outputs(len(batch))))
        assert out[0][0] == out[1][0], (step, out[0][0], out[1][0])
        for a, b in zip(out[0][1], out[1][1]):
            assert np.array_equal(a, b), step
    assert _same(graphed.get_params(), direct.get_params())
    for t in c.graph.n_nodes:
        assert np.array_equal(graphed.embeddings(t), direct.embeddings(t))
    graphed.close()
    direct.close()
