"""The drop-in surface driven exactly like the reference's callers drive TF
(DecagonDataSet._getPlaceholdersDict, DecagonTrainableBuilder.build, DecagonTrainer.train,
DecagonAccuracyEvaluator._computePredictions), checked against the oracle."""
import numpy as np
import pytest

import common
from common import rel_err
from decagon_b200 import datasets
from decagon_b200 import tf_compat as tf
from decagon_b200.deep import inits
from decagon_b200.deep.minibatch import EdgeMinibatchIterator
from decagon_b200.deep.model import DecagonModel
from decagon_b200.deep.optimizer import DecagonOptimizer
from oracle import decagon_oracle as O

pytestmark = pytest.mark.gpu
SEED = 1234


def construct_placeholders(edge_types):
    """DecagonDataSet.py:84-120 / main.py:93-108 with tf -> tf_compat."""
    ph = {
        'batch': tf.placeholder(tf.int32, name='batch'),
        'batch_edge_type_idx': tf.placeholder(tf.int32, shape=(), name='batch_edge_type_idx'),
        'batch_row_edge_type': tf.placeholder(tf.int32, shape=(), name='batch_row_edge_type'),
        'batch_col_edge_type': tf.placeholder(tf.int32, shape=(), name='batch_col_edge_type'),
        'degrees': tf.placeholder(tf.int32),
        'dropout': tf.placeholder_with_default(0., shape=()),
    }
    ph.update({'adj_mats_%d,%d,%d' % (i, j, k): tf.sparse_placeholder(tf.float32)
               for i, j in edge_types for k in range(edge_types[i, j])})
    ph.update({'feat_%d' % i: tf.sparse_placeholder(tf.float32) for i, _ in edge_types})
    return ph


def build_trainable(inputs, batch_size=512):
    placeholders = construct_placeholders(inputs.edge_types)
    np.random.seed(0)
    minibatch = EdgeMinibatchIterator(inputs.adj_mats, inputs.feat, inputs.edge_types, {}, batch_size, 0.05)
    inits.set_seed(7)
    model = DecagonModel(placeholders, inputs.num_feat, inputs.nonzero_feat, inputs.edge_types,
                         inputs.edge_type2decoder)
    with tf.name_scope('optimizer'):
        opt = DecagonOptimizer(model.embeddings, model.latent_inters, model.latent_varies, inputs.degrees,
                               inputs.edge_types, inputs.edge_type2dim, placeholders, margin=0.1,
                               neg_sample_weights=1., batch_size=batch_size)
    return placeholders, minibatch, model, opt


def oracle_params(model):
    p = {'W1': {}, 'W2': {}, 'R': {}, 'D': {}}
    for g, K in model.edge_types.items():
        p['W1'][g] = np.stack([model.layer1[g].vars['weights_%d' % k].initial for k in range(K)])
        p['W2'][g] = np.stack([model.layer2[g].vars['weights_%d' % k].initial for k in range(K)])
        dec = model.edge_type2decoder[g]
        if dec.kind == 'dedicom':
            p['R'][g] = dec.vars['global_interaction'].initial
            p['D'][g] = np.stack([dec.vars['local_variation_%d' % k].initial for k in range(K)])
        elif dec.kind in ('distmult', 'bilinear'):
            p['D'][g] = np.stack([dec.vars['relation_%d' % k].initial for k in range(K)])
    return p


def test_training_loop_like_the_reference_trainer():
    inputs = datasets.toy_graph()
    placeholders, minibatch, model, opt = build_trainable(inputs)
    sess = tf.Session(seed=SEED)
    sess.run(tf.global_variables_initializer())

    graph = O.Graph.from_iterator(minibatch, inputs.edge_type2decoder)
    p = O.cast_params(oracle_params(model), np.float64)
    adam = O.AdamTF1(p, lr=tf.FLAGS.learning_rate)

    np.random.seed(1)
    minibatch.shuffle()
    losses, ref_losses = [], []
    for step in range(6):
        feed_dict = minibatch.next_minibatch_feed_dict(placeholders)
        feed_dict = minibatch.update_feed_dict(feed_dict, 0.1, placeholders)
        outs = sess.run([opt.opt_op, opt.cost, opt.batch_edge_type_idx, opt.neg_samples], feed_dict=feed_dict)
        assert outs[0] is None and outs[1].dtype == np.float32
        r = int(outs[2])
        assert r == feed_dict[placeholders['batch_edge_type_idx']]
        g, k = graph.flat[r]
        batch = feed_dict[placeholders['batch']]
        negs = O.sample_negatives(O.sampler_thresholds(inputs.degrees[g[0]][k]), len(batch), r, step, SEED)
        assert np.array_equal(outs[3], negs)
        masks = O.masks_for(graph, 0.1, step, SEED)
        loss, _, _, grads, _ = O.train_step_grads(graph, p, g, k, batch, negs, 0.1, masks, 'hinge')
        adam.apply(p, grads)
        losses.append(float(outs[1]))
        ref_losses.append(loss)
    assert rel_err(losses, ref_losses) <= 1e-4, (losses, ref_losses)

    # evaluator path (DecagonAccuracyEvaluator.py:115-149,188-194): dropout 0, relation (1,1,0)
    rel = (1, 1, 0)
    feed_dict[placeholders['dropout']] = 0
    feed_dict[placeholders['batch_edge_type_idx']] = minibatch.edge_type2idx[rel]
    feed_dict[placeholders['batch_row_edge_type']] = rel[0]
    feed_dict[placeholders['batch_col_edge_type']] = rel[1]
    pred = sess.run(opt.predictions, feed_dict=feed_dict)
    Z, _ = O.encoder_forward(graph, p)
    want = O.predict_all_pairs(graph, p, Z, (1, 1), 0)
    assert pred.shape == (400, 400) and pred.dtype == np.float32
    assert rel_err(pred, want) <= 1e-4  # parameters went through 6 float32 Adam steps

    # logger fetches (DecagonLogger.py:239-281)
    emb = sess.run(model.embeddings[1], feed_dict=feed_dict)
    assert rel_err(emb, Z[1]) <= 1e-4
    r = minibatch.edge_type2idx[rel]
    loc, glb = sess.run([model.latent_varies[r], model.latent_inters[r]], feed_dict=feed_dict)
    want_glb, want_loc = O.relation_matrices(graph, p, (1, 1), 0)
    assert rel_err(glb, want_glb) <= 1e-4 and rel_err(loc, want_loc) <= 1e-4
    assert np.count_nonzero(loc - np.diag(np.diag(loc))) == 0
    var = model.edge_type2decoder[1, 1].vars['global_interaction']
    assert rel_err(sess.run(var), p['R'][1, 1]) <= 1e-4
    assert set(model.vars) == {v.name for v in model._variables()} and len(model.vars) == 10 * 2 + 4 + 1 + 6


def test_variable_names_and_errors():
    inputs = datasets.toy_graph()
    placeholders, minibatch, model, opt = build_trainable(inputs)
    names = sorted(model.vars)
    assert any(n.startswith('decagonmodel/graphconvolutionsparsemulti_') and n.endswith('_vars/weights_0:0') for n in names)
    assert any('dedicomdecoder_' in n and n.endswith('_vars/global_interaction:0') for n in names)
    assert any(n.endswith('_vars/local_variation_5:0') for n in names)
    with pytest.raises(ValueError, match='Unknown decoder type'):
        bad = dict(inputs.edge_type2decoder)
        bad[1, 1] = 'nope'
        DecagonModel(placeholders, inputs.num_feat, inputs.nonzero_feat, inputs.edge_types, bad)
    with pytest.raises(AssertionError):
        from decagon_b200.deep.layers import DEDICOMDecoder
        DEDICOMDecoder(32, edge_type=(1, 1), num_types=2, bogus=1)
    sess = tf.Session(seed=1)
    feed_dict = minibatch.update_feed_dict(minibatch.next_minibatch_feed_dict(placeholders), 0.0, placeholders)
    feed_dict[placeholders['batch']] = feed_dict[placeholders['batch']][:100]
    with pytest.raises(ValueError):
        sess.run([opt.opt_op, opt.cost], feed_dict=feed_dict)


def test_adjacency_is_uploaded_once():
    inputs = datasets.toy_graph()
    placeholders, minibatch, model, opt = build_trainable(inputs)
    sess = tf.Session(seed=3)
    sess.run(tf.global_variables_initializer())
    uploads = []
    for step in range(3):
        fd = minibatch.update_feed_dict(minibatch.next_minibatch_feed_dict(placeholders), 0.1, placeholders)
        sess.run([opt.opt_op, opt.cost, opt.batch_edge_type_idx], feed_dict=fd)
        if step == 0:  # from now on count uploads
            original = model.engine.set_relation
            model.engine.set_relation = lambda r, *a: (uploads.append(r), original(r, *a))
    assert uploads == []
    # a rebuilt tuple (different object) is uploaded again
    key = placeholders['adj_mats_1,1,0']
    c, v, s_ = fd[key]
    fd[key] = (c.copy(), v.copy(), s_)
    sess.run([opt.opt_op, opt.cost, opt.batch_edge_type_idx], feed_dict=fd)
    assert uploads == [minibatch.edge_type2idx[1, 1, 0]]
