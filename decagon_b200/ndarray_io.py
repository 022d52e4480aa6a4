"""Parameters and embeddings in the reference's on-disk formats.

* ``write_ndarrays`` -- what ``CheckpointToNdarrayWriter._writeAsNdarray`` leaves in ``NpSaveDir``
  (``main/Predictor/CheckpointToNdarrayWriter.py:105-169``): ``embeddings.npy`` (drug embeddings, float32
  ``[n_drugs, hidden2]``), ``GlobalRelations.npy`` (``latent_inters`` of the first drug-drug relation: DEDICOM's R),
  one ``EmbeddingImportance-<side effect id>[-Transposed].npy`` per drug-drug relation (``latent_varies``:
  ``diag(d_k)``).  ``logger_format=True`` writes the variant of ``DecagonLogger._writeAsNdarray``
  (``main/Logger/DecagonLogger.py:232-287``) instead: a single ``EmbeddingImportance.npyz.npz`` whose ``arr_0`` stacks
  the matrices.  ``NpPredictor`` (``main/Predictor/NpPredictor.py:214-333``) reads the first form.
* ``np_predict_edges`` -- ``NpPredictor._predictEdges`` (``:304-319``) restated on the files, for interop checks.
* ``save_variables`` / ``load_variables`` -- every trainable variable under its TF-1 graph name
  (``decagonmodel/graphconvolutionsparsemulti_1_vars/weights_0:0`` ...), the key scheme of a TF checkpoint read with
  ``tf.train.load_checkpoint(...).get_tensor(name)``; a checkpoint converted to such an ``.npz`` loads into the engine.
"""
import os

import numpy as np

DRUG_GRAPH_IDX = 1
DRUG_DRUG = (1, 1)


def _drug_relation_idxs(iterator):
    return [i for i in range(len(iterator.idx2edge_type)) if tuple(iterator.idx2edge_type[i][:2]) == DRUG_DRUG]


def _feed(iterator, placeholders):
    feed = {}
    iterator.update_feed_dict(feed, dropout=0.0, placeholders=placeholders)
    return feed


def write_ndarrays(session, model, iterator, base_dir, side_effect_ids=None, logger_format=False, feed_dict=None):
    """Returns the list of files written.  ``side_effect_ids[k]`` names relation k of the un-transposed half
    (``CheckpointToNdarrayWriter.sideEffectIdx``); relations beyond its length are the transposed twins."""
    os.makedirs(base_dir, exist_ok=True)
    feed = feed_dict if feed_dict is not None else _feed(iterator, model.placeholders)
    valid = _drug_relation_idxs(iterator)
    if not valid:
        raise ValueError('the model has no drug-drug relation (1, 1, *)')
    written = []

    def save(name, arr):
        path = os.path.join(base_dir, name)
        np.save(path, np.asarray(arr, dtype=np.float32), allow_pickle=False)
        written.append(path)

    save('embeddings.npy', session.run(model.embeddings[DRUG_GRAPH_IDX], feed_dict=feed))
    importance = session.run([model.latent_varies[i] for i in valid], feed_dict=feed)
    if logger_format:
        path = os.path.join(base_dir, 'EmbeddingImportance.npyz')
        np.savez(path, np.stack(importance).astype(np.float32))
        written.append(path + '.npz')
    else:
        n_half = len(side_effect_ids) if side_effect_ids is not None else (len(valid) + 1) // 2
        ids = list(side_effect_ids) if side_effect_ids is not None else ['%d' % k for k in range(n_half)]
        for idx, mtx in enumerate(importance):
            tpose = ''
            if idx >= len(ids):
                idx -= len(ids)
                tpose = '-Transposed'
            save('EmbeddingImportance-%s%s.npy' % (ids[idx], tpose), mtx)
    save('GlobalRelations.npy', session.run(model.latent_inters[valid[0]], feed_dict=feed))
    return written


def read_ndarrays(base_dir, relation_id):
    """(embeddings, importance matrix of ``relation_id``, global interaction) as ``NpPredictor`` loads them."""
    emb = np.load(os.path.join(base_dir, 'embeddings.npy'))
    imp = np.load(os.path.join(base_dir, 'EmbeddingImportance-%s.npy' % relation_id))
    glb = np.load(os.path.join(base_dir, 'GlobalRelations.npy'))
    return emb, imp, glb


def np_predict_edges(embeddings, importance, global_interaction, edges):
    """``NpPredictor._predictEdges``: sigmoid(E D R D E^T) sampled at ``edges[:, 0] * n + edges[:, 1]``."""
    raw = embeddings @ importance @ global_interaction @ importance @ embeddings.T
    prob = 1. / (1 + np.exp(-raw))
    edges = np.asarray(edges)
    return np.take(prob, edges[:, 0] * prob.shape[1] + edges[:, 1])


def load_dedicom(engine, base_dir, relation_ids, group=DRUG_DRUG):
    """Puts ``GlobalRelations.npy`` and the diagonals of ``EmbeddingImportance-<id>.npy`` (k-th id -> relation k of
    ``group``) into the engine's DEDICOM variables, and ``embeddings.npy`` into its drug embeddings, so that the
    device scores (``predict`` / ``predict_edges`` / ``evaluate_edges``) reproduce ``NpPredictor`` on the same files."""
    from . import _lib
    glb = np.load(os.path.join(base_dir, 'GlobalRelations.npy'))
    engine.set_param(_lib.PARAM_DEC_GLOBAL, group, None, glb)
    for k, rid in enumerate(relation_ids):
        imp = np.load(os.path.join(base_dir, 'EmbeddingImportance-%s.npy' % rid))
        if np.count_nonzero(imp - np.diag(np.diag(imp))):
            raise ValueError('EmbeddingImportance-%s.npy is not diagonal' % rid)
        engine.set_param(_lib.PARAM_DEC_LOCAL, group, k, np.diag(imp).copy())
    engine.set_embeddings(group[0], np.load(os.path.join(base_dir, 'embeddings.npy')))


def save_variables(session, model, path):
    """``{TF variable name: float32 array}`` of every trainable variable -> ``path`` (.npz)."""
    variables = list(model._variables())
    values = session.run(variables)
    np.savez(path, **{v.name: np.asarray(a, dtype=np.float32) for v, a in zip(variables, values)})
    return [v.name for v in variables]


def load_variables(session, model, path, strict=True):
    """Restores variables saved by ``save_variables`` (or converted from a TF checkpoint under the same names).
    Call after the first ``session.run`` that fed the graph (the engine must exist).  Returns the names loaded."""
    eng = model.engine
    if eng is None:
        raise RuntimeError('the model has no engine yet: run the session once with the graph fed')
    if not getattr(eng, '_initialized', False):
        session._initialize(model)
        eng._initialized = True
    data = np.load(path)
    loaded = []
    for v in model._variables():
        if v.name not in data:
            if strict:
                raise KeyError('variable %s is missing from %s' % (v.name, path))
            continue
        arr = np.asarray(data[v.name], dtype=np.float32)
        if arr.shape != tuple(v.shape):
            raise ValueError('variable %s: saved shape %s, model shape %s' % (v.name, arr.shape, tuple(v.shape)))
        kind, g, k = v.slot
        eng.set_param(kind, g, k, arr)
        loaded.append(v.name)
    return loaded
