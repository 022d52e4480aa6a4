"""``Session.run(fetches, feed_dict)`` of the reference's contract, executed by the CUDA engine.

Contract: train ``[opt.opt_op, opt.cost, opt.batch_edge_type_idx]`` ->
``(None, float32, int32)`` (``DecagonTrainer.py:94-102``); eval ``opt.predictions`` -> float32
``[n_i, n_j]`` (``DecagonAccuracyEvaluator.py:122``); dumps ``model.embeddings[1]``,
``model.latent_varies[r]``, ``model.latent_inters[r]`` (``DecagonLogger.py:239-281``);
``tf.global_variables_initializer()`` (``DecagonTrainer.py:49``).

The reference re-feeds every adjacency and feature tuple on every run (``minibatch.py:259-267``,
~450 MB per step at the polypharmacy shape).  Here the tuples are uploaded ONCE: a later run that
feeds the same tuple objects (what ``update_feed_dict`` does) costs nothing; feeding different
objects re-uploads.  TF's dropout / sampler randomness is unseeded; the engine's Philox streams
are keyed by ``(seed, step)`` of this session (``Session(seed=...)``), step = number of runs so far
that consumed randomness.
"""
import os

import numpy as np

from . import _lib
from . import tf_compat as tf
from .engine import Engine


class Session(object):
    def __init__(self, config=None, seed=None, device=None):
        self.config = config
        self.seed = int.from_bytes(os.urandom(8), 'little') if seed is None else int(seed)
        self.step = 0
        self.device = int(os.environ.get('LOCAL_RANK', '0')) if device is None else device

    def close(self):
        pass

    # ------------------------------------------------------------------ engine lifetime
    def _engine(self, model, feed):
        eng = model.engine
        token = getattr(feed, 'graph_token', None)
        if eng is not None and token is not None and token == eng._graph_token and eng.finalized and eng._degrees_set:
            return eng  # the iterator vouches for its tuples (minibatch.FeedDict): nothing to compare
        adj_ph = model.adj_mats
        groups = model.groups()
        if eng is None:
            n_nodes = {}
            for g in groups:
                tup = self._fed(feed, adj_ph[g][0])
                n_nodes[g[0]], n_nodes[g[1]] = int(tup[2][0]), int(tup[2][1])
            eng = Engine(n_nodes, model.input_dim, model.edge_types, model.decoders, model.hidden1_dim,
                         model.hidden2_dim, device=self.device)
            eng._fed_ids, eng._initialized, eng._graph_token, eng._degrees_set, eng._keep = {}, False, None, False, {}
            self._dist = None
            world = int(os.environ.get('WORLD_SIZE', '1'))
            if world > 1:
                # one process per GPU (torchrun): the many-relation groups are partitioned over the ranks;
                # torch.distributed is only the control plane that carries the exchange handles
                import torch.distributed as dist
                if not dist.is_initialized():
                    raise RuntimeError('WORLD_SIZE > 1: call torch.distributed.init_process_group first')
                eng.comm_init(dist.get_rank(), dist.get_world_size())
                self._dist = dist
            model.engine = eng
        dirty = False
        for r, (g, k) in enumerate(eng.flat):
            ph = adj_ph[g][k]
            if ph in feed and eng._fed_ids.get(ph) != id(feed[ph]):
                eng.set_relation(r, *feed[ph])
                eng._fed_ids[ph] = id(feed[ph])
                eng._keep[ph] = feed[ph]  # keep the tuple alive so its id stays unique
                dirty = True
        for t, ph in model.inputs.items():
            if ph in feed and eng._fed_ids.get(ph) != id(feed[ph]):
                eng.set_features(t, *feed[ph])
                eng._fed_ids[ph] = id(feed[ph])
                eng._keep[ph] = feed[ph]
                dirty = True
        opt = getattr(model, 'optimizer', None)
        if opt is not None and not eng._degrees_set:
            for r, (g, k) in enumerate(eng.flat):
                eng.set_degrees(r, opt.degrees[g[0]][k])
            eng._degrees_set = True
        if dirty or not eng.finalized:
            eng.finalize()
            if getattr(self, '_dist', None) is not None:
                eng.connect(self._dist)
        eng._graph_token = token
        return eng

    @staticmethod
    def _fed(feed, ph):
        if ph in feed:
            return feed[ph]
        if getattr(ph, 'default', None) is not None:
            return ph.default
        raise ValueError('You must feed a value for placeholder %r' % (ph,))

    def _initialize(self, model):
        eng = model.engine
        for v in model._variables():
            kind, g, k = v.slot
            eng.set_param(kind, g, k, v.initial)
        eng.reset_optimizer()
        eng._initialized = True

    # ------------------------------------------------------------------ run
    def run(self, fetches, feed_dict=None):
        feed = feed_dict or {}
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        out = self._run(flist, feed)
        return out[0] if single else out

    def _run(self, flist, feed):
        models = []
        for f in flist:
            m = self._model_of(f)
            if m is not None and m not in models:
                models.append(m)
        results = [None] * len(flist)
        if any(isinstance(f, tf.InitOp) for f in flist):
            # every variable of every live model goes back to its initial value, optimizer slots to
            # zero; models whose engine does not exist yet are initialised when it is created
            for m in list(tf.MODELS):
                if m.engine is not None:
                    self._initialize(m)
        for idx, f in enumerate(flist):
            if isinstance(f, tf.Placeholder):
                results[idx] = np.asarray(self._fed(feed, f), dtype=f.dtype)
        for model in models:
            self._run_model(model, flist, feed, results)
        return results

    @staticmethod
    def _model_of(f):
        if isinstance(f, tf.Variable):
            return f.model
        if isinstance(f, tf.Tensor):
            owner = f.owner
            return owner if hasattr(owner, 'edge_type2decoder') else getattr(owner, 'model', None)
        return None

    def _run_model(self, model, flist, feed, results):
        mine = [(idx, f) for idx, f in enumerate(flist) if self._model_of(f) is model]
        kinds = {f.kind for _, f in mine if isinstance(f, tf.Tensor)}
        needs_graph = bool(kinds - {'latent_inter', 'latent_vary'}) or model.engine is None
        eng = self._engine(model, feed) if needs_graph or model.engine is None else model.engine
        if not eng._initialized:
            # the reference runs global_variables_initializer() before anything else
            # (DecagonTrainer.py:49); variables hold their initial values until then
            self._initialize(model)
        opt = getattr(model, 'optimizer', None)
        dropout = float(np.float32(self._fed(feed, model.dropout)))
        step_kinds = {'opt_op', 'cost', 'outputs', 'neg_outputs', 'neg_samples', 'grad'}
        step_kinds |= {'preds', 'neg_preds'}
        loss = None
        if kinds & step_kinds:
            batch = np.asarray(self._fed(feed, opt.inputs))
            if batch.ndim != 2 or batch.shape != (opt.batch_size, 2):
                raise ValueError("'batch' must have shape [%d, 2] (optimizer.py:36), got %s"
                                 % (opt.batch_size, batch.shape))
            r = int(self._fed(feed, opt.batch_edge_type_idx))
            g, _ = eng.flat[r]
            if (int(self._fed(feed, opt.batch_row_edge_type)), int(self._fed(feed, opt.batch_col_edge_type))) != g:
                raise ValueError('batch_row_edge_type / batch_col_edge_type do not match relation %d = %s' % (r, g))
            eng.keep_gradients('grad' in kinds)  # [opt_op, grads_vars]: no fused Adam, every gradient is stored
            step_kw = dict(negatives=None, loss=opt.loss_kind, margin=opt.margin, neg_weight=opt.neg_sample_weights,
                           lr=opt.learning_rate, dropout=dropout, seed=self.seed, step=self.step)
            pair_scores = {}
            if kinds & {'preds', 'neg_preds'}:
                # batch_predict (optimizer.py:51,55,63-85): the full B x B matrices of the step's embeddings and
                # PRE-update decoder variables -- only their diagonals enter the loss.  A debugging fetch, never on
                # the training path: the step is first run without the update, the pairs are scored on the device,
                # then (with opt_op) the identical step (same streams: same masks and negatives) applies the update.
                loss = eng.train_step(r, batch, apply_update=False, **step_kw)
                negs = eng.last_batch_outputs(opt.batch_size)[2]
                B = opt.batch_size
                for kind, rows in (('preds', batch[:, 0]), ('neg_preds', negs)):
                    if kind in kinds:
                        pairs = np.stack([np.repeat(rows, B), np.tile(batch[:, 1], B)], axis=1)
                        pair_scores[kind] = eng.predict_edges(r, pairs, sigmoid=False).reshape(B, B)
            if 'opt_op' in kinds or not pair_scores:
                loss = eng.train_step(r, batch, apply_update='opt_op' in kinds, **step_kw)
            self.step += 1
        elif kinds & {'predictions', 'embeddings', 'hidden1', 'layer1_group', 'layer2_group', 'decoder_scores'}:
            eng.forward(dropout, self.seed, self.step)
            if dropout > 0:
                self.step += 1
        batch_out = None
        for idx, f in mine:
            if isinstance(f, tf.Variable):
                kind, g, k = f.slot
                results[idx] = eng.get_param(kind, g, k)
                continue
            kind = f.kind
            if kind == 'opt_op':
                results[idx] = None
            elif kind == 'cost':
                results[idx] = np.float32(loss)
            elif kind in ('outputs', 'neg_outputs', 'neg_samples'):
                if batch_out is None:
                    batch_out = eng.last_batch_outputs(opt.batch_size)
                results[idx] = batch_out[{'outputs': 0, 'neg_outputs': 1, 'neg_samples': 2}[kind]]
            elif kind in ('preds', 'neg_preds'):
                results[idx] = pair_scores[kind]
            elif kind in ('row_inputs', 'col_inputs'):
                batch = np.asarray(self._fed(feed, opt.inputs))
                results[idx] = batch[:, 0 if kind == 'row_inputs' else 1].astype(np.int32)
            elif kind == 'grad':
                gk, g, k = f.index.slot
                results[idx] = eng.get_grad(gk, g, k)
            elif kind == 'predictions':
                results[idx] = eng.predict(int(self._fed(feed, opt.batch_edge_type_idx)))
            elif kind == 'embeddings':
                results[idx] = eng.embeddings(f.index)
            elif kind == 'hidden1':
                results[idx] = eng.hidden1_of(f.index)
            elif kind in ('layer1_group', 'layer2_group'):
                which = _lib.TENSOR_LAYER1_GROUP if kind == 'layer1_group' else _lib.TENSOR_LAYER2_GROUP
                results[idx] = eng.tensor(which, eng.groups.index(f.index))
            elif kind in ('latent_inter', 'latent_vary'):
                glb, loc = eng.relation_matrices(f.index)
                results[idx] = glb if kind == 'latent_inter' else loc
            elif kind == 'decoder_scores':
                dec = f.owner
                from .deep.layers import _is_identity
                scores = eng.predict(eng.flat_index[dec.edge_type, f.index])
                results[idx] = scores if _is_identity(dec.act) else np.asarray(dec.act(scores), dtype=np.float32)
            else:
                raise ValueError('cannot fetch %r' % (f,))
