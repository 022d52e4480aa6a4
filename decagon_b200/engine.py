"""``Engine``: one ``dgn_graph`` handle (one GPU) behind a small Python object.

This is the layer the TF-style shim (``decagon_b200.session``) and ``bench.py`` drive; every
method is a thin ctypes call into libdecagon_b200.so -- no arithmetic happens here.
"""
import ctypes
import math

import numpy as np

from . import _lib
from ._lib import as_f32, as_i32, check, ptr


class Engine(object):
    def __init__(self, n_nodes, feat_dim, edge_types, decoders, hidden1=64, hidden2=32, device=0):
        """n_nodes / feat_dim: ``{type: int}``; edge_types: ``{(i, j): K}`` in dict order;
        decoders: ``{(i, j): 'innerproduct' | 'distmult' | 'bilinear' | 'dedicom'}``."""
        self.lib = _lib.load()
        self.groups = list(edge_types)
        self.K = {g: int(edge_types[g]) for g in self.groups}
        self.n_types = max(max(i, j) for i, j in self.groups) + 1
        self.n_nodes = {t: int(n_nodes[t]) for t in range(self.n_types)}
        self.feat_dim = {t: int(feat_dim[t]) for t in range(self.n_types)}
        self.decoders = {}
        for g in self.groups:
            if decoders[g] not in _lib.DEC_KINDS:
                raise ValueError('Unknown decoder type')  # model.py:113-114
            self.decoders[g] = decoders[g]
        self.hidden1, self.hidden2 = int(hidden1), int(hidden2)
        self.flat = [(g, k) for g in self.groups for k in range(self.K[g])]
        self.flat_index = {gk: r for r, gk in enumerate(self.flat)}
        self.R = len(self.flat)
        self._h = ctypes.c_void_p()
        nn = as_i32([self.n_nodes[t] for t in range(self.n_types)])
        fd = as_i32([self.feat_dim[t] for t in range(self.n_types)])
        ij = as_i32([x for g in self.groups for x in g])
        kk = as_i32([self.K[g] for g in self.groups])
        dd = as_i32([_lib.DEC_KINDS[self.decoders[g]] for g in self.groups])
        check(self.lib.dgn_graph_create(ctypes.byref(self._h), device, self.n_types, ptr(nn, ctypes.c_int32),
                                        ptr(fd, ctypes.c_int32), len(self.groups), ptr(ij, ctypes.c_int32),
                                        ptr(kk, ctypes.c_int32), ptr(dd, ctypes.c_int32), self.hidden1, self.hidden2))
        self.finalized = False

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self.lib.dgn_graph_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ one process per GPU
    def comm_init(self, rank, world):
        """Partition the many-relation groups over ``world`` ranks (call before ``load_iterator``)."""
        check(self.lib.dgn_comm_init(self._h, int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)
        self.finalized = False

    def comm_handle(self):
        """64-byte CUDA IPC handle of this rank's exchange arena (after ``finalize``)."""
        buf = ctypes.create_string_buffer(64)
        check(self.lib.dgn_comm_handle(self._h, ctypes.cast(buf, ctypes.c_void_p)))
        return buf.raw

    def comm_connect(self, handles):
        """handles: the ``world`` 64-byte handles in rank order (an all-gather of ``comm_handle()``)."""
        blob = b''.join(handles)
        buf = ctypes.create_string_buffer(blob, len(blob))
        check(self.lib.dgn_comm_connect(self._h, ctypes.cast(buf, ctypes.c_void_p)))

    def connect(self, dist):
        """Exchange the arena handles through an initialised ``torch.distributed`` (any backend)."""
        handles = [None] * dist.get_world_size()
        dist.all_gather_object(handles, self.comm_handle())
        self.comm_connect(handles)
        dist.barrier()

    def relation_owner(self, r):
        o = ctypes.c_int(0)
        check(self.lib.dgn_relation_owner(self._h, int(r), ctypes.byref(o)))
        return o.value

    # ------------------------------------------------------------------ graph
    def set_relation(self, r, coords, values, shape):
        """(coords int[nnz,2], values, shape) exactly as ``adj_train[i,j][k]``; values are cast to
        float32 like TF's ``sparse_placeholder(tf.float32)`` does on feed."""
        coords = np.asarray(coords)
        rows, cols, vals = as_i32(coords[:, 0]), as_i32(coords[:, 1]), as_f32(values)
        check(self.lib.dgn_graph_set_relation(self._h, r, int(shape[0]), int(shape[1]), len(vals),
                                              ptr(rows, ctypes.c_int32), ptr(cols, ctypes.c_int32),
                                              ptr(vals, ctypes.c_float)))
        self.finalized = False

    def set_features(self, t, coords, values, shape):
        coords = np.asarray(coords)
        rows, cols, vals = as_i32(coords[:, 0]), as_i32(coords[:, 1]), as_f32(values)
        check(self.lib.dgn_graph_set_features(self._h, t, int(shape[0]), int(shape[1]), len(vals),
                                              ptr(rows, ctypes.c_int32), ptr(cols, ctypes.c_int32),
                                              ptr(vals, ctypes.c_float)))
        self.finalized = False

    def set_degrees(self, r, degrees):
        d = np.ascontiguousarray(degrees, dtype=np.float64)
        check(self.lib.dgn_sampler_set_degrees(self._h, r, ptr(d, ctypes.c_double), len(d)))

    def finalize(self):
        check(self.lib.dgn_graph_finalize(self._h))
        self.finalized = True

    def load_iterator(self, iterator, degrees=None):
        """Upload every ``adj_train`` tuple and feature tuple of an ``EdgeMinibatchIterator`` (what
        ``update_feed_dict`` would feed each step, ``minibatch.py:259-267``) and, when given, the
        sampler's degree tables (``degrees[i][k]``, ``optimizer.py:38-47``)."""
        for r, (g, k) in enumerate(self.flat):
            self.set_relation(r, *iterator.adj_train[g][k])
        for t, feat in iterator.feat.items():
            self.set_features(t, *feat)
        if degrees is not None:
            for r, (g, k) in enumerate(self.flat):
                self.set_degrees(r, degrees[g[0]][k])
        self.finalize()

    def get_csr(self, r):
        nnz = ctypes.c_int64(0)
        check(self.lib.dgn_graph_relation_nnz(self._h, r, ctypes.byref(nnz)))
        g, _ = self.flat[r]
        rowptr = np.empty(self.n_nodes[g[0]] + 1, dtype=np.int32)
        col = np.empty(nnz.value, dtype=np.int32)
        val = np.empty(nnz.value, dtype=np.float32)
        check(self.lib.dgn_graph_get_csr(self._h, r, ptr(rowptr, ctypes.c_int32), ptr(col, ctypes.c_int32),
                                         ptr(val, ctypes.c_float)))
        return rowptr, col, val

    # ------------------------------------------------------------------ parameters
    def param_shape(self, kind, g, k=None):
        K = self.K[g] if k is None else 1
        d1, d2 = self.hidden1, self.hidden2
        dec = self.decoders[g]
        if kind == _lib.PARAM_W1:
            shape = (self.feat_dim[g[1]], d1)
        elif kind == _lib.PARAM_W2:
            shape = (d1, d2)
        elif kind == _lib.PARAM_DEC_GLOBAL:
            return (d2, d2)
        else:
            shape = (d2, d2) if dec == 'bilinear' else (d2,)
        return shape if k is not None else (K,) + shape

    def _param_io(self, fn, kind, g, k, arr):
        gi = self.groups.index(g)
        check(fn(self._h, kind, gi, -1 if k is None else int(k), ptr(arr, ctypes.c_float), arr.size))

    def set_param(self, kind, g, k, values):
        arr = as_f32(values)
        if arr.shape != self.param_shape(kind, g, k):
            raise ValueError('parameter kind %d of %s[%s]: shape %s, expected %s'
                             % (kind, g, k, arr.shape, self.param_shape(kind, g, k)))
        self._param_io(self.lib.dgn_params_set, kind, g, k, arr)

    def get_param(self, kind, g, k=None):
        arr = np.empty(self.param_shape(kind, g, k), dtype=np.float32)
        self._param_io(self.lib.dgn_params_get, kind, g, k, arr)
        return arr

    def get_grad(self, kind, g, k=None):
        arr = np.empty(self.param_shape(kind, g, k), dtype=np.float32)
        self._param_io(self.lib.dgn_grads_get, kind, g, k, arr)
        return arr

    def set_params(self, p):
        """p: ``{'W1': {g: [K,F,d1]}, 'W2': {g: [K,d1,d2]}, 'R': {g: [d2,d2]}, 'D': {g: [K,...]}}``"""
        for g in self.groups:
            self.set_param(_lib.PARAM_W1, g, None, p['W1'][g])
            self.set_param(_lib.PARAM_W2, g, None, p['W2'][g])
            if g in p.get('R', {}):
                self.set_param(_lib.PARAM_DEC_GLOBAL, g, None, p['R'][g])
            if g in p.get('D', {}):
                self.set_param(_lib.PARAM_DEC_LOCAL, g, None, p['D'][g])

    def _collect(self, getter):
        out = {'W1': {}, 'W2': {}, 'R': {}, 'D': {}}
        for g in self.groups:
            out['W1'][g] = getter(_lib.PARAM_W1, g)
            out['W2'][g] = getter(_lib.PARAM_W2, g)
            if self.decoders[g] == 'dedicom':
                out['R'][g] = getter(_lib.PARAM_DEC_GLOBAL, g)
            if self.decoders[g] != 'innerproduct':
                out['D'][g] = getter(_lib.PARAM_DEC_LOCAL, g)
        return out

    def get_params(self):
        return self._collect(self.get_param)

    def get_grads(self):
        return self._collect(self.get_grad)

    def keep_gradients(self, keep=True):
        """Materialise every gradient also in steps that apply the update (no fused Adam): needed to fetch
        ``opt.grads_vars`` together with ``opt.opt_op`` (optimizer.py:111-114)."""
        check(self.lib.dgn_keep_gradients(self._h, 1 if keep else 0))

    def n_params(self):
        n = ctypes.c_int64(0)
        check(self.lib.dgn_params_count(self._h, ctypes.byref(n)))
        return n.value

    def reset_optimizer(self, beta1=0.9, beta2=0.999, epsilon=1e-8):
        check(self.lib.dgn_optimizer_reset(self._h, beta1, beta2, epsilon))

    # ------------------------------------------------------------------ compute
    def forward(self, dropout=0.0, seed=0, step=0):
        check(self.lib.dgn_encoder_forward(self._h, float(dropout), int(seed), int(step)))

    def train_step(self, r, batch, negatives=None, loss='hinge', margin=0.1, neg_weight=1.0, lr=1e-3, dropout=0.0,
                   seed=0, step=0, apply_update=True, want_loss=True):
        batch = as_i32(batch)
        if batch.ndim != 2 or batch.shape[1] != 2:
            raise ValueError("'batch' must be [B, 2]")
        neg_p = None
        if negatives is not None:
            negatives = np.ascontiguousarray(negatives, dtype=np.int64)
            if negatives.shape != (batch.shape[0],):
                raise ValueError('negatives must be [B]')
            neg_p = ptr(negatives, ctypes.c_int64)
        out = ctypes.c_float(0.0)
        check(self.lib.dgn_train_step(self._h, int(r), ptr(batch, ctypes.c_int32), batch.shape[0], neg_p,
                                      _lib.LOSS_KINDS[loss], margin, neg_weight, lr, float(dropout), int(seed), int(step),
                                      1 if apply_update else 0, ctypes.byref(out) if want_loss else None))
        return np.float32(out.value) if want_loss else None

    def last_batch_outputs(self, batch_size):
        pos = np.empty(batch_size, dtype=np.float32)
        neg = np.empty(batch_size, dtype=np.float32)
        samples = np.empty(batch_size, dtype=np.int64)
        check(self.lib.dgn_last_batch_outputs(self._h, ptr(pos, ctypes.c_float), ptr(neg, ctypes.c_float),
                                              ptr(samples, ctypes.c_int64), batch_size))
        return pos, neg, samples

    def predict(self, r):
        g, _ = self.flat[r]
        out = np.empty((self.n_nodes[g[0]], self.n_nodes[g[1]]), dtype=np.float32)
        check(self.lib.dgn_predict_all_pairs(self._h, int(r), ptr(out, ctypes.c_float)))
        return out

    def predict_relations_dev(self, r0, count, out_dev_ptr):
        check(self.lib.dgn_predict_relations_dev(self._h, int(r0), int(count), ctypes.c_void_p(out_dev_ptr)))

    def predict_edges(self, r, edges, sigmoid=True):
        edges = as_i32(np.asarray(edges).reshape(-1, 2))
        out = np.empty(len(edges), dtype=np.float32)
        check(self.lib.dgn_predict_edges(self._h, int(r), ptr(edges, ctypes.c_int32), len(edges), 1 if sigmoid else 0,
                                         ptr(out, ctypes.c_float)))
        return out

    def evaluate_edges(self, group, rel_k, edges, labels=None, sigmoid=True, want_scores=True):
        """Edges of many relations of one group in one launch; with ``labels`` also the pooled AUROC / AUPRC computed
        on the device.  Returns ``(scores or None, auroc, auprc)`` (DecagonAccuracyEvaluator.py:57-91)."""
        gi = self.groups.index(tuple(group)) if not isinstance(group, int) else group
        edges = as_i32(np.asarray(edges).reshape(-1, 2))
        rel_k = as_i32(np.asarray(rel_k).reshape(-1))
        if len(rel_k) != len(edges):
            raise ValueError('rel_k and edges differ in length')
        scores = np.empty(len(edges), dtype=np.float32) if want_scores else None
        auroc, auprc = ctypes.c_double(math.nan), ctypes.c_double(math.nan)
        lab = None
        if labels is not None:
            lab = np.ascontiguousarray(np.asarray(labels).reshape(-1) != 0, dtype=np.uint8)
            if len(lab) != len(edges):
                raise ValueError('labels and edges differ in length')
        check(self.lib.dgn_evaluate_edges(
            self._h, gi, len(edges), ptr(rel_k, ctypes.c_int32), ptr(edges, ctypes.c_int32),
            ptr(lab, ctypes.c_uint8) if lab is not None else None, 1 if sigmoid else 0,
            ptr(scores, ctypes.c_float) if want_scores else None,
            ctypes.byref(auroc) if lab is not None else None, ctypes.byref(auprc) if lab is not None else None))
        return scores, auroc.value, auprc.value

    def rank_edges(self, r, edges, top=None, sigmoid=True):
        """Indices of the ``top`` best-scoring candidates of relation r in descending score order, and their scores
        (GreedyActiveLearner._getRankedPossibilities, GreedyActiveLearner.py:84-92), scored and sorted on the device."""
        edges = as_i32(np.asarray(edges).reshape(-1, 2))
        top = len(edges) if top is None else min(int(top), len(edges))
        order = np.empty(top, dtype=np.int32)
        scores = np.empty(top, dtype=np.float32)
        check(self.lib.dgn_rank_edges(self._h, int(r), ptr(edges, ctypes.c_int32), len(edges), 1 if sigmoid else 0, top,
                                      ptr(order, ctypes.c_int32), ptr(scores, ctypes.c_float)))
        return order, scores

    def tensor(self, which, index):
        if which in (_lib.TENSOR_HIDDEN1, _lib.TENSOR_EMBEDDINGS, _lib.TENSOR_GRAD_EMBEDDINGS):
            rows = self.n_nodes[index]
        else:
            rows = self.n_nodes[self.groups[index][0]]
        d = self.hidden1 if which in (_lib.TENSOR_HIDDEN1, _lib.TENSOR_LAYER1_GROUP) else self.hidden2
        out = np.empty((rows, d), dtype=np.float32)
        check(self.lib.dgn_tensor_get(self._h, which, int(index), ptr(out, ctypes.c_float), out.size))
        return out

    def set_embeddings(self, t, values):
        arr = as_f32(values)
        check(self.lib.dgn_tensor_set(self._h, _lib.TENSOR_EMBEDDINGS, int(t), ptr(arr, ctypes.c_float), arr.size))

    def embeddings(self, t):
        return self.tensor(_lib.TENSOR_EMBEDDINGS, t)

    def hidden1_of(self, t):
        return self.tensor(_lib.TENSOR_HIDDEN1, t)

    def relation_matrices(self, r):
        glb = np.empty((self.hidden2, self.hidden2), dtype=np.float32)
        loc = np.empty((self.hidden2, self.hidden2), dtype=np.float32)
        check(self.lib.dgn_relation_matrices(self._h, int(r), ptr(glb, ctypes.c_float), ptr(loc, ctypes.c_float)))
        return glb, loc

    def sync(self):
        check(self.lib.dgn_sync(self._h))

    # ------------------------------------------------------------------ measurement
    def timing(self, enable):
        check(self.lib.dgn_timing_enable(self._h, 1 if enable else 0))

    def timing_reset(self):
        check(self.lib.dgn_timing_reset(self._h))

    def timing_get(self, name):
        ms, n = ctypes.c_double(0), ctypes.c_int64(0)
        check(self.lib.dgn_timing_get(self._h, name.encode(), ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def timeline(self):
        """[(name, lane, start_ms, stop_ms)] of the phases recorded since ``timing_reset``."""
        out, buf = [], ctypes.create_string_buffer(96)
        lane, a, b = ctypes.c_int(0), ctypes.c_double(0), ctypes.c_double(0)
        i = 0
        while self.lib.dgn_timeline_get(self._h, i, buf, 96, ctypes.byref(lane), ctypes.byref(a), ctypes.byref(b)) == 0:
            out.append((buf.value.decode(), lane.value, a.value, b.value))
            i += 1
        return out

    def launch_count(self):
        n = ctypes.c_int64(0)
        check(self.lib.dgn_launch_count(self._h, ctypes.byref(n)))
        return n.value

    def counters(self):
        """(training steps replayed as one CUDA graph, groups built by finalize so far)."""
        a, b = ctypes.c_int64(0), ctypes.c_int64(0)
        check(self.lib.dgn_counters(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def timer_start(self):
        check(self.lib.dgn_timer_start(self._h))

    def timer_stop(self):
        ms = ctypes.c_double(0)
        check(self.lib.dgn_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def memory(self):
        f, t = ctypes.c_int64(0), ctypes.c_int64(0)
        check(self.lib.dgn_memory_bytes(self._h, ctypes.byref(f), ctypes.byref(t)))
        return f.value, t.value
