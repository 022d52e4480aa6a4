"""Host side of the hot path: edge split, adjacency normalisation, minibatch schedule.

Drop-in for ``decagon/deep/minibatch.py`` of the reference (same class name,
constructor signature, attribute names and feed-dict keys).  Index work must be
bit-exact with the reference under the same ``np.random`` state, so every draw from
the legacy global numpy stream happens in the same order and with the same call as
there (cited per method).  The arithmetic that consumes these tuples runs on the GPU
(``decagon_b200/csrc``); nothing here touches the device.

Reference quirks that are kept on purpose:

* ``np.random.seed(123)`` at import (``minibatch.py:9``).
* only ONE false validation edge and ONE false test edge are drawn per relation
  (the sampling loops ``break`` after their first append, ``minibatch.py:202,216``).
* the test split is always 50 edges, the validation split ``max(50, floor(n * frac))``
  (``minibatch.py:176-177``).
* a transposed twin re-uses its original's normalised tuple with the coordinates
  flipped and the values untouched (``minibatch.py:141-149``).
"""
import os

import numpy as np
import scipy.sparse as sp

from ..utility import preprocessing

np.random.seed(123)  # minibatch.py:9

EDGES_IDX = 0
DRUG_DRUG_GRAPH_TYPE = (1, 1)
_FIXED_SCHEDULE = ((0, 0, 0), (0, 1, 0), (1, 0, 0))  # minibatch.py:283-290,338


class GraphRelationType:
    def __init__(self, graphType, relationType):
        self.graphType = graphType
        self.relationType = relationType


def normalize_adjacency(adj):
    """Normalised adjacency as ``(coords, values float64, shape)``.

    Restates ``EdgeMinibatchIterator.preprocess_graph`` (``minibatch.py:80-93``):

    * square  : ``D^-1/2 (A+I)^T D^-1/2`` with ``D = rowsum(A+I)``.  Entry ``(i, j)``
      of ``A+I`` lands at coordinate ``(j, i)`` with value ``(a * d_j) * d_i`` and the
      tuple is ordered by ``(i, j)`` -- the order scipy's CSC->COO conversion yields there.
    * rect    : ``Dr^-1/2 A Dc^-1/2`` ordered row-major, value ``(dr_i * a) * dc_j``;
      zero degrees go through ``nan_to_num`` like the reference (they never meet an entry).
    """
    a = sp.csr_matrix(sp.coo_matrix(adj), dtype=np.float64)
    n_rows, n_cols = a.shape
    if n_rows == n_cols:
        a = sp.csr_matrix(a + sp.eye(n_rows))
    a.sum_duplicates()
    a.sort_indices()
    rows = np.repeat(np.arange(n_rows, dtype=np.int32), np.diff(a.indptr))
    cols = a.indices.astype(np.int32, copy=False)
    if n_rows == n_cols:
        d = np.power(np.asarray(a.sum(1)), -0.5).ravel()
        values = (a.data * d[cols]) * d[rows]
        coords = np.stack([cols, rows], axis=1)
        return coords, values, (n_rows, n_cols)
    with np.errstate(divide='ignore'):
        dr = np.nan_to_num(np.power(np.asarray(a.sum(1)), -0.5)).ravel()
        dc = np.nan_to_num(np.power(np.asarray(a.sum(0)), -0.5)).ravel()
    values = (dr[rows] * a.data) * dc[cols]
    coords = np.stack([rows, cols], axis=1)
    return coords, values, (n_rows, n_cols)


def _contains(pair, edges):
    """``minibatch.py:95-99``: is the (row, col) pair one of ``edges``?"""
    edges = np.asarray(edges)
    if edges.size == 0:
        return False
    return bool(np.any((edges[:, 0] == pair[0]) & (edges[:, 1] == pair[1])))


_EAGER_FEED = os.environ.get('DGN_EAGER_FEED') == '1'  # A/B switch: copy the entries every step like a plain dict


class FeedDict(dict):
    """A feed dict whose ~2000 adjacency / feature entries are attached LAZILY.

    ``update_feed_dict`` (called every step by the reference's trainer) adds the same 1934 tuples to a fresh dict
    each time -- 48 us of hashing at the polypharmacy shape that nobody looks at when the session already holds the
    graph.  Here it attaches the iterator's shared ``{placeholder: tuple}`` mapping as ``_pending`` together with
    ``graph_token`` (which tells ``Session`` that every such entry is the iterator's own tuple); the entries are
    merged into the dict the first time anything could observe their absence: a lookup miss, ``in``, ``len``,
    iteration, ``keys / values / items / get / copy / pop`` and ``dict(fd)`` (which goes through ``keys`` because
    ``__iter__`` is overridden).  Replacing a sparse entry by hand clears the token."""
    graph_token = None
    _pending = None

    def _materialise(self):
        pending = self._pending
        if pending is not None:
            self._pending = None
            for k, v in pending.items():   # explicit entries win over the attached ones
                if not dict.__contains__(self, k):
                    dict.__setitem__(self, k, v)

    def attach(self, entries, token):
        self._materialise()
        self._pending = entries
        self.graph_token = token

    def __missing__(self, key):
        if self._pending is not None:
            self._materialise()
            return dict.__getitem__(self, key)
        raise KeyError(key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or (self._pending is not None and key in self._pending)

    def get(self, key, default=None):
        if dict.__contains__(self, key):
            return dict.__getitem__(self, key)
        if self._pending is not None and key in self._pending:
            return self._pending[key]
        return default

    def __setitem__(self, key, value):
        if getattr(key, 'sparse', False):
            self._materialise()
            self.graph_token = None
        dict.__setitem__(self, key, value)

    def update(self, *args, **kwargs):
        self._materialise()
        self.graph_token = None
        dict.update(self, *args, **kwargs)

    def _full(name):
        def method(self, *args, **kwargs):
            self._materialise()
            return getattr(dict, name)(self, *args, **kwargs)
        method.__name__ = name
        return method

    for _name in ('__len__', '__iter__', 'keys', 'values', 'items', 'pop', 'popitem', 'setdefault', '__delitem__',
                  '__eq__', '__ne__', '__repr__', '__reversed__', '__or__', '__ror__'):
        locals()[_name] = _full(_name)
    del _name, _full
    __hash__ = None

    def copy(self):
        self._materialise()
        out = FeedDict(dict.copy(self))
        out.graph_token = self.graph_token
        return out


class EdgeMinibatchIterator(object):
    """Iterates over batches of training edges of one relation at a time.

    adj_mats -- ``{(i, j): [matrix, ...]}``; every matrix carries ``.id``,
                ``.isTranspose`` and ``.transposedMtxLink`` (``decagon_b200.sparse``)
    feat     -- ``{i: (coords, values, shape)}``
    edge_types -- ``{(i, j): K}``
    drug_drug_test_edges -- ``{key: {'positive': int[n,2], 'negative': int[n,2]}}``,
                enumerated in dict order like the reference (``minibatch.py:33-36``)
    """
    verbose = False  # the reference prints three lines per relation (minibatch.py:67-72)

    def __init__(self, adj_mats, feat, edge_types, drug_drug_test_edges, batch_size=100, val_test_size=0.01):
        self.adj_mats = adj_mats
        self.feat = feat
        self.edge_types = edge_types
        self.batch_size = batch_size
        self.val_test_size = val_test_size
        self.num_edge_types = sum(self.edge_types.values())
        self.drug_drug_test_edges = dict(enumerate(drug_drug_test_edges.values()))

        self.iter = 0
        self.freebatch_edge_types = list(range(self.num_edge_types))
        self.batch_num = [0] * self.num_edge_types
        self.current_edge_type_idx = 0
        self.edge_type2idx = {}
        self.idx2edge_type = {}
        self.adj_mtx_to_idx = {}
        for i, j in self.edge_types:
            for k in range(self.edge_types[i, j]):
                r = len(self.idx2edge_type)
                self.edge_type2idx[i, j, k] = r
                self.idx2edge_type[r] = i, j, k
                self.adj_mtx_to_idx[self.adj_mats[i, j][k].id] = ((i, j), k)

        def per_relation():
            return {et: [None] * n for et, n in self.edge_types.items()}

        self.train_edges = per_relation()
        self.val_edges = per_relation()
        self.test_edges = per_relation()
        self.test_edges_false = per_relation()
        self.val_edges_false = per_relation()
        self.adj_train = per_relation()

        for i, j in self.edge_types:
            for k in range(self.edge_types[i, j]):
                self.mask_test_edges((i, j), k)
                if self.verbose:
                    print("Minibatch edge type:", "(%d, %d, %d)" % (i, j, k))
                    print("Train edges=", "%04d" % len(self.train_edges[i, j][k]))
                    print("Val edges=", "%04d" % len(self.val_edges[i, j][k]))
                    print("Test edges=", "%04d" % len(self.test_edges[i, j][k]))

    @property
    def graphAndRelationTypes(self):
        for graph_type, per_rel in self.val_edges.items():
            for idx in range(len(per_rel)):
                yield GraphRelationType(graph_type, idx)

    def preprocess_graph(self, adj):
        return normalize_adjacency(adj)

    # ------------------------------------------------------------------ split
    def mask_test_edges(self, edge_type, type_idx):
        """Dispatch of ``minibatch.py:120-128``."""
        mtx = self.adj_mats[edge_type][type_idx]
        if mtx.isTranspose and self._test_edges_exist_tposed_mtx(mtx):
            return self._mask_test_edges_from_tpose(edge_type, type_idx, mtx)
        if edge_type == DRUG_DRUG_GRAPH_TYPE and type_idx in self.drug_drug_test_edges:
            return self._mask_test_edges_drug_drug_precomputed(type_idx)
        return self._mask_test_edges_new(edge_type, type_idx)

    def _test_edges_exist_tposed_mtx(self, mtx):
        if not mtx.isTranspose:
            return False
        et, k = self.adj_mtx_to_idx[mtx.transposedMtxLink.id]
        return self.train_edges[et][k] is not None

    def _mask_test_edges_from_tpose(self, edge_type, type_idx, mtx):
        """Mirror the already-split twin (``minibatch.py:137-172``)."""
        if not mtx.isTranspose:
            return False
        et, k = self.adj_mtx_to_idx[mtx.transposedMtxLink.id]
        coords, values, shape = self.adj_train[et][k]
        self.adj_train[edge_type][type_idx] = (np.flip(coords, axis=1), values, (shape[1], shape[0]))
        for name in ('train_edges', 'val_edges', 'val_edges_false', 'test_edges'):
            store = getattr(self, name)
            store[edge_type][type_idx] = np.flip(store[et][k], axis=1)
        try:
            self.test_edges_false[edge_type][type_idx] = np.flip(self.test_edges_false[et][k], axis=1)
        except ValueError:
            self.test_edges_false[edge_type][type_idx] = []

    def _draw_false_edge(self, shape, edges_all):
        """One uniformly drawn non-edge; two ``np.random.randint`` per attempt
        (``minibatch.py:194-197`` / ``:208-211``)."""
        while True:
            idx_i = np.random.randint(0, shape[0])
            idx_j = np.random.randint(0, shape[1])
            if not _contains((idx_i, idx_j), edges_all):
                return [idx_i, idx_j]

    def _mask_test_edges_new(self, edge_type, type_idx):
        """Random split of one relation (``minibatch.py:174-233``)."""
        mtx = self.adj_mats[edge_type][type_idx]
        edges_all, _, _ = preprocessing.sparse_to_tuple(mtx)
        n = edges_all.shape[0]
        num_test = 50
        num_val = max(50, int(np.floor(n * self.val_test_size)))

        # np.random.shuffle draws the same interval sequence for a list and an ndarray
        order = np.arange(n)
        np.random.shuffle(order)
        val_edge_idx = order[:num_val]
        test_edge_idx = order[num_val:num_val + num_test]
        val_edges = edges_all[val_edge_idx]
        test_edges = edges_all[test_edge_idx]
        train_edges = np.delete(edges_all, np.hstack([test_edge_idx, val_edge_idx]), axis=0)

        test_edges_false = [self._draw_false_edge(mtx.shape, edges_all)] if len(test_edges) > 0 else []
        val_edges_false = [self._draw_false_edge(mtx.shape, edges_all)] if len(val_edges) > 0 else []

        adj_train = sp.csr_matrix(
            (np.ones(train_edges.shape[0]), (train_edges[:, 0], train_edges[:, 1])), shape=mtx.shape)
        self.adj_train[edge_type][type_idx] = self.preprocess_graph(adj_train)
        self.train_edges[edge_type][type_idx] = train_edges
        self.val_edges[edge_type][type_idx] = val_edges
        self.test_edges[edge_type][type_idx] = test_edges
        self.val_edges_false[edge_type][type_idx] = np.array(val_edges_false).reshape((len(val_edges_false), 2))
        self.test_edges_false[edge_type][type_idx] = np.array(test_edges_false).reshape((len(test_edges_false), 2))

    def _mask_test_edges_drug_drug_precomputed(self, type_idx):
        """Validation edges handed in by the caller (``minibatch.py:235-253``)."""
        g = DRUG_DRUG_GRAPH_TYPE
        mtx = self.adj_mats[g][type_idx]
        self.adj_train[g][type_idx] = self.preprocess_graph(mtx)
        rows, cols = mtx.nonzero()
        self.train_edges[g][type_idx] = np.stack([rows, cols], axis=1)
        self.val_edges[g][type_idx] = self.drug_drug_test_edges[type_idx]['positive']
        self.val_edges_false[g][type_idx] = self.drug_drug_test_edges[type_idx]['negative']
        self.test_edges[g][type_idx] = np.empty((0, 2))
        self.test_edges_false[g][type_idx] = np.empty((0, 2))

    # ------------------------------------------------------------ feed dicts
    def end(self):
        return len(self.freebatch_edge_types) == 0

    def _graph_feed(self, placeholders):
        """{placeholder: tuple} of every adjacency and feature tuple, built once per placeholder dict."""
        cached = getattr(self, '_graph_feed_cache', None)
        if cached is None or cached[0] is not placeholders:
            entries = {}
            for i, j in self.edge_types:
                for k in range(self.edge_types[i, j]):
                    entries[placeholders['adj_mats_%d,%d,%d' % (i, j, k)]] = self.adj_train[i, j][k]
            for i, _ in self.edge_types:
                entries[placeholders['feat_%d' % i]] = self.feat[i]
            self._graph_feed_cache = cached = (placeholders, entries)
        return cached[1]

    def update_feed_dict(self, feed_dict, dropout, placeholders):
        """Adds every adjacency tuple, the feature tuples and the dropout rate
        (``minibatch.py:259-267``).  The tuples are the SAME objects on every call, which is
        what lets ``decagon_b200.session.Session`` keep them resident on the device; a ``FeedDict``
        (what ``batch_feed_dict`` returns) also carries a token naming this iterator's graph, so the
        session does not have to compare the 1932 tuples of the polypharmacy shape on every run."""
        if isinstance(feed_dict, FeedDict):
            feed_dict.attach(self._graph_feed(placeholders), (id(self), id(placeholders)))
            feed_dict._graph_owner = self  # keeps id(self) unique while the feed dict lives
            if _EAGER_FEED:
                feed_dict._materialise()
        else:
            dict.update(feed_dict, self._graph_feed(placeholders))
        dict.__setitem__(feed_dict, placeholders['dropout'], dropout)
        return feed_dict

    def batch_feed_dict(self, batch_edges, batch_edge_type, placeholders):
        return FeedDict({
            placeholders['batch']: batch_edges,
            placeholders['batch_edge_type_idx']: batch_edge_type,
            placeholders['batch_row_edge_type']: self.idx2edge_type[batch_edge_type][0],
            placeholders['batch_col_edge_type']: self.idx2edge_type[batch_edge_type][1],
        })

    def next_minibatch_feed_dict(self, placeholders):
        """Round-robin schedule of ``minibatch.py:278-313``: the fixed relations
        (0,0,0), (0,1,0) and (when present) (1,0,0) in turn, then one relation drawn with
        ``np.random.choice`` from those that still have a full batch left."""
        period = 4 if (1, 0, 0) in self.edge_type2idx else 3
        while True:
            slot = self.iter % period
            if slot == 0:
                self.current_edge_type_idx = self.edge_type2idx[0, 0, 0]
            elif slot == 1:
                self.current_edge_type_idx = self.edge_type2idx[0, 1, 0]
            elif slot == 2 and period == 4:
                self.current_edge_type_idx = self.edge_type2idx[1, 0, 0]
            elif len(self.freebatch_edge_types) > 0:
                # np.random.choice(list) of the reference (minibatch.py:292) draws randint(0, len) on the same legacy
                # stream after converting the list to an array (115 us for 1929 relations); the direct draw is
                # bit-identical (tests/test_host.py::test_choice_is_randint) and takes 4 us
                free = self.freebatch_edge_types
                self.current_edge_type_idx = free[np.random.randint(0, len(free))]
            else:
                self.current_edge_type_idx = self.edge_type2idx[0, 0, 0]
                self.iter = 0

            i, j, k = self.idx2edge_type[self.current_edge_type_idx]
            r = self.current_edge_type_idx
            if self.batch_num[r] * self.batch_size <= len(self.train_edges[i, j][k]) - self.batch_size:
                break
            if self.iter % 4 in (0, 1, 2):  # literal 4 in the reference (minibatch.py:304)
                self.batch_num[r] = 0
            else:
                self.freebatch_edge_types.remove(r)

        self.iter += 1
        start = self.batch_num[r] * self.batch_size
        self.batch_num[r] += 1
        batch_edges = self.train_edges[i, j][k][start:start + self.batch_size]
        return self.batch_feed_dict(batch_edges, r, placeholders)

    def num_training_batches(self, edge_type, type_idx):
        return len(self.train_edges[edge_type][type_idx]) // self.batch_size + 1

    def val_feed_dict(self, edge_type, type_idx, placeholders, size=None):
        edge_list = self.val_edges[edge_type][type_idx]
        if size is None:
            return self.batch_feed_dict(edge_list, edge_type, placeholders)
        ind = np.random.permutation(len(edge_list))
        picked = [edge_list[i] for i in ind[:min(size, len(ind))]]
        return self.batch_feed_dict(picked, edge_type, placeholders)

    def shuffle(self):
        """Re-shuffle every relation's training edges and reset the schedule
        (``minibatch.py:327-345``)."""
        for edge_type in self.edge_types:
            for k in range(self.edge_types[edge_type]):
                self.train_edges[edge_type][k] = np.random.permutation(self.train_edges[edge_type][k])
                self.batch_num[self.edge_type2idx[edge_type[0], edge_type[1], k]] = 0
        self.current_edge_type_idx = 0
        self.freebatch_edge_types = list(range(self.num_edge_types))
        for fixed in _FIXED_SCHEDULE:
            if fixed in self.edge_type2idx:
                self.freebatch_edge_types.remove(self.edge_type2idx[fixed])
        self.iter = 0
