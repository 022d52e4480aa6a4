"""CPU oracle for the Decagon hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path (``decagon_b200``) never
does and fails loudly when its CUDA library is missing.

What it restates (numpy, float64 as the arbiter or float32 as the reference-precision twin):

* encoder forward     ``decagon/deep/layers.py:85-94,109-118`` + ``model.py:64-88``
* decoder parameters  ``model.py:116-137`` (glb / loc per decoder kind)
* minibatch scores    ``optimizer.py:63-85`` + ``diag_part`` ``:52,56``
* all-pairs scores    ``optimizer.py:87-106``
* hinge / xent loss   ``optimizer.py:116-127``
* reverse-mode gradients of all of the above (what ``AdamOptimizer.minimize`` differentiates,
  ``optimizer.py:111-113``), hand-derived and cross-checked against torch
  autograd in ``tests/test_host.py::test_oracle_backward_matches_autograd``
* TF-1.8 ``ApplyAdam`` update (``requirements.txt:22``; TF is a third-party dependency that is
  not under /root/reference and cannot be installed here -- its published semantics are restated:
  ``alpha = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g*g-v)(1-b2);
  p -= m*alpha/(sqrt(v)+eps)``; ``l2_normalize = x*rsqrt(max(sum x^2, 1e-12))``;
  ``dropout = x/keep*floor(keep+U)``).

PARITY UNPINNED at the TensorFlow boundary: the reference has no tests, golden vectors or
known-answer files for this arithmetic and TensorFlow cannot run here, so these functions are
pinned only by (a) the reference's own numpy statement of the DEDICOM all-pairs formula
(``main/Predictor/NpPredictor.py:293-313``) evaluated on the reference's dumped parameters
(``tests/golden/make_golden.py``), and (b) torch autograd.  The integer / index half of the path
(normalised adjacency tuples, splits, minibatch order) is NOT restated here: it is checked
against the reference's own ``minibatch.py`` imported unmodified (``tests/golden``).

Randomness: TF's dropout and negative sampler are unseeded, so the product defines its own
counter-based streams (Philox4x32-10); this file restates those integer streams bit-exactly.
"""
import math

import numpy as np
import scipy.sparse as sp

# --------------------------------------------------------------------------- Philox4x32-10
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
STREAM_DROPOUT1, STREAM_DROPOUT2, STREAM_NEGATIVES = 1, 2, 3


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11).  Counter words may be arrays; returns 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & _MASK for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


def stream_u32(n, relation, stream, step, seed):
    """Element e of a stream is word ``e & 3`` of Philox(counter=(e >> 2, relation, stream, step),
    key=(seed_lo, seed_hi))."""
    e = np.arange(n, dtype=np.uint64)
    words = philox4x32(e >> np.uint64(2), relation, stream, step, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(words, axis=1)[np.arange(n), (e & np.uint64(3)).astype(np.int64)]


def dropout_threshold(rate):
    """keep element  <=>  u32 >= threshold;  rate is the float32 value TF would be fed."""
    rate = float(np.float32(rate))
    return min(int(math.ceil(rate * 4294967296.0)), 0xFFFFFFFF)


def dropout_keep(n, relation, layer, step, seed, rate):
    """Boolean keep mask of ``n`` elements (layer 1: feature non-zeros in row-major order;
    layer 2: elements of H_j in row-major order), ``layers.py:23-31`` / ``:112``."""
    if rate == 0:
        return np.ones(n, dtype=bool)
    stream = STREAM_DROPOUT1 if layer == 1 else STREAM_DROPOUT2
    return stream_u32(n, relation, stream, step, seed) >= np.uint32(dropout_threshold(rate))


# --------------------------------------------------------------------------- negative sampler
def sampler_thresholds(degrees):
    """uint32 CDF of ``degrees ** 0.75`` (``optimizer.py:40-47``: distortion 0.75, unique=False).
    ``d**0.75`` is evaluated as ``sqrt(d) * sqrt(sqrt(d))`` so that it is correctly rounded
    everywhere (libm ``pow`` is not)."""
    d = np.asarray(degrees, dtype=np.float64)
    w = np.sqrt(d) * np.sqrt(np.sqrt(d))
    cum = np.cumsum(w)
    thr = np.floor(cum / cum[-1] * 4294967296.0)
    return np.minimum(thr, 4294967295.0).astype(np.uint32)


def sample_negatives(thresholds, batch_size, relation, step, seed):
    """``batch_size`` i.i.d. draws: index = #{v : thr[v] <= u}, clamped to the last node."""
    u = stream_u32(batch_size, relation, STREAM_NEGATIVES, step, seed)
    idx = np.searchsorted(thresholds, u, side='right')
    return np.minimum(idx, len(thresholds) - 1).astype(np.int64)


# --------------------------------------------------------------------------- model description
class Graph:
    """groups: list of (i, j) in dict order; adj[(i,j)]: list of scipy CSR (normalised);
    feat[t]: scipy CSR feature matrix; decoders[(i,j)]: kind string."""

    def __init__(self, n_nodes, groups, adj, feat, decoders, hidden1=64, hidden2=32):
        self.n_nodes, self.groups, self.adj, self.feat = dict(n_nodes), list(groups), adj, feat
        self.decoders, self.d1, self.d2 = dict(decoders), hidden1, hidden2
        self.K = {g: len(adj[g]) for g in self.groups}
        self.flat = [(g, k) for g in self.groups for k in range(self.K[g])]
        self.flat_index = {gk: r for r, gk in enumerate(self.flat)}

    @staticmethod
    def from_iterator(iterator, decoders, hidden1=64, hidden2=32, dtype=np.float64):
        """Build from an ``EdgeMinibatchIterator``'s ``adj_train`` / ``feat`` tuples."""
        groups = list(iterator.edge_types)
        adj, n_nodes = {}, {}
        for g in groups:
            adj[g] = []
            for coords, values, shape in iterator.adj_train[g]:
                # TF's sparse_placeholder(float32) casts the float64 values on feed
                v = np.asarray(values, dtype=np.float64).astype(np.float32).astype(dtype)
                adj[g].append(sp.csr_matrix((v, (coords[:, 0], coords[:, 1])), shape=shape))
                n_nodes[g[0]], n_nodes[g[1]] = shape
        feat = {}
        for t, (coords, values, shape) in iterator.feat.items():
            v = np.asarray(values, dtype=np.float64).astype(np.float32).astype(dtype)
            feat[t] = sp.csr_matrix((v, (coords[:, 0], coords[:, 1])), shape=shape)
        return Graph(n_nodes, groups, adj, feat, decoders, hidden1, hidden2)


def glorot(rng, fan_in, fan_out):
    """``inits.py:5-12``: U(-a, a), a = sqrt(6 / (in + out)), float32."""
    a = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-a, a, size=(fan_in, fan_out)).astype(np.float32)


def init_params(graph, rng):
    """Float32 parameters with the reference's shapes (``layers.py:80-83,104-107,127-133,
    156-160,181-184``).  Draw order: W1 of every relation, W2 of every relation, decoders."""
    p = {'W1': {}, 'W2': {}, 'R': {}, 'D': {}}
    for g in graph.groups:
        f_in = graph.feat[g[1]].shape[1]
        p['W1'][g] = np.stack([glorot(rng, f_in, graph.d1) for _ in range(graph.K[g])])
    for g in graph.groups:
        p['W2'][g] = np.stack([glorot(rng, graph.d1, graph.d2) for _ in range(graph.K[g])])
    d2 = graph.d2
    for g in graph.groups:
        kind = graph.decoders[g]
        if kind == 'dedicom':
            p['R'][g] = glorot(rng, d2, d2)
            p['D'][g] = np.stack([glorot(rng, d2, 1).reshape(-1) for _ in range(graph.K[g])])
        elif kind == 'distmult':
            p['D'][g] = np.stack([glorot(rng, d2, 1).reshape(-1) for _ in range(graph.K[g])])
        elif kind == 'bilinear':
            p['D'][g] = np.stack([glorot(rng, d2, d2) for _ in range(graph.K[g])])
        elif kind != 'innerproduct':
            raise ValueError('Unknown decoder type')
    return p


def cast_params(p, dtype):
    return {name: {g: np.array(a, dtype=dtype) for g, a in d.items()} for name, d in p.items()}  # always a copy


def zeros_like_params(p):
    return {name: {g: np.zeros_like(a) for g, a in d.items()} for name, d in p.items()}


def relation_matrices(graph, p, g, k):
    """(glb, loc) of ``model.py:116-137``."""
    kind, d2 = graph.decoders[g], graph.d2
    dt = p['W2'][g].dtype
    eye = np.eye(d2, dtype=dt)
    if kind == 'innerproduct':
        return eye, eye
    if kind == 'distmult':
        return np.diag(p['D'][g][k]), eye
    if kind == 'bilinear':
        return p['D'][g][k], eye
    if kind == 'dedicom':
        return p['R'][g], np.diag(p['D'][g][k])
    raise ValueError('Unknown decoder type')


# --------------------------------------------------------------------------- encoder
def _l2norm(s):
    n = np.sqrt(np.maximum((s * s).sum(axis=1, keepdims=True), s.dtype.type(1e-12)))
    return s / n, n


def _l2norm_bwd(y, n, dy):
    return (dy - y * (y * dy).sum(axis=1, keepdims=True)) / n


def masks_for(graph, rate, step, seed):
    """Keep masks of both layers for every flat relation (None when rate == 0)."""
    if rate == 0:
        return None
    m1, m2 = {}, {}
    for r, (g, k) in enumerate(graph.flat):
        j = g[1]
        m1[g, k] = dropout_keep(graph.feat[j].nnz, r, 1, step, seed, rate)
        # the device keeps hidden1 in 32-column panels (1, 2 or 4 of them): element (row, col) of the layer-2
        # input is bit row * stride + col of the relation's stream
        stride = 32 if graph.d1 <= 32 else 64 if graph.d1 <= 64 else 128
        m2[g, k] = dropout_keep(graph.n_nodes[j] * stride, r, 2, step, seed, rate).reshape(graph.n_nodes[j], stride)[:, :graph.d1]
    return m1, m2


def encoder_forward(graph, p, rate=0.0, masks=None):
    """Returns (Z, cache).  Z[t]: final embedding of node type t (``model.embeddings``)."""
    dt = p['W2'][graph.groups[0]].dtype.type
    keep = np.float32(1) - np.float32(rate)
    scale = dt(np.float32(1) / keep)
    c = {'X1': {}, 'Y1': {}, 'n1': {}, 'Y2': {}, 'n2': {}, 'Hm': {}, 'scale': scale, 'rate': rate, 'masks': masks}
    acc = {}
    for g in graph.groups:
        i, j = g
        x = graph.feat[j]
        s = None
        for k in range(graph.K[g]):
            xk = x
            if masks is not None:
                xk = sp.csr_matrix((x.data * masks[0][g, k] * scale, x.indices, x.indptr), shape=x.shape)
            c['X1'][g, k] = xk
            t = graph.adj[g][k] @ (xk @ p['W1'][g][k])
            s = t if s is None else s + t
        c['Y1'][g], c['n1'][g] = _l2norm(np.asarray(s))
        acc[i] = c['Y1'][g] if i not in acc else acc[i] + c['Y1'][g]
    H = {i: np.maximum(a, 0) for i, a in acc.items()}
    c['H'] = H
    acc = {}
    for g in graph.groups:
        i, j = g
        s = None
        for k in range(graph.K[g]):
            hk = H[j]
            if masks is not None:
                hk = H[j] * (masks[1][g, k] * scale)
            c['Hm'][g, k] = hk
            t = graph.adj[g][k] @ (hk @ p['W2'][g][k])
            s = t if s is None else s + t
        c['Y2'][g], c['n2'][g] = _l2norm(np.asarray(s))
        acc[i] = c['Y2'][g] if i not in acc else acc[i] + c['Y2'][g]
    return acc, c


def encoder_backward(graph, p, cache, dZ, relu_mask=None):
    """Gradients of W1 / W2 given dL/dZ (SURVEY.md section 9 'Backward').  relu_mask[t] (bool [n_t, d1], optional):
    the activation pattern to differentiate through instead of ``H > 0`` -- ReLU is not differentiable at 0, and a
    float32 implementation may round a pre-activation of ~1e-9 to the other side than float64 does; a parity test
    passes the pattern of the implementation under test after checking that it differs only at such entries."""
    scale, masks = cache['scale'], cache['masks']
    gW1 = {g: np.zeros_like(a) for g, a in p['W1'].items()}
    gW2 = {g: np.zeros_like(a) for g, a in p['W2'].items()}
    dH = {t: np.zeros_like(h) for t, h in cache['H'].items()}
    for g in graph.groups:
        i, j = g
        if i not in dZ:
            continue
        dS = _l2norm_bwd(cache['Y2'][g], cache['n2'][g], dZ[i])
        for k in range(graph.K[g]):
            G = graph.adj[g][k].T @ dS
            gW2[g][k] = cache['Hm'][g, k].T @ G
            back = G @ p['W2'][g][k].T
            if masks is not None:
                back = back * (masks[1][g, k] * scale)
            dH[j] += back
    for g in graph.groups:
        i, j = g
        dY = dH[i] * ((cache['H'][i] > 0) if relu_mask is None else relu_mask[i])
        dS = _l2norm_bwd(cache['Y1'][g], cache['n1'][g], dY)
        for k in range(graph.K[g]):
            G = graph.adj[g][k].T @ dS
            gW1[g][k] = np.asarray(cache['X1'][g, k].T @ G)
    return gW1, gW2


# --------------------------------------------------------------------------- decoder / loss
def batch_scores(graph, p, Z, g, k, rows, cols):
    """``optimizer.py:63-85``: s_b = ((Z_i[u_b] loc) glb) loc . Z_j[v_b]."""
    glb, loc = relation_matrices(graph, p, g, k)
    return ((((Z[g[0]][rows] @ loc) @ glb) @ loc) * Z[g[1]][cols]).sum(axis=1)


def predict_all_pairs(graph, p, Z, g, k):
    """``optimizer.py:87-106``: Z_i loc glb loc Z_j^T."""
    glb, loc = relation_matrices(graph, p, g, k)
    return (((Z[g[0]] @ loc) @ glb) @ loc) @ Z[g[1]].T


def loss_and_dscores(pos, neg, kind='hinge', margin=0.1, neg_weight=1.0):
    """hinge ``optimizer.py:116-120`` / xent ``:122-127``; returns (loss, dL/dpos, dL/dneg)."""
    dt = pos.dtype.type
    if kind == 'hinge':
        diff = neg - (pos - dt(margin))
        active = (diff > 0).astype(pos.dtype)
        return np.maximum(diff, 0).sum(), -active, active
    if kind == 'xent':
        softplus = lambda x: np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))
        sigmoid = lambda x: 1.0 / (1.0 + np.exp(-x))
        loss = softplus(-pos).sum() + dt(neg_weight) * softplus(neg).sum()
        return loss, (-sigmoid(-pos)).astype(pos.dtype), (dt(neg_weight) * sigmoid(neg)).astype(pos.dtype)
    raise ValueError('Unknown loss kind')


def decode_backward(graph, p, Z, g, k, rows, cols, negs, dpos, dneg):
    """dZ (dict by node type) and decoder-parameter gradients for relation (g, k)."""
    i, j = g
    glb, loc = relation_matrices(graph, p, g, k)
    M = loc @ glb @ loc
    zu, zn, zv = Z[i][rows], Z[i][negs], Z[j][cols]
    a = zv @ M.T                      # a_b = M z_v
    dZ = {t: np.zeros_like(z) for t, z in Z.items()}
    np.add.at(dZ[i], rows, dpos[:, None] * a)
    np.add.at(dZ[i], negs, dneg[:, None] * a)
    np.add.at(dZ[j], cols, dpos[:, None] * (zu @ M) + dneg[:, None] * (zn @ M))
    dM = (zu * dpos[:, None]).T @ zv + (zn * dneg[:, None]).T @ zv
    grads = {}
    kind = graph.decoders[g]
    if kind == 'bilinear':
        grads['D'] = dM
    elif kind == 'distmult':
        grads['D'] = np.diag(dM).copy()
    elif kind == 'dedicom':
        d, R = p['D'][g][k], p['R'][g]
        grads['R'] = np.outer(d, d) * dM
        grads['D'] = (dM * R) @ d + (dM * R).T @ d
    return dZ, grads


def train_step_grads(graph, p, g, k, batch, negs, rate=0.0, masks=None, loss_kind='hinge', margin=0.1,
                     neg_weight=1.0, relu_mask=None, want_cache=False):
    """One ``session.run([opt_op, cost, ...])`` worth of forward + backward (no update).
    Returns (loss, pos, neg, grads, Z)."""
    Z, cache = encoder_forward(graph, p, rate, masks)
    rows, cols = np.asarray(batch[:, 0], dtype=np.int64), np.asarray(batch[:, 1], dtype=np.int64)
    negs = np.asarray(negs, dtype=np.int64)
    pos = batch_scores(graph, p, Z, g, k, rows, cols)
    neg = batch_scores(graph, p, Z, g, k, negs, cols)
    loss, dpos, dneg = loss_and_dscores(pos, neg, loss_kind, margin, neg_weight)
    dZ, dec = decode_backward(graph, p, Z, g, k, rows, cols, negs, dpos, dneg)
    gW1, gW2 = encoder_backward(graph, p, cache, dZ, relu_mask)
    grads = zeros_like_params(p)
    grads['W1'], grads['W2'] = gW1, gW2
    if 'R' in dec:
        grads['R'][g] = dec['R']
    if 'D' in dec:
        grads['D'][g][k] = dec['D']
    if want_cache:
        return loss, pos, neg, grads, Z, cache
    return loss, pos, neg, grads, Z


# --------------------------------------------------------------------------- TF1 Adam
class AdamTF1:
    """``tf.train.AdamOptimizer`` (TF 1.8 ``ApplyAdam``), dense update of EVERY variable each step
    -- variables with a zero gradient still move through their decaying moments."""

    def __init__(self, params, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
        self.dt = params['W2'][next(iter(params['W2']))].dtype.type
        self.lr, self.b1, self.b2, self.eps = (self.dt(x) for x in (lr, beta1, beta2, eps))
        self.b1p, self.b2p = self.b1, self.b2
        self.m, self.v = zeros_like_params(params), zeros_like_params(params)

    def alpha(self):
        one = self.dt(1)
        return self.lr * np.sqrt(one - self.b2p) / (one - self.b1p)

    def apply(self, params, grads):
        one, alpha = self.dt(1), self.alpha()
        for name in params:
            for g in params[name]:
                gr, m, v = grads[name][g], self.m[name][g], self.v[name][g]
                m += (gr - m) * (one - self.b1)
                v += (gr * gr - v) * (one - self.b2)
                params[name][g] -= (m * alpha) / (np.sqrt(v) + self.eps)
        self.b1p, self.b2p = self.b1p * self.b1, self.b2p * self.b2


# --------------------------------------------------------------------------- evaluation
def sigmoid(x):
    """``main/Utils/MathUtils.py:3-4``."""
    return 1.0 / (1.0 + np.exp(-x))


def sampled_scores(pred, edges):
    """``DecagonAccuracyEvaluator.py:151-186``: sigma(P)[u * n_cols + v]."""
    e = np.asarray(edges).astype(np.int64)
    return np.take(sigmoid(pred), e[:, 0] * pred.shape[1] + e[:, 1])
