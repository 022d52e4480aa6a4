"""torch-CPU restatement of the reference's TF graph, written in the reference's own style --
TEST INFRASTRUCTURE / CPU BASELINE, NOT PRODUCT (see ``decagon_oracle.py`` for the rules).

One sparse matmul per relation, a Python-level ``add_n`` over relations, l2-normalise per
(i, j) group, ReLU after the cross-group sum, B x B scores of which the diagonal is kept,
hinge loss, autograd for the backward pass and the TF-1.8 Adam formula:
``decagon/deep/layers.py:85-118``, ``model.py:64-137``, ``optimizer.py:29-127``.

Two uses:
* ``tests/test_host.py::test_oracle_backward_matches_autograd`` -- autograd gradients as an independent check of the hand-derived
  backward pass in ``decagon_oracle.py`` (float64);
* ``bench.py`` -- the ``cpu_baseline`` / ``--impl reference`` timing ("kind": "port"): TensorFlow
  1.8 cannot be installed here, so this is the stand-in for the reference's CPU path, float32,
  all host threads.
PARITY UNPINNED at the TensorFlow boundary (no reference tests / golden vectors exist).
"""
import numpy as np
import torch


def _sparse(m, dtype):
    m = m.tocoo()
    idx = torch.from_numpy(np.stack([m.row, m.col]).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(m.data).to(dtype), m.shape).coalesce()


class TorchDecagon:
    def __init__(self, graph, params, dtype=torch.float32, lr=1e-3, all_samplers=False):
        self.g, self.dtype, self.lr = graph, dtype, lr
        self.adj = {g: [_sparse(a, dtype) for a in graph.adj[g]] for g in graph.groups}
        self.feat = {t: _sparse(x, dtype) for t, x in graph.feat.items()}
        self.p = {}
        for name, d in params.items():
            for g, a in d.items():
                if a.ndim >= 2 and name != 'R':
                    for k in range(a.shape[0]):
                        self.p[name, g, k] = torch.tensor(a[k], dtype=dtype, requires_grad=True)
                else:
                    self.p[name, g, None] = torch.tensor(a, dtype=dtype, requires_grad=True)
        self.m = {n: torch.zeros_like(v) for n, v in self.p.items()}
        self.v = {n: torch.zeros_like(v) for n, v in self.p.items()}
        self.b1p, self.b2p = 0.9, 0.999
        self.all_samplers = all_samplers

    def _l2n(self, x):
        return x * torch.rsqrt(torch.clamp((x * x).sum(dim=1, keepdim=True), min=1e-12))

    def encoder(self, rate=0.0, masks=None):
        g_ = self.g
        scale = float(np.float32(1) / (np.float32(1) - np.float32(rate)))
        hidden = {}
        for g in g_.groups:
            i, j = g
            outs = []
            for k in range(g_.K[g]):
                x = self.feat[j]
                if masks is not None:
                    keep = torch.from_numpy(masks[0][g, k].astype(np.float64)).to(self.dtype) * scale
                    x = torch.sparse_coo_tensor(x.indices(), x.values() * keep, x.shape)
                h = torch.sparse.mm(x, self.p['W1', g, k])
                outs.append(torch.sparse.mm(self.adj[g][k], h))
            hidden.setdefault(i, []).append(self._l2n(sum(outs)))
        hidden = {i: torch.relu(sum(v)) for i, v in hidden.items()}
        emb = {}
        for g in g_.groups:
            i, j = g
            outs = []
            for k in range(g_.K[g]):
                x = hidden[j]
                if masks is not None:
                    x = x * (torch.from_numpy(masks[1][g, k].astype(np.float64)).to(self.dtype) * scale)
                outs.append(torch.sparse.mm(self.adj[g][k], x @ self.p['W2', g, k]))
            emb.setdefault(i, []).append(self._l2n(sum(outs)))
        return {i: sum(v) for i, v in emb.items()}

    def glb_loc(self, g, k):
        kind, d2 = self.g.decoders[g], self.g.d2
        eye = torch.eye(d2, dtype=self.dtype)
        if kind == 'innerproduct':
            return eye, eye
        if kind == 'distmult':
            return torch.diag(self.p['D', g, k]), eye
        if kind == 'bilinear':
            return self.p['D', g, k], eye
        return self.p['R', g, None], torch.diag(self.p['D', g, k])

    def batch_predict(self, Z, g, k, rows, cols):
        glb, loc = self.glb_loc(g, k)
        preds = Z[g[0]][rows] @ loc @ glb @ loc @ Z[g[1]][cols].T  # B x B, optimizer.py:81-84
        return torch.diagonal(preds)

    def loss(self, g, k, batch, negs, rate=0.0, masks=None, kind='hinge', margin=0.1, neg_weight=1.0):
        Z = self.encoder(rate, masks)
        rows = torch.as_tensor(np.asarray(batch[:, 0], dtype=np.int64))
        cols = torch.as_tensor(np.asarray(batch[:, 1], dtype=np.int64))
        negs = torch.as_tensor(np.asarray(negs, dtype=np.int64))
        pos = self.batch_predict(Z, g, k, rows, cols)
        neg = self.batch_predict(Z, g, k, negs, cols)
        if kind == 'hinge':
            cost = torch.relu(neg - (pos - margin)).sum()
        else:
            sp_ = torch.nn.functional.softplus
            cost = sp_(-pos).sum() + neg_weight * sp_(neg).sum()
        return cost, pos, neg, Z

    def grads(self, *args, **kw):
        for v in self.p.values():
            v.grad = None
        cost, pos, neg, Z = self.loss(*args, **kw)
        cost.backward()
        return cost, pos, neg, Z

    def train_step(self, g, k, batch, negs, rate=0.0, masks=None, **kw):
        """One reference ``session.run([opt_op, cost, batch_edge_type_idx])``."""
        cost, _, _, _ = self.grads(g, k, batch, negs, rate, masks, **kw)
        alpha = self.lr * np.sqrt(1 - self.b2p) / (1 - self.b1p)
        with torch.no_grad():
            for n, p in self.p.items():
                gr = p.grad if p.grad is not None else torch.zeros_like(p)
                self.m[n] += (gr - self.m[n]) * 0.1
                self.v[n] += (gr * gr - self.v[n]) * 0.001
                p -= (self.m[n] * alpha) / (self.v[n].sqrt() + 1e-8)
        self.b1p *= 0.9
        self.b2p *= 0.999
        return float(cost)

    def grad_dict(self):
        """Gradients in the layout of ``decagon_oracle.train_step_grads``."""
        out = {}
        for (name, g, k), p in self.p.items():
            gr = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
            if k is None:
                out.setdefault(name, {})[g] = gr
            else:
                out.setdefault(name, {}).setdefault(g, {})[k] = gr
        return {name: {g: (np.stack([v[k] for k in sorted(v)]) if isinstance(v, dict) else v)
                       for g, v in d.items()} for name, d in out.items()}
