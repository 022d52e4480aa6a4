"""Shared builders for the parity tests: one seeded case = graph + iterator + oracle graph +
parameters (+ a loaded Engine on the GPU box)."""
import numpy as np

from decagon_b200 import datasets
from decagon_b200.deep.minibatch import EdgeMinibatchIterator
from oracle import decagon_oracle as O

PLACEHOLDER_KEYS = ['batch', 'batch_edge_type_idx', 'batch_row_edge_type', 'batch_col_edge_type', 'dropout']
MIXED_DECODERS = {(0, 0): 'innerproduct', (0, 1): 'distmult', (1, 1): 'dedicom', (1, 0): 'bilinear'}


def rel_err(a, b):
    """max |a - b| / max |b|: error relative to the scale of the reference tensor."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)) if b.size else 0.0


def mini_poly(n_types=12, seed=0, features=None):
    """A small graph with the structure of config #3 (many small drug-drug relations), big
    enough to take the staged SpMM path (K >= 8)."""
    return datasets.polypharmacy_graph(n_types=n_types, seed=seed, n_proteins=300, n_drugs=97, n_ppi=2500,
                                       n_targets=400, n_pairs=1500, n_ddi=9000, min_size=100, max_size=1400,
                                       features=features)


def rel_l2(a, b):
    """||a - b||_2 / ||b||_2 over the whole tensor."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30)) if b.size else 0.0


def assert_close(a, b, tol, what=None):
    """Both metrics of the parity contract: max|a-b| / max|b| and ||a-b||_2 / ||b||_2, each <= tol."""
    e_max, e_l2 = rel_err(a, b), rel_l2(a, b)
    assert e_max <= tol and e_l2 <= tol, (what, 'max-norm rel-err %.3e' % e_max, 'l2 rel-err %.3e' % e_l2)
    return max(e_max, e_l2)


class Case(object):
    def __init__(self, inputs, iterator_seed=0, param_seed=1, batch_size=512, val_test_size=0.05, hidden1=64, hidden2=32):
        self.inputs = inputs
        np.random.seed(iterator_seed)
        self.it = EdgeMinibatchIterator(inputs.adj_mats, inputs.feat, inputs.edge_types, {}, batch_size=batch_size,
                                        val_test_size=val_test_size)
        self.graph = O.Graph.from_iterator(self.it, inputs.edge_type2decoder, hidden1=hidden1, hidden2=hidden2)
        self.p32 = O.init_params(self.graph, np.random.RandomState(param_seed))
        self.p64 = O.cast_params(self.p32, np.float64)
        self.batch_size = batch_size
        self.placeholders = {k: k for k in PLACEHOLDER_KEYS}
        self.hidden1, self.hidden2 = hidden1, hidden2

    def engine(self):
        from decagon_b200.engine import Engine
        eng = Engine(self.inputs.n_nodes, self.inputs.num_feat, self.inputs.edge_types, self.inputs.edge_type2decoder,
                     hidden1=self.hidden1, hidden2=self.hidden2)
        eng.load_iterator(self.it, self.inputs.degrees)
        eng.set_params(self.p32)
        return eng

    def batches(self, n, seed=5):
        """First n minibatches of an epoch: [(r, batch int[B,2])]."""
        np.random.seed(seed)
        self.it.shuffle()
        out = []
        for _ in range(n):
            fd = self.it.next_minibatch_feed_dict(self.placeholders)
            out.append((int(fd['batch_edge_type_idx']), np.array(fd['batch'])))
        return out

    def thresholds(self, r):
        g, k = self.graph.flat[r]
        return O.sampler_thresholds(self.inputs.degrees[g[0]][k])
