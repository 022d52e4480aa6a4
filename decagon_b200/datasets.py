"""Synthetic graphs of the shapes BASELINE.json names.

* ``toy_graph``            -- config #1: the generator of the reference's ``main.py:134-183``
                              (500 genes, 400 drugs, 3 drug-drug types).
* ``polypharmacy_graph``   -- config #3 / #5: a seeded graph with the published shape of the
                              Decagon polypharmacy data (the reference ships no
                              generator for it): 19 085 proteins, 645 drugs, 715 612 PPI edges,
                              18 596 drug-target pairs, 964 side-effect types over a base set of
                              63 473 drug pairs with ~4.6 M (pair, type) edges.

Both return a ``GraphInputs`` whose fields are exactly the arguments the reference's
``DecagonTrainableBuilder`` (``DecagonTrainableBuilder.py:70-118``) hands to the iterator,
model and optimizer: dict order (0,0), (0,1), (1,1), (1,0) and transposed twins appended as
``DecagonDataSet._augmentAdjMtxDictWithTranspose`` does (``DecagonDataSet.py:212-231``).
"""
from collections import namedtuple

import numpy as np
import scipy.sparse as sp

from .sparse import RelationCsrMatrix
from .utility import preprocessing

GraphInputs = namedtuple('GraphInputs', [
    'adj_mats', 'feat', 'num_feat', 'nonzero_feat', 'edge_types', 'degrees',
    'edge_type2dim', 'edge_type2decoder', 'n_nodes'])

DEFAULT_DECODERS = {(0, 0): 'bilinear', (0, 1): 'bilinear', (1, 0): 'bilinear', (1, 1): 'dedicom'}


def multi_hot_features(n_nodes, n_feat, per_row=12, seed=0):
    """Multi-hot node features shaped like the reference's drug features (one column per mono side effect, a 1
    for every side effect of the drug, ``DecagonPublicDataNodeFeaturesBuilder.py:34-51``): ``per_row`` distinct
    columns on average, heavy-tailed column popularity, every row non-empty."""
    rng = np.random.RandomState(seed)
    pop = 1.0 / np.sqrt(np.arange(1, n_feat + 1))
    pop /= pop.sum()
    rows, cols = [], []
    for r in range(n_nodes):
        c = np.unique(rng.choice(n_feat, size=max(1, rng.poisson(per_row)), p=pop))
        rows.append(np.full(len(c), r))
        cols.append(c)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    return sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(n_nodes, n_feat))


def assemble(gene_adj, gene_drug_adj, drug_drug_adj_list, decoders=None, transpose=True, features=None):
    """Raw scipy matrices -> the dicts of the drop-in surface.  ``features``: {node type: scipy sparse [n, F]}
    replaces the identity features of that type (``DecagonDataSet.py:110-131`` builds the same tuples)."""
    adj = {
        (0, 0): [RelationCsrMatrix(gene_adj)],
        (0, 1): [RelationCsrMatrix(gene_drug_adj)],
        (1, 1): [RelationCsrMatrix(m) for m in drug_drug_adj_list],
    }
    if transpose:
        augmented = {}
        for et, mtxs in adj.items():
            twins = [m.transpose(copy=True, setId=True) for m in mtxs]
            if et == (0, 1):
                augmented[et] = mtxs
                augmented[(1, 0)] = twins
            else:
                augmented[et] = mtxs + twins
        adj.update(augmented)

    n_genes, n_drugs = gene_drug_adj.shape
    feat = {0: preprocessing.sparse_to_tuple(sp.identity(n_genes).tocoo()),
            1: preprocessing.sparse_to_tuple(sp.identity(n_drugs).tocoo())}
    for t, x in (features or {}).items():
        feat[t] = preprocessing.sparse_to_tuple(sp.coo_matrix(x))
    num_feat = {t: f[2][1] for t, f in feat.items()}
    nonzero_feat = {t: int(f[1].sum()) for t, f in feat.items()}

    def column_sums(mtxs):  # DecagonDataSet.py:276-292
        return [np.array(m.sum(axis=0)).squeeze() for m in mtxs]

    degrees = {0: column_sums(adj[0, 0]), 1: column_sums(adj[1, 1])}
    edge_types = {et: len(m) for et, m in adj.items()}
    edge_type2dim = {et: [m.shape for m in mtxs] for et, mtxs in adj.items()}
    dec = dict(DEFAULT_DECODERS if decoders is None else decoders)
    dec = {et: dec[et] for et in adj}
    return GraphInputs(adj, feat, num_feat, nonzero_feat, edge_types, degrees,
                       edge_type2dim, dec, {0: n_genes, 1: n_drugs})


def toy_graph(decoders=None, seed=0, features=None):
    """Config #1.  ``np.random.seed(seed)`` then the draws of ``main.py:134-156``."""
    return dummy_graph(500, 400, 3, decoders, seed, features)


def dummy_graph(n_genes=200, n_drugs=250, n_types=3, decoders=None, seed=0, features=None):
    """The reference's dummy-data generator (``DecagonDummyDataAdjacencyMatricesBuilder.py:36-67``; the same
    algorithm as ``main.py:134-156``): planted-partition PPI graph of ``n_genes // 10`` groups of 10, gene-drug
    edges where ``10 * randn > 15``, drug-drug relation t = pairs sharing exactly ``t + 4`` targets.  The defaults
    are ``NumProteins`` / ``NumDrugs`` / ``NumDrugDrugRelationTypes`` of the reference's ``configuration.json:6-8``
    (the run that produced ``decagon_iteration_results_0.csv``)."""
    import networkx as nx
    np.random.seed(seed)
    gene_net = nx.planted_partition_graph(n_genes // 10, 10, 0.2, 0.05, seed=42)
    gene_adj = sp.csr_matrix(nx.adjacency_matrix(gene_net))
    gene_drug_adj = sp.csr_matrix((10 * np.random.randn(n_genes, n_drugs) > 15).astype(int))
    shared = (gene_drug_adj.T @ gene_drug_adj).toarray()
    np.fill_diagonal(shared, -1)
    drug_drug = [sp.csr_matrix((shared == t + 4).astype(np.float64)) for t in range(n_types)]
    return assemble(gene_adj, gene_drug_adj, drug_drug, decoders, features=features)


def _unique_pairs(rng, n, p, count):
    """``count`` distinct unordered pairs (u < v), endpoints drawn i.i.d. from ``p``."""
    keys = np.empty(0, dtype=np.int64)
    while keys.size < count:
        u = rng.choice(n, size=int(count * 1.5) + 16, p=p)
        v = rng.choice(n, size=u.size, p=p)
        lo, hi = np.minimum(u, v), np.maximum(u, v)
        new = (lo.astype(np.int64) * n + hi)[lo != hi]
        merged = np.concatenate([keys, new])
        _, first = np.unique(merged, return_index=True)
        keys = merged[np.sort(first)]
    keys = keys[:count]
    return keys // n, keys % n


def _symmetric(n, u, v):
    data = np.ones(2 * u.size)
    return sp.csr_matrix((data, (np.concatenate([u, v]), np.concatenate([v, u]))), shape=(n, n))


def polypharmacy_graph(scale=1, n_types=964, decoders=None, seed=0,
                       n_proteins=19085, n_drugs=645, n_ppi=715612, n_targets=18596,
                       n_pairs=63473, n_ddi=4651131, min_size=500, max_size=28568, features=None):
    """Config #3 (``scale=1``) / #5 (``scale=10``): node and edge counts scale, the number of
    side-effect types does not."""
    rng = np.random.RandomState(seed)
    n0, n1 = n_proteins * scale, n_drugs * scale

    w = (np.arange(n0) + 1.0) ** -0.5  # Chung-Lu weights
    u, v = _unique_pairs(rng, n0, w / w.sum(), n_ppi * scale)
    gene_adj = _symmetric(n0, u, v)

    wd = (np.arange(n1) + 1.0) ** -0.7
    cnt = n_targets * scale
    keys = np.empty(0, dtype=np.int64)
    while keys.size < cnt:
        d = rng.choice(n1, size=int(cnt * 1.3), p=wd / wd.sum())
        p = rng.randint(0, n0, size=d.size)
        merged = np.concatenate([keys, p.astype(np.int64) * n1 + d])
        _, first = np.unique(merged, return_index=True)
        keys = merged[np.sort(first)]
    keys = keys[:cnt]
    gene_drug_adj = sp.csr_matrix((np.ones(cnt), (keys // n1, keys % n1)), shape=(n0, n1))

    base = n_pairs * scale * scale if scale > 1 else n_pairs
    base = min(base, n1 * (n1 - 1) // 2)
    wp = (np.arange(n1) + 1.0) ** -0.3
    bu, bv = _unique_pairs(rng, n1, wp / wp.sum(), base)
    sizes = np.exp(rng.uniform(np.log(min_size), np.log(max_size), size=n_types))
    sizes = sizes * (n_ddi * scale / sizes.sum())
    sizes = np.clip(np.round(sizes).astype(np.int64), min_size, base)
    drug_drug = []
    for c in sizes:
        pick = rng.choice(base, size=int(c), replace=False)
        drug_drug.append(_symmetric(n1, bu[pick], bv[pick]))
    return assemble(gene_adj, gene_drug_adj, drug_drug, decoders, features=features)
