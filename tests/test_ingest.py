"""SNAP-CSV ingest (decagon_b200/ingest.py) against the reference's own parsers, which were run UNMODIFIED on the
committed CSV files by tests/golden/make_golden_ingest.py (node lists, relation order, every matrix)."""
import os

import numpy as np
import scipy.sparse as sp

from decagon_b200 import ingest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SNAP = os.path.join(GOLDEN, 'snap')


def _same_matrix(m, coords, values, shape):
    m = sp.coo_matrix(m)
    order = np.lexsort((m.col, m.row))
    assert tuple(m.shape) == tuple(shape)
    assert np.array_equal(np.stack([m.row[order], m.col[order]], axis=1), coords)
    assert np.array_equal(np.asarray(m.data, dtype=np.float64)[order], values)


def test_format_id_follows_the_reference_including_its_quirk():
    assert ingest.format_id('CID000012314') == 12314 and ingest.format_id('C0051234') == 51234
    assert ingest.format_id('SID123') == 123 and ingest.format_id('7157') == 7157
    assert ingest.format_id('CID000004170') == 0 and ingest.format_id('70') == 0 and ingest.format_id('0') == 0


def test_ingest_matches_the_reference_parsers():
    g = np.load(os.path.join(GOLDEN, 'snap_ingest.npz'))
    data = ingest.load_public_data(os.path.join(SNAP, 'combo.csv'), os.path.join(SNAP, 'ppi.csv'),
                                   os.path.join(SNAP, 'targets.csv'), os.path.join(SNAP, 'mono.csv'))
    assert np.array_equal(data.proteins, g['proteins']) and np.array_equal(data.drugs, g['drugs'])
    assert 0 in data.drugs and 0 in data.proteins            # the ids ending in 0
    assert np.array_equal(data.relation_ids, g['relation_ids'])   # order of the relation matrices = flat relation index
    assert len(data.relation_ids) == 3                        # 499- and 120-edge types are dropped (>= 500 rule)
    for i, m in enumerate(data.drug_drug):
        _same_matrix(m, g['dd%d_coords' % i], g['dd%d_values' % i], g['dd%d_shape' % i])
        assert (m != m.T).nnz == 0
    _same_matrix(data.drug_protein, g['dp_coords'], g['dp_values'], g['dp_shape'])
    _same_matrix(data.ppi, g['ppi_coords'], g['ppi_values'], g['ppi_shape'])
    _same_matrix(data.protein_features, g['feat_protein_coords'], g['feat_protein_values'], g['feat_protein_shape'])
    _same_matrix(data.drug_features, g['feat_drug_coords'], g['feat_drug_values'], g['feat_drug_shape'])


def test_ingested_graph_feeds_the_iterator():
    """The matrices go straight into the drop-in surface: dict order, transposed twins, multi-hot drug features."""
    from decagon_b200.deep.minibatch import EdgeMinibatchIterator
    data = ingest.load_public_data(os.path.join(SNAP, 'combo.csv'), os.path.join(SNAP, 'ppi.csv'),
                                   os.path.join(SNAP, 'targets.csv'), os.path.join(SNAP, 'mono.csv'))
    inputs = ingest.graph_inputs(data)
    assert list(inputs.edge_types.items()) == [((0, 0), 2), ((0, 1), 1), ((1, 1), 6), ((1, 0), 1)]
    assert inputs.num_feat == {0: len(data.proteins), 1: len(data.side_effects)}
    np.random.seed(0)
    it = EdgeMinibatchIterator(inputs.adj_mats, inputs.feat, inputs.edge_types, {}, batch_size=64, val_test_size=0.05)
    assert it.num_edge_types == 10 and len(it.train_edges[1, 1][0]) > 64


def test_relation_order_is_the_multigraph_walk():
    """A hand-checkable case: the walk visits node a's edges first (a was inserted first), in the order its
    neighbours were first connected to it, so relation 7 (line 3, a-c) precedes relation 9 (line 2, b-c)."""
    a, b, c = 11, 22, 33
    u = np.array([a, b, b, a])
    v = np.array([b, a, c, c])
    rel = np.array([5, 5, 9, 7])
    order, counts = ingest.relation_order(u, v, rel)
    assert list(order) == [5, 7, 9] and list(counts) == [2, 1, 1]
