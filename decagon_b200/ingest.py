"""SNAP / Decagon public-data CSV files -> the sparse matrices of the hot path, without networkx.

Reference (the callers on the data side of the path):

* ``main/DataSetParsers/NodeLists/DecagonPublicDataNodeListsBuilder.py:37-78``  -- sorted node lists
* ``main/DataSetParsers/AdjacencyMatrices/DecagonPublicDataAdjacencyMatricesBuilder.py:54-152`` -- one symmetric 0/1
  matrix per side-effect type with at least 500 edges, the protein x drug target matrix, the PPI matrix
* ``main/DataSetParsers/NodeFeatures/DecagonPublicDataNodeFeaturesBuilder.py:34-79`` -- identity protein features,
  multi-hot drug features (one column per mono side effect)
* ``main/Dtos/NodeIds.py:27-45`` -- the id normalisation, INCLUDING its quirk: every id whose text ends in ``0``
  becomes node 0.  It is reproduced because node lists, matrix shapes and relation order all depend on it.

The reference reads the files with ``networkx.read_edgelist`` and walks ``MultiGraph.edges`` edge by edge in Python
(minutes for the 4.6 M rows of ``bio-decagon-combo.csv``); the ORDER of the relation matrices -- hence the flat
relation index r of the whole model -- is the order in which that walk first meets each side-effect type.  Here the
files are parsed once into integer arrays and the same order is obtained with sorts: an edge (u, v, key) is visited
at position (insertion rank of its earlier-inserted endpoint n, first line on which the pair {n, nbr} occurs, line),
which is what ``MultiEdgeView.__iter__`` yields.  Outputs are bit-identical to the reference's parsers
(``tests/golden/make_golden_ingest.py`` runs them unmodified; ``tests/test_ingest.py``).

File formats (the preprocessed files the reference's ``configuration.json:10-13`` points at): combo
``CID,CID,C<side effect>``; ppi ``gene,gene``; targets ``CID,gene`` (either order); mono
``STITCH,Individual Side Effect,Side Effect Name`` with ONE header line.  Lines starting with ``#`` are comments
(``read_edgelist``).
"""
import csv
from collections import namedtuple

import numpy as np
import scipy.sparse as sp

from . import datasets

MIN_EDGES_PER_RELATION = 500  # DecagonPublicDataAdjacencyMatricesBuilder._isEdgeListValid (:122-123)

PublicData = namedtuple('PublicData', ['proteins', 'drugs', 'relation_ids', 'drug_drug', 'drug_protein', 'ppi',
                                       'protein_features', 'drug_features', 'side_effects'])


def format_id(text):
    """``BaseNodeId._formatStr`` + ``int`` (``NodeIds.py:27-45``): ``CID000012314 -> 12314``, ``C0051234 -> 51234``;
    any id ending in ``0`` -> 0 (sic)."""
    if text == '0' or text[-1] == '0':
        return 0
    digits = ''.join(ch for ch in text if ch.isdigit()).lstrip('0')
    return int(digits)


def _format_ids(texts):
    """format_id over an array of strings, one evaluation per distinct string."""
    uniq, inverse = np.unique(np.asarray(texts, dtype=object).astype(str), return_inverse=True)
    return np.array([format_id(t) for t in uniq], dtype=np.int64)[inverse]


def _read_columns(path, n_cols, skip_header=False):
    cols = [[] for _ in range(n_cols)]
    with open(path) as f:
        if skip_header:
            next(f)
        for line in f:
            line = line.strip()
            if not line or line[0] == '#':
                continue
            parts = line.split(',')
            for c in range(n_cols):
                cols[c].append(parts[c])
    return [np.array(c, dtype=object) for c in cols]


def _first_rank(values):
    """Rank of every distinct value by first appearance in ``values`` (insertion order of a dict)."""
    uniq, first = np.unique(values, return_index=True)
    order = np.argsort(first, kind='stable')
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[order] = np.arange(len(uniq))
    return uniq, rank


def _symmetric_binary(n, rows, cols):
    """``nx.adjacency_matrix`` of an undirected simple graph: 1 at (u, v) and (v, u), duplicates collapse."""
    r = np.concatenate([rows, cols])
    c = np.concatenate([cols, rows])
    m = sp.csr_matrix((np.ones(len(r)), (r, c)), shape=(n, n))
    m.data[:] = 1.0
    m.sort_indices()
    return m


def relation_order(u, v, rel):
    """Side-effect types in the order ``MultiGraph.edges`` first meets them
    (``DecagonPublicDataAdjacencyMatricesBuilder._buildAllEdgeSets``, :88-96), and for every type the number of
    multi-edges the walk yields (duplicated rows count, as they do in the reference's ``len(edgeList) >= 500``)."""
    n_lines = len(u)
    flat = np.stack([u, v], axis=1).ravel()
    nodes, rank = _first_rank(flat)                     # node insertion order: u then v, line by line
    ru, rv = rank[np.searchsorted(nodes, u)], rank[np.searchsorted(nodes, v)]
    lo, hi = np.minimum(ru, rv), np.maximum(ru, rv)     # the walk yields an edge from its earlier-inserted endpoint
    pair = lo * (len(nodes) + 1) + hi
    _, inverse, = np.unique(pair, return_inverse=True)
    first_line = np.full(inverse.max() + 1, n_lines, dtype=np.int64)
    np.minimum.at(first_line, inverse, np.arange(n_lines))
    visit = np.lexsort((np.arange(n_lines), first_line[inverse], lo))
    rels, rel_rank = _first_rank(rel[visit])
    order = rels[np.argsort(rel_rank, kind='stable')]
    counts = np.array([(rel == r).sum() for r in order], dtype=np.int64)
    return order, counts


def load_public_data(combo_path, ppi_path, targets_path, mono_path=None, min_edges=MIN_EDGES_PER_RELATION):
    """Parse the four files.  Returns ``PublicData``: sorted node lists, the kept side-effect types in the
    reference's order, one symmetric csr matrix per type [n_drugs, n_drugs], the target matrix
    [n_proteins, n_drugs], the PPI matrix [n_proteins, n_proteins] and the feature matrices."""
    d1, d2, se = _read_columns(combo_path, 3)
    u, v = _format_ids(d1), _format_ids(d2)
    rel = np.array([int(s[1:]) for s in np.unique(se.astype(str))], dtype=np.int64)[np.unique(se.astype(str), return_inverse=True)[1]]
    g1, g2 = _read_columns(ppi_path, 2)
    p1, p2 = _format_ids(g1), _format_ids(g2)
    t1, t2 = _read_columns(targets_path, 2)
    t1s, t2s = t1.astype(str), t2.astype(str)
    first_is_drug = np.char.startswith(t1s, 'CID')
    second_is_drug = np.char.startswith(t2s, 'CID')
    # node lists (DecagonPublicDataNodeListsBuilder.py:44-78): every node of the target file counts, edge by edge
    target_drugs = np.concatenate([_format_ids(t1s[first_is_drug]), _format_ids(t2s[second_is_drug])])
    target_proteins = np.concatenate([_format_ids(t1s[~first_is_drug]), _format_ids(t2s[~second_is_drug])])
    drugs = np.unique(np.concatenate([u, v, target_drugs]))
    proteins = np.unique(np.concatenate([p1, p2, target_proteins]))

    # drug-drug relation matrices
    order, counts = relation_order(u, v, rel)
    keep = order[counts >= min_edges]
    ui, vi = np.searchsorted(drugs, u), np.searchsorted(drugs, v)
    drug_drug = [_symmetric_binary(len(drugs), ui[rel == r], vi[rel == r]) for r in keep]

    # protein x drug targets (_buildDrugProteinRelationMtx, :125-133; _extractDrugProtein: the CID end is the drug)
    ok = first_is_drug != second_is_drug
    drug_txt = np.where(first_is_drug, t1s, t2s)[ok]
    prot_txt = np.where(first_is_drug, t2s, t1s)[ok]
    di, pi = np.searchsorted(drugs, _format_ids(drug_txt)), np.searchsorted(proteins, _format_ids(prot_txt))
    drug_protein = sp.csr_matrix((np.ones(len(di)), (pi, di)), shape=(len(proteins), len(drugs)))
    drug_protein.data[:] = 1.0
    drug_protein.sort_indices()

    ppi = _symmetric_binary(len(proteins), np.searchsorted(proteins, p1), np.searchsorted(proteins, p2))

    protein_features = sp.identity(len(proteins), format='csr')
    drug_features, side_effects = None, None
    if mono_path is not None:
        rows = []
        with open(mono_path) as f:
            reader = csv.reader(f)
            next(reader)  # header (DecagonPublicDataNodeFeaturesBuilder.py:66-67)
            for row in reader:
                rows.append((row[0], row[1]))
        fd = _format_ids(np.array([r[0] for r in rows], dtype=object))
        fs = _format_ids(np.array([r[1] for r in rows], dtype=object))
        side_effects = np.unique(fs)                   # every side effect of the file, also of drugs that are no node
        known = np.isin(fd, drugs)
        drug_features = sp.csr_matrix((np.ones(known.sum()), (np.searchsorted(drugs, fd[known]),
                                                              np.searchsorted(side_effects, fs[known]))),
                                      shape=(len(drugs), len(side_effects)))
        drug_features.data[:] = 1.0
        drug_features.sort_indices()
    return PublicData(proteins, drugs, keep, drug_drug, drug_protein, ppi, protein_features, drug_features, side_effects)


def graph_inputs(data, decoders=None, transpose=True, use_drug_features=True):
    """``PublicData`` -> the dicts the iterator / model / optimizer take (``datasets.assemble``: dict order
    (0,0), (0,1), (1,1), (1,0), transposed twins appended as ``DecagonDataSet._augmentAdjMtxDictWithTranspose``)."""
    features = {1: data.drug_features} if use_drug_features and data.drug_features is not None else None
    return datasets.assemble(data.ppi, data.drug_protein, data.drug_drug, decoders, transpose=transpose, features=features)
