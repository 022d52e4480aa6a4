"""The device-side pieces of the reference's active-learning loop.

``RandomMaskingActiveLearner`` (``main/ActiveLearner/RandomMaskingActiveLearner.py:150-200``) keeps a 0/1 mask per
drug-drug relation and un-masks ``floor(dataSetSize * (2^i - 2^(i-1)) / 100)`` candidate coordinates per iteration;
``GreedyActiveLearner`` (``main/ActiveLearner/GreedyActiveLearner.py:68-92``) picks them by rank: sigmoid of the
all-pairs ``predictions`` of relation (1, 1, 0), ``np.take`` at the candidates' ``row * n_cols + col``, ``np.argsort``
descending.  Here the candidates are scored and sorted on the device (``dgn_rank_edges``) -- no ``[n, n]`` matrix,
no host argsort over every remaining possibility.
"""
import numpy as np
import scipy.sparse as sp


def num_to_unmask(data_set_size, num_iters):
    """``RandomMaskingActiveLearner._updateMask`` (``:166-171``): how many candidates iteration ``num_iters`` reveals."""
    last = 2 ** (num_iters - 1) if num_iters > 0 else 0
    this = min(2 ** num_iters, 100)
    return int(np.floor(data_set_size * ((this - last) / 100)))


class GreedyCandidateRanker(object):
    """``possibilities``: int array ``[n, 3]`` of ``(relation id, row, col)`` as the reference keeps them
    (``_getPossibilitiesAndTestEdges``, ``:44-75``).  ``ranking_relation``: flat index of the relation whose
    predictions rank them -- the reference always uses (1, 1, 0) (``GreedyActiveLearner._updateFeedDict``, ``:94-98``)."""

    def __init__(self, engine, possibilities, ranking_relation):
        self.engine = engine
        self.possibilities = np.asarray(possibilities).reshape(-1, 3)
        self.ranking_relation = int(ranking_relation)

    def ranked_possibilities(self, top=None):
        """``_getRankedPossibilities``: indices into ``possibilities``, best first (ties keep their order)."""
        order, _ = self.engine.rank_edges(self.ranking_relation, self.possibilities[:, 1:3], top=top, sigmoid=True)
        return order.astype(np.int64)

    def get_new_sample_idxs(self, num):
        """``_getNewSampleIdxs`` (``:68-82``) after the first iteration: the ``num`` best candidates."""
        return self.ranked_possibilities(top=num)

    def unmask(self, masks, idxs):
        """``_updateMask`` (``:173-177``): set the chosen coordinates in the per-relation masks and drop them from the
        candidate list."""
        for rel, row, col in self.possibilities[idxs]:
            masks[rel][row, col] = 1
        self.possibilities = np.delete(self.possibilities, idxs, axis=0)
        return masks


class SparseRelationMasks(object):
    """``RandomMaskingActiveLearner.adjMtxMasks`` + ``_applyMask`` (``RandomMaskingActiveLearner.py:24-27,173-200``)
    without the dense arrays: the reference keeps one dense ``n x n`` float mask per relation (964 x 645^2 x 8 B =
    3.2 GB at the polypharmacy shape) and multiplies it into ``mtx.toarray()`` every round, although a mask entry
    only matters where the adjacency is non-zero.  Here a relation is its canonical CSR plus ONE keep bit per
    non-zero; ``unmask`` looks the coordinates up with a sorted search, ``apply`` returns exactly the matrices
    ``RelationCsrMatrix(np.multiply(mask, mtx.toarray()))`` would hold.  Feeding the resulting tuples re-uploads
    only the drug-drug group: ``dgn_graph_finalize`` rebuilds the groups whose relations changed and nothing else."""

    def __init__(self, matrices):
        self.full, self.keep, self._lin = {}, {}, {}
        for rel, m in matrices.items():
            m = sp.csr_matrix(m).copy()
            m.eliminate_zeros()
            m.sort_indices()
            self.full[rel] = m
            self.keep[rel] = np.zeros(m.nnz, dtype=bool)
            rows = np.repeat(np.arange(m.shape[0], dtype=np.int64), np.diff(m.indptr))
            self._lin[rel] = rows * m.shape[1] + m.indices

    def unmask(self, coords):
        """``coords``: int ``[n, 3]`` rows ``(relation id, row, col)`` (``self.possibilities[idxsToUnmask]``)."""
        coords = np.asarray(coords).reshape(-1, 3)
        for rel in np.unique(coords[:, 0]):
            sel = coords[coords[:, 0] == rel]
            lin = sel[:, 1].astype(np.int64) * self.full[rel].shape[1] + sel[:, 2]
            nz = self._lin[rel]
            pos = np.searchsorted(nz, lin)
            hit = (pos < len(nz)) & (nz[np.minimum(pos, len(nz) - 1)] == lin) if len(nz) else np.zeros(len(lin), bool)
            self.keep[rel][pos[hit]] = True

    def apply(self):
        """{relation id: csr} of the un-masked part of every relation."""
        out = {}
        for rel, m in self.full.items():
            k = self.keep[rel]
            counts = np.add.reduceat(k.astype(np.int64), m.indptr[:-1]) if m.nnz else np.zeros(m.shape[0], np.int64)
            counts[np.diff(m.indptr) == 0] = 0
            indptr = np.concatenate([[0], np.cumsum(counts)]).astype(m.indptr.dtype)
            out[rel] = sp.csr_matrix((m.data[k], m.indices[k], indptr), shape=m.shape)
        return out
