"""The device-side pieces of the reference's active-learning loop.

``RandomMaskingActiveLearner`` (``main/ActiveLearner/RandomMaskingActiveLearner.py:150-200``) keeps a 0/1 mask per
drug-drug relation and un-masks ``floor(dataSetSize * (2^i - 2^(i-1)) / 100)`` candidate coordinates per iteration;
``GreedyActiveLearner`` (``main/ActiveLearner/GreedyActiveLearner.py:68-92``) picks them by rank: sigmoid of the
all-pairs ``predictions`` of relation (1, 1, 0), ``np.take`` at the candidates' ``row * n_cols + col``, ``np.argsort``
descending.  Here the candidates are scored and sorted on the device (``dgn_rank_edges``) -- no ``[n, n]`` matrix,
no host argsort over every remaining possibility.
"""
import numpy as np


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
