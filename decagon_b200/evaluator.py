"""``DecagonAccuracyEvaluator`` with the reference's call surface
(``main/AccuracyEvaluators/Tensorflow/DecagonAccuracyEvaluator.py:46-196``, ``main/Utils/MathUtils.py:3-4``).

``evaluate`` / ``evaluateAll`` return ``AccuracyScores(auroc, auprc, apk)``: sigmoid of the all-pairs
scores of a relation sampled at the positive / negative validation edges (``row * n_cols + col``,
``:169-186``), labels 1 / 0, ``sklearn.metrics.roc_auc_score`` / ``average_precision_score``;
``evaluateAll`` pools every drug-drug relation ``(1, 1, *)`` (``:57-91``), ``apk`` is 0 as in the reference.

Two execution paths give the same numbers:

* ``fast=False`` -- the reference's own sequence: ``session.run(predictions)`` per relation (a whole encoder
  forward and an ``[n_i, n_j]`` matrix to the host each time), host sigmoid, ``np.take``;
* ``fast=True`` (default) -- SURVEY.md 8(f) rank 1: ONE encoder forward, then only the sampled coordinates are
  scored on the device (gather + bilinear form + sigmoid), no score matrix is materialised.  ``evaluateAll``
  scores the pooled edges of every drug-drug relation in ONE launch (``dgn_evaluate_edges``) and, with
  ``device_metrics=True`` (default), sorts them on the device and evaluates sklearn's AUROC / AUPRC definitions
  there; ``device_metrics=False`` brings the scores back and calls sklearn like the reference.
"""
import collections
import math

import numpy as np
from sklearn import metrics

AccuracyScores = collections.namedtuple('AccuracyScores', ['auroc', 'auprc', 'apk'])


def sigmoid(x):
    """main/Utils/MathUtils.py:3-4"""
    return 1. / (1 + np.exp(-x))


class LossElementsContainer(object):
    def __init__(self, predictions, labels):
        self.predictions = predictions
        self.labels = labels

    @staticmethod
    def reduce(containers):
        containers = list(containers)
        preds = np.hstack([c.predictions for c in containers]) if containers else np.zeros(0)
        labels = np.hstack([c.labels for c in containers]) if containers else np.zeros(0)
        return LossElementsContainer(preds, labels)


class DecagonAccuracyEvaluator(object):
    def __init__(self, session, placeholdersDict, predictionsTensor, relCoordToIdx, config=None, fast=True,
                 device_metrics=True):
        self.session = session
        self.placeholdersDict = placeholdersDict
        self.predictionsTensor = predictionsTensor
        self.relCoordToIdx = relCoordToIdx
        apk = 0
        if config is not None:
            apk = config.getSetting('ApkRank') if hasattr(config, 'getSetting') else config
        self.apkRank = int(apk)
        self.fast = fast
        self.device_metrics = device_metrics
        self._forward_done = False

    # ------------------------------------------------------------------ reference surface
    def evaluateAll(self, feedDict, positiveEdgeSamples, negativeEdgeSamples):
        drug_rels = [rc for rc in self.relCoordToIdx.keys() if tuple(rc[:2]) == (1, 1)]
        self._forward_done = False
        if self.fast and drug_rels:
            return self._evaluateAllBatched(feedDict, drug_rels, positiveEdgeSamples, negativeEdgeSamples)
        parts = []
        for relCoord in drug_rels:
            self._updateFeedDictForEval(feedDict, relCoord)
            parts.append(self._computePredictions(feedDict, relCoord, positiveEdgeSamples, negativeEdgeSamples))
        self._forward_done = False
        loss = LossElementsContainer.reduce(parts)
        auroc = auprc = math.nan
        try:
            auroc = metrics.roc_auc_score(loss.labels, loss.predictions)
        except ValueError:
            pass
        try:
            auprc = metrics.average_precision_score(loss.labels, loss.predictions)
        except ValueError:
            pass
        return AccuracyScores(auroc, auprc, 0)

    def evaluate(self, feedDict, relCoord, positiveEdgeSamples, negativeEdgeSamples):
        self._updateFeedDictForEval(feedDict, relCoord)
        self._forward_done = False
        loss = self._computePredictions(feedDict, relCoord, positiveEdgeSamples, negativeEdgeSamples)
        self._forward_done = False
        auroc = metrics.roc_auc_score(loss.labels, loss.predictions)
        auprc = metrics.average_precision_score(loss.labels, loss.predictions)
        return AccuracyScores(auroc, auprc, 0)

    # ------------------------------------------------------------------ internals
    def _evaluateAllBatched(self, feedDict, drug_rels, positiveEdgeSamples, negativeEdgeSamples):
        """One encoder forward, one scoring launch for every (1, 1, k), pooled in the reference's order
        (per relation: positives then negatives, ``:115-149``)."""
        self._updateFeedDictForEval(feedDict, drug_rels[0])
        eng = self._engine(feedDict)
        self._forward_done = False
        ks, edges, labels = [], [], []
        for rc in drug_rels:
            pos, neg = self._samples(positiveEdgeSamples, rc), self._samples(negativeEdgeSamples, rc)
            edges += [pos, neg]
            ks.append(np.full(len(pos) + len(neg), rc[2], dtype=np.int32))
            labels += [np.ones(len(pos), dtype=np.uint8), np.zeros(len(neg), dtype=np.uint8)]
        ks, edges, labels = np.concatenate(ks), np.concatenate(edges), np.concatenate(labels)
        if self.device_metrics:
            _, auroc, auprc = eng.evaluate_edges((1, 1), ks, edges, labels, sigmoid=True, want_scores=False)
            return AccuracyScores(auroc, auprc, 0)
        scores, _, _ = eng.evaluate_edges((1, 1), ks, edges, None, sigmoid=True)
        auroc = auprc = math.nan
        try:
            auroc = metrics.roc_auc_score(labels, scores)
        except ValueError:
            pass
        try:
            auprc = metrics.average_precision_score(labels, scores)
        except ValueError:
            pass
        return AccuracyScores(auroc, auprc, 0)

    def _engine(self, feedDict):
        """The CUDA engine behind the predictions tensor; runs the encoder once per evaluate / evaluateAll."""
        model = self.predictionsTensor.owner.model
        eng = self.session._engine(model, feedDict)
        if not eng._initialized:
            self.session._initialize(model)
        if not self._forward_done:
            eng.forward(0.0, self.session.seed, self.session.step)
            self._forward_done = True
        return eng

    def _computePredictions(self, feedDict, relCoord, positiveEdgeSamples, negativeEdgeSamples):
        if self.fast:
            eng = self._engine(feedDict)
            r = self.relCoordToIdx[relCoord]
            pos = eng.predict_edges(r, self._samples(positiveEdgeSamples, relCoord), sigmoid=True)
            neg = eng.predict_edges(r, self._samples(negativeEdgeSamples, relCoord), sigmoid=True)
        else:
            decoderOutput = self.session.run(self.predictionsTensor, feed_dict=feedDict)
            predictions = sigmoid(decoderOutput)
            pos = self._getSampledPredictions(predictions, positiveEdgeSamples, relCoord)
            neg = self._getSampledPredictions(predictions, negativeEdgeSamples, relCoord)
        return LossElementsContainer(predictions=np.hstack([pos, neg]),
                                     labels=np.hstack([np.ones(len(pos)), np.zeros(len(neg))]))

    @staticmethod
    def _samples(edgeSamples, relCoord):
        two_dim = np.asarray(edgeSamples[tuple(relCoord[:2])][relCoord[2]])
        if not np.issubdtype(two_dim.dtype, np.integer):
            two_dim = two_dim.astype(np.int64)
        return two_dim.reshape(-1, 2)

    def _getSampledPredictions(self, predictions, edgeSamples, relCoord):
        linear = self._linearizeSampleIdxs(edgeSamples, relCoord, predictions.shape[1])
        return np.take(predictions, linear)  # == predictions.ravel()[linear]

    def _linearizeSampleIdxs(self, edgeSamples, relCoord, numCols):
        two_dim = self._samples(edgeSamples, relCoord)
        return (two_dim[:, 0] * numCols) + two_dim[:, 1]

    def _updateFeedDictForEval(self, feedDict, relCoord):
        feedDict[self.placeholdersDict['dropout']] = 0
        feedDict[self.placeholdersDict['batch_edge_type_idx']] = self.relCoordToIdx[relCoord]
        feedDict[self.placeholdersDict['batch_row_edge_type']] = relCoord[0]
        feedDict[self.placeholdersDict['batch_col_edge_type']] = relCoord[1]
