"""Top-k ranking metrics, numpy, per-rank SUMS over users exactly like the reference
(REC/evaluator/metrics.py:115-178, base_metric.py:43-67).  Registered names mirror register.py."""
import numpy as np

from ..utils.enum_type import EvaluatorType


def recall_(pos_index, pos_len):
    return np.cumsum(pos_index, axis=1) / pos_len.reshape(-1, 1)


def ndcg_(pos_index, pos_len):
    n, K = pos_index.shape
    ranks = np.tile(np.arange(1, K + 1, dtype=np.float64), (n, 1))
    idcg_len = np.minimum(pos_len, K)
    idcg = np.cumsum(1.0 / np.log2(ranks + 1), axis=1)
    for row, m in enumerate(idcg_len):
        idcg[row, m:] = idcg[row, m - 1]
    dcg = np.cumsum(np.where(pos_index, 1.0 / np.log2(ranks + 1), 0), axis=1)
    return dcg / idcg


def hit_(pos_index, pos_len):
    return (np.cumsum(pos_index, axis=1) > 0).astype(int)


def mrr_(pos_index, pos_len):
    idxs = pos_index.argmax(axis=1)
    result = np.zeros_like(pos_index, dtype=np.float64)
    for row, idx in enumerate(idxs):
        if pos_index[row, idx] > 0:
            result[row, idx:] = 1 / (idx + 1)
    return result


def precision_(pos_index, pos_len):
    return pos_index.cumsum(axis=1) / np.arange(1, pos_index.shape[1] + 1)


METRICS = {"recall": recall_, "ndcg": ndcg_, "hit": hit_, "mrr": mrr_, "precision": precision_}
metric_types = {k: EvaluatorType.RANKING for k in METRICS}
smaller_metrics = ["rmse", "mae", "logloss", "averagepopularity", "giniindex"]
