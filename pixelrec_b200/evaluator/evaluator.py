"""Evaluator: metric name -> per-rank sums (reference REC/evaluator/evaluator.py:6-32)."""
from collections import OrderedDict

import torch

from .metrics import METRICS


class Evaluator:
    def __init__(self, config):
        self.config = config
        self.metrics = [m.lower() for m in config["metrics"]]
        self.topk = config["topk"]

    def evaluate(self, dataobject):
        rec = dataobject.get("rec.topk")
        topk_idx, pos_len = torch.split(rec, [max(self.topk), 1], dim=1)
        pos_index = topk_idx.to(torch.bool).numpy()
        pos_len = pos_len.squeeze(-1).numpy()
        result = OrderedDict()
        for m in self.metrics:
            val = METRICS[m](pos_index, pos_len).sum(axis=0)
            for k in self.topk:
                result[f"{m}@{k}"] = val[k - 1]
        return result
