"""Collects per-batch top-k hit matrices for the evaluator (reference REC/evaluator/collector.py:115-160:
torch.topk(scores, max(topk)) -> hit matrix against (positive_u, positive_i) -> cat on CPU).
`eval_batch_collect_topk` takes top-k indices directly, so a fused score+top-k kernel never has to
materialise the [B_e, N] score matrix."""
import torch


class DataStruct:
    def __init__(self):
        self._d = {}

    def get(self, name):
        if name not in self._d:
            raise IndexError("Can not load the data without registration !")
        return self._d[name]

    def set(self, name, value):
        self._d[name] = value

    def __contains__(self, k):
        return k in self._d

    def update_tensor(self, name, value):
        v = value.detach().cpu().clone()
        self._d[name] = v if name not in self._d else torch.cat((self._d[name], v), dim=0)


class Collector:
    def __init__(self, config):
        self.config = config
        self.topk = config["topk"]
        self.data_struct = DataStruct()

    def eval_batch_collect(self, scores_tensor, positive_u, positive_i):
        _, topk_idx = torch.topk(scores_tensor, max(self.topk), dim=-1)
        self.eval_batch_collect_topk(topk_idx, positive_u, positive_i)

    def eval_batch_collect_topk(self, topk_idx, positive_u, positive_i):
        n = topk_idx.shape[0]
        positive_u = positive_u.to(topk_idx.device)
        positive_i = positive_i.to(topk_idx.device)
        pos_idx = torch.zeros_like(topk_idx, dtype=torch.int)
        hit = (topk_idx[positive_u] == positive_i.view(-1, 1)).int()
        pos_idx.index_put_((positive_u,), hit, accumulate=True)
        pos_len = torch.zeros(n, dtype=torch.int, device=topk_idx.device)
        pos_len.index_put_((positive_u,), torch.ones_like(positive_u, dtype=torch.int), accumulate=True)
        self.data_struct.update_tensor("rec.topk", torch.cat((pos_idx.clamp_(max=1), pos_len.view(-1, 1)), dim=1))

    def get_data_struct(self):
        out = self.data_struct
        self.data_struct = DataStruct()
        return out
