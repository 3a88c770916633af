"""Datasets of the SEQ plugin family (reference REC/data/dataset/trainset.py:22-75, evalset.py:4-36,
collate_fn.py:6-32).  Output formats are the hot path's input formats:
   train: items int64 [2, L+1] (positives | aligned negatives, left-padded with 0), masked_index int64 [L]
   eval : (history, item_seq [L] left-padded, target)  ->  collate: item_seq [B,L], (history_u, history_i),
          positive_u [B], positive_i [B]
"""
import numpy as np
import torch
from torch.utils.data import Dataset


class SEQTrainDataset(Dataset):
    def __init__(self, config, dataload):
        self.dataload = dataload
        self.config = config
        self.item_num = dataload.item_num
        self.train_seq = dataload.train_feat["item_seq"]
        self.length = len(self.train_seq)
        self.max_seq_length = config["MAX_ITEM_LIST_LENGTH"] + 1
        W = self.max_seq_length
        # pre-padded copy so batches can be assembled without Python loops (sample_batch)
        self.padded = np.zeros((self.length, W), dtype=np.int64)
        self.lens = np.zeros(self.length, dtype=np.int64)
        for i, s in enumerate(self.train_seq):
            s = np.asarray(s)[-W:]
            self.padded[i, W - len(s):] = s
            self.lens[i] = len(s)

    def __len__(self):
        return self.length

    def _neg_sample(self, item_set, rng):
        item = rng.randint(1, self.item_num - 1)
        while item in item_set:
            item = rng.randint(1, self.item_num - 1)
        return item

    def __getitem__(self, index):
        """trainset.py:52-75: neg[t] pairs with pos[t]; neg[0] == 0; mask has len(seq)-1 ones."""
        import random
        seq = list(self.train_seq[index])
        W = self.max_seq_length
        sset = set(seq)
        negs = [self._neg_sample(sset, random) for _ in range(len(seq) - 1)]
        pad = lambda xs, n: ([0] * (n - len(xs)) + list(xs))[-n:]
        items = torch.tensor([pad(seq, W), pad(negs, W)], dtype=torch.long)
        mask = torch.tensor(pad([1] * (len(seq) - 1), W - 1), dtype=torch.long)
        return items, mask

    def sample_batch(self, indices, rng):
        """Vectorised equivalent of collating __getitem__ over `indices` (numpy Generator `rng`):
        uniform negatives on [1, item_num-1] rejected against the sequence's own items."""
        pos = self.padded[indices]                               # [B, W]
        B, W = pos.shape
        valid = pos != 0
        neg_valid = valid.copy()
        first = W - self.lens[indices]
        neg_valid[np.arange(B), np.minimum(first, W - 1)] = False   # neg[first valid slot] == 0 (no transition)
        neg = rng.integers(1, self.item_num, size=(B, W))
        for _ in range(100):
            clash = (neg[:, :, None] == pos[:, None, :]).any(-1) & neg_valid
            if not clash.any():
                break
            neg[clash] = rng.integers(1, self.item_num, size=int(clash.sum()))
        neg = np.where(neg_valid, neg, 0)
        items = np.stack([pos, neg], 1)
        mask = neg_valid[:, 1:].astype(np.int64)
        return torch.from_numpy(items), torch.from_numpy(mask)


class SeqEvalDataset(Dataset):
    def __init__(self, config, dataload, phase="valid"):
        self.dataload = dataload
        self.max_item_list_length = config["MAX_ITEM_LIST_LENGTH"]
        self.user_seq = list(dataload.user_seq.values())
        self.phase = phase
        self.length = len(self.user_seq)
        self.item_num = dataload.item_num

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        seq = self.user_seq[index]
        cut = -2 if self.phase == "valid" else -1
        history = np.asarray(seq[:cut])
        L = self.max_item_list_length
        tail = list(history[-L:])
        item_seq = [0] * (L - len(tail)) + tail
        return torch.as_tensor(history, dtype=torch.long), item_seq, int(seq[cut])


def seq_eval_collate(batch):
    hist = [b[0] for b in batch]
    history_u = torch.cat([torch.full_like(h, i) for i, h in enumerate(hist)])
    history_i = torch.cat(hist)
    item_seq = torch.tensor([b[1] for b in batch], dtype=torch.long)
    item_target = torch.tensor([b[2] for b in batch], dtype=torch.long)
    positive_u = torch.arange(item_seq.shape[0])
    return item_seq, (history_u, history_i), positive_u, item_target
