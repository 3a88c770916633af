"""Interaction data -> per-user sequences, the feeder of the hot path.

`Data` follows the reference's REC/data/dataload.py:16-150: `<data_path>/<dataset>.csv` with columns
item_id,user_id,timestamp (:30-40); ids re-mapped with [PAD] = 0 so item_num = #items + 1 (:42-63);
build(): stable sort by timestamp, per-user sequences, the last two interactions held out for
valid/test (:66-91); training windows of MAX_ITEM_LIST_LENGTH + 1 items, a longer history is cut into
consecutive windows after dropping its oldest (n mod (L+1)) items (:103-150).

`SyntheticData` produces the same object for benchmark shapes: long-tail item
popularity p(r) ~ 1/(r+c)^0.8 with ids randomly permuted, sequence lengths uniform in [3, L+1] (+2 held out).
"""
import os
from logging import getLogger

import numpy as np

from ..utils.enum_type import InputType


def split_windows(seq, window):
    """dataload.py:125-136: one window if len <= window, else drop the oldest (n % window) items and chunk."""
    n = len(seq)
    if n > window:
        off = n % window
        return [seq[i:i + window] for i in range(off, n, window)]
    return [seq]


class Data:
    def __init__(self, config):
        self.config = config
        self.dataset_path = config["data_path"]
        self.dataset_name = config["dataset"]
        self.logger = getLogger()
        self.uid_field, self.iid_field = "user_id", "item_id"
        self.user_seq = None
        self.train_feat = None
        self._load()

    def _load(self):
        import pandas as pd
        path = os.path.join(self.dataset_path, f"{self.dataset_name}.csv")
        if not os.path.isfile(path):
            raise ValueError(f"File {path} not exist.")
        df = pd.read_csv(path, delimiter=",", dtype={"item_id": str, "user_id": str, "timestamp": int}, header=0,
                         names=["item_id", "user_id", "timestamp"])
        self.id2token, self.token2id = {}, {}
        cols = {}
        for feat in ("user_id", "item_id"):
            codes, uniques = pd.factorize(df[feat])
            mp = np.array(["[PAD]"] + list(uniques))
            self.id2token[feat] = mp
            self.token2id[feat] = {t: i for i, t in enumerate(mp)}
            cols[feat] = codes.astype(np.int64) + 1
        cols["timestamp"] = df["timestamp"].values
        self.inter_feat = cols
        self.user_num = len(self.id2token["user_id"])
        self.item_num = len(self.id2token["item_id"])
        self.inter_num = len(df)

    def build(self):
        # dataload.py:66-76: sort by timestamp, then group by user IN ORDER OF FIRST APPEARANCE (dict insertion order of
        # _grouped_index) -- the order of user_seq is the order of the eval users and of the training windows, i.e. what a seeded
        # sampler permutes; tests/golden/dataload_ref.npz pins it to the reference.  (The reference sorts a DataFrame with pandas'
        # default unstable sort: equal timestamps have no defined order there; here they keep file order.)
        order = np.argsort(self.inter_feat["timestamp"], kind="stable")
        users = self.inter_feat["user_id"][order]
        items = self.inter_feat["item_id"][order]
        uorder = np.argsort(users, kind="stable")          # group by user, keep time order inside a user
        us, it = users[uorder], items[uorder]
        bounds = np.flatnonzero(np.r_[True, us[1:] != us[:-1], True])
        first_pos = uorder[bounds[:-1]]                    # where (in time order) each user appears first
        self.user_seq = {}
        for gi in np.argsort(first_pos, kind="stable"):
            s, e = bounds[gi], bounds[gi + 1]
            self.user_seq[int(us[s])] = it[s:e]
        self._build_train()

    def _build_train(self):
        window = self.config["MAX_ITEM_LIST_LENGTH"] + 1
        uid_list, seqs = [], []
        for uid, seq in self.user_seq.items():
            tr = seq[:-2]
            if len(tr) == 0:
                continue
            if self.config["MODEL_INPUT_TYPE"] in (InputType.SEQ, None):
                for w in split_windows(tr, window):
                    uid_list.append(uid)
                    seqs.append(np.asarray(w))
            else:
                raise NotImplementedError("only InputType.SEQ models are on the hot path (SURVEY section 2)")
        self.train_feat = {"user_id": np.array(uid_list), "item_seq": seqs}

    def __str__(self):
        return f"{self.dataset_name}: users {self.user_num - 1}, items {self.item_num - 1}, interactions {self.inter_num}"


class SyntheticData(Data):
    """Pixel200K-shaped synthetic interactions (no CSV).  config keys: synthetic_users, synthetic_items,
    seed; sequences get 2 extra held-out interactions so valid/test exist."""

    def __init__(self, config):
        self.config = config
        self.dataset_name = "synthetic"
        self.logger = getLogger()
        self.uid_field, self.iid_field = "user_id", "item_id"
        n_users = int(config["synthetic_users"] or 2000)
        n_items = int(config["synthetic_items"] or 10000)
        L = config["MAX_ITEM_LIST_LENGTH"]
        g = np.random.default_rng(config["seed"] or 2020)
        self.user_num, self.item_num = n_users + 1, n_items + 1
        r = np.arange(1, n_items + 1, dtype=np.float64)
        p = 1.0 / (r + 10.0) ** 0.8
        p /= p.sum()
        perm = g.permutation(n_items) + 1
        lens = g.integers(3, L + 2, size=n_users) + 2
        flat = perm[g.choice(n_items, size=int(lens.sum()), p=p)].astype(np.int64)
        offs = np.r_[0, np.cumsum(lens)]
        self.user_seq = {u + 1: flat[offs[u]:offs[u + 1]] for u in range(n_users)}
        self.inter_num = int(lens.sum())
        self.train_feat = None

    def build(self):
        self._build_train()
