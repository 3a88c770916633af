"""load_data / bulid_dataloader with the reference's plugin registration table for the SEQ family
(REC/data/utils.py:15-114; the misspelt name is the reference's public name and is kept)."""
import math
import random
from functools import partial
from logging import getLogger

import numpy as np
import torch
from torch.utils.data import DataLoader

from ..utils.utils import get_rank, get_world_size
from .dataload import Data, SyntheticData
from .dataset import SEQTrainDataset, SeqEvalDataset, seq_eval_collate

DATASET_TABLE = {   # model -> (train set, eval set, eval collate)   (data/utils.py:24-53, hot-path rows)
    "SASRec": (SEQTrainDataset, SeqEvalDataset, seq_eval_collate),
    "GRU4Rec": (SEQTrainDataset, SeqEvalDataset, seq_eval_collate),
}


def load_data(config):
    if config["dataset"] == "synthetic" or config["synthetic_users"]:
        return SyntheticData(config)
    return Data(config)


class NonConsecutiveSequentialDistributedSampler(torch.utils.data.Sampler):
    """rank r evaluates users r, r+W, r+2W, ... (data/utils.py:134-159)."""

    def __init__(self, dataset, rank=None, num_replicas=None):
        self.dataset = dataset
        self.num_replicas = get_world_size() if num_replicas is None else num_replicas
        self.rank = get_rank() if rank is None else rank
        self.total_size = len(dataset)
        self.num_samples = math.ceil((self.total_size - self.rank) / self.num_replicas)

    def __iter__(self):
        return iter(range(self.rank, self.total_size, self.num_replicas))

    def __len__(self):
        return self.num_samples


class _EpochSampler:
    def __init__(self, dataset):
        self.dataset = dataset
        self.epoch = 0

    def set_epoch(self, epoch):
        self.epoch = epoch


class DeviceSeqLoader:
    """Train loader whose batches are assembled ON THE GPU (pr_seq_batch_build): the padded training windows stay
    resident in HBM, an epoch is a seeded permutation partitioned over ranks exactly like DistributedSampler
    (pad to a multiple of world, rank r takes r, r+W, ...), and every step is one kernel launch instead of
    B python `__getitem__` calls in DataLoader workers (REC/data/utils.py:88-112, trainset.py:40-75).
    Enabled with `device_sampler: True` in the yaml (SEQ models)."""

    def __init__(self, dataset, batch_size, device, rank=0, world=1, seed=0):
        from .. import ops
        self._ops = ops
        self.dataset = dataset
        self.batch_size = int(batch_size)
        self.device = device
        self.rank, self.world, self.seed = rank, world, int(seed)
        self.sampler = _EpochSampler(dataset)
        self.padded = torch.from_numpy(dataset.padded).to(device)
        n = len(dataset)
        self.num_samples = math.ceil(n / world)
        self._step = 0

    def __len__(self):
        return math.ceil(self.num_samples / self.batch_size)

    def __iter__(self):
        n = len(self.dataset)
        g = torch.Generator()
        g.manual_seed(self.seed + self.sampler.epoch)
        perm = torch.randperm(n, generator=g)
        total = self.num_samples * self.world
        if total > n:
            perm = torch.cat([perm, perm[: total - n]])
        mine = perm[self.rank:total:self.world].to(self.device)
        for i in range(0, self.num_samples, self.batch_size):
            self._step += 1
            sel = mine[i:i + self.batch_size].contiguous()
            yield self._ops.seq_batch_build(self.padded, sel, self.dataset.item_num,
                                            (self.seed * 1000003 + self.sampler.epoch) * 1000003 + self._step * self.world + self.rank)


class DeviceSeqEvalLoader:
    """Eval loader whose batches are VIEWS of tensors resident in HBM: what SeqEvalDataset.__getitem__ + seq_eval_collate
    (REC/data/dataset/evalset.py:24-37, collate_fn.py:6-32) rebuild per batch in Python -- the left-padded history window, the
    target, and the (history_u, history_i) pairs that the scoring kernel masks -- is laid out ONCE for this rank's users in
    their evaluation order (rank r: users r, r+W, ...: data/utils.py:134-159) as a CSR over batches; a step is four slices, no
    kernel, no host loop.  Output is element-for-element the DataLoader's (`device_sampler: True`, SEQ models)."""

    def __init__(self, dataset, batch_size, device, rank=0, world=1):
        self.dataset = dataset
        self.batch_size = B = int(batch_size)
        self.device = device
        self.sampler = NonConsecutiveSequentialDistributedSampler(dataset, rank=rank, num_replicas=world)
        L = dataset.max_item_list_length
        cut = -2 if dataset.phase == "valid" else -1
        users = range(rank, len(dataset), world)
        n = len(users)
        seqs = np.zeros((n, L), dtype=np.int64)
        tgt = np.zeros(n, dtype=np.int64)
        lens = np.zeros(n + 1, dtype=np.int64)
        hist = []
        for j, u in enumerate(users):
            s = dataset.user_seq[u]
            h = np.asarray(s[:cut], dtype=np.int64)
            tail = h[-L:]
            if len(tail):
                seqs[j, L - len(tail):] = tail
            tgt[j] = s[cut]
            lens[j + 1] = len(h)
            hist.append(h)
        offs = np.cumsum(lens)
        hist_i = np.concatenate(hist) if hist else np.zeros(0, np.int64)
        hist_u = np.repeat(np.arange(n, dtype=np.int64) % B, lens[1:])          # row of the user inside ITS batch
        self.num_samples = n
        self._batch_offs = [int(offs[min(i, n)]) for i in range(0, n + B, B)]  # pair range of batch k: [offs[k*B], offs[(k+1)*B])
        to = lambda a: torch.from_numpy(a).to(device)
        self.item_seq, self.target, self.hist_u, self.hist_i = to(seqs), to(tgt), to(hist_u), to(hist_i)
        self.positive_u = torch.arange(B, device=device)

    def __len__(self):
        return math.ceil(self.num_samples / self.batch_size)

    def __iter__(self):
        B = self.batch_size
        for k in range(len(self)):
            lo, hi = k * B, min((k + 1) * B, self.num_samples)
            p0, p1 = self._batch_offs[k], self._batch_offs[k + 1]
            yield (self.item_seq[lo:hi], (self.hist_u[p0:p1], self.hist_i[p0:p1]), self.positive_u[:hi - lo], self.target[lo:hi])


def _worker_init(worker_id, num_workers, rank, seed):
    s = (num_workers * rank + worker_id + seed) % (2 ** 32)
    np.random.seed(s)
    random.seed(s)


def bulid_dataloader(config, dataload):
    model_name = config["model"]
    if model_name not in DATASET_TABLE:
        raise KeyError(f"model {model_name!r} has no dataset binding (hot-path models: {sorted(DATASET_TABLE)})")
    dataload.build()
    train_cls, eval_cls, collate = DATASET_TABLE[model_name]
    train_data = train_cls(config, dataload)
    valid_data = eval_cls(config, dataload, phase="valid")
    test_data = eval_cls(config, dataload, phase="test")
    log = getLogger()
    log.info(f"[Training]: train_batch_size = [{config['train_batch_size']}]")
    log.info(f"[Evaluation]: eval_batch_size = [{config['eval_batch_size']}]")
    world, rank = get_world_size(), get_rank()
    train_sampler = torch.utils.data.distributed.DistributedSampler(train_data, num_replicas=world, rank=rank)
    workers = config["num_workers"] if config["num_workers"] is not None else 10
    init_fn = partial(_worker_init, num_workers=workers, rank=rank, seed=torch.initial_seed())
    pin = torch.cuda.is_available()
    if config["device_sampler"] and torch.cuda.is_available() and config["device"] is not None:
        train_loader = DeviceSeqLoader(train_data, config["train_batch_size"], config["device"], rank, world,
                                       seed=config["seed"] or 0)
    else:
        train_loader = DataLoader(train_data, batch_size=config["train_batch_size"], num_workers=workers, pin_memory=pin,
                                  sampler=train_sampler, worker_init_fn=init_fn)
    if config["device_sampler"] and torch.cuda.is_available() and config["device"] is not None and collate is seq_eval_collate:
        mk = lambda ds: DeviceSeqEvalLoader(ds, config["eval_batch_size"], config["device"], rank, world)
    else:
        mk = lambda ds: DataLoader(ds, batch_size=config["eval_batch_size"], num_workers=workers, pin_memory=pin,
                                   sampler=NonConsecutiveSequentialDistributedSampler(ds), collate_fn=collate)
    return train_loader, mk(valid_data), mk(test_data)
