"""train-step time vs per-GPU batch (C2 shape): python tools/step_sweep.py 1024 4096 8192"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pixelrec_b200.model.IDNet.sasrec import SASRec  # noqa: E402
from pixelrec_b200.trainer.optim import FusedAdamW  # noqa: E402

dev = torch.device("cuda", 0)
c = bench.C2


class Dl:
    item_num = c["N"]


cfg = dict(n_layers=c["layers"], n_heads=c["heads"], embedding_size=c["D"], inner_size=c["inner"], hidden_dropout_prob=0.1,
           attn_dropout_prob=0.1, hidden_act="gelu", layer_norm_eps=1e-12, initializer_range=0.02, MAX_ITEM_LIST_LENGTH=c["L"], seed=1)
model = SASRec(cfg, Dl()).to(dev).train()
opt = FusedAdamW(model.parameters(), lr=1e-4, weight_decay=0.1, tables=[model.item_embedding])
perm, p = bench.popularity(c["N"])
g = np.random.default_rng(0)
out = {}
for B in [int(x) for x in sys.argv[1:]] or [64, 1024, 4096, 8192]:
    batches = [tuple(torch.from_numpy(x).to(dev) for x in bench.synth_batch(g, B, c["N"], c["L"], perm, p)) for _ in range(4)]

    def step(i):
        opt.zero_grad()
        loss = model(batches[i % 4])
        loss.backward()
        opt.step()
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    s.record()
    for i in range(K):
        step(i)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / K
    out[B] = {"ms_per_step": ms, "seq_per_s": B / ms * 1e3, "peak_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}
    print(B, json.dumps(out[B]), flush=True)
    del batches
    torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/step_sweep.json", "w"), indent=1)
