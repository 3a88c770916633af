"""GPU parity of the CUDA-graph replay of the training step (pixelrec_b200/trainer/graph.py): K replays must leave the
model exactly where K eager steps leave it -- including the dropout masks (device-side seed offset; the default build's
-DPR_SEED_DEV) and AdamW's bias correction (device-side step count)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, B, L = 503, 32, 10


def _cfg(p):
    return dict(n_layers=2, n_heads=4, embedding_size=128, inner_size=2, hidden_dropout_prob=p, attn_dropout_prob=p,
                hidden_act="gelu", layer_norm_eps=1e-12, initializer_range=0.02, MAX_ITEM_LIST_LENGTH=L, seed=5)


def _batches(k):
    g = np.random.default_rng(3)
    out = []
    for _ in range(k):
        items = g.integers(1, N, size=(B, 2, L + 1)).astype(np.int64)
        items[:, 1, 0] = 0
        items[::4, 0, :3] = 0
        items[::4, 1, :4] = 0
        mask = (items[:, 1, 1:] != 0).astype(np.int64)
        out.append((torch.from_numpy(items).cuda(), torch.from_numpy(mask).cuda()))
    return out


def _run(p, graph_from):
    """8 steps; from step `graph_from` on they are graph replays (None: all eager)"""
    from pixelrec_b200.model.IDNet.sasrec import SASRec
    from pixelrec_b200.trainer.graph import GraphedTrainStep
    from pixelrec_b200.trainer.optim import FusedAdamW

    class Dl:
        item_num = N
    torch.manual_seed(0)
    m = SASRec(_cfg(p), Dl()).cuda().train()
    opt = FusedAdamW(m.parameters(), lr=1e-3, weight_decay=0.1, tables=[m.item_embedding])
    losses, graphed = [], None
    for i, b in enumerate(_batches(8)):
        if graph_from is not None and i == graph_from:
            graphed = GraphedTrainStep(m, opt, b)
        if graphed is not None:
            losses.append(float(graphed(b)))
        else:
            opt.zero_grad()
            loss = m(b)
            loss.backward()
            opt.step()
            losses.append(float(loss))
            del loss                         # no autograd graph of an eager step may outlive it (see trainer/graph.py)
    if graphed is not None:
        graphed.close()
        opt.zero_grad()                      # and eager execution resumes seamlessly
        loss = m(_batches(1)[0])
        loss.backward()
        opt.step()
        losses.append(float(loss))
    else:
        opt.zero_grad()
        loss = m(_batches(1)[0])
        loss.backward()
        opt.step()
        losses.append(float(loss))
    return losses, {k: v.detach().clone() for k, v in m.state_dict().items()}, opt._step


def _seed_dev_supported():
    from pixelrec_b200 import lib
    L_ = lib.load()
    t = torch.zeros(1, dtype=torch.int64, device="cuda")
    ok = L_.pr_set_seed_device(t.data_ptr()) == 0
    L_.pr_set_seed_device(None)
    return ok


@pytest.mark.parametrize("p", [0.0, 0.1])
def test_graph_replay_equals_eager_steps(p):
    if p > 0 and not _seed_dev_supported():
        pytest.skip("library built without -DPR_SEED_DEV: no device-side dropout seeds")
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        l_e, sd_e, n_e = _run(p, None)
        l_g, sd_g, n_g = _run(p, 3)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = True
    assert n_e == n_g == 9
    assert l_e == l_g                                               # same kernels, same seeds, same order: bit-identical
    for k in sd_e:
        assert torch.equal(sd_e[k], sd_g[k]), k
