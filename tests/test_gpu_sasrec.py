"""GPU parity of the whole plugin against goldens produced by the UNMODIFIED reference SASRec
(oracle/make_golden.py): loss, every parameter gradient, an AdamW step, predict() scores and the masked
top-k.  Tolerances: 1e-3 relative (north_star) for TF32 linear layers, 1e-4 with matmul_precision fp32."""
import numpy as np
import pytest
import torch

from tests.gpu_util import dev, rel, t

pytestmark = pytest.mark.gpu


class _Data:
    def __init__(self, n):
        self.item_num = n


def build(golden, dropout=0.0, seed=2020):
    from pixelrec_b200.model.IDNet.sasrec import SASRec
    c = golden["cfg"]
    cfg = dict(n_layers=c["layers"], n_heads=c["h"], embedding_size=c["D"], inner_size=2, hidden_dropout_prob=dropout,
               attn_dropout_prob=dropout, hidden_act="gelu", layer_norm_eps=1e-12, initializer_range=0.02,
               MAX_ITEM_LIST_LENGTH=c["L"], seed=seed)
    m = SASRec(cfg, _Data(c["N"]))
    missing = m.load_state_dict({k: torch.from_numpy(v) for k, v in golden["params"].items()}, strict=True)
    return m.to(dev())


@pytest.mark.parametrize("tf32,tol", [(False, 1e-4), (True, 1e-3)])
def test_forward_backward_matches_reference(golden, tf32, tol):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    m = build(golden)
    m.train()
    loss = m((t(golden["items"]), t(golden["masked_index"])))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - float(golden["loss"])) / float(golden["loss"]) < tol
    grads = {k: p.grad for k, p in m.named_parameters()}
    for k, ref in golden["grads"].items():
        assert grads[k] is not None, k
        got = grads[k].cpu().numpy()
        if k.endswith("key.bias"):        # mathematically zero; reference holds 1e-12 noise
            assert np.abs(got).max() < 1e-6
            continue
        assert rel(got, ref) < tol * 3, (k, rel(got, ref))
    assert (grads["item_embedding.weight"][0] == 0).all()        # padding_idx row
    torch.backends.cuda.matmul.allow_tf32 = True


def test_sparse_table_grad_and_fused_adamw_step(golden):
    """The product path: sparse table gradient + FusedAdamW == reference torch.optim.AdamW(lr 1e-4, wd 0.1) step."""
    from pixelrec_b200.trainer.optim import FusedAdamW
    torch.backends.cuda.matmul.allow_tf32 = False
    m = build(golden)
    m.train()
    opt = FusedAdamW(m.parameters(), lr=1e-4, weight_decay=0.1, tables=[m.item_embedding])
    for step in (1, 2):
        opt.zero_grad()
        loss = m((t(golden["items"]), t(golden["masked_index"])))
        if step == 1:
            loss.backward()
            assert m.item_embedding.weight.grad is None and len(m.item_embedding.sink.pending) == 1
            saved = [(p, r.clone()) for p, r in m.item_embedding.sink.pending]
            dense_g = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        else:   # goldens apply the SAME gradient twice: replay step-1 grads
            m.item_embedding.sink.pending[:] = saved
            m.item_embedding.sink.row2slot[saved[0][0].uniq_ids[:saved[0][0].n_uniq.item()].long()] = \
                torch.arange(saved[0][0].n_uniq.item(), dtype=torch.int32, device=dev())
            for k, p in m.named_parameters():
                if k in dense_g:
                    p.grad.copy_(dense_g[k])
        opt.step()
        torch.cuda.synchronize()
        if step == 1:
            for k in ("item_embedding.weight", "position_embedding.weight", "LayerNorm.weight",
                      "trm_encoder.layer.0.multi_head_attention.query.weight", "trm_encoder.layer.0.feed_forward.dense_2.bias"):
                got = dict(m.named_parameters())[k].detach().cpu().numpy()
                assert np.abs(got - golden["adamw1/" + k]).max() < 2e-6, k
        else:
            got = m.item_embedding.weight.detach().cpu().numpy()
            assert np.abs(got - golden["adamw2/item_embedding.weight"]).max() < 4e-6
    torch.backends.cuda.matmul.allow_tf32 = True


def test_predict_and_masked_topk(golden):
    torch.backends.cuda.matmul.allow_tf32 = False
    m = build(golden).eval()
    seqs = t(golden["eval_item_seq"])
    scores = m.predict(seqs, m.compute_item_all())
    assert scores.shape == golden["eval_scores_raw"].shape
    assert rel(scores.cpu().numpy(), golden["eval_scores_raw"]) < 1e-4
    scores[:, 0] = -np.inf                                        # trainer.py:334-336
    scores[t(golden["eval_hist_u"]), t(golden["eval_hist_i"])] = -np.inf
    _, idx = torch.topk(scores, 10, dim=-1)
    assert np.array_equal(idx.cpu().numpy(), golden["eval_topk_idx"])
    torch.backends.cuda.matmul.allow_tf32 = True


def test_training_with_dropout_is_finite_and_seeded():
    from pixelrec_b200.model.IDNet.sasrec import SASRec
    cfg = dict(n_layers=2, n_heads=4, embedding_size=128, inner_size=2, hidden_dropout_prob=0.1, attn_dropout_prob=0.1,
               hidden_act="gelu", layer_norm_eps=1e-12, initializer_range=0.02, MAX_ITEM_LIST_LENGTH=10, seed=2020)
    g = np.random.default_rng(0)
    items = g.integers(1, 500, size=(32, 2, 11)).astype(np.int64)
    mask = np.ones((32, 10), dtype=np.int64)
    losses = []
    for rep in range(2):
        torch.manual_seed(0)
        m = SASRec(cfg, _Data(500)).to(dev()).train()
        l1 = m((t(items), t(mask)))
        l2 = m((t(items), t(mask)))
        l1.backward()
        losses.append((l1.item(), l2.item()))
        assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
        assert abs(l1.item() - 10 * np.log(2)) < 0.2              # init loss ~ L * ln 2
    assert losses[0] == losses[1] and losses[0][0] != losses[0][1]   # same seed -> same masks; new mask each call
    m.eval()
    assert m((t(items), t(mask))).item() == m((t(items), t(mask))).item()


def test_clip_grad_norm_covers_the_sparse_table_rows():
    """trainer.py:123 clip_grad_norm_ over ALL parameters: FusedAdamW.clip_grad_norm_ counts the table's sparse gradient rows (the
    table has no dense .grad here) -- the norm must equal torch's over a dense-gradient replica, and one clipped step must move
    the weights exactly as torch.optim.AdamW moves the replica after torch.nn.utils.clip_grad_norm_."""
    from pixelrec_b200.model.IDNet.sasrec import SASRec
    from pixelrec_b200.trainer.optim import FusedAdamW
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = dict(n_layers=1, n_heads=2, embedding_size=64, inner_size=2, hidden_dropout_prob=0.0, attn_dropout_prob=0.0,
               hidden_act="gelu", layer_norm_eps=1e-12, initializer_range=0.02, MAX_ITEM_LIST_LENGTH=10, seed=1)
    g = np.random.default_rng(3)
    items = g.integers(1, 300, size=(16, 2, 11)).astype(np.int64)
    items[:, 1, 0] = 0
    mask = np.ones((16, 10), dtype=np.int64)
    torch.manual_seed(0)
    m = SASRec(cfg, _Data(300)).to(dev()).train()
    ref = SASRec(cfg, _Data(300)).to(dev()).train()
    ref.load_state_dict(m.state_dict())
    ref.item_embedding.sink.sparse = False                       # dense [N, D] gradient, as in the reference
    opt = FusedAdamW(m.parameters(), lr=1e-2, weight_decay=0.1, tables=[m.item_embedding])
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.1)
    batch = (t(items), t(mask))
    opt.zero_grad()
    (m(batch) * 50.0).backward()
    ropt.zero_grad()
    (ref(batch) * 50.0).backward()
    assert m.item_embedding.weight.grad is None and ref.item_embedding.weight.grad is not None
    max_norm = 1.0
    want = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm).item()
    assert want > 2 * max_norm                                   # the clip is active
    ropt.step()
    before = {k: v.detach().clone() for k, v in m.state_dict().items()}
    total = opt.clip_grad_norm_(max_norm)
    assert abs(total.item() - want) < 1e-4 * want
    opt.step()
    for (k, a), b in zip(m.state_dict().items(), ref.state_dict().values()):
        assert (a - b).abs().max().item() < 2e-5, k
        assert k.endswith("LayerNorm.bias") or not torch.equal(a, before[k]), k
    torch.backends.cuda.matmul.allow_tf32 = True


def test_gru4rec_plugin_matches_reference_golden():
    """config 5 backbone: our table gather / scatter-add / loss around cuDNN's GRU vs the reference module."""
    import os
    from pixelrec_b200.model.IDNet.gru4rec import GRU4Rec
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "gru4rec_small.npz"))
    N, D, L = int(z["cfg_N"]), int(z["cfg_D"]), int(z["cfg_L"])
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    class Dl:
        item_num, user_num = N, 0
    m = GRU4Rec(dict(embedding_size=D, hidden_size=1, num_layers=1, dropout_prob=0), Dl())
    m.load_state_dict({k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")})
    m = m.to(dev()).train()
    loss = m((t(z["items"]), t(z["masked_index"])))
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) / float(z["loss"]) < 1e-4
    for k, p in m.named_parameters():
        assert rel(p.grad.cpu().numpy(), z["grad/" + k]) < 1e-3, k
    assert (m.item_embedding.weight.grad[0] == 0).all()
    m.eval()
    sc = m.predict(t(z["eval_item_seq"]), m.compute_item_all())
    assert rel(sc.cpu().numpy(), z["eval_scores_raw"]) < 1e-4
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True


@pytest.mark.parametrize("tf32,tol", [(False, 1e-4), (True, 1e-3)])
def test_bench_shape_forward_backward_matches_reference(tf32, tol):
    """The shape bench.py times (D=512, L=20, h=4, 2 layers; K up to 1024 in dense_2) against the reference's golden:
    TF32 linear layers + tensor-core attention must hold north_star's 1e-3 there, not only on the small fixtures."""
    from oracle.make_golden import grad_subset
    from tests.conftest import load_bench_shape_golden
    g = load_bench_shape_golden()
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    m = build(g)
    m.train()
    loss = m((t(g["items"]), t(g["masked_index"])))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - float(g["loss"])) / float(g["loss"]) < tol
    grads = {k: p.grad for k, p in m.named_parameters()}
    for k, ref in g["grads"].items():
        got = grad_subset(k, grads[k].cpu().numpy(), g["touched_rows"])
        if k.endswith("key.bias"):
            assert np.abs(got).max() < 1e-6
            continue
        assert rel(got, ref) < tol * 3, (k, rel(got, ref))
    m.eval()
    with torch.no_grad():
        sc = m.predict(t(g["eval_item_seq"]), m.compute_item_all())
    assert rel(sc.cpu().numpy(), g["eval_scores_raw"]) < tol
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
