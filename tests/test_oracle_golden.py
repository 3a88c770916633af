"""Pins the CPU oracle (oracle/sasrec_np.py) to golden vectors produced by the UNMODIFIED reference
modules (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import sasrec_np as O

EPS = 1e-12


def rel(a, b, floor=1e-6):
    """max-abs error relative to max-abs of the reference; `floor` keeps mathematically-zero grads
    (e.g. key.bias: softmax is invariant to a per-row constant) from dividing noise by noise."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), floor)


def test_gather_bit_exact(golden):
    W = golden["params"]["item_embedding.weight"]
    out = O.gather_rows(W, golden["items"])
    assert out.dtype == W.dtype and out.shape == golden["items"].shape + (W.shape[1],)
    assert np.array_equal(out, W[golden["items"]])
    # pad row is returned with its real (non-zero) contents: sasrec.py:49,56
    assert np.abs(W[0]).sum() > 0 and np.array_equal(out[golden["items"] == 0][0], W[0])


@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-5), (np.float64, 2e-6)])
def test_forward_loss_and_encoder(golden, dtype, tol):
    c = golden["cfg"]
    loss, cache = O.sasrec_forward(golden["params"], golden["items"], golden["masked_index"], c["layers"], c["h"], EPS,
                                   dtype=dtype)
    assert rel(loss, golden["loss"]) < tol
    # fully masked query rows (left padding) are don't-care; compare valid rows only
    valid = golden["masked_index"].astype(bool)
    assert rel(cache["out"][valid], golden["enc_out"][valid]) < 50 * tol
    assert np.isfinite(cache["out"]).all()


def test_backward_all_grads(golden):
    c = golden["cfg"]
    loss, cache = O.sasrec_forward(golden["params"], golden["items"], golden["masked_index"], c["layers"], c["h"], EPS,
                                   dtype=np.float64)
    grads = O.sasrec_backward(cache)
    for k, ref in golden["grads"].items():
        assert grads[k].shape == ref.shape, k
        if k.endswith("key.bias"):  # mathematically zero (softmax shift invariance); reference holds fp32 noise
            assert np.abs(grads[k]).max() < 1e-9 and np.abs(ref).max() < 1e-9
            continue
        assert rel(grads[k], ref) < 2e-5, (k, rel(grads[k], ref))
    assert np.all(grads["item_embedding.weight"][0] == 0)  # padding_idx row never receives gradient


def test_backward_fp32_within_tolerance(golden):
    c = golden["cfg"]
    _, cache = O.sasrec_forward(golden["params"], golden["items"], golden["masked_index"], c["layers"], c["h"], EPS)
    grads = O.sasrec_backward(cache)
    for k, ref in golden["grads"].items():
        if k.endswith("key.bias"):
            assert np.abs(grads[k]).max() < 1e-9
            continue
        assert rel(grads[k], ref) < 1e-3, (k, rel(grads[k], ref))


def test_scatter_add_matches_reference_dense_grad(golden):
    c = golden["cfg"]
    _, cache = O.sasrec_forward(golden["params"], golden["items"], golden["masked_index"], c["layers"], c["h"], EPS)
    g = O.sasrec_backward(cache, dense_table_grad=False)
    G = O.scatter_add_rows(g["d_item_emb"], golden["items"], c["N"], 0)
    assert rel(G, golden["grads"]["item_embedding.weight"]) < 1e-4
    u, cnt = O.unique_segments(golden["items"])
    nz = np.flatnonzero(np.abs(golden["grads"]["item_embedding.weight"]).sum(1))
    assert set(nz.tolist()) <= set(u.tolist()) and 0 not in u
    assert cnt.sum() == (golden["items"] != 0).sum()


def test_adamw_two_steps(golden):
    for key in ("item_embedding.weight", "position_embedding.weight", "LayerNorm.weight",
                "trm_encoder.layer.0.multi_head_attention.query.weight", "trm_encoder.layer.0.feed_forward.dense_2.bias"):
        w = golden["params"][key].astype(np.float32)
        g = golden["grads"][key]
        m = np.zeros_like(w)
        v = np.zeros_like(w)
        w1, m, v = O.adamw_step(w, g, m, v, 1, 1e-4, 0.1)
        assert np.abs(w1 - golden["adamw1/" + key]).max() < 2e-7, key
        if key == "item_embedding.weight":
            w2, m, v = O.adamw_step(w1, g, m, v, 2, 1e-4, 0.1)
            assert np.abs(w2 - golden["adamw2/" + key]).max() < 3e-7
            # untouched rows (zero grad) still decay: dense AdamW semantics, trainer.py:102
            untouched = np.flatnonzero(np.abs(g).sum(1) == 0)
            assert len(untouched) and np.allclose(w2[untouched], w[untouched] * (1 - 1e-5) ** 2, rtol=1e-6, atol=0)


def test_predict_and_masked_topk(golden):
    c = golden["cfg"]
    scores, _ = O.sasrec_predict(golden["params"], golden["eval_item_seq"], c["layers"], c["h"], EPS)
    assert scores.shape == golden["eval_scores_raw"].shape
    assert rel(scores, golden["eval_scores_raw"]) < 1e-4
    val, idx = O.full_sort_topk(golden["eval_scores_raw"], golden["eval_hist_u"], golden["eval_hist_i"], 10)
    assert np.array_equal(idx, golden["eval_topk_idx"])
    assert np.array_equal(val, golden["eval_topk_val"])
    assert not (idx == 0).any()
    hist = set(zip(golden["eval_hist_u"].tolist(), golden["eval_hist_i"].tolist()))
    assert not any((u, int(i)) in hist for u in range(idx.shape[0]) for i in idx[u])


def test_attention_mask_semantics():
    ids = np.array([[0, 0, 5, 7], [1, 2, 3, 4]])
    m = O.attention_mask(ids)
    assert m.shape == (2, 1, 4, 4)
    assert (m[0, 0, 3] == np.array([-1e9, -1e9, 0, 0], dtype=np.float32)).all()
    assert (m[0, 0, 0] == -1e9).all()          # fully-masked query row (left pad): uniform softmax, never NaN
    assert (m[1, 0, 1] == np.array([0, 0, -1e9, -1e9], dtype=np.float32)).all()


def test_metrics_known_answer():
    # one user, positive ranked 3rd of 10: recall@5 = 1, ndcg@5 = 1/log2(4) = 0.5
    idx = np.array([[9, 8, 7, 6, 5, 4, 3, 2, 1, 11]])
    pos, plen = O.topk_hits(idx, [0], [7], 1)
    r = O.recall_ndcg(pos, plen, [2, 5, 10])
    assert r["recall@2"] == 0 and r["recall@5"] == 1 and abs(r["ndcg@5"] - 0.5) < 1e-12 and abs(r["ndcg@10"] - 0.5) < 1e-12


def test_train_sample_layout():
    import random

    rng = random.Random(3)
    items, mask = O.seq_train_sample([5, 6, 7], item_num=50, max_item_list_length=5, rng=rng)
    assert items.shape == (2, 6) and mask.tolist() == [0, 0, 0, 1, 1]
    assert items[0].tolist() == [0, 0, 0, 5, 6, 7]
    assert items[1, :4].tolist() == [0, 0, 0, 0] and all(1 <= x < 50 and x not in (5, 6, 7) for x in items[1, 4:])
    # longer than L+1: keeps the most recent L+1 (trainset.py:46-50)
    items, mask = O.seq_train_sample(list(range(1, 10)), 50, 5, rng)
    assert items[0].tolist() == [4, 5, 6, 7, 8, 9] and mask.tolist() == [1] * 5


def test_torch_port_matches_reference_golden(golden):
    """oracle/torch_port.py (the CPU baseline bench.py times) reproduces the reference's loss and gradients."""
    import torch
    from oracle import torch_port as TP
    c = golden["cfg"]
    P = {k: torch.from_numpy(v.copy()).requires_grad_() for k, v in golden["params"].items()}
    loss, out = TP.forward_loss(P, torch.from_numpy(golden["items"]), torch.from_numpy(golden["masked_index"]), c["layers"], c["h"])
    loss.backward()
    assert abs(loss.item() - float(golden["loss"])) < 1e-5
    for k, ref in golden["grads"].items():
        if k.endswith("key.bias"):
            continue
        assert rel(P[k].grad.numpy(), ref) < 1e-4, k
    tv, ti = TP.predict_topk({k: v.detach() for k, v in P.items()}, torch.from_numpy(golden["eval_item_seq"]),
                             torch.from_numpy(golden["eval_hist_u"]), torch.from_numpy(golden["eval_hist_i"]), 10, c["layers"], c["h"])
    assert np.array_equal(ti.numpy(), golden["eval_topk_idx"])


def test_full_catalog_ce_oracle_restates_torch_cross_entropy():
    """the CE extension has no reference counterpart; its oracle is pinned to torch's F.cross_entropy instead"""
    import torch
    import torch.nn.functional as F
    from oracle import sasrec_np as O
    g = np.random.default_rng(5)
    scores = g.standard_normal((17, 301))
    target = g.integers(1, 301, size=17)
    lse, tl, nll = O.full_catalog_ce(scores, target, mask_col0=False)
    ref = F.cross_entropy(torch.from_numpy(scores), torch.from_numpy(target), reduction="none").numpy()
    assert np.allclose(nll, ref, rtol=1e-12, atol=1e-12)
    lse0, _, nll0 = O.full_catalog_ce(scores, target, mask_col0=True)
    ref0 = F.cross_entropy(torch.from_numpy(scores[:, 1:]), torch.from_numpy(target - 1), reduction="none").numpy()
    assert np.allclose(nll0, ref0, rtol=1e-12, atol=1e-12) and (lse0 <= lse + 1e-12).all()


def test_bench_shape_golden_pins_the_oracle_at_d512_l20():
    """D=512 / L=20 / h=4 / 2 layers (the shape the headline is quoted on): fp64 oracle and the torch port (the CPU baseline)
    vs the reference's loss, encoder output, the stored gradient subset and the masked top-k ids."""
    import torch
    from oracle import torch_port as TP
    from oracle.make_golden import grad_subset
    from tests.conftest import load_bench_shape_golden
    g = load_bench_shape_golden()
    c = g["cfg"]
    loss, cache = O.sasrec_forward(g["params"], g["items"], g["masked_index"], c["layers"], c["h"], EPS, dtype=np.float64)
    assert rel(loss, g["loss"]) < 2e-6
    valid = g["masked_index"].astype(bool)
    assert rel(cache["out"][valid], g["enc_out"][valid]) < 1e-4
    grads = O.sasrec_backward(cache)
    for k, ref in g["grads"].items():
        got = grad_subset(k, grads[k], g["touched_rows"])
        assert got.shape == ref.shape, k
        if k.endswith("key.bias"):
            assert np.abs(got).max() < 1e-9
            continue
        assert rel(got, ref) < 5e-5, (k, rel(got, ref))
    P = {k: torch.from_numpy(v.copy()).requires_grad_() for k, v in g["params"].items()}
    tl, _ = TP.forward_loss(P, torch.from_numpy(g["items"]), torch.from_numpy(g["masked_index"]), c["layers"], c["h"])
    tl.backward()
    assert abs(tl.item() - float(g["loss"])) < 1e-5 * float(g["loss"])
    for k, ref in g["grads"].items():
        if not k.endswith("key.bias"):
            assert rel(grad_subset(k, P[k].grad.numpy(), g["touched_rows"]), ref) < 1e-4, k
    scores, _ = O.sasrec_predict(g["params"], g["eval_item_seq"], c["layers"], c["h"], EPS)
    assert rel(scores, g["eval_scores_raw"]) < 1e-4
    _, idx = O.full_sort_topk(g["eval_scores_raw"], g["eval_hist_u"], g["eval_hist_i"], 10)
    assert np.array_equal(idx, g["eval_topk_idx"])


def test_full_catalog_ce_backward_oracle_restates_autograd():
    """oracle.full_catalog_ce_bwd (extension) == float64 autograd of logsumexp - target logit with the padding column excluded"""
    import torch
    g = np.random.default_rng(11)
    B, N, D = 9, 40, 6
    seq, W = g.standard_normal((B, D)), g.standard_normal((N, D))
    target = g.integers(1, N, size=B)
    wts = g.random(B) + 0.5
    X = torch.from_numpy(seq).requires_grad_()
    Wt = torch.from_numpy(W).requires_grad_()
    sc = X @ Wt.t()
    sc = torch.cat([torch.full_like(sc[:, :1], -float("inf")), sc[:, 1:]], 1)
    nll = torch.logsumexp(sc, 1) - sc.gather(1, torch.from_numpy(target)[:, None])[:, 0]
    (nll * torch.from_numpy(wts)).sum().backward()
    dX, dW = O.full_catalog_ce_bwd(seq, W, target, wts)
    assert np.abs(dX - X.grad.numpy()).max() < 1e-12 and np.abs(dW - Wt.grad.numpy()).max() < 1e-12
    assert np.abs(dW[0]).max() == 0.0                                   # the padding item gets no gradient
    lse, tl, nll_o = O.full_catalog_ce(seq @ W.T, target)
    assert np.abs(nll_o - nll.detach().numpy()).max() < 1e-12
