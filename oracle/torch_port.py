"""TEST INFRASTRUCTURE / CPU BASELINE -- a from-scratch torch (CPU, fp32) restatement of the reference's
training step for the SASRec hot path, used (a) as the `cpu_baseline` / `--impl reference` leg of bench.py
on the GPU box (the real reference is a Python tree under /root/reference and cannot travel), (b) as a second
checker next to oracle/sasrec_np.py.  It runs exactly the ops the reference runs, in the reference's order:

    REC/model/IDNet/sasrec.py:65-92   (embedding, pos-emb + LayerNorm + dropout, mask, encoder, pairwise loss)
    REC/model/layers.py:543-759       (post-LN transformer blocks, erf-GELU)
    REC/trainer/trainer.py:116-125    (zero_grad -> forward -> backward -> torch.optim.AdamW.step)

tests/test_oracle_golden.py pins it to the goldens generated from the unmodified reference modules.
Never imported by the product path (pixelrec_b200/).
"""
import math

import torch
import torch.nn.functional as F


def init_params(N, D, L, n_layers, inner_mult=2, std=0.02, seed=2020, device="cpu"):
    """Parameters with the reference's names/shapes/initialisation (sasrec.py:51-61)."""
    g = torch.Generator().manual_seed(seed)
    P = {}

    def normal(*shape):
        return (torch.randn(*shape, generator=g) * std).to(device)

    P["item_embedding.weight"] = normal(N, D)
    P["position_embedding.weight"] = normal(L, D)
    P["LayerNorm.weight"] = torch.ones(D, device=device)
    P["LayerNorm.bias"] = torch.zeros(D, device=device)
    for i in range(n_layers):
        pre = f"trm_encoder.layer.{i}."
        for nm in ("query", "key", "value", "dense"):
            P[pre + f"multi_head_attention.{nm}.weight"] = normal(D, D)
            P[pre + f"multi_head_attention.{nm}.bias"] = torch.zeros(D, device=device)
        P[pre + "multi_head_attention.LayerNorm.weight"] = torch.ones(D, device=device)
        P[pre + "multi_head_attention.LayerNorm.bias"] = torch.zeros(D, device=device)
        P[pre + "feed_forward.dense_1.weight"] = normal(inner_mult * D, D)
        P[pre + "feed_forward.dense_1.bias"] = torch.zeros(inner_mult * D, device=device)
        P[pre + "feed_forward.dense_2.weight"] = normal(D, inner_mult * D)
        P[pre + "feed_forward.dense_2.bias"] = torch.zeros(D, device=device)
        P[pre + "feed_forward.LayerNorm.weight"] = torch.ones(D, device=device)
        P[pre + "feed_forward.LayerNorm.bias"] = torch.zeros(D, device=device)
    return P


def attention_mask(ids):
    """sasrec.py:119-126."""
    valid = (ids != 0)
    L = ids.size(-1)
    m = torch.tril(valid.unsqueeze(1).unsqueeze(2).expand(-1, -1, L, -1))
    return torch.where(m, 0.0, -1e9)


def encoder(P, x, mask, n_layers, n_heads, eps, p_hidden, p_attn, training):
    B, L, D = x.shape
    dh = D // n_heads
    for i in range(n_layers):
        pre = f"trm_encoder.layer.{i}."
        lin = lambda t, nm: F.linear(t, P[pre + nm + ".weight"], P[pre + nm + ".bias"])
        q = lin(x, "multi_head_attention.query").view(B, L, n_heads, dh).permute(0, 2, 1, 3)
        k = lin(x, "multi_head_attention.key").view(B, L, n_heads, dh).permute(0, 2, 3, 1)
        v = lin(x, "multi_head_attention.value").view(B, L, n_heads, dh).permute(0, 2, 1, 3)
        s = torch.matmul(q, k) / math.sqrt(dh) + mask                       # layers.py:595-601
        p = F.dropout(torch.softmax(s, dim=-1), p_attn, training)           # :604-608
        ctx = torch.matmul(p, v).permute(0, 2, 1, 3).contiguous().view(B, L, D)
        h = F.dropout(lin(ctx, "multi_head_attention.dense"), p_hidden, training)
        a = F.layer_norm(h + x, (D,), P[pre + "multi_head_attention.LayerNorm.weight"],
                         P[pre + "multi_head_attention.LayerNorm.bias"], eps)
        h1 = lin(a, "feed_forward.dense_1")
        h1 = h1 * 0.5 * (1.0 + torch.erf(h1 / math.sqrt(2.0)))              # :651-660
        h2 = F.dropout(lin(h1, "feed_forward.dense_2"), p_hidden, training)
        x = F.layer_norm(h2 + a, (D,), P[pre + "feed_forward.LayerNorm.weight"], P[pre + "feed_forward.LayerNorm.bias"], eps)
    return x


def forward_loss(P, items, masked_index, n_layers, n_heads, eps=1e-12, p_hidden=0.0, p_attn=0.0, training=True):
    """sasrec.py:65-92."""
    E = F.embedding(items, P["item_embedding.weight"], padding_idx=0)
    pos, neg = E[:, 0], E[:, 1]
    inp, tp, tn = pos[:, :-1], pos[:, 1:], neg[:, 1:]
    L = masked_index.size(1)
    D = inp.size(-1)
    x = inp + P["position_embedding.weight"][:L].unsqueeze(0)
    x = F.dropout(F.layer_norm(x, (D,), P["LayerNorm.weight"], P["LayerNorm.bias"], eps), p_hidden, training)
    out = encoder(P, x, attention_mask(masked_index), n_layers, n_heads, eps, p_hidden, p_attn, training)
    ps, ns = (out * tp).sum(-1), (out * tn).sum(-1)
    loss = -(torch.log((ps - ns).sigmoid() + 1e-8) * masked_index).sum(-1)
    return loss.mean(-1), out


@torch.no_grad()
def predict_topk(P, item_seq, hist_u, hist_i, k, n_layers, n_heads, eps=1e-12):
    """sasrec.py:94-113 + trainer.py:332-336 + collector.py:133."""
    W = P["item_embedding.weight"]
    L = item_seq.size(1)
    D = W.size(1)
    x = F.layer_norm(F.embedding(item_seq, W) + P["position_embedding.weight"][:L].unsqueeze(0), (D,),
                     P["LayerNorm.weight"], P["LayerNorm.bias"], eps)
    out = encoder(P, x, attention_mask(item_seq), n_layers, n_heads, eps, 0.0, 0.0, False)
    scores = out[:, -1] @ W.t()
    scores[:, 0] = -float("inf")
    if hist_u is not None:
        scores[hist_u, hist_i] = -float("inf")
    return torch.topk(scores, k, dim=-1)


class TrainStep:
    """The loop body of trainer.py:116-125 around the functional model above."""

    def __init__(self, P, n_layers, n_heads, lr=1e-4, weight_decay=0.1, p_drop=0.1, eps=1e-12):
        self.P = {k: v.detach().clone().requires_grad_() for k, v in P.items()}
        self.opt = torch.optim.AdamW(list(self.P.values()), lr=lr, weight_decay=weight_decay)
        self.cfg = (n_layers, n_heads, eps, p_drop)

    def __call__(self, items, masked_index):
        n_layers, n_heads, eps, p = self.cfg
        self.opt.zero_grad(set_to_none=False)                 # torch 1.10 semantics: dense zero fill (trainer.py:117)
        loss, _ = forward_loss(self.P, items, masked_index, n_layers, n_heads, eps, p, p, True)
        loss.backward()
        self.opt.step()
        return loss.detach()
