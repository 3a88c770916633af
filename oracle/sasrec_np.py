"""TEST INFRASTRUCTURE -- the CPU oracle for the SASRec training / scoring hot path.

A numpy restatement (forward AND hand-derived backward) of the reference's algorithm.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this
file; the product path (pixelrec_b200/) never does and fails loudly without its CUDA
extension.

Pinned: oracle/make_golden.py runs the *imported reference modules* (oracle/refload.py)
on seeded inputs and stores inputs + reference outputs/grads in tests/golden/*.npz;
tests/test_oracle_golden.py checks every function here against those files.  The
reference ships no tests / golden vectors of its own, so the
goldens generated from the reference module are the pin.

Every function cites the reference file:line (relative to /root/reference/code) it follows.
Parameter names are the reference's state_dict keys.  Arithmetic is done in `dtype`
(np.float32 by default = the reference's precision; np.float64 for a tight reference).
"""
from __future__ import annotations

import math

import numpy as np

try:  # scipy ships in the image; math.erf fallback keeps the oracle importable anywhere
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


# ----------------------------------------------------------------------------------------------
# row ops (integer-indexed): REC/model/IDNet/sasrec.py:31,68  (nn.Embedding(N, D, padding_idx=0))
# ----------------------------------------------------------------------------------------------
def gather_rows(W: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """out[..., :] = W[idx[...], :]  -- sasrec.py:68 `self.item_embedding(items)`.  Bit-exact copy.
    The pad row (id 0) is returned with its real contents (it is re-initialised non-zero by
    `self.apply(self._init_weights)`, sasrec.py:49,56)."""
    idx = np.asarray(idx)
    if idx.size and (idx.min() < 0 or idx.max() >= W.shape[0]):
        raise IndexError("index out of range in gather_rows")
    return W[idx.reshape(-1)].reshape(*idx.shape, W.shape[1])


def scatter_add_rows(dOut: np.ndarray, idx: np.ndarray, N: int, padding_idx: int = 0,
                     dtype=np.float32) -> np.ndarray:
    """Dense gradient of gather_rows: G[i] = sum_{r: idx[r]==i, i!=padding_idx} dOut[r]
    (autograd's embedding_dense_backward for sasrec.py:68, padding_idx from sasrec.py:31).

    Summation order is DEFINED (the CUDA kernel follows the same order so the comparison is
    bit-exact): rows with the same index are added in ascending flat position r, sequentially,
    starting from the first row's value (not from 0.0f + first).
    """
    idx = np.asarray(idx).reshape(-1)
    D = dOut.shape[-1]
    dO = np.asarray(dOut, dtype=dtype).reshape(-1, D)
    G = np.zeros((N, D), dtype=dtype)
    order = np.argsort(idx, kind="stable")
    sidx = idx[order]
    if len(sidx) == 0:
        return G
    starts = np.flatnonzero(np.r_[True, sidx[1:] != sidx[:-1]])
    ends = np.r_[starts[1:], len(sidx)]
    maxlen = int((ends - starts).max())
    uniq = sidx[starts]
    acc = dO[order[starts]].copy()
    for k in range(1, maxlen):  # sequential adds, vectorised over segments
        live = (starts + k) < ends
        acc[live] = acc[live] + dO[order[starts[live] + k]]
    keep = uniq != padding_idx
    G[uniq[keep]] = acc[keep]
    return G


def unique_segments(idx: np.ndarray, padding_idx: int = 0):
    """(uniq ids ascending without padding_idx, counts) -- the sparse-row format the CUDA
    scatter emits instead of the dense zero-filled [N,D] of trainer.py:117,122."""
    idx = np.asarray(idx).reshape(-1)
    u, c = np.unique(idx, return_counts=True)
    keep = u != padding_idx
    return u[keep], c[keep]


# ----------------------------------------------------------------------------------------------
# elementwise / normalisation pieces
# ----------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps):
    """nn.LayerNorm(D, eps) -- sasrec.py:44,82; layers.py:574,615,637,671 (biased variance)."""
    mu = x.mean(-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + x.dtype.type(eps))
    xhat = xc * rstd
    return xhat * gamma + beta, (xhat, rstd)


def layernorm_bwd(dy, gamma, cache):
    xhat, rstd = cache
    dxhat = dy * gamma
    dx = rstd * (dxhat - dxhat.mean(-1, keepdims=True) - xhat * (dxhat * xhat).mean(-1, keepdims=True))
    red = tuple(range(dy.ndim - 1))
    return dx, (dy * xhat).sum(red), dy.sum(red)


def gelu_fwd(x):
    """layers.py:651-660: x * 0.5 * (1 + erf(x / sqrt(2)))  (erf form, not tanh)."""
    return (x * 0.5 * (1.0 + _erf(x / math.sqrt(2.0)))).astype(x.dtype)


def gelu_bwd(x, dy):
    cdf = 0.5 * (1.0 + _erf(x / math.sqrt(2.0)))
    pdf = np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)
    return (dy * (cdf + x * pdf)).astype(x.dtype)


def attention_mask(key_ids: np.ndarray, dtype=np.float32) -> np.ndarray:
    """sasrec.py:119-126 get_attention_mask(bidirectional=False):
    mask[b,0,i,j] = 0 if (key_ids[b,j] != 0 and j <= i) else -1e9."""
    valid = (np.asarray(key_ids) != 0)
    L = valid.shape[1]
    tri = np.tril(np.ones((L, L), dtype=bool))
    m = valid[:, None, None, :] & tri[None, None, :, :]
    return np.where(m, 0.0, -1e9).astype(dtype)


def softmax_lastdim(s):
    s = s - s.max(-1, keepdims=True)
    e = np.exp(s)
    return e / e.sum(-1, keepdims=True)


# ----------------------------------------------------------------------------------------------
# attention core (the part the CUDA kernel pr_sasrec_attn_* replaces): layers.py:590-612
# ----------------------------------------------------------------------------------------------
def attn_core_fwd(q, k, v, mask, n_heads, drop_p=None):
    """q,k,v: [B,L,D] (outputs of the query/key/value Linear, layers.py:586-588).
    softmax(q k^T / sqrt(dh) + mask) v per head, heads concatenated back to [B,L,D]
    (layers.py:590-612).  drop_p: optional multiplicative dropout mask on the probabilities
    [B,h,L,L] (already scaled by 1/(1-p)), layers.py:608."""
    B, L, D = q.shape
    dh = D // n_heads
    qh = q.reshape(B, L, n_heads, dh).transpose(0, 2, 1, 3)
    kh = k.reshape(B, L, n_heads, dh).transpose(0, 2, 1, 3)
    vh = v.reshape(B, L, n_heads, dh).transpose(0, 2, 1, 3)
    s = np.matmul(qh, kh.transpose(0, 1, 3, 2)) / q.dtype.type(math.sqrt(dh))
    s = s + mask.astype(q.dtype)
    p = softmax_lastdim(s)
    pd = p if drop_p is None else p * drop_p
    ctx = np.matmul(pd, vh)  # [B,h,L,dh]
    out = ctx.transpose(0, 2, 1, 3).reshape(B, L, D)
    return out, (qh, kh, vh, p, pd, drop_p)


def attn_core_bwd(dout, cache):
    qh, kh, vh, p, pd, drop_p = cache
    B, h, L, dh = qh.shape
    dctx = dout.reshape(B, L, h, dh).transpose(0, 2, 1, 3)
    dvh = np.matmul(pd.transpose(0, 1, 3, 2), dctx)
    dpd = np.matmul(dctx, vh.transpose(0, 1, 3, 2))
    dp = dpd if drop_p is None else dpd * drop_p
    ds = p * (dp - (dp * p).sum(-1, keepdims=True))
    ds = ds / qh.dtype.type(math.sqrt(dh))
    dqh = np.matmul(ds, kh)
    dkh = np.matmul(ds.transpose(0, 1, 3, 2), qh)
    back = lambda t: t.transpose(0, 2, 1, 3).reshape(B, L, h * dh)
    return back(dqh), back(dkh), back(dvh)


def linear_fwd(x, W, b):
    return x @ W.T + b


def linear_bwd(x, W, dy):
    D_in = x.shape[-1]
    x2 = x.reshape(-1, D_in)
    dy2 = dy.reshape(-1, dy.shape[-1])
    return dy @ W, dy2.T @ x2, dy2.sum(0)


# ----------------------------------------------------------------------------------------------
# loss: sasrec.py:88-92 (identical in gru4rec.py:63-67, mosasrec.py:89-93)
# ----------------------------------------------------------------------------------------------
def bpr_loss_fwd(out, tgt_pos, tgt_neg, masked_index):
    """loss = mean_b( - sum_t log(sigmoid(<out,pos> - <out,neg>) + 1e-8) * mask[b,t] )."""
    dt = out.dtype
    pos = (out * tgt_pos).sum(-1)
    neg = (out * tgt_neg).sum(-1)
    sig = (1.0 / (1.0 + np.exp(-(pos - neg)))).astype(dt)
    m = masked_index.astype(dt)
    per_b = -(np.log(sig + dt.type(1e-8)) * m).sum(-1)
    return per_b.mean(-1).astype(dt), (sig, m, pos, neg)


def bpr_loss_bwd(out, tgt_pos, tgt_neg, cache, dloss=1.0):
    sig, m, _, _ = cache
    B = out.shape[0]
    dt = out.dtype
    ds = (-(m / dt.type(B)) * sig * (1 - sig) / (sig + dt.type(1e-8)) * dt.type(dloss)).astype(dt)
    ds = ds[..., None]
    return ds * (tgt_pos - tgt_neg), ds * out, -ds * out


# ----------------------------------------------------------------------------------------------
# the encoder: layers.py:543-759
# ----------------------------------------------------------------------------------------------
def _lp(i, name):
    return f"trm_encoder.layer.{i}.{name}"


def encoder_fwd(params, x, mask, n_layers, n_heads, eps, drops=None):
    """TransformerEncoder(output_all_encoded_layers=False)[-1], layers.py:740-759.
    drops: optional dict of multiplicative dropout masks keyed
    'layer{i}.attn' [B,h,L,L], 'layer{i}.out' [B,L,D], 'layer{i}.ffn' [B,L,D]."""
    caches = []
    drops = drops or {}
    for i in range(n_layers):
        c = {}
        c["x"] = x
        q = linear_fwd(x, params[_lp(i, "multi_head_attention.query.weight")], params[_lp(i, "multi_head_attention.query.bias")])
        k = linear_fwd(x, params[_lp(i, "multi_head_attention.key.weight")], params[_lp(i, "multi_head_attention.key.bias")])
        v = linear_fwd(x, params[_lp(i, "multi_head_attention.value.weight")], params[_lp(i, "multi_head_attention.value.bias")])
        ctx, c["attn"] = attn_core_fwd(q, k, v, mask, n_heads, drops.get(f"layer{i}.attn"))
        c["ctx"] = ctx
        h = linear_fwd(ctx, params[_lp(i, "multi_head_attention.dense.weight")], params[_lp(i, "multi_head_attention.dense.bias")])
        c["drop_out"] = drops.get(f"layer{i}.out")
        if c["drop_out"] is not None:
            h = h * c["drop_out"]
        a, c["ln1"] = layernorm_fwd(h + x, params[_lp(i, "multi_head_attention.LayerNorm.weight")],
                                    params[_lp(i, "multi_head_attention.LayerNorm.bias")], eps)  # layers.py:615
        c["a"] = a
        h1 = linear_fwd(a, params[_lp(i, "feed_forward.dense_1.weight")], params[_lp(i, "feed_forward.dense_1.bias")])
        c["h1"] = h1
        g = gelu_fwd(h1)
        c["g"] = g
        h2 = linear_fwd(g, params[_lp(i, "feed_forward.dense_2.weight")], params[_lp(i, "feed_forward.dense_2.bias")])
        c["drop_ffn"] = drops.get(f"layer{i}.ffn")
        if c["drop_ffn"] is not None:
            h2 = h2 * c["drop_ffn"]
        x, c["ln2"] = layernorm_fwd(h2 + a, params[_lp(i, "feed_forward.LayerNorm.weight")],
                                    params[_lp(i, "feed_forward.LayerNorm.bias")], eps)  # layers.py:671
        caches.append(c)
    return x, caches


def encoder_bwd(params, dx, caches, n_heads, grads):
    for i in reversed(range(len(caches))):
        c = caches[i]
        dres, dg, db = layernorm_bwd(dx, params[_lp(i, "feed_forward.LayerNorm.weight")], c["ln2"])
        grads[_lp(i, "feed_forward.LayerNorm.weight")] = dg
        grads[_lp(i, "feed_forward.LayerNorm.bias")] = db
        dh2 = dres if c["drop_ffn"] is None else dres * c["drop_ffn"]
        dgel, dW, dB = linear_bwd(c["g"], params[_lp(i, "feed_forward.dense_2.weight")], dh2)
        grads[_lp(i, "feed_forward.dense_2.weight")] = dW
        grads[_lp(i, "feed_forward.dense_2.bias")] = dB
        dh1 = gelu_bwd(c["h1"], dgel)
        da, dW, dB = linear_bwd(c["a"], params[_lp(i, "feed_forward.dense_1.weight")], dh1)
        grads[_lp(i, "feed_forward.dense_1.weight")] = dW
        grads[_lp(i, "feed_forward.dense_1.bias")] = dB
        da = da + dres
        dres1, dg, db = layernorm_bwd(da, params[_lp(i, "multi_head_attention.LayerNorm.weight")], c["ln1"])
        grads[_lp(i, "multi_head_attention.LayerNorm.weight")] = dg
        grads[_lp(i, "multi_head_attention.LayerNorm.bias")] = db
        dh = dres1 if c["drop_out"] is None else dres1 * c["drop_out"]
        dctx, dW, dB = linear_bwd(c["ctx"], params[_lp(i, "multi_head_attention.dense.weight")], dh)
        grads[_lp(i, "multi_head_attention.dense.weight")] = dW
        grads[_lp(i, "multi_head_attention.dense.bias")] = dB
        dq, dk, dv = attn_core_bwd(dctx, c["attn"])
        dxx = dres1.copy()
        for nm, d in (("query", dq), ("key", dk), ("value", dv)):
            dxi, dW, dB = linear_bwd(c["x"], params[_lp(i, f"multi_head_attention.{nm}.weight")], d)
            grads[_lp(i, f"multi_head_attention.{nm}.weight")] = dW
            grads[_lp(i, f"multi_head_attention.{nm}.bias")] = dB
            dxx = dxx + dxi
        dx = dxx
    return dx


# ----------------------------------------------------------------------------------------------
# SASRec.forward / backward: sasrec.py:65-92
# ----------------------------------------------------------------------------------------------
def sasrec_forward(params, items, masked_index, n_layers, n_heads, eps, drops=None, dtype=np.float32):
    """Returns (loss, cache).  items int64 [B,2,L+1], masked_index int64 [B,L]."""
    P = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    drops = drops or {}
    W = P["item_embedding.weight"]
    item_emb = gather_rows(W, items)               # sasrec.py:68
    pos = item_emb[:, 0]
    neg = item_emb[:, 1]
    inp = pos[:, :-1]                              # sasrec.py:72
    tgt_pos = pos[:, 1:]
    tgt_neg = neg[:, 1:]
    L = masked_index.shape[1]
    x0 = inp + P["position_embedding.weight"][:L][None]      # sasrec.py:77-81
    x, ln0 = layernorm_fwd(x0, P["LayerNorm.weight"], P["LayerNorm.bias"], eps)
    d0 = drops.get("emb")
    if d0 is not None:
        x = x * d0
    mask = attention_mask(masked_index, dtype)     # sasrec.py:84 (note: built from masked_index)
    out, caches = encoder_fwd(P, x, mask, n_layers, n_heads, eps, drops)
    loss, lc = bpr_loss_fwd(out, tgt_pos, tgt_neg, masked_index)
    cache = dict(P=P, items=items, ln0=ln0, d0=d0, caches=caches, out=out, tgt_pos=tgt_pos, tgt_neg=tgt_neg,
                 lc=lc, n_heads=n_heads, L=L, x_in=x, pos_score=lc[2], neg_score=lc[3])
    return loss, cache


def sasrec_backward(cache, dense_table_grad=True):
    """Gradients w.r.t. every parameter (dict keyed like state_dict).  'item_embedding.weight'
    is the DENSE [N,D] grad autograd would produce (row 0 == 0)."""
    P = cache["P"]
    grads = {}
    d_out, d_tp, d_tn = bpr_loss_bwd(cache["out"], cache["tgt_pos"], cache["tgt_neg"], cache["lc"])
    dx = encoder_bwd(P, d_out, cache["caches"], cache["n_heads"], grads)
    if cache["d0"] is not None:
        dx = dx * cache["d0"]
    dx0, dg, db = layernorm_bwd(dx, P["LayerNorm.weight"], cache["ln0"])
    grads["LayerNorm.weight"] = dg
    grads["LayerNorm.bias"] = db
    L = cache["L"]
    gp = np.zeros_like(P["position_embedding.weight"])
    gp[:L] = dx0.sum(0)
    grads["position_embedding.weight"] = gp
    B = dx0.shape[0]
    D = dx0.shape[-1]
    d_item = np.zeros((B, 2, L + 1, D), dtype=dx0.dtype)
    d_item[:, 0, :-1] += dx0
    d_item[:, 0, 1:] += d_tp
    d_item[:, 1, 1:] += d_tn
    grads["d_item_emb"] = d_item
    if dense_table_grad:
        grads["item_embedding.weight"] = scatter_add_rows(d_item, cache["items"], P["item_embedding.weight"].shape[0], 0,
                                                          dtype=dx0.dtype)
    return grads


# ----------------------------------------------------------------------------------------------
# optimizer: trainer.py:100-103,125  torch.optim.AdamW(lr, weight_decay) defaults betas/eps
# ----------------------------------------------------------------------------------------------
def adamw_step(w, g, m, v, step, lr, weight_decay, beta1=0.9, beta2=0.999, eps=1e-8):
    """One torch.optim.AdamW step (decoupled decay first, then the Adam update; `step` counts from 1).
    Returns new (w, m, v); dtype follows w."""
    dt = w.dtype.type
    w = w * dt(1.0 - lr * weight_decay)
    m = m * dt(beta1) + g * dt(1.0 - beta1)
    v = v * dt(beta2) + (g * g) * dt(1.0 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v) / dt(math.sqrt(bc2)) + dt(eps)
    w = w - dt(step_size) * (m / denom)
    return w.astype(w.dtype), m, v


# ----------------------------------------------------------------------------------------------
# evaluation: sasrec.py:94-113, trainer.py:327-337, evaluator/collector.py:131-139, metrics.py
# ----------------------------------------------------------------------------------------------
def sasrec_predict(params, item_seq, n_layers, n_heads, eps, item_feature=None, dtype=np.float32):
    """scores[B_e, N] = encoder(item_seq)[:, -1] @ item_feature.T   (sasrec.py:94-113)."""
    P = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    W = P["item_embedding.weight"]
    feat = W if item_feature is None else np.asarray(item_feature, dtype=dtype)
    L = item_seq.shape[1]
    x0 = gather_rows(W, item_seq) + P["position_embedding.weight"][:L][None]
    x, _ = layernorm_fwd(x0, P["LayerNorm.weight"], P["LayerNorm.bias"], eps)
    mask = attention_mask(item_seq, dtype)
    out, _ = encoder_fwd(P, x, mask, n_layers, n_heads, eps)
    seq_out = out[:, -1]
    return seq_out @ feat.T, seq_out


def full_sort_topk(scores, hist_u, hist_i, k):
    """trainer.py:334-336 (col 0 and history -> -inf) then torch.topk(scores, k) (collector.py:133).
    Returns (values [B,k], idx [B,k]) sorted descending; ties broken by lower item id."""
    s = np.array(scores, copy=True)
    s[:, 0] = -np.inf
    if hist_u is not None and len(hist_u):
        s[np.asarray(hist_u), np.asarray(hist_i)] = -np.inf
    order = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(s, order, 1), order


def full_catalog_ce(scores, target, mask_col0=True):
    """EXTENSION, parity unpinned by the reference (it trains with sampled negatives, sasrec.py:88-92): softmax
    cross-entropy of `target` [B] over the whole catalog, restating torch's F.cross_entropy(scores, target, reduction='none')
    with column 0 ([PAD]) removed from the softmax as trainer.py:334 removes it from the ranking.
    Returns (lse [B], target logit [B], nll [B] = lse - target logit)."""
    s = np.array(scores, dtype=np.float64, copy=True)
    if mask_col0:
        s[:, 0] = -np.inf
    m = s.max(1, keepdims=True)
    lse = (m + np.log(np.exp(s - m).sum(1, keepdims=True)))[:, 0]
    tl = s[np.arange(s.shape[0]), np.asarray(target)]
    return lse, tl, lse - tl


def full_catalog_ce_bwd(seq_out, W, target, dnll, mask_col0=True):
    """Backward of full_catalog_ce(seq_out @ W.T, target) (same extension; what pixelrec_b200.ops.ScoreCEFn.backward computes in
    column chunks on pr_gemm_tf32 + pr_ce_grad_chunk_f32):  dS = dnll[:, None] * (softmax(S) - onehot(target)) with the padding
    column excluded from the softmax (its dS is 0);  d seq_out = dS @ W,  d W = dS.T @ seq_out.  Returns (d seq_out, d W)."""
    X = np.asarray(seq_out, dtype=np.float64)
    Wd = np.asarray(W, dtype=np.float64)
    S = X @ Wd.T
    if mask_col0:
        S[:, 0] = -np.inf
    P = np.exp(S - S.max(1, keepdims=True))
    P /= P.sum(1, keepdims=True)
    P[np.arange(S.shape[0]), np.asarray(target)] -= 1.0
    dS = P * np.asarray(dnll, dtype=np.float64)[:, None]
    return dS @ Wd, dS.T @ X


def topk_hits(topk_idx, positive_u, positive_i, n_users):
    """collector.py:134-139: pos_idx[u, r] = 1 iff topk_idx[u, r] is u's positive item; pos_len = #positives."""
    pos = np.zeros(topk_idx.shape, dtype=np.int32)
    plen = np.zeros((n_users,), dtype=np.int32)
    for u, i in zip(np.asarray(positive_u), np.asarray(positive_i)):
        pos[u] |= (topk_idx[u] == i)
        plen[u] += 1
    return pos, plen


def recall_ndcg(pos_idx, pos_len, topk):
    """metrics.py:115-136 (Recall), :139-178 (NDCG), base_metric.py:43-67: per-rank SUMS over users."""
    pos_idx = pos_idx.astype(bool)
    K = pos_idx.shape[1]
    rec = np.cumsum(pos_idx, axis=1) / pos_len.reshape(-1, 1)
    len_rank = np.full_like(pos_len, K)
    idcg_len = np.where(pos_len > len_rank, len_rank, pos_len)
    ranks = np.tile(np.arange(1, K + 1, dtype=np.float64), (pos_idx.shape[0], 1))
    idcg = np.cumsum(1.0 / np.log2(ranks + 1), axis=1)
    for row, n in enumerate(idcg_len):
        idcg[row, n:] = idcg[row, n - 1]
    dcg = np.cumsum(np.where(pos_idx, 1.0 / np.log2(ranks + 1), 0), axis=1)
    ndcg = dcg / idcg
    res = {}
    for kk in topk:
        res[f"recall@{kk}"] = rec.sum(0)[kk - 1]
        res[f"ndcg@{kk}"] = ndcg.sum(0)[kk - 1]
    return res


# ----------------------------------------------------------------------------------------------
# batch construction: data/dataset/trainset.py:40-75 (SEQTrainDataset), evalset.py:4-36
# ----------------------------------------------------------------------------------------------
def seq_train_sample(item_seq, item_num, max_item_list_length, rng):
    """One SEQTrainDataset.__getitem__ (trainset.py:65-75): positives left-padded to L+1, one
    uniform negative per transition drawn by rejection against the sequence (trainset.py:40-44),
    aligned so neg[t] pairs with pos[t] and neg[0] == 0; mask has len(seq)-1 ones, left-padded to L.
    `rng` needs .randint(a, b) inclusive like random.randint."""
    Lp1 = max_item_list_length + 1
    seq = list(item_seq)
    s = set(seq)
    negs, mask = [], []
    for _ in range(len(seq) - 1):
        it = rng.randint(1, item_num - 1)
        while it in s:
            it = rng.randint(1, item_num - 1)
        negs.append(it)
        mask.append(1)
    pad = lambda xs, n: ([0] * (n - len(xs)) + list(xs))[-n:]
    items = np.array([pad(seq, Lp1), pad(negs, Lp1)], dtype=np.int64)
    return items, np.array(pad(mask, Lp1 - 1), dtype=np.int64)


def seq_batch_build(padded, sel, item_num, seed):
    """Restates pixelrec_b200/csrc/sampler.cu (the on-device form of trainset.py:52-75) with the same Philox stream, so the
    GPU batch builder is checked bit for bit; statistical properties are those of seq_train_sample above."""
    from oracle.philox_np import philox4x32_10
    padded = np.asarray(padded)
    sel = np.asarray(sel)
    W = padded.shape[1]
    B = len(sel)
    items = np.zeros((B, 2, W), dtype=np.int64)
    mask = np.zeros((B, W - 1), dtype=np.int64)
    rng_range = np.uint64(item_num - 1)
    for b in range(B):
        row = padded[sel[b]]
        items[b, 0] = row
        nz = np.flatnonzero(row != 0)
        first = int(nz[0]) if len(nz) else W
        present = set(row[first:].tolist())
        for t in range(first + 1, W):
            for attempt in range(64):
                ctr = np.array([np.uint64((b * W + t) * 64 + attempt)], dtype=np.uint64)
                x = np.uint64(philox4x32_10(ctr, 0x5eed, seed)[0, 0])
                neg = 1 + int((x * rng_range) >> np.uint64(32))
                if neg not in present:
                    break
            items[b, 1, t] = neg
            mask[b, t - 1] = 1
    return items, mask
