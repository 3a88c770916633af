"""Transformer blocks of the SASRec hot path, same module tree / parameter names as the reference
(REC/model/layers.py:543-759) so checkpoints interchange, but every non-GEMM op is one of our sm_100a
kernels (pixelrec_b200/ops.py); the Linear layers are cuBLAS calls (TF32 by default, as torch 1.10 ran
them on Ampere -- set matmul_precision: fp32 in the yaml for strict fp32).
"""
import copy
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


# One hand-written autograd node per encoder layer (ops.TransformerLayerFn) instead of ~12 nodes; PR_FUSED_LAYER=0 selects
# the op-by-op composition (identical kernels and numerics up to summation order of the bias gradients).
FUSED_LAYER = os.environ.get("PR_FUSED_LAYER", "1") != "0"


class TableGradSink:
    """Receives the sparse table gradient (ScatterPlan + reduced rows) from GatherFn.backward.
    The fused optimizer (pixelrec_b200/trainer/optim.py) consumes and clears it."""

    def __init__(self, table):
        self.table = table
        self.sparse = False        # switched on by the fused optimizer; default = dense reference semantics
        self.row2slot = None
        self.pending = []

    def enable_sparse(self):
        w = self.table.weight
        self.row2slot = torch.full((w.shape[0],), -1, device=w.device, dtype=torch.int32)
        self.sparse = True

    def deposit(self, plan, rows):
        self.pending.append((plan, rows))


class TableEmbedding(nn.Module):
    """nn.Embedding(N, D, padding_idx) replacement (sasrec.py:31): weight [N,D] fp32, gather through
    pr_gather_rows_f32, gradient through pr_scatter_plan + pr_scatter_add_rows_f32."""

    def __init__(self, num_embeddings, embedding_dim, padding_idx=None):
        super().__init__()
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.padding_idx = padding_idx
        self.weight = nn.Parameter(torch.empty(num_embeddings, embedding_dim))
        nn.init.normal_(self.weight)
        if padding_idx is not None:
            with torch.no_grad():
                self.weight[padding_idx].fill_(0)
        self.sink = TableGradSink(self)
        self.gather_impl = 0

    def forward(self, idx):
        return ops.GatherFn.apply(self.weight, idx.contiguous(), self.padding_idx, self.sink, self.gather_impl)

    def prefetch(self, idx):
        """no-op on one GPU (dist.ShardedTableEmbedding pre-computes its exchange plan here)"""

    def extra_repr(self):
        return f"{self.num_embeddings}, {self.embedding_dim}, padding_idx={self.padding_idx}"


class MultiHeadAttention(nn.Module):
    """layers.py:543-617."""

    def __init__(self, n_heads, hidden_size, hidden_dropout_prob, attn_dropout_prob, layer_norm_eps):
        super().__init__()
        if hidden_size % n_heads != 0:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention "
                             "heads (%d)" % (hidden_size, n_heads))
        self.num_attention_heads = n_heads
        self.attention_head_size = hidden_size // n_heads
        self.all_head_size = hidden_size
        self.query = nn.Linear(hidden_size, hidden_size)
        self.key = nn.Linear(hidden_size, hidden_size)
        self.value = nn.Linear(hidden_size, hidden_size)
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.LayerNorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)
        self.attn_dropout_prob = attn_dropout_prob
        self.hidden_dropout_prob = hidden_dropout_prob
        self.layer_norm_eps = layer_norm_eps

    def forward(self, x, key_ids, causal, seed, site):
        p_attn = self.attn_dropout_prob if self.training else 0.0
        p_hid = self.hidden_dropout_prob if self.training else 0.0
        w = torch.cat([self.query.weight, self.key.weight, self.value.weight], 0)
        b = torch.cat([self.query.bias, self.key.bias, self.value.bias], 0)
        qkv = F.linear(x, w, b)                                                 # layers.py:586-588 (one GEMM)
        ctx = ops.attention(qkv, key_ids, self.num_attention_heads, causal, p_attn, seed, site)   # :590-612
        h = self.dense(ctx)                                                     # :613
        return ops.add_ln(h, x, self.LayerNorm.weight, self.LayerNorm.bias, self.layer_norm_eps,
                          p_pre=p_hid, seed=seed, stream_pre=site + 1)          # :614-615


class FeedForward(nn.Module):
    """layers.py:620-673."""

    def __init__(self, hidden_size, inner_size, hidden_dropout_prob, hidden_act, layer_norm_eps):
        super().__init__()
        if hidden_act not in ops.ACT_IDS:
            raise KeyError(hidden_act)
        self.dense_1 = nn.Linear(hidden_size, inner_size)
        self.dense_2 = nn.Linear(inner_size, hidden_size)
        self.LayerNorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)
        self.hidden_act = hidden_act
        self.hidden_dropout_prob = hidden_dropout_prob
        self.layer_norm_eps = layer_norm_eps

    def forward(self, x, seed, site):
        p_hid = self.hidden_dropout_prob if self.training else 0.0
        h = ops.activation(self.dense_1(x), self.hidden_act)                    # :666-667
        h = self.dense_2(h)                                                     # :669
        return ops.add_ln(h, x, self.LayerNorm.weight, self.LayerNorm.bias, self.layer_norm_eps,
                          p_pre=p_hid, seed=seed, stream_pre=site + 2)          # :670-671


class TransformerLayer(nn.Module):
    def __init__(self, n_heads, hidden_size, intermediate_size, hidden_dropout_prob, attn_dropout_prob, hidden_act,
                 layer_norm_eps):
        super().__init__()
        self.multi_head_attention = MultiHeadAttention(n_heads, hidden_size, hidden_dropout_prob, attn_dropout_prob,
                                                       layer_norm_eps)
        self.feed_forward = FeedForward(hidden_size, intermediate_size, hidden_dropout_prob, hidden_act, layer_norm_eps)

    def forward(self, x, key_ids, causal, seed, site):
        if FUSED_LAYER and x.is_cuda:
            m, f = self.multi_head_attention, self.feed_forward
            p_attn = m.attn_dropout_prob if self.training else 0.0
            p_hid = m.hidden_dropout_prob if self.training else 0.0
            return ops.TransformerLayerFn.apply(
                x.contiguous(), key_ids, m.query.weight, m.query.bias, m.key.weight, m.key.bias, m.value.weight, m.value.bias,
                m.dense.weight, m.dense.bias, m.LayerNorm.weight, m.LayerNorm.bias, f.dense_1.weight, f.dense_1.bias,
                f.dense_2.weight, f.dense_2.bias, f.LayerNorm.weight, f.LayerNorm.bias, m.num_attention_heads, bool(causal),
                float(m.layer_norm_eps), float(p_attn), float(p_hid), ops.ACT_IDS[f.hidden_act], int(seed), int(site))
        return self.feed_forward(self.multi_head_attention(x, key_ids, causal, seed, site), seed, site)


class TransformerEncoder(nn.Module):
    """layers.py:706-759.  Takes the key ids the mask derives from (sasrec.py:119-126) instead of a
    materialised [B,1,L,L] additive mask; the attention kernel builds the mask in registers."""

    def __init__(self, n_layers=2, n_heads=2, hidden_size=64, inner_size=256, hidden_dropout_prob=0.5,
                 attn_dropout_prob=0.5, hidden_act="gelu", layer_norm_eps=1e-12):
        super().__init__()
        layer = TransformerLayer(n_heads, hidden_size, inner_size, hidden_dropout_prob, attn_dropout_prob, hidden_act,
                                 layer_norm_eps)
        self.layer = nn.ModuleList([copy.deepcopy(layer) for _ in range(n_layers)])

    def forward(self, hidden_states, key_ids, output_all_encoded_layers=True, causal=True, seed=0):
        outs = []
        for i, layer_module in enumerate(self.layer):
            hidden_states = layer_module(hidden_states, key_ids, causal, seed, 1 + 3 * i)
            if output_all_encoded_layers:
                outs.append(hidden_states)
        if not output_all_encoded_layers:
            outs.append(hidden_states)
        return outs
