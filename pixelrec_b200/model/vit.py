"""CLIP ViT item encoder for PixelNet (reference: REC/model/load.py:90-117 builds HF `CLIPVisionModel(
'openai/clip-vit-base-patch32')` and REC/model/layers.py:100-128 wraps it in MeanItemEncoder).

Same module tree / parameter names / parameter ORDER as HF's CLIPVisionModel, so (a) HF / reference checkpoints
load with load_state_dict, (b) the reference's index-based freezing (`tune_scale: 165` = embeddings + pre-LN + the
first 10 of 12 layers, load.py:97-99) selects the same tensors.  The forward runs on our kernels: LayerNorm via
pr_add_ln_*, attention core via pr_sasrec_attn_* (bidirectional, no key padding, L = 50 tokens, dh = 64),
quick-GELU via pr_act_* or the GEMM epilogue; every Linear (q/k/v, out_proj, fc1 + quick-GELU, fc2, rec_fc + ReLU) runs forward,
input-gradient and weight-gradient on pr_gemm_tf32 (ops.linear) when TF32 matmuls are allowed; only the patch-embedding
convolution (frozen, forward only) stays a cuDNN call.  Frozen leading layers run under
no_grad (nothing is saved for a backward that never happens).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


class CLIPVisionConfig:
    def __init__(self, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                 image_size=224, patch_size=32, layer_norm_eps=1e-5, hidden_act="quick_gelu"):
        self.hidden_size, self.intermediate_size = hidden_size, intermediate_size
        self.num_hidden_layers, self.num_attention_heads = num_hidden_layers, num_attention_heads
        self.image_size, self.patch_size = image_size, patch_size
        self.layer_norm_eps, self.hidden_act = layer_norm_eps, hidden_act


class CLIPVisionEmbeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.class_embedding = nn.Parameter(torch.randn(c.hidden_size))
        self.patch_embedding = nn.Conv2d(3, c.hidden_size, kernel_size=c.patch_size, stride=c.patch_size, bias=False)
        self.num_positions = (c.image_size // c.patch_size) ** 2 + 1
        self.position_embedding = nn.Embedding(self.num_positions, c.hidden_size)

    def forward(self, pixel_values):
        n = pixel_values.shape[0]
        patches = self.patch_embedding(pixel_values).flatten(2).transpose(1, 2)          # [n, T-1, H]
        cls = self.class_embedding.expand(n, 1, -1)
        return torch.cat([cls, patches], dim=1) + self.position_embedding.weight[None]


class CLIPAttention(nn.Module):
    def __init__(self, c):
        super().__init__()
        H = c.hidden_size
        self.num_heads = c.num_attention_heads
        self.k_proj = nn.Linear(H, H)
        self.v_proj = nn.Linear(H, H)
        self.q_proj = nn.Linear(H, H)
        self.out_proj = nn.Linear(H, H)

    def forward(self, x):
        w = torch.cat([self.q_proj.weight, self.k_proj.weight, self.v_proj.weight], 0)
        b = torch.cat([self.q_proj.bias, self.k_proj.bias, self.v_proj.bias], 0)
        qkv = ops.linear(x, w, b)
        ctx = ops.attention(qkv, None, self.num_heads, causal=False)     # softmax(q k^T / sqrt(dh)) v, no mask
        return ops.linear(ctx, self.out_proj.weight, self.out_proj.bias)


class CLIPMLP(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.fc1 = nn.Linear(c.hidden_size, c.intermediate_size)
        self.fc2 = nn.Linear(c.intermediate_size, c.hidden_size)
        self.act = c.hidden_act

    def forward(self, x):
        h = ops.linear(x, self.fc1.weight, self.fc1.bias, act=self.act)   # bias + quick-GELU in the GEMM epilogue
        return ops.linear(h, self.fc2.weight, self.fc2.bias)


def _ln(x, ln):
    return ops.add_ln(x.contiguous(), None, ln.weight, ln.bias, ln.eps)


class CLIPEncoderLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.self_attn = CLIPAttention(c)
        self.layer_norm1 = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.mlp = CLIPMLP(c)
        self.layer_norm2 = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)

    def forward(self, x):
        x = x + self.self_attn(_ln(x, self.layer_norm1))            # pre-LN blocks
        return x + self.mlp(_ln(x, self.layer_norm2))

    def trainable(self):
        return any(p.requires_grad for p in self.parameters())


class CLIPEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layers = nn.ModuleList([CLIPEncoderLayer(c) for _ in range(c.num_hidden_layers)])


class CLIPVisionTransformer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.embeddings = CLIPVisionEmbeddings(c)
        self.pre_layrnorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)      # (sic) HF's spelling
        self.encoder = CLIPEncoder(c)
        self.post_layernorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)

    def forward(self, pixel_values):
        layers = list(self.encoder.layers)
        first_trainable = next((i for i, l in enumerate(layers) if l.trainable()), len(layers))
        stem_trainable = any(p.requires_grad for p in list(self.embeddings.parameters()) + list(self.pre_layrnorm.parameters()))
        frozen_prefix = 0 if stem_trainable else first_trainable

        def run(lo, hi, x):
            for l in layers[lo:hi]:
                x = l(x)
            return x
        if frozen_prefix > 0 and not pixel_values.requires_grad:
            with torch.no_grad():                                       # frozen-layer fast path (ViT.yaml tune_scale)
                x = _ln(self.embeddings(pixel_values), self.pre_layrnorm)
                x = run(0, frozen_prefix, x)
            x = run(frozen_prefix, len(layers), x)
        else:
            x = run(0, len(layers), _ln(self.embeddings(pixel_values), self.pre_layrnorm))
        pooled = self.post_layernorm(x[:, 0]) if isinstance(self.post_layernorm, nn.LayerNorm) else x[:, 0]
        return x, pooled                                                 # (last_hidden_state, pooler_output)


class CLIPVisionModel(nn.Module):
    def __init__(self, config=None):
        super().__init__()
        self.config = config or CLIPVisionConfig()
        self.vision_model = CLIPVisionTransformer(self.config)

    def forward(self, pixel_values):
        return self.vision_model(pixel_values)


class Identity(nn.Module):
    def forward(self, x):
        return x


class MeanItemEncoder(nn.Module):
    """REC/model/layers.py:100-128: rec_fc(Linear + activation) on every token, then the mean over tokens."""

    def __init__(self, item_encoder, input_dim, output_dim, act_name="relu"):
        super().__init__()
        self.item_encoder = item_encoder
        self.rec_fc = nn.Sequential(nn.Linear(input_dim, output_dim), nn.ReLU() if act_name == "relu" else nn.Identity())
        self.act_name = act_name

    def forward(self, x):
        h = self.item_encoder(x)[0]                                      # [n, T, H]
        fc = self.rec_fc[0]
        y = ops.linear(h, fc.weight, fc.bias, act=self.act_name if self.act_name in ops.ACT_IDS else None)
        return y.mean(dim=1)
