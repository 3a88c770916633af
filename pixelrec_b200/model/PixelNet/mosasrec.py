"""MOSASRec model plugin (PixelNet) -- drop-in for REC/model/PixelNet/mosasrec.py:9-128: the SASRec body fed by item
vectors computed from raw pixels with the CLIP ViT item encoder (model/vit.py, our attention / LayerNorm kernels)
instead of an ID table.  Same constructor, forward / predict / compute_item contract and parameter names
(`visual_encoder.item_encoder.vision_model...`, `visual_encoder.rec_fc.0...`, `trm_encoder...`)."""
import torch
from torch import nn

from ... import ops
from ...utils.enum_type import InputType
from ..basemodel import BaseModel
from ..IDNet.sasrec import _opt
from ..layers import TransformerEncoder
from ..load import load_model


class MOSASRec(BaseModel):
    input_type = InputType.SEQ

    def __init__(self, config, dataload):
        super().__init__()
        self.pretrain_weights = _opt(config, "pretrain_path")
        self.n_layers = config["n_layers"]
        self.n_heads = config["n_heads"]
        self.embedding_size = config["embedding_size"]
        self.inner_size = config["inner_size"] * self.embedding_size
        self.hidden_dropout_prob = config["hidden_dropout_prob"]
        self.attn_dropout_prob = config["attn_dropout_prob"]
        self.hidden_act = config["hidden_act"]
        self.layer_norm_eps = config["layer_norm_eps"]
        self.initializer_range = config["initializer_range"]
        self.max_seq_length = config["MAX_ITEM_LIST_LENGTH"]
        self.item_num = dataload.item_num

        self.visual_encoder = load_model(config=config)
        if self.pretrain_weights:
            self.load_weights(self.pretrain_weights)
        self.position_embedding = nn.Embedding(self.max_seq_length, self.embedding_size)
        self.LayerNorm = nn.LayerNorm(self.embedding_size, eps=self.layer_norm_eps)
        self.trm_encoder = TransformerEncoder(
            n_layers=self.n_layers, n_heads=self.n_heads, hidden_size=self.embedding_size, inner_size=self.inner_size,
            hidden_dropout_prob=self.hidden_dropout_prob, attn_dropout_prob=self.attn_dropout_prob,
            hidden_act=self.hidden_act, layer_norm_eps=self.layer_norm_eps)
        self.rng = ops.DropoutRng(_opt(config, "seed", 0))
        self.position_embedding.weight.data.normal_(mean=0.0, std=self.initializer_range)      # mosasrec.py:49-52
        self.trm_encoder.apply(self._init_weights)
        self.LayerNorm.bias.data.zero_()
        self.LayerNorm.weight.data.fill_(1.0)

    def _init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.initializer_range)
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def _embed(self, E, L, seq_stride, B, seed, slab):
        p = self.hidden_dropout_prob if self.training else 0.0
        return ops.add_ln(E, self.position_embedding.weight, self.LayerNorm.weight, self.LayerNorm.bias, self.layer_norm_eps,
                          p_post=p, seed=seed, stream_post=0, layout=(L, seq_stride, B), res_period=L, slab=slab)

    def forward(self, interaction):
        items, masked_index = interaction                     # images [B, 2(L+1), 3, H, W] pos/neg interleaved, mask [B,L]
        B, L = masked_index.shape
        D = self.embedding_size
        seed = self.rng.next_seed()
        item_emb = self.visual_encoder(items.flatten(0, 1)).view(B, -1, 2, D)        # mosasrec.py:69  [B, L+1, 2, D]
        E = item_emb.permute(0, 2, 1, 3).contiguous()                                # -> [B, 2, L+1, D] (our kernels' layout)
        slab = ops.GradSlab()
        masked_index = masked_index.contiguous()
        x = self._embed(E, L, 2 * (L + 1) * D, B, seed, slab)
        out = self.trm_encoder(x, masked_index, output_all_encoded_layers=False, causal=True, seed=seed)[-1]
        return ops.bpr_loss(out, E, masked_index, slab)                               # mosasrec.py:89-93

    @torch.no_grad()
    def encode_last(self, item_seq, item_feature):
        item_seq = item_seq.contiguous()
        B, L = item_seq.shape
        E = ops.gather_rows(item_feature.contiguous(), item_seq)                      # item_feature[item_seq], mosasrec.py:103
        x = self._embed(E, L, L * self.embedding_size, B, 0, None)
        out = self.trm_encoder(x, item_seq, output_all_encoded_layers=False, causal=True, seed=0)[-1]
        return out[:, -1].contiguous()

    @torch.no_grad()
    def predict(self, item_seq, item_feature):
        return torch.matmul(self.encode_last(item_seq, item_feature), item_feature.t())

    @torch.no_grad()
    def compute_item(self, item):
        return self.visual_encoder(item)
