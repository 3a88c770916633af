"""SASRec model plugin -- drop-in for REC/model/IDNet/sasrec.py:9-126 (same class name, constructor,
forward / predict / compute_item_all contract, parameter names), running on hand-written sm_100a kernels.
"""
import torch
from torch import nn

from ... import ops
from ...utils.enum_type import InputType
from ..basemodel import BaseModel
from ...dist import ShardedTableEmbedding, make_table
from ..layers import TableEmbedding, TransformerEncoder


def _opt(config, key, default=None):
    """config[key] with the reference Config's missing-key-is-None behaviour, for plain dicts too."""
    try:
        v = config[key]
    except KeyError:
        v = None
    return default if v is None else v


class SASRec(BaseModel):
    input_type = InputType.SEQ

    def __init__(self, config, dataload):
        super().__init__()
        self.n_layers = config["n_layers"]
        self.n_heads = config["n_heads"]
        self.hidden_size = config["embedding_size"]
        self.inner_size = config["inner_size"] * self.hidden_size          # sasrec.py:19-21 (multiplier)
        self.hidden_dropout_prob = config["hidden_dropout_prob"]
        self.attn_dropout_prob = config["attn_dropout_prob"]
        self.hidden_act = config["hidden_act"]
        self.layer_norm_eps = config["layer_norm_eps"]
        self.initializer_range = config["initializer_range"]
        self.max_seq_length = config["MAX_ITEM_LIST_LENGTH"]
        self.item_num = dataload.item_num

        self.item_embedding = make_table(self.item_num, self.hidden_size, padding_idx=0,
                                         sharding=_opt(config, "table_sharding", "auto"))
        self.position_embedding = nn.Embedding(self.max_seq_length, self.hidden_size)
        self.trm_encoder = TransformerEncoder(
            n_layers=self.n_layers, n_heads=self.n_heads, hidden_size=self.hidden_size, inner_size=self.inner_size,
            hidden_dropout_prob=self.hidden_dropout_prob, attn_dropout_prob=self.attn_dropout_prob,
            hidden_act=self.hidden_act, layer_norm_eps=self.layer_norm_eps)
        self.LayerNorm = nn.LayerNorm(self.hidden_size, eps=self.layer_norm_eps)
        self.rng = ops.DropoutRng(_opt(config, "seed", 0))
        self.apply(self._init_weights)

    def _init_weights(self, module):
        """sasrec.py:51-61.  NB: like the reference this re-initialises the pad row (id 0) to N(0, std)."""
        if isinstance(module, ShardedTableEmbedding):
            module.init_normal_(0.0, self.initializer_range)       # rows rank::world of the logical table (iid across ranks)
        elif isinstance(module, (nn.Linear, nn.Embedding, TableEmbedding)):
            module.weight.data.normal_(mean=0.0, std=self.initializer_range)
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def _embed(self, E, rows_per_seq, seq_stride, n_seq, seed, slab):
        p = self.hidden_dropout_prob if self.training else 0.0
        return ops.add_ln(E, self.position_embedding.weight, self.LayerNorm.weight, self.LayerNorm.bias,
                          self.layer_norm_eps, p_post=p, seed=seed, stream_post=0,
                          layout=(rows_per_seq, seq_stride, n_seq), res_period=rows_per_seq, slab=slab)

    def forward(self, interaction):
        items, masked_index = interaction                       # [B,2,L+1] int64, [B,L] int64
        B, L = masked_index.shape
        D = self.hidden_size
        seed = self.rng.next_seed()
        E = self.item_embedding(items)                           # sasrec.py:68  [B,2,L+1,D]
        slab = ops.GradSlab()
        masked_index = masked_index.contiguous()
        # input_emb = E[:,0,:-1] + P ; LayerNorm ; dropout        sasrec.py:72,77-83 (read in place from E)
        x = self._embed(E, L, 2 * (L + 1) * D, B, seed, slab)
        out = self.trm_encoder(x, masked_index, output_all_encoded_layers=False, causal=True, seed=seed)[-1]
        # pos/neg dot products + log-sigmoid loss                  sasrec.py:88-92 (targets read in place from E)
        return ops.bpr_loss(out, E, masked_index, slab)

    def prefetch(self, interaction):
        """Optional hint with the NEXT batch (same tensors that will be passed to forward): lets a row-sharded table
        overlap its index exchange with the current step (Trainer._train_epoch and bench.py call it)."""
        self.item_embedding.prefetch(interaction[0])

    @torch.no_grad()
    def predict(self, item_seq, item_feature):
        """sasrec.py:94-113: scores [B_e, N] = encoder(item_seq)[:, -1] @ item_feature.T"""
        seq_output = self.encode_last(item_seq, item_feature)
        return torch.matmul(seq_output, item_feature.t())

    @torch.no_grad()
    def encode_last(self, item_seq, item_feature=None):
        """Encoder output at the last position.  With a row-sharded table pass the all-gathered `item_feature`
        (compute_item_all()): the rows are then read locally -- the sharded lookup issues collectives, and ranks may hold
        different numbers of eval batches (data/utils.py strided sampler), which would deadlock."""
        item_seq = item_seq.contiguous()
        B, L = item_seq.shape
        D = self.hidden_size
        if item_feature is not None and isinstance(self.item_embedding, ShardedTableEmbedding):
            E = ops.gather_rows(item_feature.contiguous(), item_seq)
        elif isinstance(self.item_embedding, ShardedTableEmbedding) and self.item_embedding.exchange == "p2p" and not self.training:
            E = self.item_embedding.lookup_static(item_seq, getattr(self, "train_lookups_hint", 0))   # sharded evaluation: no collective
        else:
            E = self.item_embedding(item_seq)                     # [B,L,D]
        x = self._embed(E, L, L * D, B, 0, None)
        out = self.trm_encoder(x, item_seq, output_all_encoded_layers=False, causal=True, seed=0)[-1]
        return out[:, -1].contiguous()

    @torch.no_grad()
    def compute_item_all(self):
        """sasrec.py:115-117; a row-sharded table is all-gathered into the reference's [N, D] layout."""
        if isinstance(self.item_embedding, ShardedTableEmbedding):
            return self.item_embedding.full_weight()
        return self.item_embedding.weight
