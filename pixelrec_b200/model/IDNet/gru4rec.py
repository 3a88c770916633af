"""GRU4Rec model plugin -- drop-in for REC/model/IDNet/gru4rec.py:10-84 (BASELINE config 5: alt backbone that
stresses the table gather / scatter-add and the pairwise loss; the recurrent body stays cuDNN's nn.GRU, as SURVEY
section 2 row 13 scopes it).  Table lookup, gradient scatter-add, loss and optimizer are our kernels.
"""
import torch
import torch.nn as nn
from torch.nn.init import xavier_normal_, xavier_uniform_

from ... import ops
from ...dist import ShardedTableEmbedding, make_table
from ...utils.enum_type import InputType
from ..basemodel import BaseModel
from ..layers import TableEmbedding
from .sasrec import _opt


class GRU4Rec(BaseModel):
    input_type = InputType.SEQ

    def __init__(self, config, data):
        super().__init__()
        self.embedding_size = config["embedding_size"]
        self.hidden_size = config["hidden_size"] * config["embedding_size"]       # gru4rec.py:17 (multiplier)
        self.num_layers = config["num_layers"]
        self.dropout_prob = config["dropout_prob"]
        self.user_num = getattr(data, "user_num", 0)
        self.item_num = data.item_num
        self.item_embedding = make_table(self.item_num, self.embedding_size, padding_idx=0,
                                         sharding=_opt(config, "table_sharding", "auto"))
        self.emb_dropout = nn.Dropout(self.dropout_prob)
        self.gru_layers = nn.GRU(input_size=self.embedding_size, hidden_size=self.hidden_size,
                                 num_layers=self.num_layers, bias=False, batch_first=True)
        self.dense = nn.Linear(self.hidden_size, self.embedding_size)
        self.apply(self._init_weights)

    def _init_weights(self, module):
        """gru4rec.py:40-45 (note: xavier on the table leaves the pad row non-zero, like the reference)."""
        if isinstance(module, (nn.Embedding, TableEmbedding, ShardedTableEmbedding)):
            xavier_normal_(module.weight)
        elif isinstance(module, nn.GRU):
            xavier_uniform_(module.weight_hh_l0)
            xavier_uniform_(module.weight_ih_l0)

    def forward(self, inputs):
        items, masked_index = inputs
        E = self.item_embedding(items)                                  # [B,2,L+1,D]   gru4rec.py:51
        x = self.emb_dropout(E[:, 0, :-1])                              # :55,59
        out, _ = self.gru_layers(x)
        out = self.dense(out).contiguous()                              # :61
        return ops.bpr_loss(out, E, masked_index.contiguous(), None)   # :63-67 (targets read in place from E)

    @torch.no_grad()
    def predict(self, item_seq, item_feature):
        """gru4rec.py:70-80: the history is embedded from `item_feature` itself."""
        return torch.matmul(self.encode_last(item_seq, item_feature), item_feature.t())

    @torch.no_grad()
    def encode_last(self, item_seq, item_feature=None):
        tab = self.item_embedding
        if item_feature is None and isinstance(tab, ShardedTableEmbedding) and tab.exchange == "p2p" and not self.training:
            rows = tab.lookup_static(item_seq, getattr(self, "train_lookups_hint", 0))              # sharded evaluation: no collective
        else:
            feat = self.compute_item_all() if item_feature is None else item_feature
            rows = ops.gather_rows(feat.contiguous(), item_seq.contiguous())
        x = self.emb_dropout(rows)
        out, _ = self.gru_layers(x)
        return self.dense(out)[:, -1].contiguous()

    @torch.no_grad()
    def compute_item_all(self):
        if isinstance(self.item_embedding, ShardedTableEmbedding):
            return self.item_embedding.full_weight()
        return self.item_embedding.weight
