"""Plugin base class (reference: REC/model/basemodel.py:10-33)."""
import numpy as np
import torch
import torch.nn as nn


class BaseModel(nn.Module):
    def __init__(self):
        super().__init__()

    def load_weights(self, path):
        """basemodel.py:17-21 -- loads checkpoint['state_dict'] non-strictly, renaming the legacy
        'item_embedding.rec_fc' prefix."""
        checkpoint = torch.load(path, map_location="cpu")
        sd = {k.replace("item_embedding.rec_fc", "visual_encoder.item_encoder.fc"): v
              for k, v in checkpoint["state_dict"].items()}
        self.load_state_dict(sd, strict=False)

    def __str__(self):
        n = sum(int(np.prod(p.size())) for p in self.parameters() if p.requires_grad)
        return super().__str__() + f"\nTrainable parameters: {n}"
