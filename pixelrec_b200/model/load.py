"""Image-encoder factory (reference: REC/model/load.py:9-164).  On the hot path named by BASELINE.json only the
HF-CLIP ViT branch exists (`encoder_source: transformers`, `encoder_name: clip-vit-base-patch32`, overall/ViT.yaml:17-18;
`clip-vit-base-patch16` is accepted too -- BASELINE config 4 names ViT-B/16).  The torchvision / openai-clip / Swin /
BEiT / MAE branches of the reference are out of scope."""
import torch

from .vit import CLIPVisionConfig, CLIPVisionModel, Identity, MeanItemEncoder

_PATCH = {"clip-vit-base-patch32": 32, "clip-vit-base-patch16": 16}


def load_model(config):
    source = config["encoder_source"]
    name = config["encoder_name"]
    if source != "transformers" or name not in _PATCH:
        raise NotImplementedError(f"image encoder {source}/{name} is outside the B200 hot path; supported: {sorted(_PATCH)}")
    ft = config["fine_tune_arg"] or {}
    tune_scale = ft.get("tune_scale", 0)
    method = ft.get("method", "mean")
    act = ft.get("activation", "relu")
    vcfg = config["vit_config"] or {}
    model = CLIPVisionModel(CLIPVisionConfig(patch_size=_PATCH[name], **vcfg))
    weights = config["encoder_weights"]          # offline: no hub download; optional local state-dict
    if weights:
        model.load_state_dict(torch.load(weights, map_location="cpu"), strict=True)
    for index, (_, p) in enumerate(model.named_parameters()):          # load.py:97-99
        if index < tune_scale:
            p.requires_grad = False
        elif not ft.get("pre_trained", True):
            p.data.normal_(mean=0.0, std=0.02)
    if method != "mean":
        raise NotImplementedError("fine_tune_arg.method: only 'mean' (overall/ViT.yaml:30) is on the hot path")
    model.vision_model.post_layernorm = Identity()                      # load.py:114-115
    return MeanItemEncoder(model, model.config.hidden_size, config["embedding_size"], act)
