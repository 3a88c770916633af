from .optim import FusedAdamW
from .trainer import Trainer

__all__ = ["Trainer", "FusedAdamW"]
