"""Glue utilities with the reference's behaviour (REC/utils/utils.py): model lookup by name (:35-61),
early stopping (:65-106), seeding (:138-156)."""
import datetime
import importlib
import importlib.util
import os
import random

import numpy as np
import torch


def dist_ready():
    return torch.distributed.is_available() and torch.distributed.is_initialized()


def get_rank():
    return torch.distributed.get_rank() if dist_ready() else 0


def get_world_size():
    return torch.distributed.get_world_size() if dist_ready() else 1


def barrier():
    if dist_ready():
        torch.distributed.barrier()


def get_local_time():
    barrier()
    return datetime.datetime.now().strftime("%b-%d-%Y_%H-%M-%S")


def ensure_dir(dir_path):
    os.makedirs(dir_path, exist_ok=True)


def get_model(model_name):
    """`model: SASRec` -> pixelrec_b200.model.IDNet.sasrec.SASRec, then PixelNet (utils.py:35-61)."""
    fname = model_name.lower()
    for family in ("IDNet", "PixelNet"):
        path = f"pixelrec_b200.model.{family}.{fname}"
        if importlib.util.find_spec(path) is not None:
            return getattr(importlib.import_module(path), model_name)
    raise ValueError("`model_name` [{}] is not the name of an existing model.".format(model_name))


def early_stopping(value, best, cur_step, max_step, bigger=True):
    """(best, cur_step, stop_flag, update_flag) -- utils.py:65-106."""
    better = value >= best if bigger else value <= best
    if better:
        return value, 0, False, True
    cur_step += 1
    return best, cur_step, cur_step > max_step, False


def calculate_valid_score(valid_result, valid_metric=None):
    return valid_result[valid_metric] if valid_metric else valid_result["Recall@10"]


def dict2str(result_dict):
    return "    ".join(f"{k} : {v}" for k, v in result_dict.items())


def init_seed(seed, reproducibility):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = not reproducibility
    torch.backends.cudnn.deterministic = bool(reproducibility)


def set_color(log, color=None, highlight=True):
    return log
