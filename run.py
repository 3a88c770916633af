"""Per-rank driver -- same role and flow as the reference's code/run.py:18-82:
Config -> seed -> logger -> data -> model plugin -> Trainer.fit -> Trainer.evaluate.

Differences (B200-native): no DistributedDataParallel wrapper -- data parallelism is one all_reduce of the fused
optimizer's flat gradient buffer, and the item table is row-sharded (pixelrec_b200/dist.py) instead of being
replicated and all-reduced densely; works without torchrun too (single process, single GPU).
"""
import argparse
import os
from logging import getLogger

import torch
import torch.distributed as dist

from pixelrec_b200.config import Config
from pixelrec_b200.data import bulid_dataloader, load_data
from pixelrec_b200.dist import broadcast_dense_params
from pixelrec_b200.trainer import Trainer
from pixelrec_b200.utils import get_model, init_seed
from pixelrec_b200.utils.logger import init_logger


def run_loop(local_rank, config_file=None, saved=True, config_dict=None):
    config = Config(config_file_list=config_file, config_dict=config_dict)
    if not torch.cuda.is_available():
        raise RuntimeError("pixelrec_b200 runs on CUDA devices only (no CPU fallback)")
    device = torch.device("cuda", local_rank)
    config["device"] = device
    init_seed(config["seed"], config["reproducibility"])
    init_logger(config)
    logger = getLogger()
    prec = (config["matmul_precision"] or "tf32").lower()
    torch.backends.cuda.matmul.allow_tf32 = prec == "tf32"
    torch.backends.cudnn.allow_tf32 = prec == "tf32"

    dataload = load_data(config)
    train_loader, valid_loader, test_loader = bulid_dataloader(config, dataload)
    model = get_model(config["model"])(config, dataload).to(device)
    broadcast_dense_params(model)
    world_size = dist.get_world_size() if dist.is_initialized() else 1
    logger.info(f"\nWorld_Size = {world_size} \n")
    logger.info(config)
    logger.info(dataload)
    logger.info(model)

    trainer = Trainer(config, model)
    best_valid_score, best_valid_result = trainer.fit(train_loader, valid_loader, saved=saved,
                                                      show_progress=config["show_progress"])
    test_result = trainer.evaluate(test_loader, load_best_model=saved, show_progress=config["show_progress"])
    logger.info(f"best valid : {best_valid_result}")
    logger.info(f"test result: {test_result}")
    return {"best_valid_score": best_valid_score, "valid_score_bigger": config["valid_metric_bigger"],
            "best_valid_result": best_valid_result, "test_result": test_result}


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--config_file", nargs="+", type=str)
    args = parser.parse_args()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    run_loop(local_rank=local_rank, config_file=args.config_file)
    if dist.is_initialized():
        dist.destroy_process_group()
