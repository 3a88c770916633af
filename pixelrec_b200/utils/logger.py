"""Logger: stream + ./log/<Model>/<time>.log, rank != 0 at WARN (reference REC/utils/logger.py:41-101)."""
import logging
import os

from .utils import ensure_dir, get_local_time, get_rank, barrier


def init_logger(config):
    rank = get_rank()
    root = "./log/"
    if rank == 0:
        ensure_dir(os.path.join(root, str(config["model"])))
    barrier()
    path = os.path.join(root, config["log_path"] or "{}/{}.log".format(config["model"], get_local_time()))
    level = {"debug": logging.DEBUG, "error": logging.ERROR, "warning": logging.WARNING,
             "critical": logging.CRITICAL}.get((config["state"] or "info").lower(), logging.INFO)
    fmt = logging.Formatter("%(asctime)-15s %(levelname)s  %(message)s", "%a %d %b %Y %H:%M:%S")
    handlers = [logging.StreamHandler()]
    try:
        handlers.append(logging.FileHandler(path))
    except OSError:
        pass
    for h in handlers:
        h.setLevel(level)
        h.setFormatter(fmt)
    logging.basicConfig(level=level if rank in (-1, 0) else logging.WARN, handlers=handlers, force=True)
