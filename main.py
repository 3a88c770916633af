"""Launcher with the reference's CLI (code/main.py:9-31):

    python main.py --device 0,1,2,3 --config_file configs/IDNet/sasrec.yaml configs/overall/ID.yaml

spawns one run.py per listed GPU with torch.distributed.run on 127.0.0.1.  Unlike the reference it does NOT set
CUDA_LAUNCH_BLOCKING=1 (main.py:6 serialises every launch) and does not force OMP_NUM_THREADS=1.
"""
import argparse
import os
import random
import subprocess
import sys

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--device", default="0", type=str)
    parser.add_argument("--config_file", nargs="+")
    args = parser.parse_args()
    if not args.config_file or len(args.config_file) > 2:
        parser.error("--config_file takes one or two yaml files (model yaml [overall yaml])")
    nproc = len(args.device.split(","))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=args.device, TOKENIZERS_PARALLELISM="false")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(random.randint(10002, 19999)),
           os.path.join(here, "run.py"), "--config_file", *args.config_file]
    sys.exit(subprocess.call(cmd, env=env))
