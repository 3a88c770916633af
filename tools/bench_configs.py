"""Throughput of the BASELINE.json configurations that bench.py does not cover (bench.py measures configs[1], the one the
headline metric is quoted on).  Same step definition (zero_grad -> forward -> backward -> fused AdamW, dropout as shipped),
same timing rules (CUDA events, barrier + synchronize on both sides, max over ranks), one JSON line per run.

    c3  IDNet/sasrec Pixel8M-shape: N=408375 items, emb_dim=2048, seq_len=20, table row-sharded over the launched ranks
    c4  PixelNet/sasrec + CLIP ViT item encoder on synthetic 224x224 pixels (--patch 32: the reference yaml; 16: BASELINE text)
    c5  IDNet/gru4rec Pixel1M-shape: N=100001 items, emb_dim=4096, seq_len=10 (gather / scatter stress; GRU body = cuDNN)

    python tools/bench_configs.py --config c5 [--batch B] [--steps K] [--warmup W]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_configs.py --config c3
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

SHAPES = {
    "c3": dict(model="SASRec", N=408375, D=2048, L=20, batch=1024),
    "c4": dict(model="MOSASRec", N=97001, D=512, L=10, batch=16),
    "c5": dict(model="GRU4Rec", N=100001, D=4096, L=10, batch=1024),
}


def cpu_port(sh, N, D, L, g, perm, p, B_cpu=64):
    """The same step on the host cores: oracle/torch_port.py (SASRec) or a torch-CPU restatement of gru4rec.py:47-67 +
    trainer.py:116-125 (GRU4Rec); ONE timed step of B_cpu sequences after one warm-up (the dense AdamW over the table alone
    moves 7*N*D*4 bytes: 23 GB at C3)."""
    import time
    torch.set_num_threads(os.cpu_count() or 1)
    items, mask = bench.synth_batch(g, B_cpu, N, L, perm, p)
    items, mask = torch.from_numpy(items), torch.from_numpy(mask)
    if sh["model"] == "SASRec":
        from oracle import torch_port as TP
        P = TP.init_params(N, D, L, 2, 2, seed=2020)
        step = TP.TrainStep(P, 2, 4, lr=1e-4, weight_decay=0.1, p_drop=0.1)
    elif sh["model"] == "GRU4Rec":
        import torch.nn as nn
        import torch.nn.functional as F
        emb = nn.Embedding(N, D, padding_idx=0)
        gru = nn.GRU(D, D, 1, bias=False, batch_first=True)
        dense = nn.Linear(D, D)
        params = list(emb.parameters()) + list(gru.parameters()) + list(dense.parameters())
        opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01)

        def step(it, mk):
            opt.zero_grad(set_to_none=False)
            E = emb(it)
            out = dense(gru(E[:, 0, :-1])[0])
            ps, ns = (out * E[:, 0, 1:]).sum(-1), (out * E[:, 1, 1:]).sum(-1)
            loss = -(torch.log((ps - ns).sigmoid() + 1e-8) * mk).sum(-1).mean(-1)
            loss.backward()
            opt.step()
            return loss.detach()
    else:
        return None                      # PixelNet: the ViT forward of 352 images per sample is minutes on CPU; not timed
    step(items, mask)
    t0 = time.perf_counter()
    step(items, mask)
    dt = time.perf_counter() - t0
    return {"value": B_cpu / dt, "unit": "sequences/s", "cores": torch.get_num_threads(), "kind": "port", "ms_per_step": dt * 1e3,
            "sample": f"1 timed step x {B_cpu} sequences (1 warm-up), torch CPU fp32 on {torch.get_num_threads()} threads"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=sorted(SHAPES))
    ap.add_argument("--batch", type=int, default=0, help="sequences per GPU per step (default: per config)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--patch", type=int, default=32, choices=[16, 32], help="c4: ViT-B/<patch>")
    ap.add_argument("--exchange", default=None, choices=["nccl", "p2p"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.exchange:
        os.environ["PR_EXCHANGE"] = args.exchange
    from pixelrec_b200.dist import broadcast_dense_params
    from pixelrec_b200.trainer.optim import FusedAdamW
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    sh = SHAPES[args.config]
    B = args.batch or sh["batch"]
    N, D, L = sh["N"], sh["D"], sh["L"]
    torch.manual_seed(2020)

    class Dl:
        item_num = N
    g = np.random.default_rng(1000 + rank)
    perm, p = bench.popularity(N)
    if sh["model"] == "SASRec":
        from pixelrec_b200.model.IDNet.sasrec import SASRec
        cfg = dict(n_layers=2, n_heads=4, embedding_size=D, inner_size=2, hidden_dropout_prob=0.1, attn_dropout_prob=0.1,
                   hidden_act="gelu", layer_norm_eps=1e-12, initializer_range=0.02, MAX_ITEM_LIST_LENGTH=L, seed=2020 + rank)
        model = SASRec(cfg, Dl()).to(dev).train()
        opt = FusedAdamW(model.parameters(), lr=1e-4, weight_decay=0.1, tables=[model.item_embedding])
        what = f"IDNet/sasrec Pixel8M-shape: N={N} items, emb_dim={D}, seq_len={L}, 4 heads (dh={D // 4}), 2 layers, dropout 0.1"
    elif sh["model"] == "GRU4Rec":
        from pixelrec_b200.model.IDNet.gru4rec import GRU4Rec
        cfg = dict(embedding_size=D, hidden_size=1, num_layers=1, dropout_prob=0.0, MAX_ITEM_LIST_LENGTH=L, seed=2020 + rank)
        model = GRU4Rec(cfg, Dl()).to(dev).train()
        opt = FusedAdamW(model.parameters(), lr=1e-4, weight_decay=0.01, tables=[model.item_embedding])
        what = f"IDNet/gru4rec Pixel1M-shape: N={N} items, emb_dim={D}, seq_len={L}, 1 GRU layer (cuDNN)"
    else:
        from pixelrec_b200.config import Config
        files = [os.path.join(ROOT, "configs/PixelNet/sasrec.yaml"), os.path.join(ROOT, "configs/overall/ViT.yaml")]
        c = Config(files, config_dict=dict(encoder_name=f"clip-vit-base-patch{args.patch}", seed=2020 + rank))
        model = c.model_class(c, Dl()).to(dev).train()
        modal = [q for n, q in model.named_parameters() if q.requires_grad and "visual_encoder" in n]
        rec = [q for n, q in model.named_parameters() if q.requires_grad and "visual_encoder" not in n]
        opt = FusedAdamW([dict(params=modal, lr=1e-4, weight_decay=0.0), dict(params=rec, lr=1e-4, weight_decay=0.1)])
        what = (f"PixelNet/sasrec + CLIP ViT-B/{args.patch} (random init, first 165 parameters frozen as overall/ViT.yaml) on synthetic "
                f"224x224 pixels, emb_dim={D}, seq_len={L}")
    broadcast_dense_params(model)

    POOL = 3
    batches = []
    for _ in range(POOL):
        items, mask = bench.synth_batch(g, B, N, L, perm, p)
        if sh["model"] == "MOSASRec":
            imgs = torch.randn(B, 2 * (L + 1), 3, 224, 224, device=dev)
            batches.append((imgs, torch.from_numpy(mask).to(dev)))
        else:
            batches.append((torch.from_numpy(items).to(dev), torch.from_numpy(mask).to(dev)))

    def step(i):
        opt.zero_grad()
        loss = model(batches[i % POOL])
        loss.backward()
        opt.step()
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        loss = step(i)
    del loss
    eager_step, graphed = step, None
    if world == 1 and not args.no_graph:
        # one CUDA-graph launch per step (trainer/graph.py; yaml `cuda_graph: True`): these configurations are launch-bound in eager
        # mode (C4: ~400 kernels of 20-100 us per step)
        from pixelrec_b200.trainer.graph import GraphedTrainStep
        try:
            graphed = GraphedTrainStep(model, opt, batches[0])

            def step(i):  # noqa: F811
                return graphed(batches[i % POOL])
            for i in range(3):
                step(i)
        except Exception as ex:  # noqa: BLE001
            print(f"graph capture failed, staying eager: {type(ex).__name__}: {ex}", file=sys.stderr)
            graphed, step = None, eager_step
    sync_all()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(args.steps):
        loss = step(i)
    e.record()
    sync_all()
    ms = torch.tensor([s.elapsed_time(e)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    loss = float(loss)
    if graphed is not None:
        graphed.close()
        step = eager_step
    # per-kernel pass (CUDA events around every launch of ours) -> achieved GB/s against the measured HBM peak
    from pixelrec_b200 import ops
    ops.PROFILE.update(on=True, names=None, events={})
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    prof = ops.profile_summary()
    ops.PROFILE.update(on=False, events={})
    hbm, _, peak_src = bench.peaks()
    R_u = B * (2 * L + 1)
    n_local = (N + world - 1) // world
    alg = {"gather_rows": R_u * (8 * D + 8) if world == 1 else None, "adamw_rows": 6 * n_local * D * 4,
           "bpr_fwd": None, "add_ln_fwd": 3 * B * L * D * 4, "add_ln_bwd": 5 * B * L * D * 4,
           "attn_fwd": 16 * B * L * D, "attn_bwd": 32 * B * L * D}
    kernels = {}
    for name, (n, mean_ms) in sorted(prof.items()):
        ent = {"launches_per_step": n / 3.0, "ms": mean_ms}
        if alg.get(name) and sh["model"] != "MOSASRec":
            ent["GBps"] = alg[name] / mean_ms / 1e6
            ent["frac_of_hbm_peak"] = ent["GBps"] / hbm
        kernels[name] = ent
    # CPU port of the same step on the host cores (bounded sample), rank 0 at N = 1 only
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            cpu = cpu_port(sh, N, D, L, g, perm, p)
        except Exception as ex:  # noqa: BLE001
            cpu = {"error": f"{type(ex).__name__}: {ex}"}
    xs = getattr(getattr(model, "item_embedding", None), "exchange_status", lambda: 0)()
    if rank == 0:
        print(json.dumps({
            "metric": f"sequences/sec {args.config}", "value": B * world * args.steps / (ms / 1e3), "unit": "sequences/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "dtype": "f32 (tf32 tensor-core linear layers)", "data": "synthetic",
            "config": {"workload": what, "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": ("single GPU" if world == 1 else
                                       f"dp{world}" + ("" if sh["model"] == "MOSASRec" else f" + item table row-sharded {world}-way"))},
            "roofline_kernels": kernels, "hbm_peak_GBps": hbm, "peak_source": peak_src, "cpu_baseline": cpu,
            "loss": loss, "cuda_graph": graphed is not None, "exchange_status": xs,
            "max_memory_GB": torch.cuda.max_memory_allocated(dev) / 2 ** 30}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
