"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz from the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

For each case it instantiates REC.model.IDNet.sasrec.SASRec (reference, CPU, dropout 0, fixed
seed), feeds seeded synthetic (items, masked_index), and stores inputs, every parameter, the
reference loss, the encoder output, every gradient produced by loss.backward(), the result of
one torch.optim.AdamW step (trainer.py:100-103,125), predict() scores, and the masked top-k of
trainer.py:334-336 + collector.py:133.  tests/test_oracle_golden.py pins oracle/sasrec_np.py,
and oracle/torch_port.py to these files; the GPU parity tests compare the CUDA
path with the same files.
"""
import os
import sys

import numpy as np
import torch

from oracle.refload import ref_sasrec

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: N, D, L, B, heads, layers
    "sasrec_c1_small": dict(N=401, D=128, L=10, B=8, h=4, layers=2, seed=2020),
    "sasrec_l20_d64": dict(N=257, D=64, L=20, B=6, h=4, layers=2, seed=7),
    "sasrec_dh128": dict(N=101, D=256, L=7, B=5, h=2, layers=1, seed=11),
}


def synth_batch(g, N, L, B, full=False):
    """Left-padded positives with ragged lengths, uniform negatives not in the sequence, neg[:,0]=0
    -- the layout SEQTrainDataset produces (data/dataset/trainset.py:52-75)."""
    items = np.zeros((B, 2, L + 1), dtype=np.int64)
    mask = np.zeros((B, L), dtype=np.int64)
    for b in range(B):
        n = L + 1 if (full or b == 0) else int(g.integers(2, L + 2))
        seq = g.integers(1, N, size=n)
        if b == 1 and n >= 3:
            seq[1] = seq[0]  # duplicate ids inside one sequence
        items[b, 0, L + 1 - n:] = seq
        s = set(seq.tolist())
        for t in range(1, n):
            it = int(g.integers(1, N))
            while it in s:
                it = int(g.integers(1, N))
            items[b, 1, L + 1 - n + t] = it
        mask[b, L - (n - 1):] = 1
    return items, mask


def make_case(name, c):
    torch.manual_seed(c["seed"])
    g = np.random.default_rng(c["seed"])
    cfg = dict(n_layers=c["layers"], n_heads=c["h"], embedding_size=c["D"], inner_size=2,
               hidden_dropout_prob=0.0, attn_dropout_prob=0.0, hidden_act="gelu", layer_norm_eps=1e-12,
               initializer_range=0.02, MAX_ITEM_LIST_LENGTH=c["L"])
    model = ref_sasrec(cfg, c["N"])
    # biases / LN params are zeros/ones at init; perturb so that their grads & use are exercised
    with torch.no_grad():
        for n_, p in model.named_parameters():
            if p.ndim == 1:
                p.add_(0.05 * torch.randn_like(p))
    model.train()
    items, mask = synth_batch(g, c["N"], c["L"], c["B"])
    t_items, t_mask = torch.from_numpy(items), torch.from_numpy(mask)

    cap = {}
    hk = model.trm_encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("enc_out", o[-1].detach().clone()))
    out = {"cfg_" + k: np.array(v) for k, v in c.items()}
    for k, v in model.state_dict().items():
        out["param/" + k] = v.detach().numpy().copy()
    out["items"] = items
    out["masked_index"] = mask

    # ---- eval path first (same initial parameters as the train step below) ----
    model.eval()
    Be = c["B"] + 3
    seqs = np.zeros((Be, c["L"]), dtype=np.int64)
    hist_u, hist_i = [], []
    for u in range(Be):
        n = int(g.integers(1, c["L"] + 6))
        full = g.integers(1, c["N"], size=n)
        seq = full[-c["L"]:]
        seqs[u, c["L"] - len(seq):] = seq
        hist_u += [u] * n
        hist_i += full.tolist()
    hist_u, hist_i = np.array(hist_u, dtype=np.int64), np.array(hist_i, dtype=np.int64)
    with torch.no_grad():
        scores = model.predict(torch.from_numpy(seqs), model.compute_item_all())
        raw = scores.numpy().copy()
        scores = scores.view(-1, c["N"])
        scores[:, 0] = -np.inf                                   # trainer.py:334
        scores[(torch.from_numpy(hist_u), torch.from_numpy(hist_i))] = -np.inf   # trainer.py:335-336
        tv, ti = torch.topk(scores, 10, dim=-1)                  # collector.py:133
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.1)  # overall/ID.yaml:20-23
    opt.zero_grad()
    loss = model((t_items, t_mask))
    loss.backward()
    hk.remove()
    out["loss"] = loss.detach().numpy()
    out["enc_out"] = cap["enc_out"].numpy()
    for k, p in model.named_parameters():
        out["grad/" + k] = p.grad.detach().numpy().copy()
    opt.step()
    for k in ("item_embedding.weight", "position_embedding.weight", "LayerNorm.weight",
              "trm_encoder.layer.0.multi_head_attention.query.weight",
              "trm_encoder.layer.0.feed_forward.dense_2.bias"):
        out["adamw1/" + k] = dict(model.named_parameters())[k].detach().numpy().copy()
    # second step with the SAME grads exercises m/v carry-over and bias correction at step 2
    opt.step()
    out["adamw2/item_embedding.weight"] = model.item_embedding.weight.detach().numpy().copy()

    out["eval_item_seq"] = seqs
    out["eval_hist_u"] = hist_u
    out["eval_hist_i"] = hist_i
    out["eval_scores_raw"] = raw
    out["eval_topk_val"] = tv.numpy()
    out["eval_topk_idx"] = ti.numpy()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", float(loss.detach()), "size", os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    torch.set_num_threads(1)
    torch.use_deterministic_algorithms(True)
    for n, c in CASES.items():
        if len(sys.argv) > 1 and n not in sys.argv[1:]:
            continue
        make_case(n, c)


def make_gru_case():
    """GRU4Rec golden (REC/model/IDNet/gru4rec.py): loss + table gradient from the unmodified reference module."""
    from oracle.refload import ref_gru4rec
    torch.manual_seed(5)
    g = np.random.default_rng(5)
    N, D, L, B = 151, 64, 6, 7
    model = ref_gru4rec(dict(embedding_size=D, hidden_size=1, num_layers=1, dropout_prob=0), N)
    model.train()
    items, mask = synth_batch(g, N, L, B)
    out = {"items": items, "masked_index": mask, "cfg_N": np.array(N), "cfg_D": np.array(D), "cfg_L": np.array(L)}
    for k, v in model.state_dict().items():
        out["param/" + k] = v.detach().numpy().copy()
    loss = model((torch.from_numpy(items), torch.from_numpy(mask)))
    loss.backward()
    out["loss"] = loss.detach().numpy()
    for k, p in model.named_parameters():
        out["grad/" + k] = p.grad.detach().numpy().copy()
    seqs = g.integers(1, N, size=(5, L)).astype(np.int64)
    seqs[0, :3] = 0
    with torch.no_grad():
        model.eval()
        out["eval_item_seq"] = seqs
        out["eval_scores_raw"] = model.predict(torch.from_numpy(seqs), model.compute_item_all()).numpy()
    np.savez_compressed(os.path.join(OUT, "gru4rec_small.npz"), **out)
    print("gru4rec_small loss", float(loss.detach()))


if __name__ == "__main__" and (len(sys.argv) == 1 or "gru4rec_small" in sys.argv[1:]):
    make_gru_case()


def make_vit_case(case="vit_small", image_size=96, patch_size=32, n_img=5):
    """CLIP ViT item encoder golden: HF transformers CLIPVisionModel (the class the reference instantiates,
    REC/model/load.py:94; random init -- no hub access) + the reference's own MeanItemEncoder wrapper
    (REC/model/layers.py:121-128) on a small config: output vectors and parameter gradients.
    vit_small: 10 tokens (patch 32 on 96x96); vit_long197: 197 tokens (patch 8 on 112x112) -- the sequence length of
    ViT-B/16 at 224x224 (BASELINE.json configs[3]) at a fixture-sized width."""
    from transformers import CLIPVisionConfig, CLIPVisionModel
    from oracle.refload import load_reference
    load_reference()
    from REC.model.layers import Identity, MeanItemEncoder
    torch.manual_seed(3)
    cfg = CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4,
                           image_size=image_size, patch_size=patch_size)
    m = CLIPVisionModel(cfg)
    with torch.no_grad():
        for p in m.parameters():
            if p.ndim == 1:
                p.add_(0.05 * torch.randn_like(p))
    for index, (name, p) in enumerate(m.named_parameters()):
        if index < 5 + 16:                      # freeze embeddings, pre-LN and the first layer (the reference's tune_scale rule)
            p.requires_grad = False
    m.vision_model.post_layernorm = Identity()
    enc = MeanItemEncoder(item_encoder=m, input_dim=64, output_dim=48, act_name="relu", dnn_layers=[])
    x = torch.randn(n_img, 3, image_size, image_size)
    out = enc(x)
    g = torch.randn_like(out)
    out.backward(g)
    d = {"x": x.numpy(), "out": out.detach().numpy(), "gout": g.numpy()}
    for k, v in enc.state_dict().items():
        d["param/" + k] = v.numpy().copy()
    for k, p in enc.named_parameters():
        if p.grad is not None:
            d["grad/" + k] = p.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, case + ".npz"), **d)
    print(case, out.shape, len([k for k in d if k.startswith("grad/")]), "grads")


if __name__ == "__main__" and (len(sys.argv) == 1 or "vit_small" in sys.argv[1:]):
    make_vit_case()
if __name__ == "__main__" and (len(sys.argv) == 1 or "vit_long197" in sys.argv[1:]):
    make_vit_case("vit_long197", image_size=112, patch_size=8, n_img=3)


# --------------------------------------------------------------------------------------------------------------------
# Bench-shape golden (BASELINE.json configs[1] / bench.py: D=512, L=20, h=4, 2 layers).  The parameters of that shape are
# 4.2 M encoder floats + the table -- too large for a fixture -- so they are a pure function of a seed
# (`bench_shape_params`, numpy's frozen legacy RandomState stream) that the tests re-create; the file holds inputs, the
# reference's loss / encoder output / predict scores and a FIXED SUBSET of every gradient (all 1-D tensors and the position
# table in full, the touched table rows, rows [::16] of every weight matrix).
BENCH_SHAPE = dict(N=3001, D=512, L=20, B=24, h=4, layers=2, seed=512)


def bench_shape_params(shapes, seed):
    """{name: fp32 array} for the given {name: shape}; N(0, 0.02^2) weights (sasrec.py:51-61), LayerNorm weights 1 + N(0, 0.05^2),
    biases N(0, 0.05^2).  Order-independent: every tensor has its own stream keyed by a stable hash of its name."""
    import zlib
    out = {}
    for name in sorted(shapes):
        rs = np.random.RandomState((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 32))
        x = rs.standard_normal(size=tuple(shapes[name])).astype(np.float32)
        if len(shapes[name]) == 1:
            x = (0.05 * x + (1.0 if name.endswith("LayerNorm.weight") else 0.0)).astype(np.float32)
        else:
            x = (0.02 * x).astype(np.float32)
        out[name] = x
    return out


def grad_subset(name, g, touched_rows):
    """the slice of a gradient the bench-shape golden keeps (see above)"""
    if name == "item_embedding.weight":
        return g[touched_rows]
    if g.ndim == 1 or name == "position_embedding.weight":
        return g
    return g[::16]


def make_bench_shape_case(name="sasrec_bench_shape"):
    c = BENCH_SHAPE
    torch.manual_seed(c["seed"])
    g = np.random.default_rng(c["seed"])
    cfg = dict(n_layers=c["layers"], n_heads=c["h"], embedding_size=c["D"], inner_size=2,
               hidden_dropout_prob=0.0, attn_dropout_prob=0.0, hidden_act="gelu", layer_norm_eps=1e-12,
               initializer_range=0.02, MAX_ITEM_LIST_LENGTH=c["L"])
    model = ref_sasrec(cfg, c["N"])
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    params = bench_shape_params(shapes, c["seed"])
    model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    model.train()
    items, mask = synth_batch(g, c["N"], c["L"], c["B"])
    out = {"cfg_" + k: np.array(v) for k, v in c.items()}
    for k, s in shapes.items():
        out["shape/" + k] = np.array(s, dtype=np.int64)
    out["items"], out["masked_index"] = items, mask
    cap = {}
    hk = model.trm_encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("enc_out", o[-1].detach().clone()))
    loss = model((torch.from_numpy(items), torch.from_numpy(mask)))
    loss.backward()
    hk.remove()
    out["loss"] = loss.detach().numpy()
    out["enc_out"] = cap["enc_out"].numpy()
    touched = np.unique(items)
    out["touched_rows"] = touched
    for k, p in model.named_parameters():
        out["grad/" + k] = grad_subset(k, p.grad.detach().numpy(), touched).copy()
    untouched = np.setdiff1d(np.arange(c["N"]), touched)
    assert not model.item_embedding.weight.grad[untouched].any()
    # eval: predict + mask + top-k at the same shape (trainer.py:327-337, collector.py:133)
    model.eval()
    Be = 48
    seqs = np.zeros((Be, c["L"]), dtype=np.int64)
    hist_u, hist_i = [], []
    for u in range(Be):
        n = int(g.integers(1, c["L"] + 6))
        full = g.integers(1, c["N"], size=n)
        seq = full[-c["L"]:]
        seqs[u, c["L"] - len(seq):] = seq
        hist_u += [u] * n
        hist_i += full.tolist()
    hist_u, hist_i = np.array(hist_u, dtype=np.int64), np.array(hist_i, dtype=np.int64)
    with torch.no_grad():
        scores = model.predict(torch.from_numpy(seqs), model.compute_item_all())
        out["eval_scores_raw"] = scores.numpy().copy()
        scores[:, 0] = -np.inf
        scores[(torch.from_numpy(hist_u), torch.from_numpy(hist_i))] = -np.inf
        tv, ti = torch.topk(scores, 10, dim=-1)
    out.update(eval_item_seq=seqs, eval_hist_u=hist_u, eval_hist_i=hist_i, eval_topk_val=tv.numpy(), eval_topk_idx=ti.numpy())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", float(loss.detach()), "size", os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024, "KiB")


def make_seqtrain_case(name="seqtrain_ref"):
    """A1 pin: the reference's own SEQTrainDataset (data/dataset/trainset.py:22-75) under a seeded `random` -- the stream its
    _neg_sample draws from (trainset.py:40-44) -- on ragged sequences incl. ones longer than L+1 and a small catalog (so that
    rejections happen).  tests/test_oracle_sampler.py replays oracle.seq_train_sample with random.Random(seed)."""
    import random
    from oracle.refload import load_reference
    load_reference()
    from REC.data.dataset.trainset import SEQTrainDataset
    g = np.random.default_rng(99)
    out = {}
    for ci, (item_num, L, n_seq) in enumerate([(40, 10, 64), (5000, 20, 64), (12, 5, 32)]):
        seqs = []
        for _ in range(n_seq):
            # dataload.py:125-132 caps windows at L+1; case 0 also feeds longer ones (_padding_sequence keeps the tail)
            ln = int(g.integers(2, min(L + (6 if ci == 0 else 2), item_num - 1)))
            seqs.append(g.choice(np.arange(1, item_num), size=ln, replace=False).astype(np.int64))

        class Dl:
            pass
        dl = Dl()
        dl.item_num = item_num
        dl.train_feat = {"item_seq": seqs}
        ds = SEQTrainDataset({"MAX_ITEM_LIST_LENGTH": L, "device": "cpu"}, dl)
        seed = 1234 + ci
        random.seed(seed)
        items, masks = [], []
        for i in range(len(ds)):
            it, m = ds[i]
            items.append(it.numpy())
            masks.append(m.numpy())
        out[f"c{ci}_meta"] = np.array([item_num, L, n_seq, seed], dtype=np.int64)
        out[f"c{ci}_flat"] = np.concatenate(seqs)
        out[f"c{ci}_offs"] = np.cumsum([0] + [len(s) for s in seqs]).astype(np.int64)
        out[f"c{ci}_items"] = np.stack(items)
        out[f"c{ci}_mask"] = np.stack(masks)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if k.endswith("items")})


def make_seqeval_case(name="seqeval_ref"):
    """A11 data side pin: the reference's own SeqEvalDataset.__getitem__ (data/dataset/evalset.py:24-37) + seq_eval_collate
    (data/dataset/collate_fn.py:6-32) through a DataLoader with the reference's strided sampler order (rank r: users r, r+W, ...:
    data/utils.py:134-159), for both phases, histories shorter and longer than L, ragged last batch.
    tests/test_host_plumbing.py replays pixelrec_b200's SeqEvalDataset / seq_eval_collate / DeviceSeqEvalLoader against it."""
    import torch
    from torch.utils.data import DataLoader
    from oracle.refload import load_reference
    load_reference()
    from REC.data.dataset.collate_fn import seq_eval_collate
    from REC.data.dataset.evalset import SeqEvalDataset
    g = np.random.default_rng(7)
    n_users, item_num, L, bs = 23, 60, 6, 4
    user_seq = {}
    for u in range(n_users):
        user_seq[u] = g.integers(1, item_num, size=int(g.integers(3, 15))).astype(np.int64)

    class Dl:
        pass
    dl = Dl()
    dl.item_num = item_num
    dl.user_seq = user_seq
    out = {"meta": np.array([n_users, item_num, L, bs], dtype=np.int64),
           "flat": np.concatenate([user_seq[u] for u in range(n_users)]),
           "offs": np.cumsum([0] + [len(user_seq[u]) for u in range(n_users)]).astype(np.int64)}
    for phase in ("valid", "test"):
        ds = SeqEvalDataset({"MAX_ITEM_LIST_LENGTH": L}, dl, phase=phase)
        for world in (1, 2, 3):
            for rank in range(world):
                order = list(range(rank, len(ds), world))
                loader = DataLoader(ds, batch_size=bs, sampler=order, collate_fn=seq_eval_collate)
                for bi, (item_seq, (hu, hi), pu, tgt) in enumerate(loader):
                    key = f"{phase}_w{world}_r{rank}_b{bi}"
                    out[key + "_seq"] = item_seq.numpy()
                    out[key + "_hu"] = hu.numpy()
                    out[key + "_hi"] = hi.numpy()
                    out[key + "_pu"] = pu.numpy()
                    out[key + "_tgt"] = tgt.numpy()
                out[f"{phase}_w{world}_r{rank}_nb"] = np.array([bi + 1], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, len(out), "arrays", os.path.getsize(os.path.join(OUT, name + ".npz")) // 1024, "KiB")


def make_evalmetrics_case(name="evalmetrics_ref"):
    """A11 metric side pin: the reference's own Collector.eval_batch_collect (evaluator/collector.py:113-139) + Evaluator /
    Recall / NDCG (evaluator/metrics.py:115-178, base_metric.py:43-67) on three batches of masked random scores (per-rank SUMS over
    users, as trainer.py:402-408 divides them afterwards).  tests/test_host_plumbing.py replays pixelrec_b200's evaluator on the
    same scores, through both of its entry points (scores, and the fused kernel's top-k ids)."""
    import torch
    from oracle.refload import load_reference
    load_reference()
    from REC.evaluator import Collector, Evaluator

    class Cfg(dict):
        def __getitem__(self, k):
            return self.get(k)
    cfg = Cfg(topk=[5, 10], metrics=["Recall", "NDCG"], device="cpu", metric_decimal_place=7)
    col, ev = Collector(cfg), Evaluator(cfg)
    g = torch.Generator().manual_seed(3)
    out = {"topk": np.array([5, 10], dtype=np.int64)}
    for bi, n in enumerate((7, 7, 3)):
        scores = torch.randn(n, 60, generator=g)
        scores[:, 0] = -float("inf")
        hu = torch.randint(0, n, (4 * n,), generator=g)
        hi = torch.randint(1, 60, (4 * n,), generator=g)
        pi = torch.randint(1, 60, (n,), generator=g)
        scores[hu, hi] = -float("inf")
        scores[torch.arange(n), pi] = torch.randn(n, generator=g) + 1.5      # positives are never masked (they are held out)
        col.eval_batch_collect(scores, torch.arange(n), pi)
        out[f"b{bi}_scores"] = scores.numpy()
        out[f"b{bi}_pi"] = pi.numpy()
    res = ev.evaluate(col.get_data_struct())
    out["names"] = np.array(list(res.keys()))
    out["values"] = np.array([float(v) for v in res.values()], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, dict(res))


def make_dataload_case(name="dataload_ref"):
    """A1 upstream pin: the reference's own Data (data/dataload.py:16-150) on a toy <dataset>.csv whose rows are NOT in time
    order: id re-mapping with [PAD] = 0 (:42-57), user_seq in order of first appearance after the timestamp sort (:66-76), training
    windows of L+1 items with the oldest n mod (L+1) dropped (:103-150).  Timestamps are distinct (the reference's DataFrame sort
    is unstable for ties).  tests/test_host_plumbing.py rebuilds the CSV from the stored rows and replays pixelrec_b200's Data."""
    import tempfile
    from oracle.refload import load_reference
    load_reference()
    from REC.data.dataload import Data
    from REC.utils.enum_type import InputType
    g = np.random.default_rng(5)
    n, L = 400, 5
    ts = g.permutation(n) + 1000
    item_tok = g.integers(0, 40, size=n)
    user_tok = g.integers(0, 14, size=n)
    user_tok[:3] = 99                    # a user with 3 interactions (one training item) ...
    user_tok[3:5] = 98                   # ... and one with 2 (none)
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "toy.csv"), "w") as f:
        f.write("item_id,user_id,timestamp\n" + "\n".join(f"i{a},u{b},{c}" for a, b, c in zip(item_tok, user_tok, ts)))

    class Cfg(dict):
        def __getitem__(self, k):
            return self.get(k)
    data = Data(Cfg(data_path=d, dataset="toy", MAX_ITEM_LIST_LENGTH=L, MODEL_INPUT_TYPE=InputType.SEQ))
    data.build()
    keys = list(data.user_seq.keys())
    out = {"rows": np.stack([item_tok, user_tok, ts], 1).astype(np.int64), "L": np.array([L]),
           "item_num": np.array([data.item_num]), "user_num": np.array([data.user_num]),
           "user_order": np.array(keys, dtype=np.int64),
           "user_seq_flat": np.concatenate([data.user_seq[k] for k in keys]).astype(np.int64),
           "user_seq_offs": np.cumsum([0] + [len(data.user_seq[k]) for k in keys]).astype(np.int64),
           "train_uid": np.asarray(data.train_feat["user_id"], dtype=np.int64),
           "train_flat": np.concatenate(data.train_feat["item_seq"]).astype(np.int64),
           "train_offs": np.cumsum([0] + [len(s) for s in data.train_feat["item_seq"]]).astype(np.int64),
           "item_tokens": np.array(list(data.id2token["item_id"])), "user_tokens": np.array(list(data.id2token["user_id"]))}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "users", len(keys), "windows", len(data.train_feat["item_seq"]))


def make_init_case(name="init_ref"):
    """Plugin-construction pin: the state_dict the reference's own SASRec / GRU4Rec hold right after `init_seed(2020, True)` +
    construction (utils/utils.py init_seed; sasrec.py:49-61 `self.apply(self._init_weights)`; gru4rec.py:38-48) on a small config.
    The plugins here must draw the SAME initial weights from the same seed (same module order, same initialisers), so that a run of
    the reference's yaml + seed starts from the reference's model.  tests/test_host_plumbing.py replays it on the CPU."""
    import torch
    from oracle.refload import load_reference
    load_reference()
    from REC.model.IDNet.gru4rec import GRU4Rec
    from REC.model.IDNet.sasrec import SASRec
    from REC.utils import init_seed

    class Dl:
        item_num = 101
        user_num = 20
    out = {}
    init_seed(2020, True)
    m = SASRec(dict(SASREC_INIT_CFG), Dl())
    for k, v in m.state_dict().items():
        out["sasrec/" + k] = v.numpy()
    init_seed(2020, True)
    m = GRU4Rec(dict(GRU4REC_INIT_CFG), Dl())
    for k, v in m.state_dict().items():
        out["gru4rec/" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, len(out), "tensors")


SASREC_INIT_CFG = dict(n_layers=2, n_heads=2, embedding_size=32, inner_size=2, hidden_dropout_prob=0.1, attn_dropout_prob=0.1,
                       hidden_act="gelu", layer_norm_eps=1e-12, initializer_range=0.02, MAX_ITEM_LIST_LENGTH=10, seed=2020, device="cpu")
GRU4REC_INIT_CFG = dict(embedding_size=32, hidden_size=2, num_layers=1, dropout_prob=0.0, MAX_ITEM_LIST_LENGTH=10, seed=2020,
                        device="cpu")


def make_utils_case(name="utils_ref"):
    """Trainer-control pin: the reference's own early_stopping / calculate_valid_score / dict2str (utils/utils.py:65-135) on a grid
    of inputs -- they decide when the reference saves a checkpoint and when it stops (trainer.py:234-262)."""
    import json
    from collections import OrderedDict
    from oracle.refload import load_reference
    load_reference()
    from REC.utils import calculate_valid_score, dict2str, early_stopping
    cases = []
    for bigger in (True, False):
        for value, best in ((0.5, 0.4), (0.3, 0.4), (0.4, 0.4), (0.41, 0.4)):
            for cur_step, max_step in ((0, 5), (3, 5), (4, 5), (5, 5), (29, 30)):
                out = early_stopping(value, best, cur_step, max_step=max_step, bigger=bigger)
                cases.append({"in": [value, best, cur_step, max_step, bigger], "out": [float(out[0]), int(out[1]), bool(out[2]), bool(out[3])]})
    res = OrderedDict([("recall@5", 0.1234567), ("recall@10", 0.2), ("ndcg@5", 0.05), ("ndcg@10", 0.0712345)])
    out = {"early_stopping": cases,
           # the trainer passes config['valid_metric'].lower() (trainer.py:55), i.e. the lower-case keys of the result dict
           "valid_score": [[m, float(calculate_valid_score(res, m))] for m in ("ndcg@10", "recall@5", "ndcg@5")],
           "dict2str_in": list(res.items()), "dict2str_out": dict2str(res)}
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(out, f, indent=1)
    print(name, len(cases), "early_stopping cases")


def make_config_case(name="config_ref"):
    """Boundary pin: the final config dict the reference's own Config builds (config/configurator.py) from ITS yaml
    files for the three hot-path model plugins -- model yaml + overall yaml, MODEL_INPUT_TYPE / eval_type / valid_metric_bigger
    derived keys included.  tests/test_host_plumbing.py loads this repo's copies of those yaml files with pixelrec_b200.config.Config
    and compares key by key.  Stored as JSON (enums as their str())."""
    import json
    from oracle.refload import load_reference
    load_reference()
    from REC.config import Config
    here = os.getcwd()
    os.chdir("/tmp")                     # the reference's logger / config may touch relative paths
    out = {}
    try:
        for tag, files in (("sasrec", ["IDNet/sasrec.yaml", "overall/ID.yaml"]), ("gru4rec", ["IDNet/gru4rec.yaml", "overall/ID.yaml"]),
                           ("mosasrec", ["PixelNet/sasrec.yaml", "overall/ViT.yaml"])):
            c = Config(config_file_list=["/root/reference/code/" + f for f in files])
            d = {k: (v if isinstance(v, (int, float, str, list, dict, bool, type(None))) else str(v))
                 for k, v in dict(c.final_config_dict).items()}
            out[tag] = {"files": files, "final": d}
    finally:
        os.chdir(here)
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(name, {k: len(v["final"]) for k, v in out.items()})


if __name__ == "__main__" and (len(sys.argv) == 1 or "sasrec_bench_shape" in sys.argv[1:]):
    make_bench_shape_case()
if __name__ == "__main__" and (len(sys.argv) == 1 or "seqtrain_ref" in sys.argv[1:]):
    make_seqtrain_case()
if __name__ == "__main__" and (len(sys.argv) == 1 or "seqeval_ref" in sys.argv[1:]):
    make_seqeval_case()
if __name__ == "__main__" and (len(sys.argv) == 1 or "evalmetrics_ref" in sys.argv[1:]):
    make_evalmetrics_case()
if __name__ == "__main__" and (len(sys.argv) == 1 or "config_ref" in sys.argv[1:]):
    make_config_case()
if __name__ == "__main__" and (len(sys.argv) == 1 or "dataload_ref" in sys.argv[1:]):
    make_dataload_case()
if __name__ == "__main__" and (len(sys.argv) == 1 or "init_ref" in sys.argv[1:]):
    make_init_case()
if __name__ == "__main__" and (len(sys.argv) == 1 or "utils_ref" in sys.argv[1:]):
    make_utils_case()
