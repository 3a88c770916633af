"""CPU tests of the host side: C-ABI library loads and exports what include/pixelrec_b200.h declares, the
reference's yaml files load unchanged, plugin lookup / dataset binding / data pipeline / evaluator / early
stopping behave like the reference, and the product path refuses to run without CUDA (no CPU fallback)."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
YAML_DIR = os.path.join(ROOT, "configs")


def test_library_loads_and_exports_every_declared_symbol():
    from pixelrec_b200 import lib
    L = lib.load()
    declared = lib.header_symbols()
    assert len(declared) >= 20 and set(declared) == set(lib.SIGNATURES)
    for s in declared:
        assert hasattr(L, s), s
    assert L.pr_version() == 1
    # argument counts of the ctypes table match the header prototypes
    hdr = open(lib.HEADER_PATH).read()
    for name, (_, args) in lib.SIGNATURES.items():
        m = re.search(r"PR_API[^;]*?\b" + name + r"\s*\(([^;]*?)\)\s*;", hdr, re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), (name, len(params), len(args))
    # invalid arguments are rejected on the host before any launch (no GPU needed)
    assert L.pr_gather_rows_f32(None, 0, 4, None, 0, None, None, 0, None) == -1
    assert b"bad shape" in L.pr_last_error_string()
    assert L.pr_scatter_plan_workspace_bytes(1000, 50) > 0


def test_no_cpu_fallback():
    from pixelrec_b200 import ops
    from pixelrec_b200.lib import PixelRecB200Error
    with pytest.raises(PixelRecB200Error):
        ops.gather_rows(torch.zeros(4, 8), torch.zeros(2, dtype=torch.long))
    with pytest.raises(PixelRecB200Error):
        ops.bpr_loss(torch.zeros(1, 2, 8), torch.zeros(1, 2, 3, 8), torch.ones(1, 2, dtype=torch.long))


def _yaml(*names):
    return [os.path.join(YAML_DIR, n) for n in names]


def test_reference_yaml_files_load_unchanged():
    from pixelrec_b200.config import Config
    from pixelrec_b200.utils import InputType
    c = Config(_yaml("IDNet/sasrec.yaml", "overall/ID.yaml"))
    assert c["model"] == "SASRec" and c["embedding_size"] == 512 and c["n_heads"] == 4
    assert isinstance(c["layer_norm_eps"], float) and c["layer_norm_eps"] == 1e-12     # custom float resolver
    assert c["does_not_exist"] is None and "does_not_exist" not in c                    # configurator.py:148-152
    assert c["MODEL_INPUT_TYPE"] == InputType.SEQ and c["valid_metric_bigger"] is True
    assert c["optim_args"] == {"learning_rate": 0.0001, "weight_decay": 0.1} and c["topk"] == [5, 10]
    assert c.model_class.__name__ == "SASRec" and c["MAX_ITEM_LIST_LENGTH"] == 10
    c["device"] = "cuda:0"
    assert c.device == "cuda:0"
    with pytest.raises(ValueError):
        Config(config_dict=dict(model="SASRec", metrics=["Recall"], valid_metric="Recall@10", topk=[0]))
    with pytest.raises(NotImplementedError):
        Config(config_dict=dict(model="SASRec", metrics=["Nope"], valid_metric="Recall@10", topk=[5]))
    with pytest.raises(ValueError):
        Config(config_dict=dict(model="NotAModel", metrics=["Recall"], valid_metric="Recall@10", topk=[5]))


def test_plugin_constructs_with_config_and_dict():
    from pixelrec_b200.config import Config
    from pixelrec_b200.model.IDNet.sasrec import SASRec

    class Dl:
        item_num = 50
    c = Config(_yaml("IDNet/sasrec.yaml", "overall/ID.yaml"), config_dict=dict(embedding_size=64))
    m = SASRec(c, Dl())
    keys = set(m.state_dict().keys())
    assert {"item_embedding.weight", "position_embedding.weight", "LayerNorm.weight",
            "trm_encoder.layer.1.multi_head_attention.query.weight", "trm_encoder.layer.0.feed_forward.dense_2.bias"} <= keys
    assert m.item_embedding.weight.shape == (50, 64) and m.item_embedding.weight[0].abs().sum() > 0   # pad row re-initialised
    d = dict(n_layers=1, n_heads=2, embedding_size=32, inner_size=2, hidden_dropout_prob=0.1, attn_dropout_prob=0.1,
             hidden_act="gelu", layer_norm_eps=1e-12, initializer_range=0.02, MAX_ITEM_LIST_LENGTH=5)
    m2 = SASRec(d, Dl())
    assert "Trainable parameters" in str(m2)
    with pytest.raises(ValueError):
        SASRec(dict(d, n_heads=3), Dl())                   # layers.py:558-562


def test_data_pipeline_matches_reference_semantics(tmp_path):
    from pixelrec_b200.config import Config
    from pixelrec_b200.data import bulid_dataloader, load_data
    from pixelrec_b200.data.dataload import split_windows
    rows = ["item_id,user_id,timestamp"]
    g = np.random.default_rng(0)
    t = 0
    for u in range(12):
        for _ in range(int(g.integers(4, 30))):
            t += 1
            rows.append(f"i{int(g.integers(0, 40))},u{u},{t}")
    (tmp_path / "toy.csv").write_text("\n".join(rows))
    c = Config(_yaml("IDNet/sasrec.yaml", "overall/ID.yaml"),
               config_dict=dict(data_path=str(tmp_path), dataset="toy", MAX_ITEM_LIST_LENGTH=5, train_batch_size=4,
                                eval_batch_size=8, num_workers=0))
    data = load_data(c)
    assert data.item_num == len(set(r.split(",")[0] for r in rows[1:])) + 1          # [PAD] = 0
    tr, va, te = bulid_dataloader(c, data)
    for uid, seq in data.user_seq.items():
        assert (seq > 0).all()
    # windows: a history longer than L+1 drops its oldest n % (L+1) items and is cut into L+1 chunks
    assert [len(w) for w in split_windows(list(range(14)), 6)] == [6, 6] and split_windows(list(range(14)), 6)[0][0] == 2
    assert [len(w) for w in split_windows(list(range(5)), 6)] == [5]
    items, mask = next(iter(tr))
    assert items.shape == (4, 2, 6) and mask.shape == (4, 5) and items.dtype == torch.int64
    for b in range(4):
        pos, neg, m = items[b, 0], items[b, 1], mask[b]
        n = int((pos != 0).sum())
        assert (pos[:6 - n] == 0).all() and m.sum() == n - 1 and (neg[:6 - n + 1] == 0).all()
        assert all(int(x) not in set(pos.tolist()) for x in neg[6 - n + 1:])
    item_seq, (hu, hi), pu, pi = next(iter(va))
    assert item_seq.shape[1] == 5 and len(hu) == len(hi) and pu.tolist() == list(range(item_seq.shape[0]))
    first = list(data.user_seq.values())[0]
    assert pi[0].item() == first[-2] and item_seq[0].tolist()[-1] == first[-3]
    # vectorised batch builder has the same layout
    ds = tr.dataset
    it2, m2 = ds.sample_batch(np.arange(len(ds)), np.random.default_rng(1))
    assert it2.shape[1:] == (2, 6) and ((it2[:, 1] != 0).sum(1) == m2.sum(1)).all()
    assert ((it2[:, 0] != 0).sum(1) - 1 == m2.sum(1)).all()


def test_evaluator_and_early_stopping():
    from pixelrec_b200.evaluator import Collector, Evaluator
    from pixelrec_b200.utils import early_stopping
    from oracle import sasrec_np as O
    cfg = {"topk": [2, 5], "metrics": ["Recall", "NDCG"], "device": "cpu"}
    col, ev = Collector(cfg), Evaluator(cfg)
    g = np.random.default_rng(3)
    scores = torch.from_numpy(g.standard_normal((6, 30)).astype(np.float32))
    pu, pi = torch.arange(6), torch.from_numpy(g.integers(1, 30, size=6))
    col.eval_batch_collect(scores, pu, pi)
    res = ev.evaluate(col.get_data_struct())
    idx = np.argsort(-scores.numpy(), axis=1, kind="stable")[:, :5]
    pos, plen = O.topk_hits(idx, pu.numpy(), pi.numpy(), 6)
    ref = O.recall_ndcg(pos, plen, [2, 5])
    for k in ref:
        assert abs(res[k] - ref[k]) < 1e-12, k
    assert early_stopping(0.5, 0.4, 3, 5) == (0.5, 0, False, True)
    assert early_stopping(0.3, 0.4, 5, 5) == (0.4, 6, True, False)
    assert early_stopping(0.3, 0.4, 0, 5, bigger=False) == (0.3, 0, False, True)


def test_vit_encoder_state_dict_matches_hf_reference_layout():
    """Our CLIP ViT item encoder has exactly the parameter names, shapes and ORDER of HF CLIPVisionModel wrapped in
    the reference's MeanItemEncoder (golden generated from those classes), so checkpoints interchange and the
    reference's index-based freezing (tune_scale) hits the same tensors."""
    from pixelrec_b200.model.vit import CLIPVisionConfig, CLIPVisionModel, Identity, MeanItemEncoder
    z = np.load(os.path.join(ROOT, "tests", "golden", "vit_small.npz"))
    m = CLIPVisionModel(CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4,
                                         image_size=96, patch_size=32))
    m.vision_model.post_layernorm = Identity()
    enc = MeanItemEncoder(m, 64, 48, "relu")
    ref_keys = [k[len("param/"):] for k in z.files if k.startswith("param/")]
    ours = list(enc.state_dict().keys())
    assert ours == ref_keys
    for k, v in enc.state_dict().items():
        assert tuple(v.shape) == z["param/" + k].shape, k
    enc.load_state_dict({k: torch.from_numpy(z["param/" + k]) for k in ref_keys}, strict=True)
    names = [n for n, _ in m.named_parameters()]
    assert names[:5] == ["vision_model.embeddings.class_embedding", "vision_model.embeddings.patch_embedding.weight",
                         "vision_model.embeddings.position_embedding.weight", "vision_model.pre_layrnorm.weight",
                         "vision_model.pre_layrnorm.bias"]
    assert names[5].endswith("layers.0.self_attn.k_proj.weight") and names[5 + 16].endswith("layers.1.self_attn.k_proj.weight")


def test_mosasrec_plugin_constructs_from_reference_yaml_keys():
    from pixelrec_b200.config import Config
    c = Config(_yaml("PixelNet/sasrec.yaml", "overall/ViT.yaml"),
               config_dict=dict(embedding_size=64, vit_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=12,
                                                                   num_attention_heads=4, image_size=64)))

    class Dl:
        item_num = 30
    m = c.model_class(c, Dl())
    assert c.model_class.__name__ == "MOSASRec" and len(c["optim_args"]) == 4
    frozen = [n for n, p in m.named_parameters() if not p.requires_grad]
    assert len(frozen) == 165 and all("visual_encoder" in n for n in frozen)            # ViT.yaml tune_scale: 165
    trainable_vit = [n for n, p in m.named_parameters() if p.requires_grad and "visual_encoder" in n]
    assert any("layers.10." in n for n in trainable_vit) and not any("layers.9." in n for n in trainable_vit)
    assert "visual_encoder.rec_fc.0.weight" in dict(m.named_parameters())


def test_tcgen05_ptx_forms_match_the_vendored_cutlass_headers():
    """The PTX strings of csrc/score.cu that only a GPU can exercise (tcgen05 / TMA multicast / cluster commit) are compared
    with the forms NVIDIA's CUTLASS headers use (vendored with flashinfer in this image; skipped where they are absent), and the
    instruction-descriptor field values with CUTLASS's enums."""
    import importlib.util
    spec = importlib.util.find_spec("flashinfer")              # located, not imported
    if spec is None or not spec.submodule_search_locations:
        pytest.skip("flashinfer (vendored CUTLASS headers) not installed")
    inc = os.path.join(list(spec.submodule_search_locations)[0], "data", "cutlass", "include")
    if not os.path.isdir(inc):
        pytest.skip("vendored CUTLASS headers not found")
    def text(*rel):
        return re.sub(r'"\s*\n\s*"', "", open(os.path.join(inc, *rel)).read())          # join adjacent string literals
    umma = text("cute", "arch", "mma_sm100_umma.hpp")
    bar = text("cutlass", "arch", "barrier.h")
    tma90 = text("cute", "arch", "copy_sm90_tma.hpp")
    desc = text("cute", "arch", "mma_sm100_desc.hpp")
    ours = "".join(re.sub(r'"\s*\n\s*"', "", open(os.path.join(ROOT, "pixelrec_b200", "csrc", f)).read())
                   for f in ("tc_ptx.cuh", "score.cu", "gemm.cu"))
    tma100 = text("cute", "arch", "copy_sm100_tma.hpp")
    alloc = text("cute", "arch", "tmem_allocator_sm100.hpp")
    # CTA-pair forms of csrc/gemm.cu
    for form, ref in [("tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5,", umma),
                      ("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;", bar),
                      ("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes", tma100),
                      ("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;", alloc),
                      ("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;", alloc),
                      ("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group", tma90)]:
        assert form in ref, form
        assert form.split(" [")[0] in ours, form
    assert "SWIZZLE_128B_BASE32B = 1" in desc and "SWIZZLE_128B = 2" in desc          # gm_desc layout types
    assert "a_major_       : 1;  // bit [15,16)" in desc and "b_major_       : 1,  // bit [16,17)" in desc
    for form, ref in [("tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3,", umma),
                      ("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3,", umma),
                      ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;", bar),
                      ("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster", tma90)]:
        assert form in ref, form
        assert form.split(" [")[0] in ours, form
    # descriptor fields used by tf32_idesc / f16_idesc (csrc/score_kernels.cuh)
    assert re.search(r"F16\s*=\s*0,\s*BF16\s*=\s*1,\s*TF32\s*=\s*2", desc)
    assert "c_format_      : 2,  // bit [ 4, 6)" in desc and "a_format_      : 3,  // bit [ 7,10)" in desc
    assert "n_dim_         : 6,  // bit [17,23)" in desc and "m_dim_         : 5,  // bit [24,29)" in desc


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("phase", ["valid", "test"])
def test_device_eval_loader_equals_dataloader_collate(world, phase):
    """DeviceSeqEvalLoader (resident CSR, a batch = four slices) yields element for element what the reference's
    SeqEvalDataset.__getitem__ + seq_eval_collate build per batch (evalset.py:24-37, collate_fn.py:6-32), for every rank's
    strided share of the users (data/utils.py:134-159) -- histories shorter and longer than L, ragged last batch."""
    from torch.utils.data import DataLoader
    from pixelrec_b200.data.dataset import SeqEvalDataset, seq_eval_collate
    from pixelrec_b200.data.utils import DeviceSeqEvalLoader, NonConsecutiveSequentialDistributedSampler

    class Dl:
        item_num = 50
        user_seq = {}
    g = np.random.default_rng(4)
    for u in range(23):
        Dl.user_seq[f"u{u}"] = g.integers(1, 50, size=int(g.integers(3, 14))).astype(np.int64)
    ds = SeqEvalDataset(dict(MAX_ITEM_LIST_LENGTH=6), Dl, phase=phase)
    for rank in range(world):
        ref = DataLoader(ds, batch_size=4, sampler=NonConsecutiveSequentialDistributedSampler(ds, rank=rank, num_replicas=world),
                         collate_fn=seq_eval_collate)
        got = DeviceSeqEvalLoader(ds, 4, "cpu", rank, world)
        assert len(got) == len(ref) and len(got.sampler.dataset) == len(ref.sampler.dataset)
        n = 0
        for (s0, (hu0, hi0), pu0, t0), (s1, (hu1, hi1), pu1, t1) in zip(ref, got):
            assert torch.equal(s0, s1) and torch.equal(hu0, hu1) and torch.equal(hi0, hi1)
            assert torch.equal(pu0, pu1) and torch.equal(t0, t1)
            n += 1
        assert n == len(ref)


@pytest.mark.parametrize("phase", ["valid", "test"])
def test_eval_batches_equal_the_reference_dataset_and_collate_golden(phase):
    """tests/golden/seqeval_ref.npz holds what the UNMODIFIED reference builds per eval batch -- SeqEvalDataset.__getitem__
    (evalset.py:24-37) + seq_eval_collate (collate_fn.py:6-32) in the strided sampler's order (data/utils.py:134-159), generated by
    oracle/make_golden.py.  Both of our routes must reproduce it element for element: the DataLoader + collate port and the
    resident-CSR DeviceSeqEvalLoader (device_sampler: True)."""
    from torch.utils.data import DataLoader
    from pixelrec_b200.data.dataset import SeqEvalDataset, seq_eval_collate
    from pixelrec_b200.data.utils import DeviceSeqEvalLoader, NonConsecutiveSequentialDistributedSampler
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "seqeval_ref.npz"))
    n_users, item_num, L, bs = (int(x) for x in gold["meta"])
    flat, offs = gold["flat"], gold["offs"]

    class Dl:
        pass
    Dl.item_num = item_num
    Dl.user_seq = {u: flat[offs[u]:offs[u + 1]] for u in range(n_users)}
    ds = SeqEvalDataset(dict(MAX_ITEM_LIST_LENGTH=L), Dl, phase=phase)
    for world in (1, 2, 3):
        for rank in range(world):
            nb = int(gold[f"{phase}_w{world}_r{rank}_nb"][0])
            ours_dl = DataLoader(ds, batch_size=bs, collate_fn=seq_eval_collate,
                                 sampler=NonConsecutiveSequentialDistributedSampler(ds, rank=rank, num_replicas=world))
            ours_dev = DeviceSeqEvalLoader(ds, bs, "cpu", rank, world)
            assert len(ours_dl) == nb and len(ours_dev) == nb
            for route in (ours_dl, ours_dev):
                for bi, (seq, (hu, hi), pu, tgt) in enumerate(route):
                    key = f"{phase}_w{world}_r{rank}_b{bi}"
                    assert np.array_equal(seq.numpy(), gold[key + "_seq"]), key
                    assert np.array_equal(hu.numpy(), gold[key + "_hu"]) and np.array_equal(hi.numpy(), gold[key + "_hi"]), key
                    assert np.array_equal(pu.numpy(), gold[key + "_pu"]) and np.array_equal(tgt.numpy(), gold[key + "_tgt"]), key
                assert bi == nb - 1


def test_evaluator_equals_the_reference_evaluator_golden():
    """tests/golden/evalmetrics_ref.npz: Recall / NDCG sums produced by the UNMODIFIED reference Collector + Evaluator
    (evaluator/collector.py:113-139, metrics.py:115-178) on three batches of masked scores.  Our evaluator must give the same
    numbers through both entry points: full score rows (the reference contract) and top-k ids (what the fused scoring returns)."""
    from pixelrec_b200.evaluator import Collector, Evaluator
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "evalmetrics_ref.npz"))
    cfg = {"topk": [int(k) for k in gold["topk"]], "metrics": ["Recall", "NDCG"], "device": "cpu"}
    want = dict(zip([str(n) for n in gold["names"]], gold["values"]))
    for route in ("scores", "topk"):
        col, ev = Collector(cfg), Evaluator(cfg)
        for bi in range(3):
            scores = torch.from_numpy(gold[f"b{bi}_scores"])
            pi = torch.from_numpy(gold[f"b{bi}_pi"])
            pu = torch.arange(scores.shape[0])
            if route == "scores":
                col.eval_batch_collect(scores, pu, pi)
            else:
                col.eval_batch_collect_topk(torch.topk(scores, max(cfg["topk"]), dim=-1)[1], pu, pi)
        res = ev.evaluate(col.get_data_struct())
        assert set(res) == set(want)
        for k, v in want.items():
            assert abs(float(res[k]) - v) < 1e-9, (route, k, float(res[k]), v)


@pytest.mark.parametrize("tag", ["sasrec", "gru4rec", "mosasrec"])
def test_config_equals_the_reference_config_golden(tag):
    """tests/golden/config_ref.json: the final config dict the UNMODIFIED reference Config (config/configurator.py) builds from its
    own yaml files, derived keys (MODEL_INPUT_TYPE, eval_type, valid_metric_bigger) included.  This repo's copies of those yaml
    files through pixelrec_b200.config.Config must give the same value for every key the reference defines."""
    import json
    from pixelrec_b200.config import Config
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config_ref.json")) as f:
        gold = json.load(f)[tag]
    c = Config(_yaml(*gold["files"]))
    assert len(gold["final"]) >= 29
    for k, want in gold["final"].items():
        got = c[k]
        if not isinstance(got, (int, float, str, list, dict, bool, type(None))):
            got = str(got)
        assert got == want, (k, got, want)


def test_dataload_equals_the_reference_data_golden(tmp_path):
    """tests/golden/dataload_ref.npz: what the UNMODIFIED reference Data (data/dataload.py:16-150) builds from a toy CSV whose rows
    are not in time order -- token re-mapping, user_seq INCLUDING its order (first appearance after the timestamp sort: the order
    of the eval users and of the training windows that a seeded sampler permutes), training windows.  Our Data must match it."""
    from pixelrec_b200.data.dataload import Data
    from pixelrec_b200.utils.enum_type import InputType
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataload_ref.npz"))
    (tmp_path / "toy.csv").write_text("item_id,user_id,timestamp\n" + "\n".join(f"i{a},u{b},{c}" for a, b, c in gold["rows"]))

    class Cfg(dict):
        def __getitem__(self, k):
            return self.get(k)
    data = Data(Cfg(data_path=str(tmp_path), dataset="toy", MAX_ITEM_LIST_LENGTH=int(gold["L"][0]), MODEL_INPUT_TYPE=InputType.SEQ))
    data.build()
    assert data.item_num == int(gold["item_num"][0]) and data.user_num == int(gold["user_num"][0])
    assert list(data.id2token["item_id"]) == list(gold["item_tokens"]) and list(data.id2token["user_id"]) == list(gold["user_tokens"])
    keys = list(data.user_seq.keys())
    assert keys == [int(k) for k in gold["user_order"]]
    offs = gold["user_seq_offs"]
    for j, k in enumerate(keys):
        assert np.array_equal(data.user_seq[k], gold["user_seq_flat"][offs[j]:offs[j + 1]]), k
    assert np.array_equal(data.train_feat["user_id"], gold["train_uid"])
    toffs = gold["train_offs"]
    assert len(data.train_feat["item_seq"]) == len(toffs) - 1
    for j, w in enumerate(data.train_feat["item_seq"]):
        assert np.array_equal(w, gold["train_flat"][toffs[j]:toffs[j + 1]]), j


def test_plugins_draw_the_reference_initial_weights_from_the_same_seed():
    """tests/golden/init_ref.npz: state_dict of the UNMODIFIED reference SASRec / GRU4Rec right after init_seed(2020, True) +
    construction.  Our plugins (same module order, same initialisers: sasrec.py:49-61, gru4rec.py:38-48) must hold identical tensors,
    so the reference's yaml + seed starts from the reference's model."""
    # the configs oracle/make_golden.py (SASREC_INIT_CFG / GRU4REC_INIT_CFG) built the reference models with
    SASREC_INIT_CFG = dict(n_layers=2, n_heads=2, embedding_size=32, inner_size=2, hidden_dropout_prob=0.1, attn_dropout_prob=0.1,
                           hidden_act="gelu", layer_norm_eps=1e-12, initializer_range=0.02, MAX_ITEM_LIST_LENGTH=10, seed=2020,
                           device="cpu")
    GRU4REC_INIT_CFG = dict(embedding_size=32, hidden_size=2, num_layers=1, dropout_prob=0.0, MAX_ITEM_LIST_LENGTH=10, seed=2020,
                            device="cpu")
    from pixelrec_b200.model.IDNet.gru4rec import GRU4Rec
    from pixelrec_b200.model.IDNet.sasrec import SASRec
    from pixelrec_b200.utils import init_seed
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "init_ref.npz"))

    class Dl:
        item_num = 101
        user_num = 20
    for tag, cls, cfg in (("sasrec", SASRec, SASREC_INIT_CFG), ("gru4rec", GRU4Rec, GRU4REC_INIT_CFG)):
        init_seed(2020, True)
        sd = cls(dict(cfg), Dl()).state_dict()
        want = {k.split("/", 1)[1]: gold[k] for k in gold.files if k.startswith(tag + "/")}
        assert list(sd.keys()) == list(want.keys()), tag
        for k, v in sd.items():
            assert np.array_equal(v.numpy(), want[k]), (tag, k)


def test_trainer_control_helpers_equal_the_reference_golden():
    """tests/golden/utils_ref.json: outputs of the UNMODIFIED reference early_stopping / calculate_valid_score / dict2str
    (utils/utils.py:65-135) on a grid -- they decide when the reference checkpoints and stops (trainer.py:234-262)."""
    import json
    from collections import OrderedDict
    from pixelrec_b200.utils import calculate_valid_score, dict2str, early_stopping
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "utils_ref.json")) as f:
        gold = json.load(f)
    assert len(gold["early_stopping"]) == 40
    for c in gold["early_stopping"]:
        value, best, cur_step, max_step, bigger = c["in"]
        got = early_stopping(value, best, cur_step, max_step=max_step, bigger=bigger)
        assert [float(got[0]), int(got[1]), bool(got[2]), bool(got[3])] == c["out"], c
    res = OrderedDict((k, v) for k, v in gold["dict2str_in"])
    for m, want in gold["valid_score"]:
        assert calculate_valid_score(res, m) == want
    assert dict2str(res) == gold["dict2str_out"]
