"""Trainer -- the caller of the hot path.  Same public surface and epoch/eval/checkpoint behaviour as the
reference's REC/trainer/trainer.py:19-409 (`Trainer(config, model)`, `.fit`, `.evaluate`,
`.resume_checkpoint`), with the B200 changes:
  * optimizer = FusedAdamW (one launch for all dense params, sparse-gradient dense-semantics update of the
    embedding table) instead of torch.optim.AdamW (trainer.py:66-103);
  * no per-step `.item()` sync: the running loss stays on the device, NaN is checked once per epoch
    (the reference syncs every step, trainer.py:120-121);
  * works with or without an initialised process group (the reference requires one, trainer.py:39).
"""
import os
from logging import getLogger
from time import time

import numpy as np
import torch

from .. import ops
from ..evaluator import Collector, Evaluator
from ..dist import ShardedTableEmbedding, ShardedTopK, shard_rows
from ..model.layers import TableEmbedding
from ..utils.utils import (barrier, calculate_valid_score, dict2str, dist_ready, early_stopping, ensure_dir,
                           get_local_time, get_rank, get_world_size)
from .optim import FusedAdamW


def unwrap(model):
    return model.module if hasattr(model, "module") else model


class Lookahead:
    """One-batch lookahead on a side stream: the next batch's host->device copy and (for a row-sharded table) its
    index-exchange plan -- which needs a host sync for NCCL's split sizes -- run while the current step's kernels
    execute, instead of draining the GPU at the top of every step."""

    def __init__(self, model, device):
        self.model = model
        self.device = device
        self.stream = torch.cuda.Stream(device=device) if torch.cuda.is_available() else None
        self.prefetch_plans = True

    def stage(self, batch):
        """host batch -> (device batch, ready event)"""
        if self.stream is None:
            return batch, None
        tensors = batch if isinstance(batch, (tuple, list)) else (batch,)
        if any(isinstance(t, torch.Tensor) and t.is_cuda for t in tensors):
            # a batch built ON the device (device_sampler: pr_seq_batch_build on the current stream) is only complete once
            # that stream reaches this point: the side stream must not read it (prefetch builds the exchange plan from
            # the ids) before then.  `.to()` is a no-op for such tensors, so nothing else orders the two streams.
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            for t in tensors:
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(self.stream)
        with torch.cuda.stream(self.stream):
            dev = tuple(t.to(self.device, non_blocking=True) for t in batch) if isinstance(batch, (tuple, list)) \
                else batch.to(self.device, non_blocking=True)
            if self.prefetch_plans and hasattr(self.model, "prefetch"):
                self.model.prefetch(dev)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dev, ev

    def acquire(self, staged):
        dev, ev = staged
        if ev is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for t in (dev if isinstance(dev, (tuple, list)) else (dev,)):
                t.record_stream(cur)
        return dev


class Trainer:
    def __init__(self, config, model):
        self.config = config
        self.model = model
        self.logger = getLogger()
        self.optim_args = config["optim_args"]
        self.epochs = config["epochs"]
        self.eval_step = min(config["eval_step"], self.epochs)
        self.stopping_step = config["stopping_step"]
        self.clip_grad_norm = config["clip_grad_norm"]
        self.valid_metric = config["valid_metric"].lower()
        self.valid_metric_bigger = config["valid_metric_bigger"]
        self.test_batch_size = config["eval_batch_size"]
        self.device = config["device"]
        self.rank = get_rank()
        self.checkpoint_dir = config["checkpoint_dir"] or "saved"
        if self.rank == 0:
            ensure_dir(self.checkpoint_dir)
        self.saved_model_file = os.path.join(self.checkpoint_dir, "{}-{}.pth".format(config["model"], get_local_time()))
        self.use_modality = config["use_modality"]
        self.start_epoch = 0
        self.cur_step = 0
        self.best_valid_score = -np.inf if self.valid_metric_bigger else np.inf
        self.best_valid_result = None
        self.train_loss_dict = {}
        self.optimizer = self._build_optimizer()
        self.eval_collector = Collector(config)
        self.evaluator = Evaluator(config)
        self.item_feature = None
        self.item_feature_f16 = None
        self.item_norm_max = None
        self._sharded_topk = None
        self.tot_item_num = None

    # ------------------------------------------------------------------ optimizer (trainer.py:66-103)
    def _build_optimizer(self):
        m = unwrap(self.model)
        tables = [mod for mod in m.modules() if isinstance(mod, (TableEmbedding, ShardedTableEmbedding))]
        a = self.optim_args
        if len(a) == 4:   # PixelNet: two groups keyed on 'visual_encoder' in the parameter name
            modal, rec = [], []
            for name, p in m.named_parameters():
                if p.requires_grad:
                    (modal if "visual_encoder" in name else rec).append(p)
            groups = [dict(params=modal, lr=a["modal_lr"], weight_decay=a["modal_decay"]),
                      dict(params=rec, lr=a["rec_lr"], weight_decay=a["rec_decay"])]
            groups = [g for g in groups if g["params"]]
            return FusedAdamW(groups, tables=tables)
        params = [p for p in m.parameters() if p.requires_grad]
        return FusedAdamW(params, lr=a["learning_rate"], weight_decay=a["weight_decay"], tables=tables)

    # ------------------------------------------------------------------ train loop (trainer.py:105-128)
    def to_device(self, data):
        if isinstance(data, (tuple, list)):
            return tuple(d.to(self.device, non_blocking=True) for d in data)
        if isinstance(data, dict):
            return {k: v.to(self.device, non_blocking=True) for k, v in data.items()}
        return data.to(self.device, non_blocking=True)

    def _train_epoch(self, train_data, epoch_idx, loss_func=None, show_progress=False):
        self.model.train()
        total = torch.zeros((), device=self.device)
        look = Lookahead(unwrap(self.model), self.device)
        it = iter(train_data)
        nxt = next(it, None)
        staged = look.stage(nxt) if nxt is not None else None
        graphed, n_eager = None, 0
        # yaml `cuda_graph: True`: one launch per step.  N > 1 needs the peer-memory exchange (GraphedTrainStep checks and the
        # loop stays eager otherwise)
        want_graph = bool(self.config["cuda_graph"]) and not self.clip_grad_norm
        while staged is not None:
            data = look.acquire(staged)
            nxt = next(it, None)
            staged = look.stage(nxt) if nxt is not None else None     # overlaps with this step's kernels
            if want_graph and graphed is None and n_eager >= 3:       # a few eager steps first (lazy initialisations)
                from .graph import GraphedTrainStep
                try:
                    graphed = GraphedTrainStep(unwrap(self.model), self.optimizer, data)
                except ops._lib.PixelRecB200Error as e:          # e.g. dropout > 0 on a build without -DPR_SEED_DEV
                    self.logger.warning("cuda_graph: staying eager (%s)" % e)
                    want_graph = False
            if graphed is not None and graphed.matches(data):
                look.prefetch_plans = False       # the index plan is built inside the captured step
                total += graphed(data).detach()
                continue
            n_eager += 1
            self.optimizer.zero_grad()
            losses = self.model(data)
            total += losses.detach()
            losses.backward()
            del losses     # keeps no autograd graph (and its AccumulateGrad nodes, bound to THIS stream) alive into a later capture
            # clip_grad_norm (trainer.py:123): over the dense gradients AND the table's sparse rows, after the rank average
            self.optimizer.step(clip=self.clip_grad_norm or None)
        if graphed is not None:
            graphed.close()
        total_loss = float(total.item())        # ONE device->host sync per epoch
        self._check_nan(total_loss)
        for m in unwrap(self.model).modules():  # peer-memory exchange (exchange: p2p): device flags, read once per epoch
            status = getattr(m, "exchange_status", None)
            if status is not None and status():
                raise RuntimeError("row exchange flagged an error (bit 0: id out of range, bit 1: a peer receive region "
                                   "overflowed and gradient rows were dropped -- raise PR_P2P_CAP_FACTOR)")
        return total_loss

    @staticmethod
    def _check_nan(loss):
        if np.isnan(loss):
            raise ValueError("Training loss is nan")

    def _valid_epoch(self, valid_data, show_progress=False):
        barrier()
        result = self.evaluate(valid_data, load_best_model=False, show_progress=show_progress)
        score = calculate_valid_score(result, self.valid_metric)
        barrier()
        return score, result

    # ------------------------------------------------------------------ checkpoint (trainer.py:138-190)
    def _save_checkpoint(self, epoch, verbose=True):
        model_state = unwrap(self.model).state_dict()      # collective when the table is sharded: every rank calls it
        optim_state = self.optimizer.state_dict()          # likewise (the table's Adam moments are gathered to [N, D])
        if self.rank == 0:
            state = {
                "config": self.config, "epoch": epoch, "cur_step": self.cur_step,
                "best_valid_score": self.best_valid_score, "state_dict": model_state,
                "optimizer": optim_state, "rng_state": torch.get_rng_state(),
                "cuda_rng_state": torch.cuda.get_rng_state() if torch.cuda.is_available() else None,
            }
            torch.save(state, self.saved_model_file)
            if verbose:
                self.logger.info(f"Saving current: {self.saved_model_file}")
        barrier()

    def resume_checkpoint(self, resume_file):
        ck = torch.load(str(resume_file), map_location="cpu", weights_only=False)
        self.start_epoch = ck["epoch"] + 1
        self.cur_step = ck["cur_step"]
        self.best_valid_score = ck["best_valid_score"]
        if ck["config"]["model"].lower() != self.config["model"].lower():
            self.logger.warning("Architecture configuration given in config file is different from that of checkpoint.")
        unwrap(self.model).load_state_dict(ck["state_dict"])     # the reference forgets this
        self.optimizer.load_state_dict(ck["optimizer"])
        torch.set_rng_state(ck["rng_state"])
        if ck.get("cuda_rng_state") is not None and torch.cuda.is_available():
            torch.cuda.set_rng_state(ck["cuda_rng_state"])
        self.logger.info("Checkpoint loaded. Resume training from epoch {}".format(self.start_epoch))

    # ------------------------------------------------------------------ fit (trainer.py:256-326)
    def fit(self, train_data, valid_data=None, verbose=True, saved=True, show_progress=False, callback_fn=None):
        if saved and self.start_epoch >= self.epochs:
            self._save_checkpoint(-1, verbose=verbose)
        for epoch_idx in range(self.start_epoch, self.epochs):
            if self.config["need_training"] is None or self.config["need_training"]:
                if hasattr(train_data, "sampler") and hasattr(train_data.sampler, "set_epoch"):
                    train_data.sampler.set_epoch(epoch_idx)
                t0 = time()
                train_loss = self._train_epoch(train_data, epoch_idx, show_progress=show_progress)
                self.train_loss_dict[epoch_idx] = train_loss
                if verbose:
                    self.logger.info("epoch %d training [time: %.2fs, train loss: %.4f]" % (epoch_idx, time() - t0, train_loss))
            if self.eval_step <= 0 or not valid_data:
                if saved:
                    self._save_checkpoint(epoch_idx, verbose=verbose)
                continue
            if (epoch_idx + 1) % self.eval_step == 0:
                t0 = time()
                valid_score, valid_result = self._valid_epoch(valid_data, show_progress=show_progress)
                self.best_valid_score, self.cur_step, stop_flag, update_flag = early_stopping(
                    valid_score, self.best_valid_score, self.cur_step, max_step=self.stopping_step,
                    bigger=self.valid_metric_bigger)
                if verbose:
                    self.logger.info("epoch %d evaluating [time: %.2fs, valid_score: %f]" % (epoch_idx, time() - t0, valid_score))
                    self.logger.info("valid result: \n" + dict2str(valid_result))
                if update_flag:
                    if saved:
                        self._save_checkpoint(epoch_idx, verbose=verbose)
                    self.best_valid_result = valid_result
                if callback_fn:
                    callback_fn(epoch_idx, valid_score)
                if stop_flag:
                    if verbose:
                        self.logger.info("Finished training, best eval result in epoch %d" %
                                         (epoch_idx - self.cur_step * self.eval_step))
                    break
        return self.best_valid_score, self.best_valid_result

    # ------------------------------------------------------------------ evaluation (trainer.py:327-409)
    @torch.no_grad()
    def _full_sort_batch_eval(self, batched_data):
        user, history_index, positive_u, positive_i = batched_data
        scores = unwrap(self.model).predict(self.to_device(user), self.item_feature)
        scores = scores.view(-1, self.tot_item_num)
        scores[:, 0] = -np.inf
        if history_index is not None:
            hu, hi = history_index
            scores[hu.to(self.device), hi.to(self.device)] = -np.inf
        return scores, positive_u, positive_i

    @torch.no_grad()
    def _fused_topk_batch_eval(self, batched_data):
        """Same contract as _full_sort_batch_eval + Collector's torch.topk, but through pr_score_topk_f32:
        encoder -> tcgen05 scoring GEMM with the pad-column / history mask and the top-k fused in its epilogue.
        The [B_e, N] score matrix (397 MB at C2) is never written."""
        user, history_index, positive_u, positive_i = batched_data
        model = unwrap(self.model)
        seq_out = model.encode_last(self.to_device(user), self.item_feature)
        hu = hi = None
        if history_index is not None:
            hu, hi = (x.to(self.device).contiguous() for x in history_index)
        if self._scoring_mode() == "tcgen05_f16" and self.item_feature.shape[1] % 64 == 0:     # staged, opt-in
            if self.item_feature_f16 is None:
                self.item_feature_f16 = ops.score_prepare_f16(self.item_feature.contiguous())   # once per evaluate()
            _, topk_idx = ops.score_topk_f16(seq_out, self.item_feature_f16, max(self.config["topk"]), hu, hi, mask_col0=True)
        elif self._scoring_mode() == "tcgen05" and max(self.config["topk"]) <= 16 and self.item_feature.shape[0] >= 32:
            # default: ids of the reference's fp32 ranking (TF32 candidates, fp32 re-score, proven complete or re-ranked)
            if self.item_norm_max is None:
                self.item_norm_max = ops.table_norm_max(self.item_feature.contiguous())          # once per evaluate()
            _, topk_idx, _ = ops.score_topk_exact(seq_out, self.item_feature.contiguous(), max(self.config["topk"]), hu, hi,
                                                  mask_col0=True, w_norm_max=self.item_norm_max)
        else:                                            # "tcgen05_tf32": ranks the TF32 scores directly (near-ties may swap)
            _, topk_idx = ops.score_topk(seq_out, self.item_feature.contiguous(), max(self.config["topk"]), hu, hi,
                                         mask_col0=True)
        return topk_idx, positive_u, positive_i

    def _scoring_mode(self):
        return (self.config["eval_scoring"] or "tcgen05").lower()

    def _use_fused_topk(self):
        mode = self._scoring_mode()
        model = unwrap(self.model)
        D = self.item_feature.shape[1]
        return (mode in ("tcgen05", "tcgen05_tf32", "tcgen05_f16") and hasattr(model, "encode_last") and D % 32 == 0
                and max(self.config["topk"]) <= 32 and self.item_feature.is_cuda)

    def _use_sharded_eval(self):
        """yaml `eval_table`: auto (default) | sharded | gathered.  Sharded: the table stays row-sharded during evaluation and
        every batch is ranked shard by shard with a candidate merge (dist.ShardedTopK) instead of all-gathering [N, D] per rank."""
        mode = (self.config["eval_table"] or "auto").lower()
        model = unwrap(self.model)
        tab = getattr(model, "item_embedding", None)
        if mode == "gathered" or not isinstance(tab, ShardedTableEmbedding) or tab.world == 1:
            return False
        ok = (tab.exchange == "p2p" and self._scoring_mode() == "tcgen05" and hasattr(model, "encode_last")
              and max(self.config["topk"]) <= 16 and tab.embedding_dim % 32 == 0 and tab.weight.is_cuda
              and shard_rows(tab.num_embeddings, tab.world, tab.world - 1) >= 32 and (tab.padding_idx in (None, 0)))
        if mode == "sharded" and not ok:
            raise ValueError("eval_table: sharded needs the p2p exchange, eval_scoring: tcgen05, topk <= 16 and >= 32 rows per shard")
        return ok

    @torch.no_grad()
    def _evaluate_sharded(self, eval_data):
        """The evaluation loop of evaluate() with the table left sharded.  Collectives per batch: every rank runs the same number of
        steps (ranks that hold fewer eval batches -- data/utils.py strided sampler -- contribute empty, padded ones)."""
        import torch.distributed as dist
        model = unwrap(self.model)
        tab = model.item_embedding
        k = max(self.config["topk"])
        B_e = int(self.config["eval_batch_size"])
        L = int(self.config["MAX_ITEM_LIST_LENGTH"])
        model.train_lookups_hint = int(self.config["train_batch_size"]) * 2 * (L + 1)
        torch.cuda.synchronize(self.device)
        dist.barrier()                                   # every rank's last optimizer step has landed: the shards are static now
        nb = torch.tensor([len(eval_data)], dtype=torch.int64, device=self.device)
        dist.all_reduce(nb, op=dist.ReduceOp.MAX)
        if self._sharded_topk is None:
            self._sharded_topk = ShardedTopK(tab.world, tab.rank)
        self._sharded_topk.reset()
        W_local = tab.weight.detach()
        it = iter(eval_data)
        for _ in range(int(nb.item())):
            batch = next(it, None)
            seqs = torch.zeros(B_e, L, dtype=torch.int64, device=self.device)
            b, hu, hi = 0, None, None
            if batch is not None:
                user, history_index, positive_u, positive_i = batch
                b = user.shape[0]
                seqs[:b] = self.to_device(user)
                if history_index is not None:
                    hu, hi = history_index
            seq_out = model.encode_last(seqs, None)
            _, topk_idx = self._sharded_topk(seq_out, W_local, k, hu, hi, pad_id=0)
            if b:
                self.eval_collector.eval_batch_collect_topk(topk_idx[:b], positive_u, positive_i)
        dist.barrier()                                   # nobody resumes training (and updates its shard) while a peer still reads it

    @torch.no_grad()
    def compute_item_feature(self, config, data):
        self.item_feature = unwrap(self.model).compute_item_all()
        self.item_feature_f16 = None
        self.item_norm_max = None

    def distributed_concat(self, tensor, num_total_examples):
        if dist_ready():
            outs = [tensor.clone() for _ in range(get_world_size())]
            torch.distributed.all_gather(outs, tensor)
            tensor = torch.cat(outs, dim=0)
        return tensor.sum() / num_total_examples

    @torch.no_grad()
    def evaluate(self, eval_data, load_best_model=True, model_file=None, show_progress=False):
        if not eval_data:
            return
        if load_best_model:
            ck = torch.load(model_file or self.saved_model_file, map_location="cpu", weights_only=False)
            unwrap(self.model).load_state_dict(ck["state_dict"])
            self.logger.info("Loading model structure and parameters from {}".format(model_file or self.saved_model_file))
        self.model.eval()
        self.tot_item_num = eval_data.dataset.dataload.item_num
        if self._use_sharded_eval():
            self._evaluate_sharded(eval_data)
        else:
            self.compute_item_feature(self.config, eval_data.dataset.dataload)
            fused = self._use_fused_topk()
            for batched_data in eval_data:
                if fused:
                    topk_idx, positive_u, positive_i = self._fused_topk_batch_eval(batched_data)
                    self.eval_collector.eval_batch_collect_topk(topk_idx, positive_u, positive_i)
                else:
                    scores, positive_u, positive_i = self._full_sort_batch_eval(batched_data)
                    self.eval_collector.eval_batch_collect(scores, positive_u, positive_i)
        num_total_examples = len(eval_data.sampler.dataset)
        result = self.evaluator.evaluate(self.eval_collector.get_data_struct())
        places = 5 if self.config["metric_decimal_place"] is None else self.config["metric_decimal_place"]
        for k, v in result.items():
            r = self.distributed_concat(torch.tensor([v], dtype=torch.float64).to(self.device), num_total_examples).cpu()
            result[k] = round(r.item(), places)
        return result
