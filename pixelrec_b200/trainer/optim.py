"""Fused AdamW for the SASRec hot path -- replaces torch.optim.AdamW as built by the reference trainer
(REC/trainer/trainer.py:66-103) with the same update rule (decoupled weight decay, betas (0.9, 0.999),
eps 1e-8) and the same param-group semantics, but:

  * every dense (non-table) parameter of a group lives in ONE flat fp32 buffer (parameters and their
    .grad become views), so zero_grad is one memset and step is ONE pr_adamw_dense_f32 launch instead
    of ~12 elementwise kernels per tensor;
  * embedding tables (model.layers.TableEmbedding) keep a SPARSE gradient (ScatterPlan + reduced rows from
    pr_scatter_add_rows_f32) while the update stays exactly dense: pr_adamw_rows_f32 walks every row
    (weight decay and moment decay apply to untouched rows as in the reference) and looks the gradient up
    through row2slot.  No dense [N,D] zero-fill, no dense gradient tensor.
"""
import torch

from .. import ops
from ..dist import ShardedTableEmbedding, world_info
from ..model.layers import TableEmbedding


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, tables=(), grad_scale=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.world, _ = world_info()
        # data parallel: gradients are SUMMED over ranks (all_reduce of the flat buffers / row exchange of the
        # sharded table) and scaled by 1/world inside the kernels == DDP's gradient mean (run.py:40)
        self.grad_scale = float(grad_scale) / self.world
        self._step = 0
        self._step_dev = None
        self._tables = {}
        for tb in tables:
            if not isinstance(tb, (TableEmbedding, ShardedTableEmbedding)):
                raise TypeError("tables must be TableEmbedding modules")
            self._tables[id(tb.weight)] = tb
        self._flat = []          # per group: dict(w, g, m, v, params)
        self._table_state = {}   # id(weight) -> (M, V, group index)
        for gi, group in enumerate(self.param_groups):
            dense = []
            for p in group["params"]:
                if not p.is_cuda or p.dtype != torch.float32:
                    raise ops._lib.PixelRecB200Error("FusedAdamW needs fp32 CUDA parameters (no CPU fallback)")
                if id(p) in self._tables:
                    tb = self._tables[id(p)]
                    tb.sink.enable_sparse()
                    self._table_state[id(p)] = (torch.zeros_like(p), torch.zeros_like(p), gi)
                else:
                    dense.append(p)
            self._flat.append(self._flatten(dense))

    @staticmethod
    def _flatten(params):
        if not params:
            return None
        dev = params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]          # keep every view 16-byte aligned
        n = sum(sizes)
        w = torch.zeros(n, device=dev, dtype=torch.float32)
        g = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p, sz in zip(params, sizes):
                view = w[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = g[off:off + p.numel()].view_as(p)
                off += sz
        return dict(w=w, g=g, m=torch.zeros_like(w), v=torch.zeros_like(w), params=params)

    def zero_grad(self, set_to_none=False):
        """Reference semantics (torch 1.10 zero_grad): dense grads are zero-filled (one memset per group);
        tables have no dense gradient to clear."""
        for f in self._flat:
            if f is not None:
                f["g"].zero_()
                for p in f["params"]:
                    if p.grad is None or p.grad.data_ptr() < f["g"].data_ptr() or \
                            p.grad.data_ptr() >= f["g"].data_ptr() + f["g"].numel() * 4:
                        self._rebind_grads(f)
                        break
        for tb in self._tables.values():
            tb.sink.pending.clear()
            if tb.weight.grad is not None:
                tb.weight.grad = None

    @staticmethod
    def _rebind_grads(f):
        off = 0
        for p in f["params"]:
            p.grad = f["g"][off:off + p.numel()].view_as(p)
            off += (p.numel() + 3) // 4 * 4

    def use_device_step(self, on=True):
        """CUDA-graph mode: keep the step count that the bias correction uses in device memory and bump it with a kernel that
        is part of the captured step (pr_adamw_*'s step_dev argument), instead of freezing a host integer into the graph.
        Call right before capturing; the device count starts at the steps taken so far."""
        if not on:
            self._step_dev = None
            return None
        dev = next(p for g in self.param_groups for p in g["params"]).device
        self._step_dev = torch.full((1,), self._step, dtype=torch.int64, device=dev)
        return self._step_dev

    def flat_grads(self):
        """Flat dense-gradient buffers (one per group) -- the data-parallel all-reduce works on these."""
        return [f["g"] for f in self._flat if f is not None]

    @torch.no_grad()
    def clip_grad_norm_(self, max_norm, norm_type=2.0):
        """torch.nn.utils.clip_grad_norm_ over ALL parameters of the optimizer -- the flat dense gradients AND the tables' sparse
        gradient rows -- with DDP's semantics (the norm of the rank-averaged gradient, trainer.py:123 after run.py:40).  Call it
        through step(clip=...): the dense gradients must already be summed over ranks.  No host synchronisation."""
        if float(norm_type) != 2.0:
            raise NotImplementedError("FusedAdamW.clip_grad_norm_: only the 2-norm is implemented")
        dev = next(p for g in self.param_groups for p in g["params"]).device
        sq = torch.zeros((), device=dev, dtype=torch.float32)
        for f in self._flat:
            if f is not None:
                sq += (f["g"] * f["g"]).sum()
        tsq = torch.zeros((), device=dev, dtype=torch.float32)
        sharded = False
        for tb in self._tables.values():
            if self.world > 1 and not isinstance(tb, ShardedTableEmbedding):
                raise NotImplementedError("clip_grad_norm with a replicated table on several GPUs")
            sharded = sharded or isinstance(tb, ShardedTableEmbedding)
            for plan, rows in tb.sink.pending:                      # rows >= n_uniq are undefined memory: mask, do not multiply
                live = torch.arange(rows.shape[0], device=dev) < plan.n_uniq
                tsq += torch.where(live, (rows * rows).sum(1), torch.zeros((), device=dev)).sum()
        if self.world > 1 and sharded:
            torch.distributed.all_reduce(tsq)                       # every rank owns different rows
        total = torch.sqrt(sq + tsq) * self.grad_scale              # grad_scale = 1 / world: the mean over ranks
        coef = torch.clamp(float(max_norm) / (total + 1e-6), max=1.0)
        for f in self._flat:
            if f is not None:
                f["g"].mul_(coef)
        for tb in self._tables.values():
            for plan, rows in tb.sink.pending:
                rows.mul_(coef)
        return total

    @torch.no_grad()
    def step(self, closure=None, clip=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._step += 1
        sd = self._step_dev                      # CUDA-graph mode (use_device_step): the kernels read the count from device memory
        if sd is not None:
            sd.add_(1)
        if self.world > 1:
            for f in self._flat:
                if f is not None:
                    with ops._prof("dense_grad_allreduce", f["g"]):
                        torch.distributed.all_reduce(f["g"])
        if clip:
            self.clip_grad_norm_(**clip)
        for gi, group in enumerate(self.param_groups):
            b1, b2 = group["betas"]
            f = self._flat[gi]
            if f is not None:
                ops.adamw_dense(f["w"], f["g"], f["m"], f["v"], group["lr"], b1, b2, group["eps"],
                                group["weight_decay"], self._step, self.grad_scale, step_dev=sd)
        for pid, (M, V, gi) in self._table_state.items():
            tb = self._tables[pid]
            group = self.param_groups[gi]
            b1, b2 = group["betas"]
            pend = tb.sink.pending
            W = tb.weight.data
            if len(pend) == 1:
                plan, rows = pend[0]
                ops.adamw_rows(W, M, V, rows, tb.sink.row2slot, group["lr"], b1, b2, group["eps"],
                               group["weight_decay"], self._step, self.grad_scale, step_dev=sd)
            elif len(pend) == 0:
                ops.adamw_rows(W, M, V, None, None, group["lr"], b1, b2, group["eps"], group["weight_decay"],
                               self._step, self.grad_scale, step_dev=sd)
            else:   # several lookups of one table in a step: merge through a dense buffer (rare path)
                tb.sink.row2slot.fill_(-1)
                G = torch.zeros_like(W)
                for plan, rows in pend:
                    U = int(plan.n_uniq.item())
                    G.index_add_(0, plan.uniq_ids[:U].long(), rows[:U])
                ops.adamw_dense(W.view(-1), G.view(-1), M.view(-1), V.view(-1), group["lr"], b1, b2, group["eps"],
                                group["weight_decay"], self._step, self.grad_scale)
            pend.clear()
        return loss

    # ---- checkpoint interchange (trainer.py:153 stores optimizer.state_dict()) ----
    def state_dict(self):
        """Collective when a table is row-sharded (every rank must call it, like model.state_dict()): the table's Adam moments are
        gathered into the reference's full [N, D] layout, so a checkpoint written by rank 0 restores every rank's shard."""
        sd = super().state_dict()
        tables = []
        for pid, (M, V, _) in self._table_state.items():
            tb = self._tables[pid]
            if isinstance(tb, ShardedTableEmbedding):
                tables.append((tb.gather_rows_full(M), tb.gather_rows_full(V)))
            else:
                tables.append((M.clone(), V.clone()))
        sd["fused"] = dict(step=self._step,
                           flat=[None if f is None else dict(m=f["m"].clone(), v=f["v"].clone()) for f in self._flat],
                           tables=tables)
        return sd

    def load_state_dict(self, sd):
        fused = sd.get("fused")
        super().load_state_dict({k: v for k, v in sd.items() if k != "fused"})
        if fused:
            self._step = fused["step"]
            for f, s in zip(self._flat, fused["flat"]):
                if f is not None and s is not None:
                    f["m"].copy_(s["m"]); f["v"].copy_(s["v"])
            for (pid, (M, V, _)), (m2, v2) in zip(self._table_state.items(), fused["tables"]):
                tb = self._tables[pid]
                if isinstance(tb, ShardedTableEmbedding) and m2.shape[0] == tb.num_embeddings and tb.world > 1:
                    m2, v2 = tb.shard_of_full(m2), tb.shard_of_full(v2)          # full [N, D] moments -> rows rank::world
                M.copy_(m2); V.copy_(v2)
