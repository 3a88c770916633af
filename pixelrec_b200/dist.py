"""Multi-GPU plumbing of the hot path: one process per GPU, `torch.distributed` (NCCL over
NVLink/NVSwitch) ONLY at the shard boundary.

  * batch: data parallel -- every rank runs the full model on its own sequences (the reference's DDP, run.py:40);
    dense (non-table) gradients are summed with ONE all_reduce of the optimizer's flat gradient buffer and scaled
    by 1/world inside the fused AdamW kernel (== DDP's mean);
  * item table: ROW-SHARDED, owner(i) = i % world, local row = i // world (interleaving spreads the popular head
    of the long-tail).  Forward: bucket the step's indices by owner -> all_to_all_single(indices) -> the owner
    gathers its rows with pr_gather_rows_f32 straight into the send buffer -> all_to_all_single(rows) ->
    un-permute with pr_gather_rows_f32.  Backward is the mirror image and ends in pr_scatter_plan +
    pr_scatter_add_rows_f32 on the owner (duplicates across ranks are reduced there) feeding the sharded
    dense-semantics AdamW.  There is no dense [N,D] gradient and no dense all-reduce of the table anywhere
    (the reference all-reduces N*D*4 bytes every step).

  * `exchange="p2p"` (the default; PR_EXCHANGE=nccl selects the all_to_all path above): the two row all_to_alls,
    the owner-side gather into a send buffer, the index all_to_all and its host sync are replaced by peer-memory
    kernels (csrc/peer.cu).  Every rank maps its peers' shards and receive buffers (CUDA IPC); the lookup reads each
    distinct row straight out of its owner's shard over NVLink (pr_gather_rows_peers_f32), the backward writes the
    locally reduced gradient rows straight into the owner's receive region (pr_push_rows_peers_f32).  NCCL is left
    with two 4-byte all_reduces per step that order those kernels against the owners' optimizer step.

The row kernels are reached through `ROWS` (and the peer-memory ones through `PEER`) so that the exchange logic can be
unit-tested on CPU/gloo with oracle-backed stand-ins (tests/test_dist_gloo.py); the product default is the CUDA
library and nothing else.
"""
from __future__ import annotations

import math
import os

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops


class CudaRows:
    """Default row backend: our sm_100a kernels."""

    @staticmethod
    def gather(W, idx):
        return ops.gather_rows(W, idx)

    @staticmethod
    def plan(idx, N, padding_idx, row2slot=None):
        return ops.ScatterPlan(idx, N, padding_idx, row2slot=row2slot)

    @staticmethod
    def scatter(dOut, plan):
        return ops.scatter_add_rows(dOut, plan)

    @staticmethod
    def scatter_slots(dOut, slot_idx, U, pad_slot):
        """[U, D] with row u = sum of dOut rows whose slot is u (slot `pad_slot` is skipped and left zero)."""
        plan = ops.ScatterPlan(slot_idx, U, pad_slot if pad_slot >= 0 else None)
        G = torch.empty(U, dOut.shape[-1], device=dOut.device, dtype=dOut.dtype)
        if pad_slot >= 0:
            G[pad_slot].zero_()
        ops.scatter_add_rows(dOut, plan, dense_G=G, out_rows=False)
        return G


ROWS = CudaRows


class CudaPeer:
    """Default peer-memory backend: CUDA IPC allocations + the kernels of csrc/peer.cu."""

    @staticmethod
    def alloc(nbytes, device):
        return ops.SharedBuffer(nbytes, device)          # .ref (device address), .handle, .tensor(shape, dtype)

    @staticmethod
    def open(handle, device):
        return ops.shared_open(handle, device)           # device address of the peer's buffer, mapped here

    @staticmethod
    def table(refs, device):
        return torch.tensor(refs, dtype=torch.int64, device=device)

    @staticmethod
    def gather(shard_table, G, N, D, idx):
        return ops.gather_rows_peers(shard_table, G, N, D, idx)

    @staticmethod
    def push(rows, ids, G, rank, cap, skip_id, rows_table, ids_table, counters, status):
        ops.push_rows_peers(rows, ids, G, rank, cap, skip_id, rows_table, ids_table, counters, status)

    @staticmethod
    def barrier(flag_table, G, rank, epoch, status, epoch_dev=None):
        ops.peer_barrier(flag_table, G, rank, epoch, status, epoch_dev)

    device_plan = True          # the plan variants below exist (CPU stand-ins in tests/ keep the host-side PeerPlan)


PEER = CudaPeer


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_rows(N, world, rank):
    """Number of table rows owned by `rank` under owner(i) = i % world."""
    return (N - rank + world - 1) // world


class ExchangePlan:
    """Index bucketing for one lookup.  The step's indices are DE-DUPLICATED first (long-tail ids repeat: ~70 K unique
    of 172 K lookups at C2/B=4096), so each distinct row crosses NVLink once per direction; `inverse` expands the
    received unique rows back to request order and reduces the gradient before it is sent."""

    def __init__(self, idx, world, group=None, padding_idx=None):
        flat = idx.reshape(-1)
        self.R = flat.numel()
        uniq, self.inverse = torch.unique(flat, sorted=True, return_inverse=True)
        self.U = uniq.numel()
        owner = uniq % world
        self.perm = torch.argsort(owner, stable=True)                 # unique slots sorted by owner
        inv_perm = torch.empty_like(self.perm)
        inv_perm[self.perm] = torch.arange(self.U, device=flat.device)
        self.expand = inv_perm[self.inverse].contiguous()              # request position -> row of the receive buffer
        send_counts = torch.bincount(owner, minlength=world)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        # slot of the padding id among the unique ids (-1 if absent): its gradient is dropped by the owner anyway, and
        # left-padding makes it by far the longest run (40 % of all lookups), so the local reduction skips it
        if padding_idx is None:
            pad_slot = torch.full((1,), -1, dtype=send_counts.dtype, device=flat.device)
        else:
            hit = (uniq == padding_idx).nonzero()
            pad_slot = hit[0] if hit.numel() else torch.full((1,), -1, dtype=send_counts.dtype, device=flat.device)
        both = torch.cat([send_counts, recv_counts, pad_slot.view(1).to(send_counts.dtype)]).cpu()   # the one host sync
        self.send_splits = both[:world].tolist()
        self.recv_splits = both[world:2 * world].tolist()
        self.pad_slot = int(both[2 * world])
        local_rows = torch.div(uniq, world, rounding_mode="floor")[self.perm].contiguous()
        self.recv_rows = torch.empty(sum(self.recv_splits), dtype=torch.int64, device=flat.device)
        dist.all_to_all_single(self.recv_rows, local_rows, self.recv_splits, self.send_splits, group=group)

    def tensors(self):
        return (self.inverse, self.perm, self.expand, self.recv_rows)


class ShardedGatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, W_local, idx, table):
        plan = table.take_plan(idx)
        D = W_local.shape[1]
        rows_send = ROWS.gather(W_local, plan.recv_rows)                              # owner-side gather
        rows_recv = torch.empty(plan.U, D, dtype=W_local.dtype, device=W_local.device)
        with ops._prof("exchange_fwd_a2a", rows_send):                                # event-timed by bench.py at N > 1
            dist.all_to_all_single(rows_recv, rows_send, plan.send_splits, plan.recv_splits, group=table.group)
        out = ROWS.gather(rows_recv, plan.expand)                                     # unique rows -> request order
        ctx.plan = plan
        ctx.table = table
        ctx.D = D
        return out.view(*idx.shape, D)

    @staticmethod
    def backward(ctx, dE):
        plan, table, D = ctx.plan, ctx.table, ctx.D
        dE = dE.contiguous().view(plan.R, D)
        # reduce duplicates locally (slot order == ascending unique id), then owner order
        d_u = ROWS.scatter_slots(dE, plan.inverse.contiguous(), plan.U, plan.pad_slot)
        d_send = ROWS.gather(d_u, plan.perm)
        d_recv = torch.empty(len(plan.recv_rows), D, dtype=dE.dtype, device=dE.device)
        with ops._prof("exchange_bwd_a2a", d_send):
            dist.all_to_all_single(d_recv, d_send, plan.recv_splits, plan.send_splits, group=table.group)
        splan = ROWS.plan(plan.recv_rows, table.n_local, table.local_padding_idx, row2slot=table.sink.row2slot)
        rows = ROWS.scatter(d_recv, splan)                                            # duplicates ACROSS ranks reduced here
        table.sink.deposit(splan, rows)
        return None, None, None


class PeerPlan:
    """Index plan of one lookup for the peer-memory exchange: the distinct ids (each row crosses NVLink once per
    direction) and the map back to request order.  No collective and no split sizes -- it can be built on any stream."""

    def __init__(self, idx, padding_idx=None):
        flat = idx.reshape(-1)
        self.R = flat.numel()
        self.uniq, inverse = torch.unique(flat, sorted=True, return_inverse=True)
        self.inverse = inverse.contiguous()
        self.U = self.uniq.numel()
        self.pad_slot = -1
        if padding_idx is not None and self.U:
            hit = (self.uniq == padding_idx).nonzero()
            if hit.numel():
                self.pad_slot = int(hit[0])

    def tensors(self):
        return (self.uniq, self.inverse)


class PeerPlanDev:
    """Device-side index plan of one lookup for the peer-memory exchange: ONE pr_scatter_plan of the step's ids (padding id
    dropped) gives the distinct ids, the runs that the backward reduces, and -- through pr_plan_inverse -- the map back to request
    order.  The number of distinct ids never leaves the device: no host synchronisation, so the whole step (lookup, exchange,
    backward, optimizer) can be captured into a CUDA graph and costs one launch per replay."""

    def __init__(self, idx, N, padding_idx=None):
        flat = idx.reshape(-1)
        self.R = flat.numel()
        self.splan = ops.ScatterPlan(flat, N, padding_idx)
        self.pad_id = padding_idx
        self.pad_slot = self.splan.max_uniq if padding_idx is not None else -1      # the pull's extra row holds W[pad]
        self.inverse = ops.plan_inverse(self.splan, self.pad_slot)

    def tensors(self):
        p = self.splan
        return (p.perm, p.uniq_ids, p.seg_start, p.n_uniq, p._ws, self.inverse)


class PeerExchange:
    """Peer-mapped state of one ShardedTableEmbedding: the shard itself (moved into shareable memory), this rank's
    receive buffers, and device tables of every rank's addresses."""

    def __init__(self, table, R):
        G, rank, D = table.world, table.rank, table.embedding_dim
        dev = table.weight.device
        factor = float(os.environ.get("PR_P2P_CAP_FACTOR", "2.0"))
        r_all = torch.tensor([int(R)], dtype=torch.int64, device=dev)
        dist.all_reduce(r_all, op=dist.ReduceOp.MAX, group=table.group)       # ranks may start on batches of different size
        R = int(r_all.item())
        self.R = R
        self.cap = int(max(1, min(R, max(64, math.ceil(factor * R / G)))))    # rows one source may send one owner per step
        self.G, self.rank = G, rank
        # the shard moves into shareable memory; the Parameter object (and its optimizer state) stay the same
        self._w = PEER.alloc(max(table.n_local, 1) * D * 4, dev)
        w = self._w.tensor((table.n_local, D), torch.float32)
        with torch.no_grad():
            w.copy_(table.weight.data)
            table.weight.data = w
        self._rows = PEER.alloc(G * self.cap * D * 4, dev)
        self._ids = PEER.alloc(G * self.cap * 8, dev)
        self.recv_rows = self._rows.tensor((G * self.cap, D), torch.float32)
        self.recv_ids = self._ids.tensor((G * self.cap,), torch.int64)
        self.recv_ids.fill_(-1)                                                # -1 = unused slot (dropped by the plan)
        # barrier flags: one uint64 per peer, written remotely by that peer (pr_peer_barrier); PR_P2P_BARRIER=nccl keeps the
        # 4-byte all_reduce instead
        self.flag_barrier = os.environ.get("PR_P2P_BARRIER", "flags").lower() == "flags" and hasattr(PEER, "barrier")
        self._flags = PEER.alloc(max(G, 1) * 8, dev)
        self._flags.tensor((G,), torch.int64).zero_()
        self.epoch = 0
        self.epoch_dev = None
        mine = (self._w.handle, self._rows.handle, self._ids.handle, self.cap, self._flags.handle)
        everyone = [None] * G
        dist.all_gather_object(everyone, mine, group=table.group)
        if any(e[3] != self.cap for e in everyone):
            raise ops._lib.PixelRecB200Error("peer exchange: ranks disagree on the receive capacity (unequal batch sizes?)")
        own = (self._w.ref, self._rows.ref, self._ids.ref, None, self._flags.ref)
        refs = [[own[k] if r == rank else PEER.open(everyone[r][k], dev) for r in range(G)] for k in (0, 1, 2, 4)]
        self.w_table, self.rows_table, self.ids_table, self.flag_table = (PEER.table(x, dev) for x in refs)
        self.counters = torch.zeros(G, dtype=torch.int32, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self._sync = torch.zeros(1, dtype=torch.float32, device=dev)
        self.group = table.group
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        dist.barrier(group=table.group)          # every shard copied and every id buffer initialised before the first pull

    def use_device_epoch(self, on=True):
        """CUDA-graph mode: the barrier's epoch is counted in device memory (a captured kernel cannot take a new host value per
        replay).  Call right before capturing / after the last replay; the count carries over in both directions."""
        if on and self.epoch_dev is None:
            self.epoch_dev = torch.full((1,), self.epoch, dtype=torch.int64, device=self.status.device)
        elif not on and self.epoch_dev is not None:
            self.epoch = int(self.epoch_dev.item())
            self.epoch_dev = None

    def barrier(self):
        """Orders the peer kernels of all ranks on their streams (asynchronous w.r.t. the host): one small kernel over peer-mapped
        flags (pr_peer_barrier), or a 4-byte NCCL all_reduce with PR_P2P_BARRIER=nccl."""
        if self.flag_barrier:
            self.epoch += 1
            if self.epoch_dev is not None:
                PEER.barrier(self.flag_table, self.G, self.rank, self.epoch, self.status, self.epoch_dev)
            else:
                PEER.barrier(self.flag_table, self.G, self.rank, self.epoch, self.status)
            return
        with ops._prof("exchange_barrier", self._sync):
            dist.all_reduce(self._sync, group=self.group)


class PeerGatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, W_local, idx, table):
        plan = table.take_plan(idx)
        px = table.peer_exchange(plan.R)
        D = table.embedding_dim
        px.barrier()                          # S1: every owner has finished the optimizer step of the previous batch
        if isinstance(plan, PeerPlanDev):
            rows_u = ops.gather_rows_peers_plan(px.w_table, table.world, table.num_embeddings, D, plan.splan, plan.pad_id,
                                                plan.pad_slot, px.status)                   # lookup + exchange, one kernel
        else:
            rows_u = PEER.gather(px.w_table, table.world, table.num_embeddings, D, plan.uniq)
        out = ROWS.gather(rows_u, plan.inverse)                                             # distinct rows -> request order
        ctx.plan, ctx.table, ctx.px = plan, table, px
        return out.view(*idx.shape, D)

    @staticmethod
    def backward(ctx, dE):
        plan, table, px = ctx.plan, ctx.table, ctx.px
        D = table.embedding_dim
        if plan.R > px.R:
            raise ops._lib.PixelRecB200Error(f"peer exchange sized for {px.R} lookups per step, got {plan.R}")
        dE = dE.contiguous().view(plan.R, D)
        px.counters.zero_()
        if isinstance(plan, PeerPlanDev):
            d_u = ROWS.scatter(dE, plan.splan)                                   # local duplicates reduced first (pad dropped)
            ops.push_rows_peers_plan(d_u, plan.splan, table.world, table.rank, px.cap, px.rows_table, px.ids_table,
                                     px.counters, px.status)
        else:
            d_u = ROWS.scatter_slots(dE, plan.inverse, plan.U, plan.pad_slot)
            PEER.push(d_u, plan.uniq, table.world, table.rank, px.cap, table.padding_idx, px.rows_table, px.ids_table,
                      px.counters, px.status)
        px.barrier()                          # S2: every rank's rows have landed in the owners' receive buffers
        splan = ROWS.plan(px.recv_ids, table.n_local, None, row2slot=table.sink.row2slot)
        rows = ROWS.scatter(px.recv_rows, splan)            # duplicates ACROSS ranks reduced here, in ascending source rank
        table.sink.deposit(splan, rows)
        px.recv_ids.fill_(-1)                 # ready for the next step (peers push again only after the next S1)
        return None, None, None


class ShardedTableEmbedding(nn.Module):
    """Row-sharded drop-in for model.layers.TableEmbedding.  `weight` holds only the local shard
    [ceil((N-rank)/world), D]; state_dict()/load_state_dict() speak the reference's full `weight` [N, D]."""

    def __init__(self, num_embeddings, embedding_dim, padding_idx=None, group=None, exchange=None):
        super().__init__()
        from .model.layers import TableGradSink
        self.group = group
        # "p2p" (peer-memory kernels over NVLink; default: r02 measured 1.03 M vs 0.84 M seq/s at N=2) | "nccl" (all_to_all)
        self.exchange = (exchange or os.environ.get("PR_EXCHANGE", "p2p")).lower()
        if self.exchange not in ("nccl", "p2p"):
            raise ValueError(f"exchange must be 'nccl' or 'p2p', got {self.exchange!r}")
        self._px = None
        self.world, self.rank = world_info(group)
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.padding_idx = padding_idx
        self.n_local = shard_rows(num_embeddings, self.world, self.rank)
        # the global padding id lives on rank (padding_idx % world) at local row padding_idx // world
        self.local_padding_idx = None
        if padding_idx is not None and padding_idx % self.world == self.rank:
            self.local_padding_idx = padding_idx // self.world
        self.weight = nn.Parameter(torch.empty(self.n_local, embedding_dim))
        self.init_normal_(0.0, 1.0)
        self.sink = TableGradSink(self)
        self.sink.sparse = True          # a sharded table has no dense-gradient mode
        self._register_state_dict_hook(self._full_on_save)
        self._register_load_state_dict_pre_hook(self._shard_on_load)

    @torch.no_grad()
    def init_normal_(self, mean=0.0, std=1.0, chunk_rows=65536):
        """Initialise the shard as rows rank::world of the LOGICAL [N, D] table drawn from the global generator
        (sasrec.py:56 draws the whole table): every rank consumes the same random stream -- the generators stay in
        lockstep and the logical table is iid and independent of the world size -- and keeps only its own rows.
        (Drawing `weight.normal_()` per rank under the common seed gave up to `world` items one embedding.)"""
        N, D, G, r = self.num_embeddings, self.embedding_dim, self.world, self.rank
        w = self.weight.data
        for r0 in range(0, N, chunk_rows):
            n = min(chunk_rows, N - r0)
            block = torch.empty(n, D, dtype=torch.float32).normal_(mean, std)       # host draw, like the reference's CPU init
            first = (r - r0) % G                                                     # first row of the block owned by this rank
            mine = block[first::G]
            if mine.shape[0]:
                lo = (r0 + first) // G
                w[lo:lo + mine.shape[0]].copy_(mine)
        return self

    def forward(self, idx):
        if self.sink.row2slot is None:
            self.sink.enable_sparse()
        fn = PeerGatherFn if self.exchange == "p2p" else ShardedGatherFn
        return fn.apply(self.weight, idx.contiguous(), self)

    def _make_plan(self, idx):
        if self.exchange == "p2p":
            if getattr(PEER, "device_plan", False) and os.environ.get("PR_P2P_PLAN", "device") == "device":
                return PeerPlanDev(idx, self.num_embeddings, self.padding_idx)
            return PeerPlan(idx, self.padding_idx)
        return ExchangePlan(idx, self.world, self.group, self.padding_idx)

    def peer_exchange(self, R):
        """Peer-mapped buffers, created at the first lookup (collective: every rank gets here in the same step)."""
        if self._px is None:
            self._px = PeerExchange(self, R)
        return self._px

    def graph_capturable(self):
        """True when a training step through this table contains no host synchronisation (peer exchange with the device plan
        and flag barriers): trainer/graph.py may capture it."""
        return (self.exchange == "p2p" and getattr(PEER, "device_plan", False)
                and os.environ.get("PR_P2P_PLAN", "device") == "device" and self._px is not None and self._px.flag_barrier)

    def exchange_status(self):
        """Host-synchronising check of the peer exchange's device flags (bit 0: bad id, bit 1: a receive region
        overflowed -- raise PR_P2P_CAP_FACTOR).  0 when the all_to_all exchange is in use."""
        return 0 if self._px is None else int(self._px.status.item())

    # ---- index-plan prefetch -----------------------------------------------------------------------------------
    # The exchange plan (unique ids, owner bucketing, split sizes, index all_to_all) depends only on the batch's
    # indices, not on the weights, and it needs a host sync for NCCL's split sizes.  Computed inline it drains the
    # GPU at the top of every step; computed one batch ahead on a side stream the sync overlaps with the previous
    # step's kernels.  Every rank must call prefetch / forward in the same order (they issue collectives).
    @staticmethod
    def _key(idx):
        return (idx.data_ptr(), tuple(idx.shape), idx._version)

    def prefetch(self, idx):
        """Build the exchange plan of a FUTURE batch on the CURRENT stream (call it inside a side-stream context,
        see trainer.Lookahead); forward() picks it up and makes its own stream wait for it."""
        if not idx.is_cuda:
            return
        idx = idx.contiguous()
        if getattr(self, "_plans", None) is None:
            self._plans = {}
        plan = self._make_plan(idx)
        plan.ready = torch.cuda.Event()
        plan.ready.record(torch.cuda.current_stream(idx.device))
        plan.keepalive = idx
        if len(self._plans) > 4:
            self._plans.clear()
        self._plans[self._key(idx)] = plan

    def take_plan(self, idx):
        plans = getattr(self, "_plans", None)
        plan = plans.pop(self._key(idx), None) if plans else None
        if plan is None:
            return self._make_plan(idx)
        torch.cuda.current_stream(idx.device).wait_event(plan.ready)
        for tns in plan.tensors():                                          # allocated on the side stream, used on this one
            tns.record_stream(torch.cuda.current_stream(idx.device))
        return plan

    @torch.no_grad()
    def lookup_static(self, idx, r_hint=0):
        """Evaluation-time lookup while NO rank is updating the table (call a barrier after the last optimizer step first): one
        pull kernel over peer memory, straight from the owners' shards -- no index plan, no barrier, no collective, so ranks may
        run different numbers of eval batches.  p2p exchange only.  r_hint: lookups per TRAINING step, used when this call is the
        one that creates the peer buffers (evaluation before the first training step)."""
        if self.exchange != "p2p":
            raise ops._lib.PixelRecB200Error("lookup_static needs the peer-memory exchange (PR_EXCHANGE=p2p)")
        idx = idx.contiguous()
        px = self.peer_exchange(max(int(r_hint), idx.numel()))
        out = PEER.gather(px.w_table, self.world, self.num_embeddings, self.embedding_dim, idx.reshape(-1))
        return out.view(*idx.shape, self.embedding_dim)

    @torch.no_grad()
    def full_weight(self):
        """All-gather the shards into the reference layout [N, D] (compute_item_all / checkpoints)."""
        return self.gather_rows_full(self.weight.detach(), clone=False)

    @torch.no_grad()
    def gather_rows_full(self, local, clone=True):
        """[n_local, D] per-rank tensor aligned with the shard (the weights, an Adam moment) -> the full [N, D] layout the
        reference's checkpoints use.  Collective: every rank calls it."""
        if self.world == 1:
            return local.clone() if clone else local
        n_max = shard_rows(self.num_embeddings, self.world, 0)
        pad = torch.zeros(n_max, self.embedding_dim, dtype=local.dtype, device=local.device)
        pad[:self.n_local] = local
        chunks = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(chunks, pad, group=self.group)
        allw = torch.stack(chunks)                                   # [world, n_max, D]; row i = shard i%world, i//world
        return allw.permute(1, 0, 2).reshape(-1, self.embedding_dim)[:self.num_embeddings].contiguous()

    def shard_of_full(self, full):
        """rows rank::world of a full [N, D] tensor (inverse of gather_rows_full)"""
        return full[self.rank::self.world].contiguous()

    @staticmethod
    def _full_on_save(module, state_dict, prefix, local_metadata):
        state_dict[prefix + "weight"] = module.full_weight()

    def _shard_on_load(self, state_dict, prefix, *args):
        key = prefix + "weight"
        if key in state_dict and state_dict[key].shape[0] == self.num_embeddings:
            state_dict[key] = state_dict[key][self.rank::self.world].contiguous()

    def extra_repr(self):
        return f"{self.num_embeddings} rows sharded {self.world}-way (local {self.n_local}), dim {self.embedding_dim}"


def make_table(num_embeddings, embedding_dim, padding_idx=None, sharding="auto"):
    """TableEmbedding on one GPU, ShardedTableEmbedding when a process group with world > 1 is up."""
    from .model.layers import TableEmbedding
    world, _ = world_info()
    if world > 1 and sharding != "replicated":
        return ShardedTableEmbedding(num_embeddings, embedding_dim, padding_idx)
    return TableEmbedding(num_embeddings, embedding_dim, padding_idx)


@torch.no_grad()
def broadcast_dense_params(model, src=0):
    """Replicated (non-table) parameters start identical on every rank (DDP does the same at wrap time)."""
    if world_info()[0] == 1:
        return
    sharded = {id(m.weight) for m in model.modules() if isinstance(m, ShardedTableEmbedding)}
    for p in model.parameters():
        if id(p) not in sharded:
            dist.broadcast(p.data, src)


# ----------------------------------------------------------------------------------------------- sharded evaluation
def merge_topk_candidates(val, idx, k):
    """[B, C] candidate scores / global item ids -> the k best per row, ordered by (score descending, id ascending): the order
    torch.topk over the full score row produces for distinct scores, with the exact kernels' tie rule (lower id first)."""
    by_id = torch.argsort(idx, dim=1, stable=True)
    v1, i1 = torch.gather(val, 1, by_id), torch.gather(idx, 1, by_id)
    by_val = torch.argsort(v1, dim=1, descending=True, stable=True)[:, :k]
    return torch.gather(v1, 1, by_val), torch.gather(i1, 1, by_val)


class ShardedTopK:
    """Full-catalog ranking with the item table LEFT row-sharded ('preferred for C3'; replaces
    REC/trainer/trainer.py:339-358 compute_item_feature + :327-337 + evaluator/collector.py:133 at N > 1): instead of
    all-gathering the [N, D] table on every rank (3.35 GB per rank at C3), every eval batch
      1. all-gathers the ranks' encoder outputs        [G, B_e, D]   (16 MB at C2 / 8 GPUs)
      2. all-gathers the (user, item) history pairs and keeps those whose item this rank owns (item % G == rank)
      3. scores ALL G*B_e users against the LOCAL shard with the fused tcgen05 scoring + exact top-k   -> [G*B_e, k] candidates
      4. all-to-alls the candidates back to the users' ranks and merges the G lists                   -> [B_e, k] global ids.
    Per-shard candidates carry fp32 scores computed by the same kernel whatever the rank, so the merged ids equal the
    single-GPU pr_score_topk_exact_f32 ranking bit for bit.  Collective: every rank calls it once per step with the same
    B_e (pad the last / missing batches; `n_valid` users of this rank's batch are real)."""

    def __init__(self, world, rank, group=None, score_fn=None):
        self.world, self.rank, self.group = world, rank, group
        self.score_fn = score_fn or self._score_cuda
        self._norm_max = None

    def reset(self):
        self._norm_max = None                         # call when the table changed (once per evaluation)

    def _score_cuda(self, seq_all, W_local, k, hu, hi, mask_col0):
        if self._norm_max is None:
            self._norm_max = ops.table_norm_max(W_local)
        val, idx, _ = ops.score_topk_exact(seq_all, W_local, k, hu, hi, mask_col0=mask_col0, w_norm_max=self._norm_max)
        return val, idx

    def __call__(self, seq_out, W_local, k, hist_u=None, hist_i=None, pad_id=0):
        G, rank = self.world, self.rank
        B_e, D = seq_out.shape
        dev = seq_out.device
        # 1. encoder outputs of every rank
        seq_all = torch.empty(G * B_e, D, device=dev, dtype=seq_out.dtype)
        dist.all_gather_into_tensor(seq_all, seq_out.contiguous(), group=self.group)
        # 2. history pairs of every rank (variable length: sizes first), user index shifted into the gathered batch
        n = 0 if hist_u is None else int(hist_u.numel())
        sizes = torch.tensor([n], dtype=torch.int64, device=dev)
        all_sizes = torch.empty(G, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_sizes, sizes, group=self.group)
        all_sizes = all_sizes.tolist()
        n_max = max(all_sizes)
        hu = hi = None
        if n_max > 0:
            mine = torch.zeros(2, n_max, dtype=torch.int64, device=dev)
            if n:
                mine[0, :n] = hist_u.to(dev)
                mine[1, :n] = hist_i.to(dev)
            everyone = torch.empty(G, 2, n_max, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(everyone.view(-1), mine.view(-1), group=self.group)
            valid = torch.arange(n_max, device=dev)[None, :] < torch.tensor(all_sizes, device=dev)[:, None]     # [G, n_max]
            gu = everyone[:, 0] + (torch.arange(G, device=dev) * B_e)[:, None]
            gi = everyone[:, 1]
            keep = valid & (gi % G == rank)
            hu, hi = gu[keep].contiguous(), torch.div(gi[keep], G, rounding_mode="floor").contiguous()
            if hu.numel() == 0:
                hu = hi = None
        # 3. all users against the local shard; the padding item lives on rank pad_id % G at local row pad_id // G == 0
        mask_col0 = pad_id is not None and pad_id % G == rank
        if mask_col0 and pad_id // G != 0:
            raise ValueError("ShardedTopK: the padding id must be the first row of its shard")
        val, idx_local = self.score_fn(seq_all, W_local, k, hu, hi, mask_col0)
        idx = idx_local.to(torch.int64) * G + rank
        # 4. candidates of rank s's users go to rank s
        val_in, idx_in = torch.empty_like(val), torch.empty_like(idx)
        dist.all_to_all_single(val_in, val.contiguous(), group=self.group)
        dist.all_to_all_single(idx_in, idx.contiguous(), group=self.group)
        cand_v = val_in.view(G, B_e, k).permute(1, 0, 2).reshape(B_e, G * k)
        cand_i = idx_in.view(G, B_e, k).permute(1, 0, 2).reshape(B_e, G * k)
        return merge_topk_candidates(cand_v, cand_i, k)
