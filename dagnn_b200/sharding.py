"""Graph-sharded data parallelism for the level sweep: one process per GPU (torch.distributed, NCCL over
NVLink on the box; gloo in the CPU tests).

Replaces the reference's single-process `DataParallel` (ogbg-code/tg/data_parallel.py:41-82 with the node-balanced
`Collater`, ogbg-code/tg/dataloader.py:17-33; D-VAE: dvae/util.py:667-683): graphs of a batch are independent
connected components, so the forward needs NO collective — each rank sweeps its own contiguous, node-balanced
range of graphs and keeps its readout rows. Training needs exactly one all-reduce (sum) of the flat fp32
gradient buffer per step (`allreduce_gradients`), which is the only data-path collective of this package.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .data import DagBatch, graph_depths, graph_node_counts, select_graphs, shard_graph_ids


def rank_world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_for_rank(B: DagBatch, rank: Optional[int] = None, world_size: Optional[int] = None, depth_aware: bool = True):
    """-> (sub-batch of this rank or None when the rank owns no graph, ascending global graph ids it owns). Every rank
    computes the same partition from the node counts and graph depths (deterministic, no communication): depth-aware by
    default (`data.shard_graph_ids`: the deepest graphs go to different ranks and take fewer nodes with them), the reference's
    contiguous node-balanced rule (ogbg-code/tg/dataloader.py:17-27) with depth_aware=False."""
    r, w = rank_world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    ids = shard_graph_ids(graph_node_counts(B), world_size, graph_depths(B) if depth_aware else None)
    mine = ids[rank]
    return (select_graphs(B, mine) if len(mine) else None), mine


def unshard_rows(rows: torch.Tensor, ids_per_rank: Sequence[Sequence[int]]) -> torch.Tensor:
    """Rows gathered in rank order (`gather_rows`) -> rows in global graph order."""
    order = torch.as_tensor(np.concatenate([np.asarray(i, dtype=np.int64) for i in ids_per_rank]), device=rows.device)
    out = torch.empty_like(rows)
    out[order] = rows
    return out


def gather_rows(local: torch.Tensor, rows_per_rank: Sequence[int], group=None) -> torch.Tensor:
    """all-gather variable-length row blocks (readouts / predictions) in rank order — used by evaluation and the
    tests; not part of the timed forward."""
    _, w = rank_world(group)
    if w == 1:
        return local
    width = local.shape[1]
    mx = max(rows_per_rank)
    pad = local.new_zeros(mx, width)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:n] for b, n in zip(bufs, rows_per_rank)])


def allreduce_gradients(params, group=None, average: bool = False) -> int:
    """One all-reduce(sum) over a flat fp32 buffer of every gradient (the reference reduce-adds replicas'
    gradients onto GPU 0 inside nn.DataParallel's backward, tg/data_parallel.py:59-62). Returns the number of
    elements reduced. Loss is a mean over graphs, so callers scale by shard size / global batch before backward
    (or pass average=True for equal shards)."""
    r, w = rank_world(group)
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(w)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off: off + n].view_as(g))
        off += n
    return int(flat.numel())


class FlatGradients(object):
    """One flat fp32 buffer holding every parameter gradient: `param.grad` of each parameter is a view into it (autograd
    accumulates in place), so the per-step gradient exchange is exactly ONE collective over the buffer — no gather into a
    temporary, no copy back. Replaces the reduce-add onto GPU 0 of the reference's DataParallel backward
    (ogbg-code/tg/data_parallel.py:59-62). Use with `optimizer.zero_grad(set_to_none=False)` (or `zero()` here)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off: off + n].view_as(p)
            off += n

    def zero(self):
        self.flat.zero_()

    def allreduce(self, group=None, average: bool = False) -> int:
        """all-reduce(sum) of the buffer over the ranks (NCCL over NVLink on the box); returns the number of elements."""
        _, w = rank_world(group)
        if w > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.flat.div_(w)
        return self.numel
