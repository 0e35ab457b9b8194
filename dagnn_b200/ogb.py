"""OGB flavour of DAGNN — drop-in for `ogbg-code/model/dagnn.py` (class DAGNN, :16-215) and the
`ASTNodeEncoder` of `ogbg-code/utils.py:7-28`, with the level sweep running in libdagnn_sm100.so.

Same constructor signature, same parameter names/shapes (checkpoints of the reference load with
`load_state_dict`), same `forward(G)` contract: `G` carries `x int64[N,2]`, `node_depth int64[N,1]`,
`edge_index int64[2,E]`, `edge_attr fp32[E,2]`, `batch int64[N]`, `_bi_layer_idx0/1`, `_bi_layer_index0/1`
on the CUDA device; returns a list of `max_seq_len` tensors `[B, num_vocab]` (or `[B, num_class]`).

Scope: the aggregators `agg="attn_h"` (default) and `"self_attn_h"` with GRU cells (`recurr=1`), uni/bidirectional,
`out_wx`, `out_pool_all`, `out_pool` in {max, mean, add, attn}. Other aggregators / `agg_x` / `recurr=0` raise
NotImplementedError at construction (§8f row 4), they never fall back to eager torch.
Training: with autograd enabled the forward runs through `dagnn_b200.autograd` (EmbedFn, SweepReadoutFn, LinearFn), whose
backward is the library's reverse-level BPTT — `loss.backward()` works as in main_pyg.py:55-65.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import autograd as ag
from . import runtime as rt

NA_ATTN_H = "attn_h"
NA_SELF_ATTN_H = "self_attn_h"
P_MAX, P_MEAN, P_ADD, P_ATTN = "max", "mean", "add", "attn"


class ASTNodeEncoder(nn.Module):
    """ogbg-code/utils.py:7-28 — three embedding tables summed; the gather-sum runs in `dagnn_embed_f32`."""

    def __init__(self, emb_dim, num_nodetypes, num_nodeattributes, max_depth):
        super().__init__()
        self.max_depth = max_depth
        self.type_encoder = nn.Embedding(num_nodetypes, emb_dim)
        self.attribute_encoder = nn.Embedding(num_nodeattributes, emb_dim)
        self.depth_encoder = nn.Embedding(self.max_depth + 1, emb_dim)

    def forward(self, x, depth):
        T, A, P = self.type_encoder.weight, self.attribute_encoder.weight, self.depth_encoder.weight
        if torch.is_grad_enabled() and (T.requires_grad or A.requires_grad or P.requires_grad):
            return ag.EmbedFn.apply(x, depth, T, A, P, self.max_depth)
        return rt.embed(x, depth, T, A, P, self.max_depth)


class AttnConv(nn.Module):
    """Parameter container with the names of the reference's AttnConv (dagnn.py:347-359). The message
    passing itself is fused into the level kernel."""

    def __init__(self, attn_q_dim, emb_dim, attn_dim=0, num_relations=1, reverse=False):
        super().__init__()
        assert attn_q_dim > 0 and emb_dim > 0
        attn_dim = attn_dim if attn_dim > 0 else emb_dim
        self.reverse = reverse
        self.wea = num_relations > 1
        if self.wea:
            self.edge_encoder = nn.Linear(num_relations, attn_dim)
        self.attn_lin = nn.Linear(attn_q_dim + attn_dim, 1)


class SelfAttnConv(nn.Module):
    """Parameter container with the names of the reference's SelfAttnConv (dagnn.py:279-313): the same additive attention
    without a query part — score = attn_lin(h_j + edge embedding). In the level kernel this is AttnConv with Dq = 0."""

    def __init__(self, emb_dim, attn_dim=0, num_relations=1, reverse=False):
        super().__init__()
        assert emb_dim > 0
        attn_dim = attn_dim if attn_dim > 0 else emb_dim
        self.reverse = reverse
        self.wea = num_relations > 1
        if self.wea:
            self.edge_encoder = nn.Linear(num_relations, attn_dim)
        self.attn_lin = nn.Linear(attn_dim, 1)


class _PackedCacheMixin(object):
    """Drops the packed-parameter cache whenever parameters may have been rewritten behind autograd's version counters."""

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._packed.invalidate()
        return out

    def train(self, mode: bool = True):
        self._packed.invalidate()
        return super().train(mode)


def _needs_grad(module: nn.Module) -> bool:
    return torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters())


class DAGNN(_PackedCacheMixin, nn.Module):

    def __init__(self, num_vocab, max_seq_len, emb_dim, hidden_dim, out_dim,
                 num_rels=2, w_edge_attr=True, num_layers=2, bidirectional=True, mapper_bias=True,
                 agg_x=False, agg=NA_ATTN_H, out_wx=True, out_pool_all=True, out_pool=P_MAX, encoder=None, dropout=0.0,
                 word_vectors=None, emb_dims=[], activation=None, num_class=0, recurr=1):
        super().__init__()
        self.num_class = num_class
        self.num_vocab = num_vocab
        self.max_seq_len = max_seq_len
        if agg_x and hidden_dim < emb_dim:
            raise ValueError('Hidden dimension too small for input.')     # dagnn.py:27-28
        if agg not in (NA_ATTN_H, NA_SELF_ATTN_H) or agg_x or not recurr:
            raise NotImplementedError("dagnn_b200 covers agg='attn_h' / 'self_attn_h', agg_x=False, recurr=1 (SURVEY.md §8f row 4); "
                                      "got agg=%r agg_x=%r recurr=%r" % (agg, agg_x, recurr))
        if out_pool not in (P_MAX, P_MEAN, P_ADD, P_ATTN):
            raise ValueError("out_pool=%r (max / mean / add / attn)" % (out_pool,))
        if encoder is None:
            raise NotImplementedError("pass encoder=ASTNodeEncoder(...) (main_pyg.py:248,396-401); EmbeddingBag "
                                      "encoders (init_encoder, dagnn.py:218-223) are not covered")
        self.agg_x, self.agg_attn, self.agg_attn_x = False, True, False
        self.bidirectional = bidirectional
        self.dirs = [0, 1] if bidirectional else [0]
        self.num_layers = num_layers
        self.out_wx = out_wx
        self.output_all = out_pool_all
        self.out_pool = out_pool
        self.recurr = recurr
        self.emb_dim = emb_dim
        self.hidden_dim = hidden_dim
        nd = len(self.dirs)
        self.out_hidden_dim = emb_dim * nd + hidden_dim * nd * num_layers if out_wx else hidden_dim * nd * num_layers
        self.encoder = encoder
        self.w_edge_attr = bool(w_edge_attr)
        num_rels = num_rels if w_edge_attr else 1
        if num_rels not in (1, 2):
            raise NotImplementedError("edge_attr must have 2 columns (utils2.py:45,68)")
        # both aggregator lists exist even when unidirectional, like the reference (dagnn.py:64-67)
        if agg == NA_SELF_ATTN_H:          # dagnn.py:56-61
            self.node_aggr_0 = nn.ModuleList([SelfAttnConv(hidden_dim, num_relations=num_rels) for _ in range(num_layers)])
            self.node_aggr_1 = nn.ModuleList([SelfAttnConv(hidden_dim, num_relations=num_rels, reverse=True) for _ in range(num_layers)])
        else:                              # dagnn.py:62-67
            self.node_aggr_0 = nn.ModuleList([AttnConv(emb_dim if l == 0 else hidden_dim, hidden_dim, num_relations=num_rels,
                                                       attn_dim=hidden_dim) for l in range(num_layers)])
            self.node_aggr_1 = nn.ModuleList([AttnConv(emb_dim if l == 0 else hidden_dim, hidden_dim, num_relations=num_rels,
                                                       attn_dim=hidden_dim, reverse=True) for l in range(num_layers)])
        for i in self.dirs:
            setattr(self, "cells_{}".format(i), nn.ModuleList(
                [nn.GRUCell(emb_dim if l == 0 else hidden_dim, hidden_dim) for l in range(num_layers)]))
        if out_pool == P_ATTN:
            # dagnn.py:88-91,114-117: Linear(d, 1) scores softmaxed over a dimension of size 1 — every weight is exactly 1, the
            # readout is a plain add-pool. The layer exists for checkpoint compatibility; its gradient is zero.
            d_ = int(self.out_hidden_dim / 2) if self.bidirectional and not self.output_all else self.out_hidden_dim
            self.self_attn_linear_out = nn.Linear(d_, 1)
        self.dropout = nn.Dropout(dropout)
        if self.num_class > 0:
            self.graph_pred_linear = nn.Linear(self.out_hidden_dim, self.num_class)
        else:
            self.graph_pred_linear_list = nn.ModuleList()
            if self.num_vocab == 1:
                self.graph_pred_linear_list.append(nn.Sequential(nn.Linear(self.out_hidden_dim, self.num_vocab), nn.ReLU()))
            else:
                for _ in range(max_seq_len):
                    self.graph_pred_linear_list.append(nn.Linear(self.out_hidden_dim, self.num_vocab))
        self._packed = rt.PackedParams()

    # ------------------------------------------------------------------ pieces of forward
    def _num_graphs(self, G) -> int:
        ng = getattr(G, "num_graphs", None)
        return int(ng) if ng is not None else int(G.batch[-1].item()) + 1

    def build_schedule(self, G, max_levels: int = 256) -> rt.Schedule:
        lv = [G._bi_layer_idx0, G._bi_layer_idx1][:len(self.dirs)]
        ids = [G._bi_layer_index0, G._bi_layer_index1][:len(self.dirs)]
        ea = G.edge_attr if self.w_edge_attr else None
        return rt.Schedule.build(G.edge_index, lv, ids, ea, G.batch, self._num_graphs(G), max_levels)

    def _pack(self, device) -> rt.PackedParams:
        cells = [getattr(self, "cells_%d" % d) for d in self.dirs]
        aggrs = [getattr(self, "node_aggr_%d" % d) for d in self.dirs]
        return self._packed.update(cells, aggrs, self.emb_dim, self.hidden_dim, 0, self.w_edge_attr, device)

    def node_states(self, G, sched=None, max_levels: int = 256):
        """encoder + level sweep: returns (X [N,D], Hs [dirs, layers, N, ldh] in position order, schedule).
        Asynchronous and unchecked: call `sched.finalize()` (or go through forward_readout) to validate."""
        X = self.encoder(G.x, G.node_depth.view(-1, ))
        if X.shape[1] != self.emb_dim:
            raise ValueError("encoder produced width %d, emb_dim is %d" % (X.shape[1], self.emb_dim))
        sched = sched if sched is not None else self.build_schedule(G, max_levels)
        packed = self._pack(X.device)
        Hs = rt.sweep(sched, X, packed, self.emb_dim, self.hidden_dim, self.num_layers, 0, self.w_edge_attr)
        return X, Hs, sched

    # ------------------------------------------------------------------ hooks of autograd.SweepReadoutFn
    def _sweep_dims(self):
        return self.emb_dim, self.hidden_dim, self.num_layers, 0, self.w_edge_attr

    def _cell_params(self):
        out = []
        for d in self.dirs:
            for i in range(self.num_layers):
                cell, ag_ = getattr(self, "cells_%d" % d)[i], getattr(self, "node_aggr_%d" % d)[i]
                out += [cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh, ag_.attn_lin.weight, ag_.attn_lin.bias]
                if self.w_edge_attr:
                    out += [ag_.edge_encoder.weight, ag_.edge_encoder.bias]
        return out

    def _readout_blocks(self, G, X, Hs):
        """dagnn.py:184-202 as a list of pooled column blocks."""
        H, Lr, blocks = self.hidden_dim, self.num_layers, []
        lvl = [G._bi_layer_idx0, G._bi_layer_idx1]
        col = 0
        if self.bidirectional and not self.output_all:
            for d in (0, 1):
                filt = dict(filter=rt.FILTER_LVL0, filter_lvl=lvl[1 - d])   # d=0: sinks, d=1: sources (:119-126)
                if self.out_wx:
                    blocks.append(dict(src=X, width=self.emb_dim, index_mode=0, out_col=col, **filt)); col += self.emb_dim
                for l in range(Lr):
                    blocks.append(dict(src=Hs[d, l], width=H, index_mode=1, dir=d, out_col=col, **filt)); col += H
        else:
            filt = dict(filter=rt.FILTER_ALL) if self.output_all else dict(filter=rt.FILTER_LVL0, filter_lvl=lvl[1])
            if self.out_wx:
                blocks.append(dict(src=X, width=self.emb_dim, index_mode=0, out_col=col, **filt)); col += self.emb_dim
            for d in self.dirs:
                for l in range(Lr):
                    blocks.append(dict(src=Hs[d, l], width=H, index_mode=1, dir=d, out_col=col, **filt)); col += H
        return blocks, (P_ADD if self.out_pool == P_ATTN else self.out_pool), col

    def readout(self, G, X, Hs, sched) -> torch.Tensor:
        """dagnn.py:184-202."""
        blocks, pool, width = self._readout_blocks(G, X, Hs)
        return rt.readout(sched, blocks, pool, width, X.device)

    def forward_readout(self, G):
        """The north-star hot path: encoder -> schedule -> level sweeps (both directions) -> pooled readout.
        A batch without nodes (a rank that owns no graph) gives an empty [0, out_hidden_dim] readout, like the reference's
        collater, which drops empty shards (tg/dataloader.py:29-31)."""
        if G.x.shape[0] == 0:
            return torch.zeros(0, self.out_hidden_dim, device=G.x.device, dtype=torch.float32)
        if _needs_grad(self):            # training: the same stages as autograd Functions (backward = reverse-level BPTT)
            X = self.encoder(G.x, G.node_depth.view(-1, ))
            if X.shape[1] != self.emb_dim:
                raise ValueError("encoder produced width %d, emb_dim is %d" % (X.shape[1], self.emb_dim))
            return ag.SweepReadoutFn.apply(self, G, X, *self._cell_params())

        def run(max_levels):
            X, Hs, sched = self.node_states(G, None, max_levels)
            return self.readout(G, X, Hs, sched), sched
        return rt.run_checked(run)

    def _head(self, layer, out):
        """nn.Linear (or Sequential(Linear, ReLU) when num_vocab == 1, dagnn.py:106-108) through dagnn_linear_f32."""
        if isinstance(layer, nn.Sequential):
            return layer[1](ag.linear(out, layer[0]))
        return ag.linear(out, layer)

    def forward(self, G):
        out = self.forward_readout(G)
        out = self.dropout(out)
        if self.num_class > 0:
            return self._head(self.graph_pred_linear, out)
        return [self._head(self.graph_pred_linear_list[i], out) for i in range(self.max_seq_len)]
