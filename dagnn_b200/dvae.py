"""D-VAE flavours of DAGNN — drop-in for `dvae/dagnn.py` (class DAGNN, NA / ENAS graphs, :18-184) and
`dvae/dagnn_bn.py` (class DAGNN_BN, Bayesian networks, :19-177): same constructors, same parameter names and
shapes as the reference (incl. the decoder parameters of `DVAE_PYG`, dvae/models_pyg.py:17-79, so reference
checkpoints load), `forward(G) -> [B, hs]` and `encode(list_of_graphs) -> (mu, logvar)` running the level
sweep in libdagnn_sm100.so.

Scope: `agg="attn_h"`, `out_pool_all=False` (last-/first-node readout). The teacher-forced
decoder (`loss`, `decode`, `_ipropagate_to`) is §8f row 2 and raises NotImplementedError.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import autograd as ag
from . import runtime as rt
from .data import DagBatch, collate_dvae
from .ogb import AttnConv, NA_ATTN_H, P_MAX, _PackedCacheMixin, _needs_grad


class _DVAEParams(nn.Module):
    """Every parameter of DVAE_PYG (dvae/models_pyg.py:17-79) under the reference's names."""

    def __init__(self, max_n, nvt, START_TYPE, END_TYPE, hs=501, nz=56, bidirectional=False, vid=True, num_layers=1,
                 bn_variant=False):
        super().__init__()
        self.max_n, self.nvt, self.START_TYPE, self.END_TYPE = max_n, nvt, START_TYPE, END_TYPE
        self.hs, self.nz, self.gs = hs, nz, hs
        self.bidir, self.vid = bidirectional, vid
        self.device = None
        self.vs = hs + max_n if vid else hs
        self.num_layers = num_layers
        cells = lambda: nn.ModuleList([nn.GRUCell(nvt, hs) if l == 0 else nn.GRUCell(hs, hs) for l in range(num_layers)])
        self.grue_forward = cells()
        self.grue_backward = cells()
        self.fc1 = nn.Linear(self.gs, nz)
        self.fc2 = nn.Linear(self.gs, nz)
        self.grud = cells()
        self.fc3 = nn.Linear(nz, hs)
        self.add_vertex = nn.Sequential(nn.Linear(hs, hs * 2), nn.ReLU(), nn.Linear(hs * 2, nvt))
        self.add_edge = nn.Sequential(nn.Linear(hs * 2, hs * 4), nn.ReLU(), nn.Linear(hs * 4, 1))
        self.gate_forward = nn.ModuleList([nn.Sequential(nn.Linear(self.vs, hs), nn.Sigmoid()) for _ in range(num_layers)])
        self.gate_backward = nn.ModuleList([nn.Sequential(nn.Linear(self.vs, hs), nn.Sigmoid()) for _ in range(num_layers)])
        self.mapper_forward = nn.ModuleList([nn.Sequential(nn.Linear(self.vs, hs, bias=False)) for _ in range(num_layers)])
        self.mapper_backward = nn.ModuleList([nn.Sequential(nn.Linear(self.vs, hs, bias=False)) for _ in range(num_layers)])
        if self.bidir:
            self.hv_unify = nn.Sequential(nn.Linear(hs * 2, hs))
            self.hg_unify = nn.Sequential(nn.Linear(self.gs * 2 * num_layers, self.gs))
        if bn_variant:   # DVAE_BN_PYG with aggx=0 (dvae/models_pyg.py:539-560, dagnn_bn.py:25)
            w = lambda l: nvt if l == 0 else hs
            self.mapper_forward = nn.ModuleList([nn.Sequential(nn.Linear(w(l), hs, bias=False)) for l in range(num_layers)])
            self.mapper_backward = nn.ModuleList([nn.Sequential(nn.Linear(w(l), hs, bias=False)) for l in range(num_layers)])
            self.gate_forward = nn.ModuleList([nn.Sequential(nn.Linear(w(l), hs), nn.Sigmoid()) for l in range(num_layers)])
            self.gate_backward = nn.ModuleList([nn.Sequential(nn.Linear(w(l), hs), nn.Sigmoid()) for l in range(num_layers)])
            self.add_edge = nn.Sequential(nn.Linear(hs * 3, hs), nn.ReLU(), nn.Linear(hs, 1))

    def get_device(self):
        if self.device is None:
            self.device = next(self.parameters()).device
        return self.device

    def reparameterize(self, mu, logvar, eps_scale=0.01):
        """dvae/models_pyg.py:324-331."""
        if self.training:
            std = logvar.mul(0.5).exp_()
            eps = torch.randn_like(std) * eps_scale
            return eps.mul(std).add_(mu)
        return mu

    def loss(self, *a, **k):
        raise NotImplementedError("teacher-forced decode (dvae/models_pyg.py:398-456) is SURVEY.md §8f row 2")

    def decode(self, *a, **k):
        raise NotImplementedError("sampling decode (dvae/models_pyg.py:338-396) is SURVEY.md §8f row 2")


class _DagnnDvaeBase(_PackedCacheMixin, _DVAEParams):
    _VID = True   # NA: one-hot vertex ids on keys (dvae/dagnn.py:130-139); BN: none

    def _init_dagnn(self, emb_dim, hidden_dim, out_dim, num_layers, bidirectional, agg, out_wx, out_pool_all, out_pool,
                    dropout, num_nodes):
        if agg != NA_ATTN_H:
            raise NotImplementedError("dagnn_b200 covers agg='attn_h' (SURVEY.md §8f row 4); got %r" % (agg,))
        if out_wx:
            raise NotImplementedError("out_wx: the reference concatenates G.x in front of the states and then feeds a layer sized "
                                      "for the states alone (dvae/dagnn.py:163-168) — a shape error there, nothing to reproduce")
        if out_pool_all and out_pool not in ("max", "mean", "add"):
            raise NotImplementedError("out_pool=%r: the D-VAE files never define the attention readout (dvae/dagnn.py:86)" % (out_pool,))
        if hidden_dim != self.hs:
            raise ValueError("hidden_dim must equal hs (the GRU cells are grue_forward/backward, dagnn.py:73-75)")
        if emb_dim != self.nvt:
            raise ValueError("emb_dim must equal nvt: the first GRU cell is GRUCell(nvt, hs) and reads G.x (dvae/models_pyg.py:37, "
                             "dvae/dagnn.py:140-144)")
        self.num_nodes = num_nodes
        self.agg, self.agg_attn, self.agg_attn_x = agg, True, False
        self.bidirectional = bidirectional
        self.dirs = [0, 1] if bidirectional else [0]
        self.out_wx, self.output_all, self.out_pool = out_wx, out_pool_all, out_pool
        self.emb_dim, self.hidden_dim = emb_dim, hidden_dim
        self.out_hidden_dim = hidden_dim * num_layers
        nv = num_nodes if self._VID else 0
        attn_dim = hidden_dim + nv
        self.node_aggr_0 = nn.ModuleList([AttnConv(emb_dim if l == 0 else attn_dim, attn_dim, num_relations=1,
                                                   attn_dim=attn_dim) for l in range(num_layers)])
        self.node_aggr_1 = nn.ModuleList([AttnConv(emb_dim if l == 0 else attn_dim, attn_dim, num_relations=1,
                                                   attn_dim=attn_dim, reverse=True) for l in range(num_layers)])
        self.cells_0 = self.grue_forward                # aliases, like dagnn.py:73-75
        if bidirectional:
            self.cells_1 = self.grue_backward
        self.dropout = nn.Dropout(dropout)
        self.out_linear = nn.Linear(self.out_hidden_dim, out_dim) if num_layers > 1 else None
        self._packed = rt.PackedParams()

    def build_schedule(self, G, max_levels: int = 256) -> rt.Schedule:
        bi = G.bi_layer_index
        nd = len(self.dirs)
        lv = [bi[d][0] for d in range(nd)]
        ids = [bi[d][1] for d in range(nd)]
        ng = getattr(G, "num_graphs", None)
        ng = int(ng) if ng is not None else int(G.batch[-1].item()) + 1
        return rt.Schedule.build(G.edge_index, lv, ids, None, G.batch, ng, max_levels)

    # ------------------------------------------------------------------ hooks of autograd.SweepReadoutFn
    def _sweep_dims(self):
        return self.emb_dim, self.hidden_dim, self.num_layers, (self.num_nodes if self._VID else 0), False

    def _cell_params(self):
        out = []
        for d in self.dirs:
            for i in range(self.num_layers):
                cell, ag_ = getattr(self, "cells_%d" % d)[i], getattr(self, "node_aggr_%d" % d)[i]
                out += [cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh, ag_.attn_lin.weight, ag_.attn_lin.bias]
        return out

    def _pack(self, device) -> rt.PackedParams:
        cells = [getattr(self, "cells_%d" % d) for d in self.dirs]
        aggrs = [getattr(self, "node_aggr_%d" % d) for d in self.dirs]
        nv = self.num_nodes if self._VID else 0
        return self._packed.update(cells, aggrs, self.emb_dim, self.hidden_dim, nv, False, device)

    def _readout_blocks(self, G, X, Hs):
        """last node of every graph (forward states) [|| first node (backward states)] over all layers (dvae/dagnn.py:147-161)."""
        H, blocks, col = self.hidden_dim, [], 0
        for l in range(self.num_layers):
            blocks.append(dict(src=Hs[0, l], width=H, index_mode=1, dir=0, filter=rt.FILTER_LAST, out_col=col)); col += H
        if self.bidirectional:
            for l in range(self.num_layers):
                blocks.append(dict(src=Hs[1, l], width=H, index_mode=1, dir=1, filter=rt.FILTER_FIRST, out_col=col)); col += H
        return blocks, "add", col

    def node_states(self, G, sched=None, max_levels: int = 256):
        sched = sched if sched is not None else self.build_schedule(G, max_levels)
        nv = self.num_nodes if self._VID else 0
        packed = self._pack(G.x.device)
        X = G.x.float().contiguous()
        Hs = rt.sweep(sched, X, packed, self.emb_dim, self.hidden_dim, self.num_layers, nv, False)
        return X, Hs, sched

    def forward(self, G):
        """dvae/dagnn.py:99-175 / dvae/dagnn_bn.py:98-168 with out_pool_all=False: last node of every graph
        (forward states) [‖ first node (backward states)] over all layers -> out_linear / hg_unify."""
        G = G.to(self.get_device())
        if self.output_all:
            return self._forward_pool_all(G)
        if _needs_grad(self):
            hcat = ag.SweepReadoutFn.apply(self, G, G.x.float().contiguous(), *self._cell_params())
        else:
            def run(max_levels):
                X, Hs, sched = self.node_states(G, None, max_levels)
                blocks, pool, width = self._readout_blocks(G, X, Hs)
                return rt.readout(sched, blocks, pool, width, X.device), sched
            hcat = rt.run_checked(run)
        if self.bidirectional:
            return ag.linear(hcat, self.hg_unify[0])
        return ag.linear(hcat, self.out_linear) if self.num_layers > 1 else hcat

    def _forward_pool_all(self, G):
        """out_pool_all=True (dvae/dagnn.py:162-172): the states of every node, all directions and layers side by side in node
        order -> hg_unify / out_linear PER NODE -> pooled over all nodes of a graph. Forward only."""
        if _needs_grad(self):
            raise NotImplementedError("out_pool_all=True is covered forward-only: call under torch.no_grad()")

        def run(max_levels):
            X, Hs, sched = self.node_states(G, None, max_levels)
            N, H, nd, nl = int(X.shape[0]), self.hidden_dim, len(self.dirs), self.num_layers
            cat = torch.empty(N, nd * nl * H, device=X.device, dtype=torch.float32)
            for d in range(nd):
                for l in range(nl):
                    col = (d * nl + l) * H
                    rt.check(rt.lib().dagnn_states_to_node_order_f32(rt.C.byref(sched.c), d, Hs[d, l].data_ptr(), Hs.stride(2), H,
                                                                     cat.data_ptr() + 4 * col, cat.stride(0), rt._stream()),
                             "dagnn_states_to_node_order_f32")
            if self.bidirectional:
                per_node = ag.linear(cat, self.hg_unify[0])
            elif nl > 1:
                per_node = ag.linear(cat, self.out_linear)
            else:
                per_node = cat
            blocks = [dict(src=per_node, width=int(per_node.shape[1]), index_mode=0, filter=rt.FILTER_ALL, out_col=0)]
            return rt.readout(sched, blocks, self.out_pool, int(per_node.shape[1]), X.device), sched
        return rt.run_checked(run)

    def encode(self, G):
        """dvae/dagnn.py:177-184: list of graphs -> (mu, logvar)."""
        if type(G) != list:
            G = [G]
        b = G[0] if (len(G) == 1 and hasattr(G[0], "batch")) else collate_dvae(G)
        Hg = self(b)
        return ag.linear(Hg, self.fc1), ag.linear(Hg, self.fc2)

    def _collate_fn(self, G):
        return [g.clone() if isinstance(g, DagBatch) else g for g in G]


class DAGNN(_DagnnDvaeBase):
    """dvae/dagnn.py:18 (NA)."""
    _VID = True

    def __init__(self, emb_dim, hidden_dim, out_dim, max_n, nvt, START_TYPE, END_TYPE, hs, nz,
                 num_layers=2, bidirectional=False, agg=NA_ATTN_H, out_wx=False, out_pool_all=False, out_pool=P_MAX,
                 dropout=0.0, num_nodes=8):
        super().__init__(max_n, nvt, START_TYPE, END_TYPE, hs, nz, bidirectional=bidirectional, num_layers=num_layers)
        self._init_dagnn(emb_dim, hidden_dim, out_dim, num_layers, bidirectional, agg, out_wx, out_pool_all, out_pool,
                         dropout, num_nodes)


class DAGNN_BN(_DagnnDvaeBase):
    """dvae/dagnn_bn.py:19 (BN)."""
    _VID = False

    def __init__(self, emb_dim, hidden_dim, out_dim, max_n, nvt, START_TYPE, END_TYPE, hs, nz, num_layers=2,
                 bidirectional=True, agg=NA_ATTN_H, out_wx=False, out_pool_all=False, out_pool=P_MAX, dropout=0.0,
                 num_nodes=8):
        super().__init__(max_n, nvt, START_TYPE, END_TYPE, hs, nz, bidirectional=bidirectional, vid=False,
                         num_layers=num_layers, bn_variant=True)
        self._init_dagnn(emb_dim, hidden_dim, out_dim, num_layers, bidirectional, agg, out_wx, out_pool_all, out_pool,
                         dropout, num_nodes)
