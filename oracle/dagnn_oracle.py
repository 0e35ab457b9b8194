"""ORACLE — TEST INFRASTRUCTURE ONLY. Not product code, never on the product path.

CPU restatement (torch CPU fp32 + numpy for the integer parts) of the DAGNN layer-wise forward of
vthost/DAGNN @ b065cd5. Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this file, and only as the checker / the CPU baseline.

Parity status: the reference ships NO tests / golden vectors for this path and its
arithmetic lives in un-vendored third-party packages (torch 1.5 `nn.GRUCell`/`nn.Linear`/`nn.Embedding`,
torch_geometric 1.6.0 `MessagePassing.propagate`, `utils.softmax`, `nn.global_*_pool`, torch_scatter).
This restatement is therefore pinned against OUTPUTS OF THE REFERENCE'S OWN MODEL FILES, imported
unchanged from /root/reference on top of the test-only PyG shim in `oracle/shim/` (script:
`oracle/gen_golden.py`, fixtures: `tests/golden/*.npz`, check: `tests/test_oracle_golden.py`). The PyG
semantics themselves are restated from upstream documentation ("[PyG-upstream]"), pinned by the
hand-computed cases in `tests/test_shim.py`. Anything beyond that is "parity unpinned".

The structure deliberately follows the reference (same loops, same per-node edge scan, same
per-(level, layer) message passing that scatters into a full [N, H] buffer) so that timing it is a fair
"port" of the reference's CPU path.

Reference files restated (relative to /root/reference):
  src/utils_dag.py:8-35,39-52,70-76           -> top_sort, add_order_info_01, add_order_info
  ogbg-code/utils.py:21-28                     -> ast_node_encoder
  ogbg-code/model/dagnn.py:128-215             -> ogb_forward (level loop :144-182, readout :184-202, heads :209-215)
  ogbg-code/model/dagnn.py:347-376             -> attn_conv  (+ PyG propagate / softmax / add-aggregate)
  dvae/dagnn.py:99-175, dvae/dagnn_bn.py:98-168 -> dvae_forward (vertex-id columns :130-139, readouts :147-172)
  dvae/dagnn.py:177-184                        -> dvae_encode
  torch.nn.GRUCell                              -> gru_cell (gate order r, z, n)
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- integer pre-pass
def top_sort(edge_index, graph_size: int) -> torch.Tensor:
    """src/utils_dag.py:8-35 — frontier peeling: a node is evaluated in round n iff none of its
    parents is still unevaluated at the start of round n."""
    ei = np.asarray(edge_index)
    parents, children = ei[0], ei[1]
    ids = np.arange(graph_size, dtype=int)
    order = np.zeros(graph_size, dtype=int)
    pending = np.ones(graph_size, dtype=bool)
    rnd = 0
    while pending.any():
        blocked = children[pending[parents]]
        ready = pending & ~np.isin(ids, blocked)
        if not ready.any():
            raise ValueError("cycle")
        order[ready] = rnd
        pending[ready] = False
        rnd += 1
    return torch.from_numpy(order).long()


def add_order_info_01(edge_index: torch.Tensor, num_nodes: int):
    """src/utils_dag.py:39-52 — (levels fwd, node ids, levels on reversed edges, node ids)."""
    l0 = top_sort(edge_index, num_nodes)
    rev = torch.stack([edge_index[1], edge_index[0]])
    l1 = top_sort(rev, num_nodes)
    ns = torch.arange(num_nodes, dtype=torch.long)
    return l0, ns, l1, ns.clone()


def add_order_info(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """src/utils_dag.py:70-76 — bi_layer_index int64[2, 2, n]: [dir][0]=level, [dir][1]=node id."""
    l0, ns, l1, _ = add_order_info_01(edge_index, num_nodes)
    return torch.stack([torch.stack([l0, ns]), torch.stack([l1, ns])])


def assert_order(edge_index, order, ns) -> None:
    """src/utils_dag.py:55-67 — every predecessor sits in a strictly earlier level."""
    done = set()
    for lvl in range(int(order.max()) + 1):
        here = ns[order == lvl].tolist()
        for n in here:
            for p in edge_index[0][edge_index[1] == n].tolist():
                assert p in done
        done.update(here)


# ----------------------------------------------------------------------------- float building blocks
def ast_node_encoder(x, depth, p: Dict[str, torch.Tensor], prefix="encoder.", max_depth=20):
    """ogbg-code/utils.py:26-28 (depth clamp, then sum of three embedding rows)."""
    depth = depth.clamp(max=max_depth)
    return (p[prefix + "type_encoder.weight"][x[:, 0]] + p[prefix + "attribute_encoder.weight"][x[:, 1]]
            + p[prefix + "depth_encoder.weight"][depth])


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.GRUCell: r=σ(W_ir x+b_ir+W_hr h+b_hr); z=σ(..); n=tanh(W_in x+b_in+r*(W_hn h+b_hn));
    h'=(1-z)*n+z*h ; h=None means zeros."""
    if h is None:
        h = x.new_zeros(x.shape[0], w_hh.shape[1])
    gi = F.linear(x, w_ih, b_ih)
    gh = F.linear(h, w_hh, b_hh)
    i_r, i_z, i_n = gi.chunk(3, 1)
    h_r, h_z, h_n = gh.chunk(3, 1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return n + z * (h - n)


def segment_softmax(a, index, num_nodes):
    """[PyG-upstream utils/softmax.py] exp(a - segmax) / (segsum + 1e-16), segments = equal `index`."""
    mx = a.new_full((num_nodes,) + tuple(a.shape[1:]), float("-inf"))
    idx = index.view(-1, 1).expand_as(a)
    mx = mx.scatter_reduce(0, idx, a, reduce="amax", include_self=True)
    e = (a - mx[index]).exp()
    s = a.new_zeros((num_nodes,) + tuple(a.shape[1:])).scatter_add(0, idx, e)
    return e / (s[index] + 1e-16)


def attn_conv(h, edge_index, h_attn_q, h_attn, attn_w, attn_b, edge_attr=None, edge_w=None, edge_b=None,
              reverse=False):
    """ogbg-code/model/dagnn.py:362-373 through PyG propagate: aggregate at i=edge_index[1] from
    j=edge_index[0] (reverse: i=edge_index[0], j=edge_index[1]); returns a full [N, H] buffer."""
    i, j = (0, 1) if reverse else (1, 0)
    tgt, nbr = edge_index[i], edge_index[j]
    key = h_attn.index_select(0, nbr)
    if edge_w is not None:
        key = key + F.linear(edge_attr, edge_w, edge_b)
    if h_attn_q is None:      # SelfAttnConv, dagnn.py:296-310: no query part
        a = F.linear(key, attn_w, attn_b)
    else:
        a = F.linear(torch.cat([h_attn_q.index_select(0, tgt), key], dim=-1), attn_w, attn_b)
    a = segment_softmax(a, tgt, h.shape[0])
    msg = h.index_select(0, nbr) * a
    out = h.new_zeros(h.shape[0], h.shape[1])
    return out.scatter_add(0, tgt.view(-1, 1).expand_as(msg), msg)


# ----------------------------------------------------------------------------- the level sweep
def level_sweep(x, edge_index, bi_layer_index, p: Dict[str, torch.Tensor], num_layers: int, dirs: List[int],
                hidden_dim: int, edge_attr=None, vid_nodes: int = 0, trace: Optional[dict] = None,
                cell_prefix="cells_{}.{}.", self_attn: bool = False):
    """ogbg-code/model/dagnn.py:141-182 / dvae/dagnn.py:106-145 / dvae/dagnn_bn.py:105-136.

    Returns H[d][i] (float32 [N, hidden]) for every direction/layer. `vid_nodes` > 0 appends the D-VAE
    one-hot vertex id (node index mod vid_nodes) to keys (always) and to the query (layers > 0)
    (dvae/dagnn.py:130-139). `trace`, if given, records the per-level node and edge lists.
    """
    n = x.shape[0]
    num_levels = int(bi_layer_index[0][0].max()) + 1
    H = [[x.new_zeros(n, hidden_dim) for _ in range(num_layers)] for _ in dirs]
    has_ea = edge_attr is not None
    for d in dirs:
        for l_idx in range(num_levels):
            layer = bi_layer_index[d][1][bi_layer_index[d][0] == l_idx]
            inp = x[layer]
            if l_idx > 0:
                le_idx = [(edge_index[1 - d] == nd).nonzero().squeeze(-1) for nd in layer]
                le_idx = torch.cat(le_idx, dim=-1) if le_idx else torch.zeros(0, dtype=torch.long)
                lp_edge_index = edge_index[:, le_idx]
            if trace is not None:
                trace.setdefault("nodes", {})[(d, l_idx)] = layer.clone()
                if l_idx > 0:
                    trace.setdefault("edges", {})[(d, l_idx)] = le_idx.clone()
            for i in range(num_layers):
                if l_idx == 0:
                    ps_h = None
                else:
                    if vid_nodes:
                        vids = F.one_hot(torch.arange(n).fmod(vid_nodes), vid_nodes).to(x.dtype)
                        keys = torch.cat([H[d][i], vids], dim=-1)
                        q = torch.cat([H[d][i - 1], vids], dim=-1) if i > 0 else x
                    else:
                        keys = H[d][i]
                        q = H[d][i - 1] if i > 0 else x
                    ap = "node_aggr_{}.{}.".format(d, i)
                    ps_h = attn_conv(H[d][i], lp_edge_index, None if self_attn else q, keys, p[ap + "attn_lin.weight"],
                                     p[ap + "attn_lin.bias"],
                                     edge_attr[le_idx] if has_ea else None,
                                     p.get(ap + "edge_encoder.weight") if has_ea else None,
                                     p.get(ap + "edge_encoder.bias") if has_ea else None,
                                     reverse=(d == 1))[layer]
                cp = cell_prefix.format(d, i)
                inp = gru_cell(inp, ps_h, p[cp + "weight_ih"], p[cp + "weight_hh"], p[cp + "bias_ih"],
                               p[cp + "bias_hh"])
                H[d][i][layer] += inp
    return H


def _pool(h, batch, kind, num_graphs):
    if kind == "attn":
        # dagnn.py:114-117: softmax over a dimension of size 1 — every weight is exactly 1.0 — then global_add_pool
        kind = "add"
    idx = batch.view(-1, 1).expand_as(h)
    if kind == "max":
        out = h.new_full((num_graphs, h.shape[1]), float("-inf")).scatter_reduce(0, idx, h, reduce="amax",
                                                                                   include_self=True)
        return torch.where(torch.isinf(out) & (out < 0), torch.zeros_like(out), out)
    s = h.new_zeros(num_graphs, h.shape[1]).scatter_add(0, idx, h)
    if kind == "add":
        return s
    cnt = h.new_zeros(num_graphs).index_add(0, batch, torch.ones_like(batch, dtype=h.dtype)).clamp(min=1)
    return s / cnt.view(-1, 1)


def ogb_readout(G, X, H, num_layers, bidirectional=True, out_wx=False, out_pool_all=False, out_pool="max"):
    """ogbg-code/model/dagnn.py:119-126,184-202."""
    lvl = [G._bi_layer_idx0, G._bi_layer_idx1]
    ids = [G._bi_layer_index0, G._bi_layer_index1]
    nb = int(G.batch.max()) + 1
    if bidirectional and not out_pool_all:
        outs = []
        for d in (0, 1):
            index = ids[1 - d][lvl[1 - d] == 0]          # d=0: sinks (reverse level 0); d=1: sources
            hd = torch.cat(([X] if out_wx else []) + [H[d][l] for l in range(num_layers)], dim=-1)
            outs.append(_pool(hd[index], G.batch[index], out_pool, nb))
        return torch.cat(outs, dim=-1)
    dirs = [0, 1] if bidirectional else [0]
    h = torch.cat(([X] if out_wx else []) + [H[d][l] for d in range(len(dirs)) for l in range(num_layers)], dim=-1)
    b = G.batch
    if not out_pool_all:
        index = ids[1][lvl[1] == 0]
        h, b = h[index], b[index]
    return _pool(h, b, out_pool, nb)


def ogb_forward(p: Dict[str, torch.Tensor], G, num_layers=2, bidirectional=True, out_wx=False,
                out_pool_all=False, out_pool="max", max_seq_len=5, num_class=0, w_edge_attr=True,
                heads=True, trace: Optional[dict] = None, agg="attn_h"):
    """ogbg-code/model/dagnn.py:128-215 with agg='attn_h', recurr=1, encoder=ASTNodeEncoder.
    Returns (pred_list | logits, readout [B, out_hidden], H)."""
    dirs = [0, 1] if bidirectional else [0]
    bi = torch.stack([torch.stack([G._bi_layer_idx0, G._bi_layer_index0]),
                      torch.stack([G._bi_layer_idx1, G._bi_layer_index1])])
    X = ast_node_encoder(G.x, G.node_depth.view(-1), p)
    hidden = p["cells_0.0.weight_hh"].shape[1]
    H = level_sweep(X, G.edge_index, bi, p, num_layers, dirs, hidden,
                    edge_attr=G.edge_attr if w_edge_attr else None, trace=trace, self_attn=(agg == "self_attn_h"))
    out = ogb_readout(G, X, H, num_layers, bidirectional, out_wx, out_pool_all, out_pool)
    if not heads:
        return None, out, H
    if num_class > 0:
        return F.linear(out, p["graph_pred_linear.weight"], p["graph_pred_linear.bias"]), out, H
    preds = [F.linear(out, p["graph_pred_linear_list.%d.weight" % k], p["graph_pred_linear_list.%d.bias" % k])
             for k in range(max_seq_len)]
    return preds, out, H


def dvae_forward(p: Dict[str, torch.Tensor], G, num_layers=2, bidirectional=False, num_nodes=8, vid=True,
                 trace: Optional[dict] = None, out_pool_all=False, out_pool="max"):
    """dvae/dagnn.py:99-175 (vid=True, NA) / dvae/dagnn_bn.py:98-168 (vid=False, BN); out_pool_all=True: :162-172."""
    dirs = [0, 1] if bidirectional else [0]
    hidden = p["grue_forward.0.weight_hh"].shape[1]
    q = dict(p)
    for l in range(num_layers):                       # cells_0/1 alias grue_forward/backward (dagnn.py:73-75)
        for nm in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
            q["cells_0.%d.%s" % (l, nm)] = p["grue_forward.%d.%s" % (l, nm)]
            if bidirectional:
                q["cells_1.%d.%s" % (l, nm)] = p["grue_backward.%d.%s" % (l, nm)]
    H = level_sweep(G.x, G.edge_index, G.bi_layer_index, q, num_layers, dirs, hidden,
                    vid_nodes=num_nodes if vid else 0, trace=trace)
    n = G.x.shape[0]
    if out_pool_all:
        cat = torch.cat([H[d][l] for d in range(len(dirs)) for l in range(num_layers)], dim=-1)
        if bidirectional:
            cat = F.linear(cat, p["hg_unify.0.weight"], p["hg_unify.0.bias"])
        elif num_layers > 1:
            cat = F.linear(cat, p["out_linear.weight"], p["out_linear.bias"])
        return _pool(cat, G.batch, out_pool, int(G.batch.max()) + 1), H
    first = torch.arange(0, n, num_nodes)
    last = first + (num_nodes - 1)
    if bidirectional:
        h0 = torch.cat([H[0][l][last] for l in range(num_layers)], dim=-1)
        h1 = torch.cat([H[1][l][first] for l in range(num_layers)], dim=-1)
        out = F.linear(torch.cat([h0, h1], dim=-1), p["hg_unify.0.weight"], p["hg_unify.0.bias"])
    else:
        hcat = torch.cat([H[0][l][last] for l in range(num_layers)], dim=-1)
        out = F.linear(hcat, p["out_linear.weight"], p["out_linear.bias"]) if num_layers > 1 else hcat
    return out, H


def dvae_encode(p, G, **kw):
    """dvae/dagnn.py:177-184 — (mu, logvar) = fc1/fc2 of the graph embedding."""
    out, _ = dvae_forward(p, G, **kw)
    return F.linear(out, p["fc1.weight"], p["fc1.bias"]), F.linear(out, p["fc2.weight"], p["fc2.bias"])
