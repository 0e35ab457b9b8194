"""Host-side batch containers and input builders for the DAGNN level sweep.

Nothing here is on the GPU hot path: these are the *inputs* either side of it:

* `DagBatch`            – duck-typed stand-in for a PyG `Batch` carrying exactly the attributes the
                          reference `forward(G)` reads (ogbg-code/model/dagnn.py:128-139, dvae/dagnn.py:99-114).
* `dag_levels_host`     – longest-path level of every node (what `top_sort` computes,
                          src/utils_dag.py:8-35) for a whole disconnected batch at once, numpy.
* `augment_edge2_batch` – `augment_edge2` (ogbg-code/utils2.py:31-79) on a collated batch, device-agnostic torch ops.
* `make_code2_batch`    – seeded synthetic "ogbg-code2-shaped" AST batches:
                          ogbg-code2 itself is not available offline, so these are shape proxies.
* `decode_enas_row` / `decode_bn_row` – the NA / BN text-row decoders (dvae/util.py:343-385 / :290-339).
* `collate_dvae`        – what `dvae/batch.py:26-145` does to the keys the forward reads.
"""
from __future__ import annotations

import ast
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch


class DagBatch(object):
    """Attribute bag with the tensor attributes of a PyG Batch that DAGNN.forward reads."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if torch.is_tensor(v)]

    def _map(self, fn):
        out = DagBatch()
        for k, v in self.__dict__.items():
            setattr(out, k, fn(v) if torch.is_tensor(v) else v)
        return out

    def to(self, device, non_blocking: bool = False):
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())

    def clone(self):
        return self._map(lambda t: t.clone())

    def nbytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.__dict__.values() if torch.is_tensor(v))


# --------------------------------------------------------------------------------------
# levels
# --------------------------------------------------------------------------------------
def dag_levels_host(src: np.ndarray, dst: np.ndarray, n: int) -> np.ndarray:
    """Longest-path depth from a source for every node of a DAG (= `top_sort`, src/utils_dag.py:8-35).

    Relaxation `lvl[v] = max(lvl[v], lvl[u]+1)` over all edges until a fixed point; on a DAG that takes
    (depth+1) rounds. Works on a whole batch (disjoint union) at once.
    """
    lvl = np.zeros(n, dtype=np.int64)
    if len(src) == 0:
        return lvl
    for _ in range(n + 1):
        new = lvl.copy()
        np.maximum.at(new, dst, lvl[src] + 1)
        if np.array_equal(new, lvl):
            return lvl
        lvl = new
    raise ValueError("edge list is not acyclic")


# --------------------------------------------------------------------------------------
# augment_edge2 on a whole batch
# --------------------------------------------------------------------------------------
def augment_edge2_batch(edge_index_ast: torch.Tensor, node_is_attributed: torch.Tensor, batch: torch.Tensor):
    """`augment_edge2` (ogbg-code/utils2.py:31-79) applied to every graph of an already collated batch, on whatever device
    the tensors live on (index plumbing with torch ops, no host loop): next-token edges chain the consecutive attributed
    nodes of each graph (nodes are in DFS order), `edge_attr` = [is next-token, is inverse] = [0, 0] for AST edges and
    [1, 0] for next-token edges. The edges come back in the order per-graph augmentation followed by PyG collation gives:
    graph by graph, AST edges first, then the graph's next-token edges (the order only fixes the summation order inside a
    softmax). Returns (edge_index int64 [2, E'], edge_attr fp32 [E', 2])."""
    dev = edge_index_ast.device
    idx = torch.where(node_is_attributed.view(-1) == 1)[0]
    if idx.numel() >= 2:
        same = batch[idx[:-1]] == batch[idx[1:]]
        nt = torch.stack([idx[:-1][same], idx[1:][same]], 0)
    else:
        nt = torch.zeros(2, 0, dtype=torch.long, device=dev)
    ei = torch.cat([edge_index_ast, nt], 1)
    kind = torch.cat([torch.zeros(edge_index_ast.shape[1], dtype=torch.long, device=dev),
                      torch.ones(nt.shape[1], dtype=torch.long, device=dev)])
    order = torch.argsort(batch[ei[0]] * 2 + kind, stable=True)
    ei, kind = ei[:, order], kind[order]
    ea = torch.stack([kind.float(), torch.zeros_like(kind, dtype=torch.float32)], 1)
    return ei, ea


# --------------------------------------------------------------------------------------
# synthetic ogbg-code2-shaped batches
# --------------------------------------------------------------------------------------
CODE2_NUM_NODETYPES = 98
CODE2_NUM_NODEATTRS = 10030
CODE2_MAX_DEPTH = 20
CODE2_NUM_VOCAB = 5002
CODE2_MAX_SEQ_LEN = 5


def _random_ast(rng: np.random.Generator, n: int):
    """Random ordered tree on n nodes in DFS pre-order: node v hangs off the node k steps up the current
    right spine, k ~ Geom(0.5)-1 clipped to the spine length. Returns (parent[1:], depth)."""
    parent = np.zeros(n, dtype=np.int64)
    depth = np.zeros(n, dtype=np.int64)
    spine = [0]
    ks = rng.geometric(0.5, size=n) - 1
    for v in range(1, n):
        k = min(int(ks[v]), len(spine) - 1)
        if k:
            del spine[len(spine) - k:]
        parent[v] = spine[-1]
        depth[v] = len(spine)
        spine.append(v)
    return parent, depth


def make_code2_batch(num_graphs: int, seed: int, mean_nodes: float = 105.0, sigma: float = 0.6,
                     min_nodes: int = 11, max_nodes: int = 1000) -> DagBatch:
    """Seeded synthetic batch with the attributes `main_pyg.py` feeds to DAGNN.forward.

    Per graph: n ~ round(LogNormal(ln mean_nodes, sigma)) clipped to [min_nodes, max_nodes]; AST edges
    parent->child; `node_depth` = tree depth; `x[:,0] ~ U{0..97}`, `x[:,1] ~ U{0..10029}`; a node is
    "attributed" w.p. 0.9 if leaf else 0.1 and consecutive attributed nodes are chained by next-token edges
    appended after the AST edges with `edge_attr = [1, 0]` (AST edges `[0, 0]`) exactly as `augment_edge2`
    does (ogbg-code/utils2.py:31-79). Levels (`_bi_layer_idx0/1`) come from the AST edges only
    (ogb/io/read_graph_pyg.py:51 runs before the augmentation, main_pyg.py:235).
    """
    rng = np.random.default_rng(seed)
    xs, depths, eis, eas, l0s, l1s, batch = [], [], [], [], [], [], []
    off = 0
    for g in range(num_graphs):
        n = int(np.clip(np.rint(rng.lognormal(np.log(mean_nodes), sigma)), min_nodes, max_nodes))
        parent, depth = _random_ast(rng, n)
        child = np.arange(1, n, dtype=np.int64)
        ast_src, ast_dst = parent[1:], child
        is_leaf = np.ones(n, dtype=bool)
        is_leaf[ast_src] = False
        u = rng.random(n)
        attributed = np.where(is_leaf, u < 0.9, u < 0.1)
        att = np.nonzero(attributed)[0]
        nt_src, nt_dst = att[:-1], att[1:]
        x = np.stack([rng.integers(0, CODE2_NUM_NODETYPES, n), rng.integers(0, CODE2_NUM_NODEATTRS, n)], 1)
        lvl0 = dag_levels_host(ast_src, ast_dst, n)
        lvl1 = dag_levels_host(ast_dst, ast_src, n)
        src = np.concatenate([ast_src, nt_src])
        dst = np.concatenate([ast_dst, nt_dst])
        ea = np.zeros((len(src), 2), dtype=np.float32)
        ea[len(ast_src):, 0] = 1.0
        xs.append(x); depths.append(depth); l0s.append(lvl0); l1s.append(lvl1)
        eis.append(np.stack([src, dst]) + off); eas.append(ea)
        batch.append(np.full(n, g, dtype=np.int64))
        off += n
    ids = torch.arange(off, dtype=torch.long)
    return DagBatch(
        x=torch.from_numpy(np.concatenate(xs)).long(),
        node_depth=torch.from_numpy(np.concatenate(depths)).long().view(-1, 1),
        edge_index=torch.from_numpy(np.concatenate(eis, 1)).long().contiguous(),
        edge_attr=torch.from_numpy(np.concatenate(eas)),
        batch=torch.from_numpy(np.concatenate(batch)),
        _bi_layer_idx0=torch.from_numpy(np.concatenate(l0s)),
        _bi_layer_index0=ids.clone(),
        _bi_layer_idx1=torch.from_numpy(np.concatenate(l1s)),
        _bi_layer_index1=ids.clone(),
        num_graphs=num_graphs,
    )


def make_random_dag_batch(num_graphs: int, seed: int, n_lo: int = 3, n_hi: int = 24, p_edge: float = 0.25,
                          p_extra: float = 0.3, with_attr: bool = True) -> DagBatch:
    """Small random DAG batches for parity tests: general DAGs (not trees), duplicate edges, isolated
    nodes, and "extra" edges that are *not* part of the level computation (like next-token edges, they may
    point from a node at the same or a higher level: SURVEY.md §9-Q1/Q2)."""
    rng = np.random.default_rng(seed)
    xs, depths, eis, eas, l0s, l1s, batch = [], [], [], [], [], [], []
    off = 0
    for g in range(num_graphs):
        n = int(rng.integers(n_lo, n_hi + 1))
        iu = np.triu_indices(n, 1)
        keep = rng.random(len(iu[0])) < p_edge
        src, dst = iu[0][keep].astype(np.int64), iu[1][keep].astype(np.int64)
        lvl0 = dag_levels_host(src, dst, n)
        lvl1 = dag_levels_host(dst, src, n)
        n_extra = int(rng.binomial(n, p_extra))
        ex_src = rng.integers(0, n, n_extra).astype(np.int64)
        ex_dst = rng.integers(0, n, n_extra).astype(np.int64)
        if len(src) and rng.random() < 0.5:     # a duplicate of an existing edge
            ex_src = np.append(ex_src, src[0]); ex_dst = np.append(ex_dst, dst[0])
        ea = np.zeros((len(src) + len(ex_src), 2), dtype=np.float32)
        ea[len(src):, 0] = 1.0
        ea[:, 1] = (rng.random(len(ea)) < 0.2).astype(np.float32)
        xs.append(np.stack([rng.integers(0, CODE2_NUM_NODETYPES, n), rng.integers(0, CODE2_NUM_NODEATTRS, n)], 1))
        depths.append(rng.integers(0, 30, n))
        l0s.append(lvl0); l1s.append(lvl1)
        eis.append(np.stack([np.concatenate([src, ex_src]), np.concatenate([dst, ex_dst])]) + off)
        eas.append(ea)
        batch.append(np.full(n, g, dtype=np.int64))
        off += n
    ids = torch.arange(off, dtype=torch.long)
    b = DagBatch(
        x=torch.from_numpy(np.concatenate(xs)).long(),
        node_depth=torch.from_numpy(np.concatenate(depths)).long().view(-1, 1),
        edge_index=torch.from_numpy(np.concatenate(eis, 1)).long().contiguous(),
        edge_attr=torch.from_numpy(np.concatenate(eas)),
        batch=torch.from_numpy(np.concatenate(batch)),
        _bi_layer_idx0=torch.from_numpy(np.concatenate(l0s)),
        _bi_layer_index0=ids.clone(),
        _bi_layer_idx1=torch.from_numpy(np.concatenate(l1s)),
        _bi_layer_index1=ids.clone(),
        num_graphs=num_graphs,
    )
    if not with_attr:
        b.edge_attr = None
    return b


# --------------------------------------------------------------------------------------
# D-VAE rows (NA = ENAS architectures, BN = Bayesian networks)
# --------------------------------------------------------------------------------------
def _one_hot_rows(types: Sequence[int], width: int) -> torch.Tensor:
    x = torch.zeros(len(types), width)
    x[torch.arange(len(types)), torch.tensor(list(types))] = 1.0
    return x


def _graph_from_adj(adj: np.ndarray, types: List[int], width: int) -> DagBatch:
    # `nx.DiGraph(adj).edges` lists the non-zeros of adj row-major (dvae/util.py:321-330, :368-372)
    src, dst = np.nonzero(adj)
    n = len(types)
    ids = np.arange(n, dtype=np.int64)
    l0 = dag_levels_host(src, dst, n)
    l1 = dag_levels_host(dst, src, n)
    bi = torch.from_numpy(np.stack([np.stack([l0, ids]), np.stack([l1, ids])]))   # add_order_info, utils_dag.py:70-76
    return DagBatch(x=_one_hot_rows(types, width),
                    edge_index=torch.from_numpy(np.stack([src, dst]).astype(np.int64)),
                    bi_layer_index=bi, vs=[{"type": t} for t in types])


def decode_enas_row(row, n_types: int = 6) -> DagBatch:
    """ENAS row -> graph (dvae/util.py:343-385): node 0 = start (type 0), node i+1 has type row[i][0]+2 and
    an edge from its predecessor in the chain plus one per set flag, last node = end (type 1)."""
    if isinstance(row, str):
        row = ast.literal_eval(row)
    width = n_types + 2
    n = len(row)
    adj = np.zeros((width, width))
    types = [0]
    for i, node in enumerate(row):
        types.append(node[0] + 2)
        adj[i, i + 1] = 1
        for j, e in enumerate(node[1:]):
            if e == 1:
                adj[j, i + 1] = 1
    types.append(1)
    adj[n, n + 1] = 1
    return _graph_from_adj(adj, types, width)


def decode_bn_row(row, n_types: int = 8) -> DagBatch:
    """BN row -> graph (dvae/util.py:290-339): start node feeds every parent-less variable, every variable
    without children feeds the end node."""
    if isinstance(row, str):
        row = ast.literal_eval(row)
    width = n_types + 2
    n = len(row)
    adj = np.zeros((width, width))
    end_vertices = [True] * n
    types = [0]
    for i, node in enumerate(row):
        types.append(node[0] + 2)
        if sum(node[1:]) == 0:
            adj[0, i + 1] = 1
        else:
            for j, e in enumerate(node[1:]):
                if e == 1:
                    adj[j + 1, i + 1] = 1
                    end_vertices[j] = False
    types.append(1)
    for j, flag in enumerate(end_vertices):
        if flag:
            adj[j + 1, n + 1] = 1
    return _graph_from_adj(adj, types, width)


def read_dvae_rows(path: str, start: int, count: int) -> list:
    """Rows [start, start+count) of final_structures6.txt / asia_200k.txt as (row, y) tuples."""
    out = []
    with open(path, "r") as f:
        for i, line in enumerate(f):
            if i < start:
                continue
            if i >= start + count:
                break
            out.append(ast.literal_eval(line.strip()))
    return out


def collate_dvae(graphs: Iterable[DagBatch]) -> DagBatch:
    """Concatenate D-VAE graphs into one batch: `edge_index` and row 1 of `bi_layer_index` (node ids) are
    offset by the running node count, row 0 (levels) is not (dvae/batch.py:54-59)."""
    xs, eis, bis, bv = [], [], [], []
    off = 0
    for i, g in enumerate(graphs):
        n = g.x.shape[0]
        xs.append(g.x)
        eis.append(g.edge_index + off)
        bi = g.bi_layer_index.clone()
        bi[:, 1] += off
        bis.append(bi)
        bv.append(torch.full((n,), i, dtype=torch.long))
        off += n
    return DagBatch(x=torch.cat(xs), edge_index=torch.cat(eis, 1).contiguous(),
                    bi_layer_index=torch.cat(bis, -1).contiguous(), batch=torch.cat(bv), num_graphs=len(xs))


def make_random_dvae_batch(num_graphs: int, seed: int, kind: str = "NA") -> DagBatch:
    """Random but well-formed ENAS (8-node) / BN (10-node) rows — same decoders as the real data."""
    rng = np.random.default_rng(seed)
    gs = []
    for _ in range(num_graphs):
        if kind == "NA":
            row = [[int(rng.integers(0, 6))] + [int(b) for b in rng.integers(0, 2, i)] for i in range(6)]
            gs.append(decode_enas_row(row))
        else:
            perm = rng.permutation(8)
            row = [[int(perm[i])] + [int(b) for b in (rng.random(i) < 0.3)] for i in range(8)]
            gs.append(decode_bn_row(row))
    return collate_dvae(gs)


def shard_graph_ranges(nodes_per_graph: Sequence[int], world_size: int) -> List[range]:
    """Contiguous node-balanced split of a graph list into `world_size` shards — the rule of the reference's
    multi-GPU collater (ogbg-code/tg/dataloader.py:17-27): graph g goes to the device whose index is
    floor(world_size * midpoint_g / total_nodes), midpoint_g = nodes before g + half of g's nodes."""
    cnt = np.asarray(nodes_per_graph, dtype=np.float64)
    cum = np.cumsum(cnt)
    mid = cum - 0.5 * cnt
    dev = np.minimum((world_size * mid / cum[-1]).astype(np.int64), world_size - 1) if len(cnt) else np.zeros(0, int)
    out, start = [], 0
    for r in range(world_size):
        stop = int(np.searchsorted(dev, r, side="right"))
        out.append(range(start, stop))
        start = stop
    return out


def shard_graph_ids(nodes_per_graph: Sequence[int], world_size: int, depths: Optional[Sequence[int]] = None,
                    level_cost_nodes: float = 200.0) -> List[np.ndarray]:
    """Graph ids of every shard. Without `depths`: the reference's contiguous node-balanced rule (`shard_graph_ranges`).
    With `depths` (levels of every graph): depth-aware balance. The sweep walks the levels of a shard one after the other, so
    a shard costs about  levels * t_level + nodes * t_node  — the deepest graph of a shard sets the first term (measured:
    SCALE_r01.json, 97 levels on one rank against 66 on rank 0 cost 21 % at equal node counts; round 2, cluster sweep, 8 ranks:
    t_level = 8.9 us, t_node = 0.0437 us, i.e. one level costs as much as ~200 nodes — the default of `level_cost_nodes`). Graphs are dealt in order of
    decreasing depth, each to the shard whose modelled cost  level_cost_nodes * max depth + nodes  stays smallest: the deepest
    graphs land on different shards and a shard with a deep graph gets fewer nodes. Ids inside a shard are ascending.
    Shards may be empty when there are fewer graphs than ranks."""
    cnt = np.asarray(nodes_per_graph, dtype=np.int64)
    if depths is None:
        return [np.arange(r.start, r.stop, dtype=np.int64) for r in shard_graph_ranges(cnt, world_size)]
    dep = np.asarray(depths, dtype=np.int64)
    order = np.lexsort((np.arange(len(cnt)), -cnt, -dep))            # depth desc, then nodes desc, then id
    nodes = np.zeros(world_size, dtype=np.int64)
    deep = np.zeros(world_size, dtype=np.int64)
    out = [[] for _ in range(world_size)]
    for gidx in order:
        cost = level_cost_nodes * np.maximum(deep, dep[gidx]) + nodes + cnt[gidx]
        r = int(np.argmin(cost))
        out[r].append(int(gidx))
        nodes[r] += cnt[gidx]
        deep[r] = max(deep[r], dep[gidx])
    return [np.asarray(sorted(ids), dtype=np.int64) for ids in out]


def deterministic_init_(module: torch.nn.Module, seed: int) -> torch.nn.Module:
    """Fill every parameter from a numpy generator (platform/torch-version independent), with the scale of
    the default initialisers: embeddings ~ N(0,1), everything else ~ U(-1/sqrt(fan), 1/sqrt(fan)) where fan is
    the last dimension (GRUCell uses 1/sqrt(hidden); biases use the fan of their weight)."""
    names = sorted(n for n, _ in module.named_parameters())
    params = dict(module.named_parameters())
    for k, name in enumerate(names):
        p = params[name]
        rng = np.random.default_rng([seed, k])
        if "encoder.weight" in name and p.dim() == 2 and "edge_encoder" not in name:
            v = rng.standard_normal(p.shape)
        else:
            if p.dim() >= 2:
                fan = p.shape[-1]
            else:
                w = params.get(name.replace("bias_ih", "weight_hh").replace("bias_hh", "weight_hh")
                               .replace("bias", "weight"))
                fan = w.shape[-1] if w is not None and w.dim() >= 2 else p.shape[0]
            b = 1.0 / np.sqrt(max(fan, 1))
            v = rng.uniform(-b, b, p.shape)
        with torch.no_grad():
            p.copy_(torch.from_numpy(np.asarray(v, dtype=np.float32)))
    return module


# --------------------------------------------------------------------------------------
# graph-level slicing of a batch (multi-GPU sharding, tests)
# --------------------------------------------------------------------------------------
def graph_node_counts(B: DagBatch) -> np.ndarray:
    ng = int(getattr(B, "num_graphs", int(B.batch.max()) + 1))
    return np.bincount(B.batch.numpy(), minlength=ng)


def graph_depths(B: DagBatch) -> np.ndarray:
    """Number of levels of every graph of a batch (max forward level + 1)."""
    lvl = (B._bi_layer_idx0 if hasattr(B, "_bi_layer_idx0") else B.bi_layer_index[0][0]).numpy()
    cnt = graph_node_counts(B)
    out = np.zeros(len(cnt), dtype=np.int64)
    np.maximum.at(out, B.batch.numpy(), lvl + 1)
    return out


def select_graphs(B: DagBatch, graph_ids) -> DagBatch:
    """New batch made of the graphs `graph_ids` of B (in that order), node ids renumbered — the same result as
    collating those graphs from scratch. Works for OGB-style (`_bi_layer_idx*`) and D-VAE-style (`bi_layer_index`)
    batches. Host-side numpy; not on the hot path."""
    graph_ids = np.asarray(list(graph_ids), dtype=np.int64)
    batch = B.batch.numpy()
    cnt = graph_node_counts(B)
    first = np.concatenate([[0], np.cumsum(cnt)])[:-1]
    new_first = np.concatenate([[0], np.cumsum(cnt[graph_ids])])[:-1]
    nodes = np.concatenate([np.arange(first[g], first[g] + cnt[g]) for g in graph_ids]) if len(graph_ids) else np.zeros(0, np.int64)
    new_id = np.full(len(batch), -1, dtype=np.int64)
    new_id[nodes] = np.arange(len(nodes))
    ei = B.edge_index.numpy()
    # edges grouped by graph in the order of graph_ids, original order inside a graph
    eg = batch[ei[0]] if ei.shape[1] else np.zeros(0, np.int64)
    rank = np.full(len(cnt), -1, dtype=np.int64)
    rank[graph_ids] = np.arange(len(graph_ids))
    keep = np.nonzero(rank[eg] >= 0)[0] if ei.shape[1] else np.zeros(0, np.int64)
    keep = keep[np.argsort(rank[eg[keep]], kind="stable")]
    nodes_t, keep_t = torch.from_numpy(nodes), torch.from_numpy(keep)
    out = DagBatch(num_graphs=len(graph_ids))
    out.batch = torch.from_numpy(np.repeat(np.arange(len(graph_ids)), cnt[graph_ids]))
    out.edge_index = torch.from_numpy(new_id[ei[:, keep]]).contiguous() if ei.shape[1] else B.edge_index.clone()
    out.x = B.x[nodes_t]
    if getattr(B, "edge_attr", None) is not None:
        out.edge_attr = B.edge_attr[keep_t]
    if hasattr(B, "node_depth"):
        out.node_depth = B.node_depth[nodes_t]
    if hasattr(B, "_bi_layer_idx0"):
        ids = torch.arange(len(nodes), dtype=torch.long)
        out._bi_layer_idx0, out._bi_layer_idx1 = B._bi_layer_idx0[nodes_t], B._bi_layer_idx1[nodes_t]
        out._bi_layer_index0, out._bi_layer_index1 = ids, ids.clone()
    if hasattr(B, "bi_layer_index"):
        bi = B.bi_layer_index[:, :, nodes_t].clone()
        bi[:, 1] = torch.arange(len(nodes), dtype=torch.long)
        out.bi_layer_index = bi.contiguous()
    return out


def split_batch(B: DagBatch, ranges) -> List[DagBatch]:
    return [select_graphs(B, r) for r in ranges]


def shard_batch(B: DagBatch, world_size: int, depth_aware: bool = True) -> List[DagBatch]:
    """One shard per rank: depth-aware balance (`shard_graph_ids`), or with depth_aware=False the contiguous node-balanced rule
    of ogbg-code/tg/dataloader.py:17-27. Empty shards come back as None (the reference collater drops them)."""
    ids = shard_graph_ids(graph_node_counts(B), world_size, graph_depths(B) if depth_aware else None)
    return [select_graphs(B, r) if len(r) else None for r in ids]
