"""CPU: host-side logic of the package — synthetic workload generators, level builder vs the oracle's top_sort, graph
selection / sharding helpers, and the world_size-2 gloo path (graph-sharded forward bookkeeping + gradient all-reduce)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dagnn_b200 import data as D, sharding
from oracle import dagnn_oracle as O


def test_code2_generator_is_deterministic_and_level_consistent():
    A, B = D.make_code2_batch(6, 77), D.make_code2_batch(6, 77)
    for k in ("x", "edge_index", "edge_attr", "batch", "_bi_layer_idx0", "_bi_layer_idx1"):
        assert torch.equal(getattr(A, k), getattr(B, k))
    assert A.num_graphs == 6 and A.x.shape[1] == 2 and A.edge_attr.shape[1] == 2
    # levels are computed on the AST edges only (edge_attr[:,0] == 0), like ogb/io/read_graph_pyg.py:51
    ast = A.edge_index[:, A.edge_attr[:, 0] == 0]
    counts = D.graph_node_counts(A)
    off = 0
    for g, n in enumerate(counts):
        m = (ast[0] >= off) & (ast[0] < off + n)
        ei = (ast[:, m] - off).numpy()
        assert np.array_equal(O.top_sort(ei, int(n)).numpy(), A._bi_layer_idx0[off:off + n].numpy())
        assert np.array_equal(O.top_sort(ei[::-1].copy(), int(n)).numpy(), A._bi_layer_idx1[off:off + n].numpy())
        off += int(n)
    # next-token edges exist and some violate the level order (they read zeros but keep softmax mass)
    nt = A.edge_index[:, A.edge_attr[:, 0] == 1]
    assert nt.shape[1] > 0 and bool((A._bi_layer_idx0[nt[0]] >= A._bi_layer_idx0[nt[1]]).any())


def test_select_and_split_roundtrip():
    B = D.make_code2_batch(9, 3)
    parts = D.split_batch(B, [range(0, 4), range(4, 9)])
    assert sum(p.num_graphs for p in parts) == 9 and sum(p.x.shape[0] for p in parts) == B.x.shape[0]
    sub = D.select_graphs(B, [2, 5])
    cnt = D.graph_node_counts(B)
    assert sub.x.shape[0] == cnt[2] + cnt[5] and int(sub.edge_index.max()) < sub.x.shape[0]
    assert torch.equal(sub._bi_layer_index0, torch.arange(sub.x.shape[0]))


def test_shard_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    for w in (1, 2, 4, 8):
        nodes = rng.integers(11, 400, size=64)
        rs = D.shard_graph_ranges(nodes, w)
        assert len(rs) == w and rs[0].start == 0 and rs[-1].stop == 64
        assert all(rs[k].stop == rs[k + 1].start for k in range(w - 1))
        loads = [int(nodes[r.start:r.stop].sum()) for r in rs]
        assert max(loads) - min(loads) <= 2 * int(nodes.max())       # contiguous node-balanced split (tg/dataloader.py:17-27)


def test_dvae_row_decoders():
    G = D.make_random_dvae_batch(5, 11, "NA")
    assert G.x.shape == (40, 8) and G.bi_layer_index.shape == (2, 2, 40)
    assert torch.equal(G.bi_layer_index[0][1], torch.arange(40))
    for g in range(5):                                              # ENAS rows: a chain 0->1->...->7 plus skip edges: 8 levels
        assert torch.equal(G.bi_layer_index[0][0][8 * g:8 * g + 8], torch.arange(8))
    Gb = D.make_random_dvae_batch(4, 12, "BN")
    assert Gb.x.shape == (40, 10)


def test_depth_aware_sharding_spreads_deep_graphs_and_balances_modelled_cost():
    """data.shard_graph_ids with depths: the w deepest graphs land on w different shards, every graph is owned exactly once,
    and the modelled cost (level_cost_nodes * max depth + nodes) is flatter than under the contiguous node-balanced rule."""
    B = D.make_code2_batch(64, 20262)
    cnt, dep = D.graph_node_counts(B), D.graph_depths(B)
    assert dep.shape == cnt.shape and int(dep.max()) == int(B._bi_layer_idx0.max()) + 1
    for w in (2, 4, 8):
        ids = D.shard_graph_ids(cnt, w, dep)
        assert sorted(int(g) for r in ids for g in r) == list(range(64))
        deepest = np.argsort(-dep, kind="stable")[:w]
        owners = {k for k, r in enumerate(ids) for g in deepest if g in set(r.tolist())}
        assert len(owners) == w
        cost = lambda shard: 100.0 * dep[shard].max() + cnt[shard].sum()
        flat = [cost(r) for r in ids]
        ref = [cost(r) for r in D.shard_graph_ids(cnt, w) if len(r)]
        assert max(flat) <= max(ref) + 1e-9
    # fewer graphs than ranks: empty shards are allowed and come back as None from shard_batch
    tiny = D.make_code2_batch(2, 3)
    parts = D.shard_batch(tiny, 4)
    assert sum(p is not None for p in parts) == 2 and sum(p.num_graphs for p in parts if p is not None) == 2


def test_dvae_rows_to_tensor_layout():
    """Text rows -> the int32 [B, n, n] layout dagnn_dvae_rows_build reads (type in column 0, flags behind it, zero padded)."""
    from dagnn_b200 import runtime as rt, _lib
    rows = [[[3], [1, 0], [5, 1, 1]], [[0], [2, 1], [4, 0, 1]]]
    t = rt.dvae_rows_to_tensor(rows, 3)
    assert t.dtype == torch.int32 and tuple(t.shape) == (2, 3, 3)
    assert t[0].tolist() == [[3, 0, 0], [1, 0, 0], [5, 1, 1]] and t[1].tolist() == [[0, 0, 0], [2, 1, 0], [4, 0, 1]]
    with pytest.raises(_lib.DagnnError):
        rt.dvae_rows_to_tensor([[[1]]], 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B = D.make_code2_batch(10, 21)
        mine, rng = sharding.shard_for_rank(B)
        # (1) every rank derives the same partition without communication; shards tile the batch
        sizes = torch.tensor([mine.num_graphs, mine.x.shape[0]])
        allsz = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(allsz, sizes)
        assert sum(int(s[0]) for s in allsz) == 10 and sum(int(s[1]) for s in allsz) == B.x.shape[0]
        # (2) a per-graph "readout" computed on the shard, gathered in rank order == the same on the whole batch
        def per_graph(b):
            return torch.zeros(b.num_graphs, 3).index_add(0, b.batch, torch.stack([b.x[:, 0].float(), b.x[:, 1].float(),
                                                                                    b._bi_layer_idx0.float()], 1))
        full = per_graph(B)
        got = sharding.gather_rows(per_graph(mine), [int(s[0]) for s in allsz])
        all_ids = [D.shard_graph_ids(D.graph_node_counts(B), world, D.graph_depths(B))[r] for r in range(world)]
        assert np.array_equal(np.asarray(rng), all_ids[rank])
        assert torch.equal(sharding.unshard_rows(got, all_ids), full)
        # (3) gradient all-reduce: sum of shard gradients of a sum-loss == full-batch gradient
        torch.manual_seed(0)
        lin = torch.nn.Linear(3, 2)
        lin(per_graph(mine)).sum().backward()
        n = sharding.allreduce_gradients(lin.parameters())
        ref = torch.nn.Linear(3, 2)
        ref.load_state_dict(lin.state_dict())
        ref(full).sum().backward()
        assert n == 8 and torch.allclose(lin.weight.grad, ref.weight.grad, atol=1e-4) and torch.allclose(lin.bias.grad, ref.bias.grad)
        # (4) the same through the flat gradient buffer (every .grad a view into it, one collective, no copies)
        lin2 = torch.nn.Linear(3, 2)
        lin2.load_state_dict(ref.state_dict())
        fg = sharding.FlatGradients(lin2.parameters())
        lin2(per_graph(mine)).sum().backward()
        assert lin2.weight.grad.data_ptr() == fg.flat.data_ptr() and fg.allreduce() == 8
        assert torch.allclose(lin2.weight.grad, ref.weight.grad, atol=1e-4) and torch.allclose(lin2.bias.grad, ref.bias.grad)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2_sharding_and_gradient_allreduce():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_augment_edge2_batch_equals_per_graph_augmentation_then_collation():
    """ogbg-code/utils2.py:31-79 restated per graph (AST edges, then next-token edges between consecutive attributed
    nodes; attrs [0,0] / [1,0]) + PyG collation (node offsets, graphs concatenated in order) == the batch-level helper."""
    g = torch.Generator().manual_seed(3)
    sizes = [1, 7, 2, 12, 5]
    eis, eas, asts, attrs, batch = [], [], [], [], []
    off = 0
    for gi, n in enumerate(sizes):
        ast = torch.stack([torch.randint(0, n, (max(n - 1, 0),), generator=g), torch.arange(1, n)]) if n > 1 else torch.zeros(2, 0, dtype=torch.long)
        is_attr = (torch.rand(n, generator=g) < 0.5).long()
        idx = torch.where(is_attr == 1)[0]
        nt = torch.stack([idx[:-1], idx[1:]]) if idx.numel() >= 2 else torch.zeros(2, 0, dtype=torch.long)
        eis.append(torch.cat([ast, nt], 1) + off)
        eas.append(torch.cat([torch.zeros(ast.shape[1], 2), torch.cat([torch.ones(nt.shape[1], 1), torch.zeros(nt.shape[1], 1)], 1)], 0))
        asts.append(ast + off); attrs.append(is_attr); batch.append(torch.full((n,), gi, dtype=torch.long))
        off += n
    ei, ea = D.augment_edge2_batch(torch.cat(asts, 1), torch.cat(attrs).view(-1, 1), torch.cat(batch))
    assert torch.equal(ei, torch.cat(eis, 1)) and torch.equal(ea, torch.cat(eas, 0))
    assert ei.dtype == torch.int64 and ea.dtype == torch.float32
