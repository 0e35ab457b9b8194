"""GPU parity (run with `pytest -m gpu` on the B200 box): the CUDA path, reached through the C ABI of
libdagnn_sm100.so, against (1) the golden fixtures produced by the reference's own model files
(oracle/gen_golden.py) and (2) the CPU oracle restatement on fresh seeded inputs.

Bars (BASELINE.json north_star): integer schedule arrays bit-exact; fp32 states / readouts / outputs within
ATOL = 1e-4 absolute.
"""
import numpy as np
import pytest
import torch

from helpers import (DVAE_GOLDEN, OGB_GOLDEN, batch_from_golden, dvae_module_from_meta, load_golden,
                     ogb_module_from_meta, state_dict_cpu)

pytestmark = pytest.mark.gpu

ATOL = 1e-4     # north_star: "within 1e-4 fp32"


@pytest.fixture(scope="module")
def dev(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _check_schedule_against_golden(sched, z, B_cpu, dirs, lvl_arrays):
    """per-level node lists == boolean-mask select of the reference; per-level edge lists == the
    lp_edge_index the reference handed to its aggregator (bit-exact)."""
    L = int(lvl_arrays[0].max()) + 1
    for d in range(dirs):
        assert sched.num_levels[d] == int(lvl_arrays[d].max()) + 1
        ids = torch.arange(lvl_arrays[d].shape[0])
        for l in range(L):
            want = ids[lvl_arrays[d] == l]
            got = sched.level_nodes(d, l).cpu()
            assert torch.equal(got, want), (d, l)
            key = "edges_%d_%d" % (d, l)
            if l > 0 and key in z.files:
                eid = sched.level_edges(d, l).cpu()
                assert np.array_equal(B_cpu.edge_index[:, eid].numpy(), z[key]), (d, l)


@pytest.mark.parametrize("name", OGB_GOLDEN)
def test_ogb_module_matches_reference_fixture(name, dev):
    from dagnn_b200 import runtime as rt
    z, meta = load_golden(name)
    B = batch_from_golden(z)
    m = ogb_module_from_meta(meta, dev)
    G = B.to(dev)
    dirs = 2 if meta["bidir"] else 1
    with torch.no_grad():
        X, Hs, sched = m.node_states(G)
        out = m.readout(G, X, Hs, sched)
        pred = m(G)
    _check_schedule_against_golden(sched, z, B, dirs, [B._bi_layer_idx0, B._bi_layer_idx1])
    states = rt.states_to_node_order(sched, Hs, meta["hid"])
    for d in range(dirs):
        for i in range(meta["layers"]):
            k = "H_%d_%d" % (d, i)
            if k in z.files:
                np.testing.assert_allclose(states[d][i].cpu().numpy(), z[k], atol=ATOL, rtol=0, err_msg=k)
    np.testing.assert_allclose(out.cpu().numpy(), z["readout"], atol=ATOL, rtol=0)
    pred = pred if meta["num_class"] > 0 else torch.stack(pred)
    np.testing.assert_allclose(pred.cpu().numpy(), z["pred"], atol=5 * ATOL, rtol=0)   # heads: 1000-term fp32 dot products


@pytest.mark.parametrize("name", DVAE_GOLDEN)
def test_dvae_module_matches_reference_fixture(name, dev):
    from dagnn_b200 import runtime as rt
    z, meta = load_golden(name)
    B = batch_from_golden(z)
    m = dvae_module_from_meta(meta, dev)
    dirs = 2 if meta["bidir"] else 1
    with torch.no_grad():
        G = B.to(dev)
        X, Hs, sched = m.node_states(G)
        out = m(B)                       # forward moves the batch itself (dvae/dagnn.py:102)
        mu, logvar = m.encode([G])
    _check_schedule_against_golden(sched, z, B, dirs, [B.bi_layer_index[0][0], B.bi_layer_index[1][0]])
    states = rt.states_to_node_order(sched, Hs, meta["hs"])
    for d in range(dirs):
        for i in range(meta["layers"]):
            k = "H_%d_%d" % (d, i)
            if k in z.files:
                np.testing.assert_allclose(states[d][i].cpu().numpy(), z[k], atol=ATOL, rtol=0, err_msg=k)
    np.testing.assert_allclose(out.cpu().numpy(), z["out"], atol=ATOL, rtol=0)
    np.testing.assert_allclose(mu.cpu().numpy(), z["mu"], atol=ATOL, rtol=0)
    np.testing.assert_allclose(logvar.cpu().numpy(), z["logvar"], atol=ATOL, rtol=0)


# ------------------------------------------------------------------ fresh seeded inputs vs the CPU oracle
OGB_ORACLE_CASES = [
    # graphs, seed, emb, hid, layers, bidir, kind
    (24, 101, 64, 64, 2, True, "code2"),
    (12, 102, 48, 100, 3, True, "rand"),       # H not a multiple of 16 or 128, Din != H
    (16, 103, 128, 130, 2, False, "rand"),     # two unit slices, second nearly empty
    (10, 104, 256, 256, 2, True, "code2"),     # config-2 model at a small batch
    (1, 105, 32, 32, 2, True, "code2"),        # single-graph batch
]


@pytest.mark.parametrize("ng,seed,emb,hid,layers,bidir,kind", OGB_ORACLE_CASES)
def test_ogb_module_matches_oracle(ng, seed, emb, hid, layers, bidir, kind, dev):
    from dagnn_b200 import data as D, ogb, runtime as rt
    from oracle import dagnn_oracle as O
    B = D.make_code2_batch(ng, seed) if kind == "code2" else D.make_random_dag_batch(ng, seed, n_hi=40)
    enc = ogb.ASTNodeEncoder(emb, D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
    m = ogb.DAGNN(50, 5, emb, hid, None, encoder=enc, num_layers=layers, bidirectional=bidir, out_wx=False,
                  out_pool_all=False)
    D.deterministic_init_(m, seed)
    m.eval()
    p = state_dict_cpu(m)
    trace = {}
    with torch.no_grad():
        _, out_ref, H_ref = O.ogb_forward(p, B, num_layers=layers, bidirectional=bidir, heads=False, trace=trace)
    m = m.to(dev)
    G = B.to(dev)
    with torch.no_grad():
        X, Hs, sched = m.node_states(G)
        out = m.readout(G, X, Hs, sched)
    dirs = 2 if bidir else 1
    for d in range(dirs):
        for l in range(sched.num_levels[0]):
            assert torch.equal(sched.level_nodes(d, l).cpu(), trace["nodes"][(d, l)]), (d, l)
            if l > 0:
                assert torch.equal(sched.level_edges(d, l).cpu(), trace["edges"][(d, l)]), (d, l)
    states = rt.states_to_node_order(sched, Hs, hid)
    for d in range(dirs):
        for i in range(layers):
            err = (states[d][i].cpu() - H_ref[d][i]).abs().max().item()
            assert err <= ATOL, "H[%d][%d] max-abs err %g" % (d, i, err)
    np.testing.assert_allclose(out.cpu().numpy(), out_ref.numpy(), atol=ATOL, rtol=0)


@pytest.mark.parametrize("kind,hs,layers,bidir", [("NA", 72, 2, False), ("NA", 56, 3, True), ("BN", 96, 2, True),
                                                  ("BN", 501, 2, True), ("NA", 501, 2, False)])
def test_dvae_module_matches_oracle(kind, hs, layers, bidir, dev):
    from dagnn_b200 import data as D, dvae
    from oracle import dagnn_oracle as O
    nvt = 8 if kind == "NA" else 10
    B = D.make_random_dvae_batch(20, 7 + hs, kind)
    cls = dvae.DAGNN if kind == "NA" else dvae.DAGNN_BN
    m = cls(nvt, hs, hs, nvt, nvt, 0, 1, hs=hs, nz=56, num_nodes=nvt, num_layers=layers, bidirectional=bidir)
    D.deterministic_init_(m, hs)
    m.eval()
    p = state_dict_cpu(m)
    with torch.no_grad():
        out_ref, _ = O.dvae_forward(p, B, num_layers=layers, bidirectional=bidir, num_nodes=nvt, vid=(kind == "NA"))
        mu_ref, lv_ref = O.dvae_encode(p, B, num_layers=layers, bidirectional=bidir, num_nodes=nvt, vid=(kind == "NA"))
    m = m.to(dev)
    with torch.no_grad():
        out = m(B)
        mu, lv = m.encode([B.to(dev)])
    np.testing.assert_allclose(out.cpu().numpy(), out_ref.numpy(), atol=ATOL, rtol=0)
    np.testing.assert_allclose(mu.cpu().numpy(), mu_ref.numpy(), atol=ATOL, rtol=0)
    np.testing.assert_allclose(lv.cpu().numpy(), lv_ref.numpy(), atol=ATOL, rtol=0)


# ------------------------------------------------------------------ full-size, size-independent properties
def test_full_size_properties_config2(dev):
    """BASELINE config 2 shape (B=128, D=H=256, 2 layers, bidirectional): the oracle takes seconds per
    forward there, so check properties instead: (1) graphs are independent — the forward of the whole batch
    equals, graph by graph, the forward of two half batches (bit-exact: same kernels, same per-node
    arithmetic order); (2) permuting the graphs permutes the readout rows (1e-6: tile boundaries move);
    (3) determinism: two runs are bit-identical; (4) all states finite, in (-1, 1) (GRU output range)."""
    from dagnn_b200 import data as D, ogb
    ng = 128
    B = D.make_code2_batch(ng, 20262)
    enc = ogb.ASTNodeEncoder(256, D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
    m = ogb.DAGNN(D.CODE2_NUM_VOCAB, 5, 256, 256, None, encoder=enc, num_layers=2, bidirectional=True, out_wx=False,
                  out_pool_all=False)
    D.deterministic_init_(m, 1)
    m = m.eval().to(dev)
    with torch.no_grad():
        full = m.forward_readout(B.to(dev))
        again = m.forward_readout(B.to(dev))
        X, Hs, sched = m.node_states(B.to(dev))
    assert full.shape == (ng, 1024)
    assert torch.equal(full, again)
    assert torch.isfinite(Hs[..., :256]).all() and Hs[..., :256].abs().max().item() < 1.0
    halves = D.split_batch(B, [range(0, 50), range(50, ng)])
    with torch.no_grad():
        parts = torch.cat([m.forward_readout(h.to(dev)) for h in halves])
    assert (parts - full).abs().max().item() <= 1e-6
    perm = np.random.default_rng(0).permutation(ng)
    Bp = D.select_graphs(B, perm)
    with torch.no_grad():
        outp = m.forward_readout(Bp.to(dev))
    assert (outp - full[torch.from_numpy(perm).to(dev)]).abs().max().item() <= 1e-6


def test_edge_cases(dev):
    """Quirks the reference relies on: an edge whose source sits at the same / a higher level
    reads zeros but keeps its softmax mass (Q1), level-0 nodes ignore in-edges (Q2), duplicate edges (Q4),
    isolated nodes, graph with a single node, batch without edges."""
    from dagnn_b200 import data as D, ogb, runtime as rt
    from oracle import dagnn_oracle as O
    # graph A: chain 0->1->2 plus "next-token" edges 2->1 (higher level -> lower), 0->0 style same-level 1->1 is
    # not allowed in a DAG level sense but extra edges are arbitrary: 3 (isolated, level 0) -> 2, duplicate 0->1
    ei = torch.tensor([[0, 1, 2, 3, 0, 1, 1], [1, 2, 1, 2, 1, 1, 3]])   # last: in-edge of a level-0 node
    l0 = torch.tensor([0, 1, 2, 0, 0])          # node 4: second graph, single node
    l1 = torch.tensor([2, 1, 0, 0, 0])
    ids = torch.arange(5)
    B = D.DagBatch(x=torch.tensor([[1, 2], [3, 4], [5, 6], [7, 8], [9, 10]]), node_depth=torch.tensor([[0], [1], [25], [3], [0]]),
                   edge_index=ei, edge_attr=torch.tensor([[0., 0.], [0., 0.], [1., 0.], [1., 1.], [0., 0.], [1., 0.], [1., 0.]]),
                   batch=torch.tensor([0, 0, 0, 0, 1]), _bi_layer_idx0=l0, _bi_layer_index0=ids.clone(), _bi_layer_idx1=l1,
                   _bi_layer_index1=ids.clone(), num_graphs=2)
    for hid in (20, 36):
        enc = ogb.ASTNodeEncoder(20, D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
        m = ogb.DAGNN(50, 5, 20, hid, None, encoder=enc, num_layers=2, bidirectional=True, out_wx=False, out_pool_all=False)
        D.deterministic_init_(m, 3)
        m.eval()
        with torch.no_grad():
            _, out_ref, H_ref = O.ogb_forward(state_dict_cpu(m), B, num_layers=2, bidirectional=True, heads=False)
        m = m.to(dev)
        with torch.no_grad():
            X, Hs, sched = m.node_states(B.to(dev))
            out = m.readout(B.to(dev), X, Hs, sched)
        st = rt.states_to_node_order(sched, Hs, hid)
        for d in range(2):
            for i in range(2):
                np.testing.assert_allclose(st[d][i].cpu().numpy(), H_ref[d][i].numpy(), atol=ATOL, rtol=0)
        np.testing.assert_allclose(out.cpu().numpy(), out_ref.numpy(), atol=ATOL, rtol=0)
    # a batch without any edge: one level, no aggregation
    B2 = D.DagBatch(x=B.x[:3], node_depth=B.node_depth[:3], edge_index=torch.zeros(2, 0, dtype=torch.long),
                    edge_attr=torch.zeros(0, 2), batch=torch.tensor([0, 1, 2]), _bi_layer_idx0=torch.zeros(3, dtype=torch.long),
                    _bi_layer_index0=torch.arange(3), _bi_layer_idx1=torch.zeros(3, dtype=torch.long),
                    _bi_layer_index1=torch.arange(3), num_graphs=3)
    m = m.cpu()
    with torch.no_grad():
        _, out_ref, _ = O.ogb_forward(state_dict_cpu(m), B2, num_layers=2, bidirectional=True, heads=False)
        out = m.to(dev).forward_readout(B2.to(dev))
    np.testing.assert_allclose(out.cpu().numpy(), out_ref.numpy(), atol=ATOL, rtol=0)


def test_errors_are_loud(dev):
    """No CPU path, bad inputs raise (error behaviour of the boundary)."""
    from dagnn_b200 import data as D, ogb, _lib
    B = D.make_code2_batch(2, 5)
    enc = ogb.ASTNodeEncoder(16, D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
    m = ogb.DAGNN(50, 5, 16, 16, None, encoder=enc, out_wx=False, out_pool_all=False)
    with pytest.raises(_lib.DagnnError):
        with torch.no_grad():
            m(B)                                        # CPU tensors: refused, not computed on the host
    m = m.to(dev)
    out = m(B.to(dev))                                  # grad mode: the autograd path (dagnn_b200.autograd), same numbers
    with torch.no_grad():
        out2 = m(B.to(dev))
    assert out[0].requires_grad and torch.equal(out[0].detach(), out2[0])
    bad = B.clone()
    bad.edge_index[0, 0] = 10 ** 6
    with pytest.raises(_lib.DagnnError):
        with torch.no_grad():
            m(bad.to(dev))
    with pytest.raises(ValueError):
        ogb.DAGNN(50, 5, 32, 16, None, encoder=enc, agg_x=True)        # dagnn.py:27-28
    n0 = _lib.launch_count()
    with torch.no_grad():
        m(B.to(dev))
    assert _lib.launch_count() > n0


def _bipartite_batch(n_src, n_dst, seed):
    """One DAG: every one of n_dst sinks has an in-edge from every one of n_src sources (level 0 -> level 1). Forward
    direction: n_dst nodes with n_src in-edges each; reverse direction: n_src nodes with n_dst in-edges each."""
    from dagnn_b200 import data as D
    g = torch.Generator().manual_seed(seed)
    n = n_src + n_dst
    src = torch.arange(n_src).repeat_interleave(n_dst)
    dst = n_src + torch.arange(n_dst).repeat(n_src)
    ei = torch.stack([src, dst])
    l0 = torch.cat([torch.zeros(n_src, dtype=torch.long), torch.ones(n_dst, dtype=torch.long)])
    l1 = 1 - l0
    ids = torch.arange(n)
    return D.DagBatch(x=torch.stack([torch.randint(0, 98, (n,), generator=g), torch.randint(0, 10030, (n,), generator=g)], 1),
                      node_depth=torch.randint(0, 25, (n, 1), generator=g), edge_index=ei,
                      edge_attr=torch.randint(0, 2, (ei.shape[1], 2), generator=g).float(), batch=torch.zeros(n, dtype=torch.long),
                      _bi_layer_idx0=l0, _bi_layer_index0=ids.clone(), _bi_layer_idx1=l1, _bi_layer_index1=ids.clone(), num_graphs=1)


@pytest.mark.parametrize("n_src,n_dst,hid", [(12, 7300, 32), (40, 300, 64), (3, 70, 520), (14, 100, 64), (11, 200, 40),
                                             (1500, 20, 32)])
def test_long_in_edge_lists_and_wide_states(n_src, n_dst, hid, dev):
    """Gate-phase corner paths: in-edge lists longer than the per-warp limit (whole-CTA aggregation), longer than one
    32-edge round, more long lists in one CTA than its cooperative queue holds (7300 sinks with 12 in-edges each over 148
    CTAs: a warp then walks the list alone), lists of thousands of edges (reverse direction), and hidden states wider than
    one 512-column pass. Small steps list their long in-edge lists a phase ahead and give each a CTA of its own: 40 such
    nodes (the other CTAs take the ordinary rows), 100 (more than half the grid: everybody takes rows too), 200 (more than the
    list holds: back to collecting them while dealing rows); 20 sinks fed by 1500 sources each in a step of 1520 rows."""
    from dagnn_b200 import data as D, ogb, runtime as rt
    from oracle import dagnn_oracle as O
    B = _bipartite_batch(n_src, n_dst, 5)
    emb = 24
    enc = ogb.ASTNodeEncoder(emb, D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
    m = ogb.DAGNN(50, 5, emb, hid, None, encoder=enc, num_layers=2, bidirectional=True, out_wx=False, out_pool_all=False)
    D.deterministic_init_(m, 9)
    m.eval()
    with torch.no_grad():
        _, out_ref, H_ref = O.ogb_forward(state_dict_cpu(m), B, num_layers=2, bidirectional=True, heads=False)
    m = m.to(dev)
    with torch.no_grad():
        X, Hs, sched = m.node_states(B.to(dev))
        out = m.readout(B.to(dev), X, Hs, sched)
    st = rt.states_to_node_order(sched, Hs, hid)
    for d in range(2):
        for i in range(2):
            err = (st[d][i].cpu() - H_ref[d][i]).abs().max().item()
            assert err <= ATOL, "H[%d][%d] max-abs err %g" % (d, i, err)
    np.testing.assert_allclose(out.cpu().numpy(), out_ref.numpy(), atol=ATOL, rtol=0)


def test_dag_levels_on_device_bit_exact(dev):
    """Input side: longest-path levels computed on the device == the reference's `top_sort` /
    `add_order_info_01` (golden level arrays of 12 graphs generated from src/utils_dag.py, tests/golden/levels.npz) and ==
    the host restatement on whole batches: code2-shaped ASTs (levels on the AST edges only), a 3000-node chain (deeper
    than the default number of passes: retried), no edges at all; a cycle raises like the host version."""
    from dagnn_b200 import data as D, runtime as rt, _lib
    import os
    from helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "levels.npz"))
    for k in range(12):
        ei, n = z["ei_%d" % k], int(z["n_%d" % k])
        l0, l1 = rt.dag_levels(torch.from_numpy(ei).to(dev), n)
        assert l0.dtype == torch.int64 and np.array_equal(l0.cpu().numpy(), z["l0_%d" % k])
        assert np.array_equal(l1.cpu().numpy(), z["l1_%d" % k])
    B = D.make_code2_batch(24, 11)
    ast = B.edge_index[:, B.edge_attr[:, 0] == 0]                    # ogb/io/read_graph_pyg.py:51: before augment_edge2
    n = int(B.x.shape[0])
    l0, l1 = rt.dag_levels(ast.to(dev), n)
    assert torch.equal(l0.cpu(), B._bi_layer_idx0) and torch.equal(l1.cpu(), B._bi_layer_idx1)
    assert np.array_equal(l0.cpu().numpy(), D.dag_levels_host(ast[0].numpy(), ast[1].numpy(), n))
    assert np.array_equal(l1.cpu().numpy(), D.dag_levels_host(ast[1].numpy(), ast[0].numpy(), n))
    chain = torch.stack([torch.arange(2999), torch.arange(1, 3000)])
    l0, l1 = rt.dag_levels(chain.to(dev), 3000)
    assert np.array_equal(l0.cpu().numpy(), np.arange(3000)) and np.array_equal(l1.cpu().numpy(), np.arange(3000)[::-1])
    l0, l1 = rt.dag_levels(torch.zeros(2, 0, dtype=torch.long, device=dev), 5)
    assert int(l0.abs().sum()) == 0 and int(l1.abs().sum()) == 0
    with pytest.raises(ValueError):
        rt.dag_levels(torch.tensor([[0, 1, 2], [1, 2, 0]], device=dev), 3)
    with pytest.raises(_lib.DagnnError):
        rt.dag_levels(torch.tensor([[0, 7], [1, 2]], device=dev), 3)
