"""GPU parity at the sizes BASELINE.json names (run with `pytest -m gpu` on the B200 box): full config 2 and config 3 batches
against the CPU oracle, every state tensor and the readout at 1e-4; sharded == unsharded forward; both sweep kernels (the
cluster-resident one and the grid-wide one) on the same inputs; a batch deeper than the default level table; no mutation of G."""
import os

import numpy as np
import pytest
import torch

from helpers import state_dict_cpu

pytestmark = pytest.mark.gpu

ATOL = 1e-4     # north_star: "within 1e-4 fp32"


@pytest.fixture(scope="module")
def dev(built_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(params=["cluster", "grid"])
def sweep_path(request):
    """DAGNN_SWEEP_PATH=grid forces the grid-wide kernel; anything else lets the library pick (cluster-resident when H <= 256)."""
    old = os.environ.get("DAGNN_SWEEP_PATH")
    os.environ["DAGNN_SWEEP_PATH"] = request.param
    yield request.param
    if old is None:
        os.environ.pop("DAGNN_SWEEP_PATH", None)
    else:
        os.environ["DAGNN_SWEEP_PATH"] = old


def _ogb_model(emb, hid, layers, seed=1, **kw):
    from dagnn_b200 import data as D, ogb
    enc = ogb.ASTNodeEncoder(emb, D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
    args = dict(num_layers=layers, bidirectional=True, out_wx=False, out_pool_all=False)
    args.update(kw)
    m = ogb.DAGNN(D.CODE2_NUM_VOCAB, 5, emb, hid, None, encoder=enc, **args)
    D.deterministic_init_(m, seed)
    return m.eval()


def _check_against_oracle(m, B, layers, dev, bidir=True):
    from dagnn_b200 import runtime as rt
    from oracle import dagnn_oracle as O
    with torch.no_grad():
        _, out_ref, H_ref = O.ogb_forward(state_dict_cpu(m), B, num_layers=layers, bidirectional=bidir, heads=False)
    m = m.to(dev)
    G = B.to(dev)
    with torch.no_grad():
        X, Hs, sched = m.node_states(G)
        out = m.readout(G, X, Hs, sched)
    sched.finalize()
    st = rt.states_to_node_order(sched, Hs, m.hidden_dim)
    worst = 0.0
    for d in range(2 if bidir else 1):
        for i in range(layers):
            err = (st[d][i].cpu() - H_ref[d][i]).abs().max().item()
            worst = max(worst, err)
            assert err <= ATOL, "H[%d][%d] max-abs err %g" % (d, i, err)
    np.testing.assert_allclose(out.cpu().numpy(), out_ref.numpy(), atol=ATOL, rtol=0)
    return worst


def test_full_config2_matches_oracle(dev, sweep_path):
    """BASELINE configs[1]: 128 code2-shaped graphs (seed 20262, the bench workload), D = H = 256, 2 layers, bidirectional."""
    from dagnn_b200 import data as D
    B = D.make_code2_batch(128, 20262)
    _check_against_oracle(_ogb_model(256, 256, 2), B, 2, dev)


def test_full_config3_matches_oracle(dev):
    """BASELINE configs[2]: 256 graphs, D = H = 300, 5 layers, bidirectional (the grid-wide kernel: H > 256)."""
    from dagnn_b200 import data as D
    B = D.make_code2_batch(256, 20262)
    _check_against_oracle(_ogb_model(300, 300, 5), B, 5, dev)


@pytest.mark.parametrize("emb,hid,layers,ng", [(256, 256, 2, 40), (64, 64, 3, 33), (48, 100, 2, 17), (200, 136, 1, 9), (16, 30, 2, 5)])
def test_both_kernels_match_oracle(emb, hid, layers, ng, dev, sweep_path):
    """Shapes on both sides of every padding rule of the cluster kernel (H % 16, H % 8, H % 4 != 0, Din != H, one layer)."""
    from dagnn_b200 import data as D
    B = D.make_code2_batch(ng, 300 + hid)
    _check_against_oracle(_ogb_model(emb, hid, layers, seed=hid), B, layers, dev)


def test_unidirectional_and_random_dags_both_kernels(dev, sweep_path):
    from dagnn_b200 import data as D
    B = D.make_random_dag_batch(30, 77, n_hi=60)
    _check_against_oracle(_ogb_model(32, 72, 3, seed=5, bidirectional=False), B, 3, dev, bidir=False)
    B = D.make_random_dag_batch(3, 78, n_hi=12, with_attr=False)
    _check_against_oracle(_ogb_model(24, 24, 2, seed=6, w_edge_attr=False), B, 2, dev)


@pytest.mark.parametrize("w", [2, 4, 8])
def test_sharded_forward_equals_unsharded(w, dev):
    """Graph sharding (sharding.py / data.shard_batch): the readouts of the shards, concatenated in shard order, are the rows
    of the unsharded readout (graphs are independent; 1e-6: nothing but the chunking of levels changes)."""
    from dagnn_b200 import data as D
    B = D.make_code2_batch(48, 4242)
    m = _ogb_model(128, 128, 2).to(dev)
    with torch.no_grad():
        full = m.forward_readout(B.to(dev))
        ranges = D.shard_graph_ids(D.graph_node_counts(B), w, D.graph_depths(B))
        assert sorted(g for r in ranges for g in r) == list(range(48))
        parts = []
        for r in ranges:
            if len(r) == 0:
                continue
            parts.append((list(r), m.forward_readout(D.select_graphs(B, r).to(dev))))
    got = torch.empty_like(full)
    for ids, out in parts:
        got[torch.tensor(ids, device=dev)] = out
    assert (got - full).abs().max().item() <= 1e-6


def test_batch_deeper_than_the_level_table(dev, sweep_path):
    """A 700-node chain next to ordinary graphs: deeper than the default 256-entry level table. The first schedule build
    flags the overflow (nothing downstream touches the unset arrays), the forward is rebuilt with a larger table and matches
    the oracle."""
    from dagnn_b200 import data as D
    n = 700
    small = D.make_code2_batch(3, 9)
    off = int(small.x.shape[0])
    g = torch.Generator().manual_seed(1)
    chain = torch.stack([torch.arange(n - 1), torch.arange(1, n)]) + off
    ids = torch.arange(off + n)
    B = D.DagBatch(
        x=torch.cat([small.x, torch.stack([torch.randint(0, 98, (n,), generator=g), torch.randint(0, 10030, (n,), generator=g)], 1)]),
        node_depth=torch.cat([small.node_depth, torch.arange(n).view(-1, 1)]),
        edge_index=torch.cat([small.edge_index, chain], 1).contiguous(),
        edge_attr=torch.cat([small.edge_attr, torch.zeros(n - 1, 2)]),
        batch=torch.cat([small.batch, torch.full((n,), 3, dtype=torch.long)]),
        _bi_layer_idx0=torch.cat([small._bi_layer_idx0, torch.arange(n)]), _bi_layer_index0=ids.clone(),
        _bi_layer_idx1=torch.cat([small._bi_layer_idx1, torch.arange(n - 1, -1, -1)]), _bi_layer_index1=ids.clone(), num_graphs=4)
    m = _ogb_model(32, 32, 2, seed=3)
    from oracle import dagnn_oracle as O
    with torch.no_grad():
        _, out_ref, _ = O.ogb_forward(state_dict_cpu(m), B, num_layers=2, bidirectional=True, heads=False)
    m = m.to(dev)
    with torch.no_grad():
        out = m.forward_readout(B.to(dev))
        pred = m(B.to(dev))
    np.testing.assert_allclose(out.cpu().numpy(), out_ref.numpy(), atol=ATOL, rtol=0)
    assert len(pred) == 5 and torch.isfinite(pred[0]).all()


def test_forward_does_not_mutate_the_batch(dev):
    """The reference overwrites G.x, sets G.h / G.bi_layer_index and clamps G.node_depth in place (dagnn.py:130-139); callers
    do not rely on it, and this forward leaves every tensor of G untouched."""
    from dagnn_b200 import data as D
    B = D.make_code2_batch(6, 31)
    B.node_depth[0, 0] = 33                       # deeper than the clamp
    G = B.to(dev)
    before = {k: getattr(G, k).clone() for k in G.keys}
    m = _ogb_model(32, 32, 2).to(dev)
    with torch.no_grad():
        m(G)
    assert sorted(G.keys) == sorted(before)
    for k, v in before.items():
        assert torch.equal(getattr(G, k), v), k


def test_out_of_vocabulary_index_is_loud(dev):
    """nn.Embedding raises IndexError for an index outside its table (ogbg-code/utils.py:27); here the row and everything that
    depends on it becomes NaN — never an out-of-bounds read, never a silently wrong number."""
    from dagnn_b200 import data as D
    B = D.make_code2_batch(3, 8)
    B.x[5, 1] = D.CODE2_NUM_NODEATTRS + 7
    m = _ogb_model(32, 32, 2).to(dev)
    with torch.no_grad():
        out = m.forward_readout(B.to(dev))
    g = int(B.batch[5])
    assert torch.isnan(out[g]).any()
    assert torch.isfinite(out[[k for k in range(3) if k != g]]).all()


@pytest.mark.parametrize("fixture,kind", [("na_real_hs501", "NA"), ("bn_real_hs501", "BN")])
def test_dvae_rows_decoded_on_device_bit_exact(fixture, kind, dev):
    """Input side: the real NA / BN rows carried by the fixtures, decoded and collated on the device
    (dagnn_dvae_rows_build) == the host decoders + collation (data.decode_*_row / collate_dvae, which the golden generator checked
    against the reference's own Batch.from_data_list) — every tensor bit-exact — and the forward on it matches the fixture."""
    import json
    from helpers import dvae_module_from_meta, load_golden
    from dagnn_b200 import data as D, runtime as rt
    z, meta = load_golden(fixture)
    rows = json.loads(str(z["rows"]))
    n = 6 if kind == "NA" else 8
    nvt = n + 2
    host = D.collate_dvae([(D.decode_enas_row if kind == "NA" else D.decode_bn_row)(r) for r in rows])
    G = rt.dvae_batch_from_rows(rt.dvae_rows_to_tensor(rows, n).to(dev), kind, nvt)
    for k in ("x", "edge_index", "bi_layer_index", "batch"):
        assert torch.equal(getattr(G, k).cpu(), getattr(host, k)), k
    m = dvae_module_from_meta(meta, dev)
    with torch.no_grad():
        out = m(G)
    np.testing.assert_allclose(out.cpu().numpy(), z["out"], atol=ATOL, rtol=0)
