"""CPU: the oracle restatement (oracle/dagnn_oracle.py) against the fixtures produced by the reference's own
model files (oracle/gen_golden.py). Pins the oracle; runs everywhere (no reference checkout needed)."""
import numpy as np
import pytest
import torch

from helpers import (DVAE_GOLDEN, OGB_GOLDEN, batch_from_golden, dvae_module_from_meta, load_golden,
                     ogb_module_from_meta, state_dict_cpu)
from oracle import dagnn_oracle as O

TOL = 2e-6   # same ops in the same order as the reference -> normally bit-identical


@pytest.mark.parametrize("name", OGB_GOLDEN)
def test_ogb_oracle_matches_reference_fixture(name):
    z, meta = load_golden(name)
    B = batch_from_golden(z)
    p = state_dict_cpu(ogb_module_from_meta(meta))
    trace = {}
    with torch.no_grad():
        pred, out, H = O.ogb_forward(p, B, num_layers=meta["layers"], bidirectional=meta["bidir"], out_wx=meta["out_wx"],
                                     out_pool_all=bool(meta["pool_all"]), out_pool=meta["pool"], max_seq_len=5,
                                     num_class=meta["num_class"], w_edge_attr=meta["wea"], trace=trace, agg=meta.get("agg", "attn_h"))
    np.testing.assert_allclose(out.numpy(), z["readout"], atol=TOL, rtol=0)
    pred = pred if meta["num_class"] > 0 else torch.stack(pred)
    np.testing.assert_allclose(pred.numpy(), z["pred"], atol=TOL, rtol=0)
    for d in range(2 if meta["bidir"] else 1):
        for i in range(meta["layers"]):
            k = "H_%d_%d" % (d, i)
            if k in z.files:
                np.testing.assert_allclose(H[d][i].numpy(), z[k], atol=TOL, rtol=0)
        # integer schedule: the per-level edge lists the reference built (bit-exact)
        lv = 1
        while "edges_%d_%d" % (d, lv) in z.files:
            ei = B.edge_index[:, trace["edges"][(d, lv)]]
            assert np.array_equal(ei.numpy(), z["edges_%d_%d" % (d, lv)])
            lv += 1
        assert lv == int(B._bi_layer_idx0.max()) + 1


@pytest.mark.parametrize("name", DVAE_GOLDEN)
def test_dvae_oracle_matches_reference_fixture(name):
    z, meta = load_golden(name)
    B = batch_from_golden(z)
    p = state_dict_cpu(dvae_module_from_meta(meta))
    nvt = 8 if meta["kind"] == "NA" else 10
    trace = {}
    with torch.no_grad():
        pool = dict(out_pool_all=meta.get("pool_all", False), out_pool=meta.get("pool", "max"))
        out, H = O.dvae_forward(p, B, num_layers=meta["layers"], bidirectional=meta["bidir"], num_nodes=nvt,
                                vid=(meta["kind"] == "NA"), trace=trace, **pool)
        mu, logvar = O.dvae_encode(p, B, num_layers=meta["layers"], bidirectional=meta["bidir"], num_nodes=nvt,
                                   vid=(meta["kind"] == "NA"), **pool)
    np.testing.assert_allclose(out.numpy(), z["out"], atol=TOL, rtol=0)
    np.testing.assert_allclose(mu.numpy(), z["mu"], atol=TOL, rtol=0)
    np.testing.assert_allclose(logvar.numpy(), z["logvar"], atol=TOL, rtol=0)
    for d in range(2 if meta["bidir"] else 1):
        for i in range(meta["layers"]):
            k = "H_%d_%d" % (d, i)
            if k in z.files:
                np.testing.assert_allclose(H[d][i].numpy(), z[k], atol=TOL, rtol=0)
        lv = 1
        while "edges_%d_%d" % (d, lv) in z.files:
            ei = B.edge_index[:, trace["edges"][(d, lv)]]
            assert np.array_equal(ei.numpy(), z["edges_%d_%d" % (d, lv)])
            lv += 1


def test_top_sort_matches_reference_fixture():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "levels.npz"))
    from dagnn_b200.data import dag_levels_host
    for k in range(12):
        ei, n = z["ei_%d" % k], int(z["n_%d" % k])
        assert np.array_equal(O.top_sort(ei, n).numpy(), z["l0_%d" % k])
        assert np.array_equal(O.top_sort(ei[::-1].copy(), n).numpy(), z["l1_%d" % k])
        # the product's host-side level builder agrees too
        assert np.array_equal(dag_levels_host(ei[0], ei[1], n), z["l0_%d" % k])
        assert np.array_equal(dag_levels_host(ei[1], ei[0], n), z["l1_%d" % k])
        bi = O.add_order_info(torch.from_numpy(ei), n)
        O.assert_order(torch.from_numpy(ei), bi[0][0], bi[0][1])


def test_gru_cell_restatement_matches_torch():
    torch.manual_seed(0)
    cell = torch.nn.GRUCell(12, 20)
    x, h = torch.randn(7, 12), torch.randn(7, 20)
    with torch.no_grad():
        ref = cell(x, h)
        got = O.gru_cell(x, h, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)
        ref0 = cell(x)
        got0 = O.gru_cell(x, None, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)
    assert torch.allclose(ref, got, atol=1e-6) and torch.allclose(ref0, got0, atol=1e-6)
