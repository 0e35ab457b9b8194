"""ORACLE — TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.npz by running the REFERENCE'S OWN model files (imported unchanged from
/root/reference on the test-only PyG shim) on seeded inputs with seeded weights. Run it in the build
container (the reference is not present on the GPU box):

    python -m oracle.gen_golden

Every fixture stores the integer inputs, the construction arguments, the weight seed
(`dagnn_b200.data.deterministic_init_`, numpy-based, so weights are regenerated bit-identically anywhere)
and the reference's outputs: forward result, every per-layer state tensor G.h[d][i] (small cases), and the
per-level edge lists the reference built (`lp_edge_index` as passed to the aggregator).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from dagnn_b200 import data as D  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

OGB_CASES = [
    # name, batch builder, kwargs
    dict(name="ogb_rand_bidir", gen=("rand", 6, 7), emb=32, hid=32, layers=2, bidir=True, out_wx=False,
         pool_all=0, pool="max", wea=True, num_class=0, wseed=3),
    dict(name="ogb_rand_unidir3", gen=("rand", 5, 11), emb=24, hid=40, layers=3, bidir=False, out_wx=False,
         pool_all=0, pool="max", wea=True, num_class=0, wseed=4),
    dict(name="ogb_rand_noattr_mean", gen=("rand", 7, 13), emb=16, hid=16, layers=2, bidir=True, out_wx=False,
         pool_all=1, pool="mean", wea=False, num_class=0, wseed=5),
    dict(name="ogb_rand_wx_add_cls", gen=("rand", 4, 17), emb=20, hid=36, layers=2, bidir=True, out_wx=True,
         pool_all=0, pool="add", wea=True, num_class=7, wseed=6),
    dict(name="ogb_code2_small", gen=("code2", 8, 20261), emb=64, hid=64, layers=2, bidir=True, out_wx=False,
         pool_all=0, pool="max", wea=True, num_class=0, wseed=1),
    dict(name="ogb_code2_h300x5", gen=("code2", 3, 20263), emb=300, hid=300, layers=5, bidir=True, out_wx=False,
         pool_all=0, pool="max", wea=True, num_class=0, wseed=2, states=False),
    # hidden width that is not a multiple of 4, a single layer, add pooling
    dict(name="ogb_rand_h30_l1_add", gen=("rand", 5, 19), emb=12, hid=30, layers=1, bidir=True, out_wx=False,
         pool_all=0, pool="add", wea=True, num_class=0, wseed=13),
    # 3 H that is not a multiple of 64, three layers, mean over all nodes (out_wx with out_pool_all does not run in the
    # reference: its head is sized for one direction's width more than the readout produces)
    dict(name="ogb_code2_h72_l3_mean_all", gen=("code2", 4, 20265), emb=48, hid=72, layers=3, bidir=True, out_wx=False,
         pool_all=1, pool="mean", wea=True, num_class=0, wseed=14),
    # out_pool="attn": Linear(d, 1) scores softmaxed over a size-1 dimension = add-pool (dagnn.py:114-117)
    # agg="self_attn_h" (SelfAttnConv, dagnn.py:279-313)
    dict(name="ogb_code2_self_attn", gen=("code2", 4, 20267), emb=40, hid=56, layers=2, bidir=True, out_wx=False,
         pool_all=0, pool="max", wea=True, num_class=0, wseed=19, agg="self_attn_h"),
    dict(name="ogb_rand_attn_pool", gen=("rand", 5, 23), emb=16, hid=24, layers=2, bidir=True, out_wx=False,
         pool_all=0, pool="attn", wea=True, num_class=0, wseed=16),
]

DVAE_CASES = [
    dict(name="na_real_hs64", kind="NA", rows=("final_structures6.txt", 1000, 32), hs=64, layers=2, bidir=False, wseed=7),
    dict(name="na_real_hs501", kind="NA", rows=("final_structures6.txt", 1000, 32), hs=501, layers=2, bidir=False, wseed=8,
         states=False),
    dict(name="na_real_bidir_hs48", kind="NA", rows=("final_structures6.txt", 1500, 16), hs=48, layers=3, bidir=True, wseed=9),
    dict(name="bn_real_hs64", kind="BN", rows=("asia_200k.txt", 0, 48), hs=64, layers=2, bidir=True, wseed=10),
    dict(name="bn_real_hs501", kind="BN", rows=("asia_200k.txt", 0, 128), hs=501, layers=2, bidir=True, wseed=11,
         states=False),
    dict(name="bn_real_unidir_hs40", kind="BN", rows=("asia_200k.txt", 300, 24), hs=40, layers=2, bidir=False, wseed=12),
    dict(name="na_real_unidir_l3_hs36", kind="NA", rows=("final_structures6.txt", 2000, 24), hs=36, layers=3, bidir=False, wseed=15),
    # out_pool_all=True: hg_unify per node, then pooled over all nodes of a graph (dvae/dagnn_bn.py:153-165)
    dict(name="bn_real_pool_all_mean_hs40", kind="BN", rows=("asia_200k.txt", 500, 20), hs=40, layers=2, bidir=True, wseed=17,
         pool_all=True, pool="mean"),
    dict(name="na_real_pool_all_max_hs32", kind="NA", rows=("final_structures6.txt", 2500, 12), hs=32, layers=2, bidir=False, wseed=18,
         pool_all=True, pool="max"),
]


def _batch_arrays(B):
    return {("in_" + k): getattr(B, k).numpy() for k in B.keys}


def _record_edges(model, store):
    """Wrap every aggregator's forward to record the lp_edge_index it receives (per call order)."""
    for name, mod in model.named_modules():
        if name.startswith("node_aggr_") and name.count(".") == 1:
            orig = mod.forward

            def fwd(h, edge_index, *a, _orig=orig, _name=name, **k):
                store.append((_name, None if edge_index is None else edge_index.clone()))
                return _orig(h, edge_index, *a, **k)
            mod.forward = fwd


def _record_cells(model, dirs, layers, store):
    """Forward hooks on the GRU cells: call k of cell (d, i) is level k (loop order d, level, layer)."""
    for d in dirs:
        cells = getattr(model, "cells_%d" % d)
        for i in range(layers):
            cells[i].register_forward_hook(
                lambda mod, inp, out, _k=(d, i): store.setdefault(_k, []).append(out.detach().clone()))


def _states_from_cells(store, lvl_by_dir, hidden):
    out = {}
    for (d, i), outs in store.items():
        lvl = lvl_by_dir[d]
        Hs = torch.zeros(lvl.shape[0], hidden)
        for l, o in enumerate(outs):
            Hs[lvl == l] = o
        out["H_%d_%d" % (d, i)] = Hs.numpy()
    return out


def gen_ogb(case):
    dag, utl = ref_loader.load_ogb()
    from torch_geometric.data import Batch
    kind, ng, seed = case["gen"]
    B = D.make_code2_batch(ng, seed) if kind == "code2" else D.make_random_dag_batch(ng, seed, with_attr=True)
    enc = utl.ASTNodeEncoder(case["emb"], D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
    m = dag.DAGNN(50, 5, case["emb"], case["hid"], None, encoder=enc, w_edge_attr=case["wea"],
                  num_layers=case["layers"], bidirectional=case["bidir"], out_wx=case["out_wx"],
                  out_pool_all=case["pool_all"], out_pool=case["pool"], num_class=case["num_class"], agg=case.get("agg", "attn_h"))
    D.deterministic_init_(m, case["wseed"])
    m.eval()
    G = Batch(batch=B.batch.clone(), x=B.x.clone(), edge_index=B.edge_index.clone(),
              edge_attr=B.edge_attr.clone(), node_depth=B.node_depth.clone(),
              _bi_layer_idx0=B._bi_layer_idx0.clone(), _bi_layer_index0=B._bi_layer_index0.clone(),
              _bi_layer_idx1=B._bi_layer_idx1.clone(), _bi_layer_index1=B._bi_layer_index1.clone())
    calls = []
    _record_edges(m, calls)
    # capture the readout (input of the heads) with a forward pre-hook on the dropout module
    ro = {}
    m.dropout.register_forward_hook(lambda mod, i, o: ro.__setitem__("out", o.detach().clone()))
    cell_out = {}
    dirs = [0, 1] if case["bidir"] else [0]
    _record_cells(m, dirs, case["layers"], cell_out)
    with torch.no_grad():
        pred = m(G)
    arrs = _batch_arrays(B)
    arrs["readout"] = ro["out"].numpy()
    if case["num_class"] > 0:
        arrs["pred"] = pred.numpy()
    else:
        arrs["pred"] = torch.stack(pred).numpy()
    if case.get("states", True):
        arrs.update(_states_from_cells(cell_out, [B._bi_layer_idx0, B._bi_layer_idx1], case["hid"]))
    # edge lists: layer-0 aggregator of each direction is called once per level > 0, in level order
    for d in ([0, 1] if case["bidir"] else [0]):
        lv = 1
        for name, ei in calls:
            if name == "node_aggr_%d.0" % d:
                arrs["edges_%d_%d" % (d, lv)] = ei.numpy()
                lv += 1
    meta = {k: v for k, v in case.items()}
    arrs["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, case["name"] + ".npz"), **arrs)
    print(case["name"], "N=%d E=%d" % (B.x.shape[0], B.edge_index.shape[1]), "readout", arrs["readout"].shape)


def gen_dvae(case):
    dagnn, dagnn_bn, batch_mod = ref_loader.load_dvae()
    from torch_geometric.data import Data
    fname, start, count = case["rows"]
    rows = [r for r, _y in D.read_dvae_rows(os.path.join(ref_loader.reference_root(), "dvae", "data", fname), start, count)]
    dec = D.decode_enas_row if case["kind"] == "NA" else D.decode_bn_row
    graphs = [dec(r) for r in rows]
    nvt = 8 if case["kind"] == "NA" else 10
    cls = dagnn.DAGNN if case["kind"] == "NA" else dagnn_bn.DAGNN_BN
    # ctor exactly as dvae/train.py:160-172
    m = cls(nvt, case["hs"], case["hs"], nvt, nvt, 0, 1, hs=case["hs"], nz=56, num_nodes=nvt,
            agg="attn_h", num_layers=case["layers"], bidirectional=case["bidir"], out_wx=False,
            out_pool_all=case.get("pool_all", False), out_pool=case.get("pool", "max"), dropout=0.0)
    D.deterministic_init_(m, case["wseed"])
    m.eval()
    data_list = [Data(x=g.x.clone(), edge_index=g.edge_index.clone(), bi_layer_index=g.bi_layer_index.clone())
                 for g in graphs]
    b = batch_mod.Batch.from_data_list(data_list)      # the reference's own collate (dvae/batch.py:26)
    mine = D.collate_dvae(graphs)
    for k in ("x", "edge_index", "bi_layer_index", "batch"):
        assert torch.equal(getattr(b, k), getattr(mine, k)), k
    calls = []
    _record_edges(m, calls)
    cell_out = {}
    _record_cells(m, [0, 1] if case["bidir"] else [0], case["layers"], cell_out)
    with torch.no_grad():
        out = m(b)
        mu, logvar = m.fc1(out), m.fc2(out)
    arrs = {"rows": np.array(json.dumps(rows)), "out": out.numpy(), "mu": mu.numpy(), "logvar": logvar.numpy()}
    arrs.update(_batch_arrays(mine))
    if case.get("states", True):
        arrs.update(_states_from_cells(cell_out, [mine.bi_layer_index[0][0], mine.bi_layer_index[1][0]], case["hs"]))
    for d in ([0, 1] if case["bidir"] else [0]):
        lv = 1
        for name, ei in calls:
            if name == "node_aggr_%d.0" % d:
                arrs["edges_%d_%d" % (d, lv)] = ei.numpy()
                lv += 1
    arrs["meta"] = np.array(json.dumps(case))
    np.savez_compressed(os.path.join(OUT, case["name"] + ".npz"), **arrs)
    print(case["name"], "N=%d E=%d" % (mine.x.shape[0], mine.edge_index.shape[1]), "out", arrs["out"].shape)


def gen_levels():
    """top_sort / add_order_info of the reference (src/utils_dag.py) on seeded random DAGs + real rows."""
    ud = ref_loader.load_utils_dag()
    arrs = {}
    rng = np.random.default_rng(99)
    for k in range(12):
        n = int(rng.integers(1, 40))
        iu = np.triu_indices(n, 1)
        keep = rng.random(len(iu[0])) < rng.uniform(0.05, 0.5)
        perm = rng.permutation(n)                        # node ids need not be topologically ordered
        ei = np.stack([perm[iu[0][keep]], perm[iu[1][keep]]]).astype(np.int64)
        l0 = ud.top_sort(ei, n).numpy()
        l1 = ud.top_sort(ei[::-1].copy(), n).numpy()
        arrs["ei_%d" % k], arrs["n_%d" % k], arrs["l0_%d" % k], arrs["l1_%d" % k] = ei, np.array(n), l0, l1
    np.savez_compressed(os.path.join(OUT, "levels.npz"), **arrs)
    print("levels: 12 graphs")


def main():
    if not ref_loader.available():
        raise SystemExit("reference checkout not found (set $DAGNN_REFERENCE)")
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])            # optional: names of the fixtures to (re)generate
    if not only:
        gen_levels()
    for c in OGB_CASES:
        if not only or c["name"] in only:
            gen_ogb(c)
    for c in DVAE_CASES:
        if not only or c["name"] in only:
            gen_dvae(c)


if __name__ == "__main__":
    main()
