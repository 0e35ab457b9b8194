"""ORACLE — TEST INFRASTRUCTURE ONLY.

Auxiliary fixtures produced by the reference's own files (imported unchanged from /root/reference, the PyG shim only where
the file needs the import to succeed). Run in the build container:

    python -m oracle.gen_aux_golden

  tests/golden/state_dict_shapes.json   parameter names -> shapes of the reference's modules
                                        (ogbg-code/model/dagnn.py DAGNN, dvae/dagnn.py DAGNN, dvae/dagnn_bn.py DAGNN_BN) for
                                        the constructor arguments stored next to them: the checkpoint contract
                                        (utils2.py:86-103, dvae/util.py:41-63).
  tests/golden/augment_edge2.npz        `augment_edge2` (ogbg-code/utils2.py:31-79) applied per graph, then collated the
                                        way PyG's Batch does (edge_index offset by the running node count): inputs + outputs.
"""
import importlib
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

OGB_CTORS = [
    dict(num_vocab=50, max_seq_len=5, emb_dim=32, hidden_dim=32, out_dim=None, num_layers=2, bidirectional=True, out_wx=False,
         out_pool_all=False),
    dict(num_vocab=50, max_seq_len=5, emb_dim=24, hidden_dim=40, out_dim=None, num_layers=3, bidirectional=False, out_wx=True,
         out_pool_all=False, num_class=7),
    dict(num_vocab=50, max_seq_len=5, emb_dim=16, hidden_dim=16, out_dim=None, num_layers=1, bidirectional=True, w_edge_attr=False,
         out_wx=False, out_pool_all=True, out_pool="mean"),
]
DVAE_CTORS = [
    dict(kind="NA", hs=64, num_layers=2, bidirectional=False),
    dict(kind="NA", hs=48, num_layers=3, bidirectional=True),
    dict(kind="BN", hs=64, num_layers=2, bidirectional=True),
    dict(kind="BN", hs=40, num_layers=2, bidirectional=False),
]


def gen_state_dicts():
    dag, utl = ref_loader.load_ogb()
    out = {"ogb": [], "dvae": []}
    for kw in OGB_CTORS:
        enc = utl.ASTNodeEncoder(kw["emb_dim"], D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
        m = dag.DAGNN(encoder=enc, **kw)
        out["ogb"].append({"ctor": kw, "state_dict": {k: list(v.shape) for k, v in m.state_dict().items()}})
    for mod in ("model.dagnn", "model", "utils"):          # dvae/ has its own top-level `dagnn` / `util` modules
        sys.modules.pop(mod, None)
    dagnn, dagnn_bn, _ = ref_loader.load_dvae()
    for kw in DVAE_CTORS:
        nvt = 8 if kw["kind"] == "NA" else 10
        cls = dagnn.DAGNN if kw["kind"] == "NA" else dagnn_bn.DAGNN_BN
        m = cls(nvt, kw["hs"], kw["hs"], nvt, nvt, 0, 1, hs=kw["hs"], nz=56, num_nodes=nvt, agg="attn_h", num_layers=kw["num_layers"],
                bidirectional=kw["bidirectional"], out_wx=False, out_pool_all=False, out_pool="max", dropout=0.0)
        out["dvae"].append({"ctor": kw, "state_dict": {k: list(v.shape) for k, v in m.state_dict().items()}})
    with open(os.path.join(OUT, "state_dict_shapes.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("state_dict_shapes.json:", [len(e["state_dict"]) for e in out["ogb"] + out["dvae"]], "entries")


class _Data(object):
    pass


def gen_augment_edge2():
    root = ref_loader.reference_root()
    sys.path.insert(0, os.path.join(root, "ogbg-code"))
    utils2 = importlib.import_module("utils2")
    rng = np.random.default_rng(77)
    arrs, off = {}, 0
    eis, eas, batch, attributed, ast = [], [], [], [], []
    ng = 9
    for g in range(ng):
        n = int(rng.integers(1, 30))
        parent, _ = D._random_ast(rng, n)
        ei_ast = torch.from_numpy(np.stack([parent[1:], np.arange(1, n)]).astype(np.int64)) if n > 1 else torch.zeros(2, 0, dtype=torch.long)
        att = torch.from_numpy((rng.random(n) < (0.0 if g == 3 else 0.5)).astype(np.int64)).view(-1, 1)   # graph 3: none attributed
        d = _Data()
        d.edge_index, d.node_is_attributed = ei_ast.clone(), att.clone()
        d = utils2.augment_edge2(d)                       # the reference, per graph (main_pyg.py:235 as a dataset transform)
        eis.append(d.edge_index + off); eas.append(d.edge_attr)
        ast.append(ei_ast + off); attributed.append(att.view(-1)); batch.append(torch.full((n,), g, dtype=torch.long))
        off += n
    arrs["in_edge_index_ast"] = torch.cat(ast, 1).numpy()
    arrs["in_node_is_attributed"] = torch.cat(attributed).numpy()
    arrs["in_batch"] = torch.cat(batch).numpy()
    arrs["edge_index"] = torch.cat(eis, 1).numpy()          # PyG collation: cat along dim -1 with node offsets
    arrs["edge_attr"] = torch.cat(eas, 0).numpy()
    np.savez_compressed(os.path.join(OUT, "augment_edge2.npz"), **arrs)
    print("augment_edge2.npz: %d graphs, %d nodes, %d -> %d edges" % (ng, off, arrs["in_edge_index_ast"].shape[1], arrs["edge_index"].shape[1]))


if __name__ == "__main__":
    if not ref_loader.available():
        raise SystemExit("reference checkout not found (set $DAGNN_REFERENCE)")
    gen_augment_edge2()
    gen_state_dicts()
