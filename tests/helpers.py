"""Shared test helpers: golden fixture loading, batch reconstruction, module <-> oracle parameter exchange."""
import json
import os

import numpy as np
import torch

from dagnn_b200 import data as D

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return z, meta


def batch_from_golden(z, num_graphs=None):
    kw = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in_")}
    b = D.DagBatch(**kw)
    b.num_graphs = int(b.batch.max()) + 1 if num_graphs is None else num_graphs
    return b


def state_dict_cpu(module):
    return {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}


def ogb_module_from_meta(meta, device=None, cls=None, enc_cls=None):
    from dagnn_b200 import ogb
    cls = cls or ogb.DAGNN
    enc_cls = enc_cls or ogb.ASTNodeEncoder
    enc = enc_cls(meta["emb"], D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
    m = cls(50, 5, meta["emb"], meta["hid"], None, encoder=enc, w_edge_attr=meta["wea"], num_layers=meta["layers"],
            bidirectional=meta["bidir"], out_wx=meta["out_wx"], out_pool_all=meta["pool_all"], out_pool=meta["pool"],
            num_class=meta["num_class"], agg=meta.get("agg", "attn_h"))
    D.deterministic_init_(m, meta["wseed"])
    m.eval()
    return m.to(device) if device is not None else m


def dvae_module_from_meta(meta, device=None):
    from dagnn_b200 import dvae
    nvt = 8 if meta["kind"] == "NA" else 10
    cls = dvae.DAGNN if meta["kind"] == "NA" else dvae.DAGNN_BN
    m = cls(nvt, meta["hs"], meta["hs"], nvt, nvt, 0, 1, hs=meta["hs"], nz=56, num_nodes=nvt, agg="attn_h",
            num_layers=meta["layers"], bidirectional=meta["bidir"], out_wx=False, out_pool_all=meta.get("pool_all", False),
            out_pool=meta.get("pool", "max"), dropout=0.0)
    D.deterministic_init_(m, meta["wseed"])
    m.eval()
    return m.to(device) if device is not None else m


OGB_GOLDEN = ["ogb_rand_bidir", "ogb_rand_unidir3", "ogb_rand_noattr_mean", "ogb_rand_wx_add_cls", "ogb_code2_small",
              "ogb_code2_h300x5", "ogb_rand_h30_l1_add", "ogb_code2_h72_l3_mean_all", "ogb_rand_attn_pool", "ogb_code2_self_attn"]
DVAE_GOLDEN = ["na_real_hs64", "na_real_hs501", "na_real_bidir_hs48", "bn_real_hs64", "bn_real_hs501",
               "bn_real_unidir_hs40", "na_real_unidir_l3_hs36", "bn_real_pool_all_mean_hs40", "na_real_pool_all_max_hs32"]
