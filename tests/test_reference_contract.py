"""CPU: the contract with the reference that does not need a GPU — checkpoint compatibility (parameter names and shapes of
the reference's own modules, recorded by oracle/gen_aux_golden.py, load with strict=True) and `augment_edge2`
(ogbg-code/utils2.py:31-79, run per graph by the reference and collated) against the batch-level helper."""
import json
import os

import numpy as np
import pytest
import torch

from dagnn_b200 import data as D, dvae, ogb
from helpers import GOLDEN


def _shapes():
    with open(os.path.join(GOLDEN, "state_dict_shapes.json")) as f:
        return json.load(f)


def _fake_state_dict(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(*s, generator=g) if len(s) else torch.randn((), generator=g) for k, s in shapes.items()}


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_reference_ogb_state_dict_loads_strict(idx):
    e = _shapes()["ogb"][idx]
    kw = dict(e["ctor"])
    enc = ogb.ASTNodeEncoder(kw["emb_dim"], D.CODE2_NUM_NODETYPES, D.CODE2_NUM_NODEATTRS, D.CODE2_MAX_DEPTH)
    m = ogb.DAGNN(encoder=enc, **kw)
    mine = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert mine == e["state_dict"]                      # same names, same shapes, nothing extra
    sd = _fake_state_dict(e["state_dict"], idx)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_reference_dvae_state_dict_loads_strict(idx):
    e = _shapes()["dvae"][idx]
    kw = e["ctor"]
    nvt = 8 if kw["kind"] == "NA" else 10
    cls = dvae.DAGNN if kw["kind"] == "NA" else dvae.DAGNN_BN
    m = cls(nvt, kw["hs"], kw["hs"], nvt, nvt, 0, 1, hs=kw["hs"], nz=56, num_nodes=nvt, agg="attn_h", num_layers=kw["num_layers"],
            bidirectional=kw["bidirectional"], out_wx=False, out_pool_all=False, out_pool="max", dropout=0.0)
    mine = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert mine == e["state_dict"]
    sd = _fake_state_dict(e["state_dict"], 10 + idx)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    # cells_0 / cells_1 are aliases of grue_forward / grue_backward (dvae/dagnn.py:73-75): one storage, two names
    assert m.cells_0[0].weight_hh.data_ptr() == m.grue_forward[0].weight_hh.data_ptr()


def test_augment_edge2_batch_equals_the_reference_per_graph():
    z = np.load(os.path.join(GOLDEN, "augment_edge2.npz"))
    ei, ea = D.augment_edge2_batch(torch.from_numpy(z["in_edge_index_ast"]), torch.from_numpy(z["in_node_is_attributed"]).view(-1, 1),
                                   torch.from_numpy(z["in_batch"]))
    assert ei.dtype == torch.int64 and ea.dtype == torch.float32
    assert np.array_equal(ei.numpy(), z["edge_index"])
    assert np.array_equal(ea.numpy(), z["edge_attr"])
