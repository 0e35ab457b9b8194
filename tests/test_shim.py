"""CPU: hand-computed cases pinning the TEST-ONLY PyG shim (oracle/shim) the reference model files run on
(SURVEY.md §8c [PyG-upstream] semantics): propagate flow / _i / _j lifting, segment softmax (+1e-16), aggregation into
N rows with empty rows = 0, global pools, Batch increments of '*index*' keys."""
import os
import sys

import pytest
import torch

SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shim")


@pytest.fixture()
def shim():
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("torch_geometric", "torch_scatter", "torch_sparse", "igraph")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, SHIM)
    try:
        import torch_geometric  # noqa: F401
        yield sys.modules["torch_geometric"]
    finally:
        sys.path.remove(SHIM)
        for k in [k for k in sys.modules if k.split(".")[0] in ("torch_geometric", "torch_scatter", "torch_sparse", "igraph")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_propagate_flow_and_lifting(shim):
    from torch_geometric.nn.conv import MessagePassing

    class Probe(MessagePassing):
        def forward(self, x, q, edge_index):
            return self.propagate(edge_index, x=x, q=q)

        def message(self, x_j, q_i, index, size_i):
            self.seen = (x_j.clone(), q_i.clone(), index.clone(), size_i)
            return x_j * 10 + q_i

    x = torch.tensor([[1.], [2.], [3.], [4.]])
    q = torch.tensor([[.1], [.2], [.3], [.4]])
    ei = torch.tensor([[0, 1, 0], [2, 2, 3]])
    p = Probe(aggr="add", flow="source_to_target")
    out = p(x, q, ei)                                   # aggregate at edge_index[1] from edge_index[0]
    assert torch.equal(p.seen[0].view(-1), torch.tensor([1., 2., 1.])) and torch.equal(p.seen[2], ei[1]) and p.seen[3] == 4
    assert torch.allclose(out.view(-1), torch.tensor([0., 0., 10.3 + 20.3, 10.4]))
    p = Probe(aggr="add", flow="target_to_source")
    out = p(x, q, ei)                                   # reverse: aggregate at edge_index[0] from edge_index[1]
    assert torch.equal(p.seen[0].view(-1), torch.tensor([3., 3., 4.])) and torch.equal(p.seen[2], ei[0])
    assert torch.allclose(out.view(-1), torch.tensor([30.1 + 40.1, 30.2, 0., 0.]))
    out = Probe(aggr="max")(x, q, ei)
    assert torch.allclose(out.view(-1), torch.tensor([0., 0., 20.3, 10.4]))
    out = Probe(aggr="mean")(x, q, ei)
    assert torch.allclose(out.view(-1), torch.tensor([0., 0., (10.3 + 20.3) / 2, 10.4]))


def test_segment_softmax(shim):
    from torch_geometric.utils import softmax
    a = torch.tensor([[0.], [1.], [5.], [2.]])
    idx = torch.tensor([1, 1, 3, 1])
    s = softmax(a, idx, None, 5).view(-1)
    e = torch.exp(torch.tensor([0., 1., 2.]) - 2.)
    assert torch.allclose(s[[0, 1, 3]], e / (e.sum() + 1e-16)) and torch.allclose(s[2], torch.tensor(1.0))
    assert abs(float(s[[0, 1, 3]].sum()) - 1.0) < 1e-6


def test_global_pools_and_batch_increment(shim):
    from torch_geometric.nn import global_add_pool, global_max_pool, global_mean_pool
    from torch_geometric.data import Batch, Data
    h = torch.tensor([[1., -5.], [3., -1.], [2., 7.]])
    b = torch.tensor([0, 0, 2])
    assert torch.equal(global_max_pool(h, b, 3), torch.tensor([[3., -1.], [0., 0.], [2., 7.]]))
    assert torch.equal(global_add_pool(h, b, 3), torch.tensor([[4., -6.], [0., 0.], [2., 7.]]))
    assert torch.equal(global_mean_pool(h, b, 3), torch.tensor([[2., -3.], [0., 0.], [2., 7.]]))
    g1 = Data(x=torch.zeros(2, 1), edge_index=torch.tensor([[0], [1]]), _bi_layer_idx0=torch.tensor([0, 1]),
              _bi_layer_index0=torch.tensor([0, 1]))
    g2 = Data(x=torch.zeros(3, 1), edge_index=torch.tensor([[0, 1], [2, 2]]), _bi_layer_idx0=torch.tensor([0, 0, 1]),
              _bi_layer_index0=torch.tensor([0, 1, 2]))
    B = Batch.from_data_list([g1, g2])
    assert torch.equal(B.edge_index, torch.tensor([[0, 2, 3], [1, 4, 4]]))          # '*index*' keys are offset by num_nodes
    assert torch.equal(B._bi_layer_index0, torch.tensor([0, 1, 2, 3, 4]))
    assert torch.equal(B._bi_layer_idx0, torch.tensor([0, 1, 0, 0, 1]))             # level values are NOT offset
    assert torch.equal(B.batch, torch.tensor([0, 0, 1, 1, 1]))
