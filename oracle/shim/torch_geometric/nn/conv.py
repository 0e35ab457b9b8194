"""TEST-ONLY shim: torch_geometric.nn.MessagePassing with PyG 1.6 propagate semantics.

[PyG-upstream nn/conv/message_passing.py]: `flow='source_to_target'` -> (i, j) = (1, 0),
`'target_to_source'` -> (i, j) = (0, 1). For every `message` parameter `<name>_i` / `<name>_j`
the kwarg `<name>` (None stays None) is `index_select`ed on dim 0 with `edge_index[i]` /
`edge_index[j]`. Special names: `index = edge_index[i]`, `ptr = None`, `size_i`/`size_j` = number
of nodes (inferred from the first lifted tensor), `edge_index`. Everything else is passed through.
`aggregate` scatters (`add`/`mean`/`max`) over `index` into `dim_size = size_i` rows; rows that
receive nothing are 0. `update(aggr_out)` gets the aggregate.
"""
import inspect
import torch


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=0):
        super().__init__()
        assert aggr in ("add", "mean", "max", None)
        assert flow in ("source_to_target", "target_to_source")
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim
        self._msg_params = [p for p in inspect.signature(self.message).parameters]

    def propagate(self, edge_index, size=None, **kwargs):
        i, j = (1, 0) if self.flow == "source_to_target" else (0, 1)
        n_nodes = None
        for name in self._msg_params:
            if name[-2:] in ("_i", "_j") and name not in ("size_i", "size_j"):
                d = kwargs.get(name[:-2])
                if torch.is_tensor(d):
                    n_nodes = d.size(0)
                    break
        if size is not None:
            n_nodes = size[i] if isinstance(size, (tuple, list)) else size
        args = {}
        for name in self._msg_params:
            if name in ("size_i", "size_j"):
                args[name] = n_nodes
            elif name[-2:] in ("_i", "_j"):
                d = kwargs.get(name[:-2])
                if torch.is_tensor(d):
                    d = d.index_select(0, edge_index[i if name[-2:] == "_i" else j])
                args[name] = d
            elif name == "index":
                args[name] = edge_index[i]
            elif name == "ptr":
                args[name] = None
            elif name == "edge_index":
                args[name] = edge_index
            else:
                args[name] = kwargs.get(name)
        out = self.message(**args)
        out = self.aggregate(out, edge_index[i], None, n_nodes)
        return self.update(out)

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        idx = index.view(-1, *([1] * (inputs.dim() - 1))).expand_as(inputs)
        shape = (dim_size,) + tuple(inputs.shape[1:])
        if self.aggr == "add":
            return inputs.new_zeros(shape).scatter_add(0, idx, inputs)
        if self.aggr == "mean":
            s = inputs.new_zeros(shape).scatter_add(0, idx, inputs)
            c = inputs.new_zeros(dim_size).index_add(0, index, torch.ones_like(index, dtype=inputs.dtype))
            return s / c.clamp(min=1).view(-1, *([1] * (inputs.dim() - 1)))
        out = inputs.new_full(shape, float("-inf")).scatter_reduce(0, idx, inputs, reduce="amax", include_self=True)
        return torch.where(torch.isinf(out) & (out < 0), torch.zeros_like(out), out)

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out
