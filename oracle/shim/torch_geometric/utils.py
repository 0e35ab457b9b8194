"""TEST-ONLY shim (see torch_geometric/__init__.py): torch_geometric.utils.softmax, PyG 1.6 semantics."""
import torch


def _seg_max(src, index, num_nodes):
    out = src.new_full((num_nodes,) + tuple(src.shape[1:]), float("-inf"))
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return out.scatter_reduce(0, idx, src, reduce="amax", include_self=True)


def _seg_sum(src, index, num_nodes):
    out = src.new_zeros((num_nodes,) + tuple(src.shape[1:]))
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return out.scatter_add(0, idx, src)


def softmax(src, index, ptr=None, num_nodes=None):
    """out = exp(src - segmax[index]) / (segsum(exp)[index] + 1e-16)   [PyG-upstream utils/softmax.py]"""
    if num_nodes is None:
        num_nodes = int(index.max()) + 1 if index.numel() > 0 else 0
    out = src - _seg_max(src, index, num_nodes)[index]
    out = out.exp()
    out_sum = _seg_sum(out, index, num_nodes)[index]
    return out / (out_sum + 1e-16)
