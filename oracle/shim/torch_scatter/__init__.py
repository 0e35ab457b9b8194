"""TEST-ONLY shim: torch_scatter.scatter_add (name imported by the reference; unused on the default path)."""
import torch


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    if dim_size is None:
        dim_size = int(index.max()) + 1
    shape = list(src.shape)
    shape[dim] = dim_size
    res = src.new_zeros(shape) if out is None else out
    idx = index
    if index.dim() != src.dim():
        view = [1] * src.dim()
        view[dim] = -1
        idx = index.view(view).expand_as(src)
    return res.scatter_add(dim, idx, src)
