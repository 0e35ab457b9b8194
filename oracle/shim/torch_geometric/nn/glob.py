"""TEST-ONLY shim: torch_geometric.nn.glob pools (scatter over the `batch` vector, PyG 1.6)."""
import torch

__all__ = ["global_add_pool", "global_mean_pool", "global_max_pool"]


def _size(batch, size):
    return int(batch.max().item() + 1) if size is None else size


def global_add_pool(x, batch, size=None):
    size = _size(batch, size)
    out = x.new_zeros((size,) + tuple(x.shape[1:]))
    return out.index_add(0, batch, x)


def global_mean_pool(x, batch, size=None):
    size = _size(batch, size)
    s = global_add_pool(x, batch, size)
    cnt = torch.zeros(size, dtype=x.dtype, device=x.device).index_add(
        0, batch, torch.ones_like(batch, dtype=x.dtype)).clamp(min=1)
    return s / cnt.view(-1, *([1] * (x.dim() - 1)))


def global_max_pool(x, batch, size=None):
    size = _size(batch, size)
    out = x.new_full((size,) + tuple(x.shape[1:]), float("-inf"))
    idx = batch.view(-1, *([1] * (x.dim() - 1))).expand_as(x)
    out = out.scatter_reduce(0, idx, x, reduce="amax", include_self=True)
    # torch_scatter fills segments that received nothing with 0
    return torch.where(torch.isinf(out) & (out < 0), torch.zeros_like(out), out)
