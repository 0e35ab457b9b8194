"""TEST-ONLY shim (see torch_geometric/__init__.py)."""
from .glob import global_add_pool, global_mean_pool, global_max_pool  # noqa: F401
from .conv import MessagePassing  # noqa: F401
from . import inits  # noqa: F401
