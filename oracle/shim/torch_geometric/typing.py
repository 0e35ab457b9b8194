"""TEST-ONLY shim (see torch_geometric/__init__.py)."""
from typing import Optional
from torch import Tensor

OptTensor = Optional[Tensor]
