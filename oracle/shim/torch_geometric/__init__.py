"""TEST-ONLY shim of the handful of torch_geometric 1.6 symbols the DAGNN reference imports.

This is NOT product code. It exists so that `oracle/gen_golden.py` (and CPU-side tests,
when `/root/reference` is mounted) can import the reference's model files *unchanged*
(`ogbg-code/model/dagnn.py`, `dvae/dagnn.py`, `dvae/dagnn_bn.py`, `dvae/models_pyg.py`,
`dvae/batch.py`) in an image that has no torch_geometric / torch_scatter / torch_sparse.
Semantics follow PyG 1.6.x as documented upstream ([PyG-upstream] in SURVEY.md §8c); the
hand-computed cases in tests/test_shim.py pin them.
"""
from . import nn, data, utils, typing  # noqa: F401

__version__ = "1.6.0-shim"


def is_debug_enabled():
    return False
