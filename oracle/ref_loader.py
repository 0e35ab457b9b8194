"""ORACLE — TEST INFRASTRUCTURE ONLY.

Imports the reference's own model files UNCHANGED from a read-only checkout (default /root/reference, or
$DAGNN_REFERENCE) on top of the test-only PyG shim in oracle/shim. Only usable where the reference is
mounted (this build container) — never on the GPU box; `tests/golden/` carries its outputs there.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "shim")


def reference_root():
    for cand in (os.environ.get("DAGNN_REFERENCE"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "ogbg-code", "model", "dagnn.py")):
            return cand
    return None


def available() -> bool:
    return reference_root() is not None


def _with_paths(paths):
    for p in reversed(paths):
        if p not in sys.path:
            sys.path.insert(0, p)


def load_ogb():
    """-> (module ogbg-code/model/dagnn.py, module ogbg-code/utils.py)"""
    root = reference_root()
    _with_paths([SHIM, root, os.path.join(root, "ogbg-code")])
    dag = importlib.import_module("model.dagnn")
    utl = importlib.import_module("utils")
    return dag, utl


def load_dvae():
    """-> (module dvae/dagnn.py, module dvae/dagnn_bn.py, module dvae/batch.py)"""
    root = reference_root()
    _with_paths([SHIM, root, os.path.join(root, "dvae")])
    return (importlib.import_module("dagnn"), importlib.import_module("dagnn_bn"),
            importlib.import_module("batch"))


def load_utils_dag():
    root = reference_root()
    _with_paths([root])
    return importlib.import_module("src.utils_dag")
