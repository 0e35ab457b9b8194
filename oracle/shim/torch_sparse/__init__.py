"""TEST-ONLY shim: names dvae/batch.py imports from torch_sparse (never instantiated on our path)."""


class SparseTensor(object):
    pass


def cat(tensors, dim):
    raise NotImplementedError("torch_sparse.cat is not part of the oracle shim")
