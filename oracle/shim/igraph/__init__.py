"""TEST-ONLY stub: dvae/models_pyg.py does `import igraph` at module scope; only loss()/decode() use it."""


class Graph(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("igraph is not available; the oracle covers forward/encode only")
