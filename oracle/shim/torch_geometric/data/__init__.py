"""TEST-ONLY shim: torch_geometric.data.{Data, Batch} — just enough of PyG 1.6 for dvae/batch.py and
a collate that mirrors `Batch.from_data_list` ([PyG-upstream data/data.py, data/batch.py]):
`__inc__`: keys matching `index|face` are offset by num_nodes; `__cat_dim__`: those keys cat on -1, else 0.
"""
import re
import torch


class Data(object):
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.x, self.edge_index, self.edge_attr, self.y, self.pos = x, edge_index, edge_attr, y, pos
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __getitem__(self, key):
        return getattr(self, key, None)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    @property
    def keys(self):
        ks = [k for k in self.__dict__.keys() if self[k] is not None]
        return [k for k in ks if k[:2] != "__" and k[-2:] != "__"]

    def __iter__(self):
        for k in sorted(self.keys):
            yield k, self[k]

    def __cat_dim__(self, key, value):
        return -1 if bool(re.search("(index|face)", key)) else 0

    def __inc__(self, key, value):
        return self.num_nodes if bool(re.search("(index|face)", key)) else 0

    @property
    def num_nodes(self):
        if hasattr(self, "__num_nodes__"):
            return self.__num_nodes__
        if torch.is_tensor(self.x):
            return self.x.size(0)
        if torch.is_tensor(self.edge_index) and self.edge_index.numel() > 0:
            return int(self.edge_index.max()) + 1
        return None

    @num_nodes.setter
    def num_nodes(self, n):
        self.__num_nodes__ = n

    def apply(self, func, *keys):
        for k in (keys if keys else self.keys):
            v = self[k]
            if torch.is_tensor(v):
                self[k] = func(v)
        return self

    def contiguous(self, *keys):
        return self.apply(lambda t: t.contiguous(), *keys)

    def to(self, device, *keys, **kwargs):
        return self.apply(lambda t: t.to(device, **kwargs), *keys)


class Batch(Data):
    def __init__(self, batch=None, **kwargs):
        super().__init__(**kwargs)
        self.batch = batch

    @staticmethod
    def from_data_list(data_list, follow_batch=()):
        keys = sorted(set().union(*[set(d.keys) for d in data_list]))
        out = Batch()
        cols = {k: [] for k in keys}
        bvec, cum = [], 0
        for i, d in enumerate(data_list):
            n = d.num_nodes
            for k in keys:
                v = d[k]
                if torch.is_tensor(v) and v.dtype != torch.bool and d.__inc__(k, v) != 0:
                    v = v + cum
                cols[k].append(v)
            bvec.append(torch.full((n,), i, dtype=torch.long))
            cum += n
        for k in keys:
            v0 = cols[k][0]
            if torch.is_tensor(v0):
                out[k] = torch.cat([v.unsqueeze(0) if v.dim() == 0 else v for v in cols[k]],
                                   data_list[0].__cat_dim__(k, v0))
            elif isinstance(v0, (int, float)):
                out[k] = torch.tensor(cols[k])
            else:
                out[k] = cols[k]
        out.batch = torch.cat(bvec)
        return out.contiguous()

    @property
    def num_graphs(self):
        return int(self.batch[-1]) + 1
