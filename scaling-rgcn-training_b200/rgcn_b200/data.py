"""Data — the attribute bag the reference builds at /root/reference/graphs/graph.py:68-69 and
fills at graphs/dataset.py:30-35,53-54; ``.to(device)`` moves every tensor attribute in place
and returns self (PyG semantics relied on at model/modelTrainer.py:43); deep-copyable
(main.py:52)."""
from __future__ import annotations

import copy

from torch import Tensor


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs) -> None:
        for k, v in (('x', x), ('edge_index', edge_index), ('edge_attr', edge_attr), ('y', y), ('pos', pos)):
            if v is not None:
                setattr(self, k, v)
        for k, v in kwargs.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith('_')]

    def to(self, device, *args, **kwargs) -> 'Data':
        for k in self.keys():
            v = self.__dict__[k]
            if isinstance(v, Tensor):
                self.__dict__[k] = v.to(device, *args, **kwargs)
        return self

    def cpu(self) -> 'Data':
        return self.to('cpu')

    def cuda(self, device=None) -> 'Data':
        return self.to('cuda' if device is None else device)

    def __contains__(self, key: str) -> bool:
        return key in self.__dict__

    def __deepcopy__(self, memo):
        new = Data()
        for k, v in self.__dict__.items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def __repr__(self) -> str:
        parts = [f'{k}={list(v.shape) if isinstance(v, Tensor) else v!r}' for k, v in self.__dict__.items()]
        return f'Data({", ".join(parts)})'
