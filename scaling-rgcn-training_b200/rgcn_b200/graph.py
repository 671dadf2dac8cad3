"""RGCNGraph — Python owner of the engine's opaque ``rgcn_graph`` (the on-device blocked
relational CSRs built once from the reference's ``edge_index`` / ``edge_type`` tensors,
/root/reference/graphs/graph.py:65-69), plus the identity-keyed cache RGCNConv uses so both
layers of a model (model/layers.py:21,23) share one build."""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np
import torch

from . import _lib

_ARRAY_DTYPES = {_lib.A_E_IDX: np.uint32, _lib.A_E_W: np.float32, _lib.A_RAW_W: np.float32}


def _env_int(name: str, default: int) -> int:
    v = os.environ.get(name)
    return int(v) if v else default


class RGCNGraph:
    def __init__(self, edge_index: torch.Tensor, edge_type: torch.Tensor, num_nodes: int, num_relations: int,
                 range_nodes: int = 0, split_threshold: int = 0, chunk_size: int = 0, own_range=None,
                 push: bool = False) -> None:
        lib = _lib.load()
        if edge_index.dtype != torch.int64 or edge_type.dtype != torch.int64:
            raise TypeError('edge_index and edge_type must be int64 (as graphs/graph.py:65 builds them)')
        if edge_index.dim() != 2 or edge_index.size(0) != 2 or edge_type.dim() != 1 \
                or edge_type.numel() != edge_index.size(1):
            raise ValueError('expected edge_index [2, E] and edge_type [E]')
        if not edge_index.is_cuda or not edge_type.is_cuda:
            raise _lib.EngineError('RGCNGraph needs CUDA tensors: the engine has no CPU path '
                                   '(move the Data object with .to("cuda") first)')
        self.device = edge_index.device
        self.num_nodes, self.num_relations, self.num_edges = int(num_nodes), int(num_relations), edge_type.numel()
        # own_range = (lo, hi): destination-partitioned graph of one rank; default = the whole graph
        self.own_lo, self.own_hi = (0, self.num_nodes) if own_range is None else (int(own_range[0]), int(own_range[1]))
        self.num_owned = self.own_hi - self.own_lo
        # push = source-partitioned forward structures (rgcn_graph_create_push): x holds the owned rows, the
        # forward output is a partial over ALL nodes that the caller reduce-scatters
        self.push = bool(push)
        range_nodes = range_nodes or _env_int('RGCN_B200_RANGE_NODES', 0)
        split_threshold = split_threshold or _env_int('RGCN_B200_SPLIT', 0)
        chunk_size = chunk_size or _env_int('RGCN_B200_CHUNK', 0)
        handle = C.c_void_p()
        src, dst = edge_index[0], edge_index[1]          # strided views are passed as they are
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            create = lib.rgcn_graph_create_push if self.push else lib.rgcn_graph_create_part
            rc = create(src.data_ptr(), src.stride(0) if src.numel() else 1,
                                            dst.data_ptr(), dst.stride(0) if dst.numel() else 1,
                                            edge_type.data_ptr(), edge_type.stride(0) if edge_type.numel() else 1,
                                            self.num_edges, self.num_nodes, self.num_relations, self.own_lo,
                                            self.own_hi, range_nodes, split_threshold, chunk_size, stream,
                                            C.byref(handle))
        _lib.check(rc, 'rgcn_graph_create')
        self._h = handle
        # source-partitioned: the rows of the full node set this rank's passes reach (destinations of the edges whose
        # source is owned + the owned rows) — what its partial output can be non-zero in and what it gathers of gout;
        # the sparse exchanges of partition.NvlComm are planned from it
        self.touched = None
        if self.push:
            t = torch.zeros(self.num_nodes, dtype=torch.bool, device=self.device)
            if self.num_edges:
                own = (src >= self.own_lo) & (src < self.own_hi)
                t[dst[own]] = True
            t[self.own_lo:self.own_hi] = True
            self.touched = t
        self._finalizer = weakref.finalize(self, lib.rgcn_graph_destroy, handle)

    @property
    def handle(self):
        return self._h

    def query(self, key: int, brc: int = _lib.BRC_FWD) -> int:
        out = C.c_int64()
        _lib.check(_lib.load().rgcn_graph_query(self._h, brc, key, C.byref(out)), 'rgcn_graph_query')
        return out.value

    def export(self, array: int, brc: int = _lib.BRC_FWD) -> np.ndarray:
        n = {
            _lib.A_PERM: self.query(_lib.Q_NUM_ENTRIES0, brc), _lib.A_RAW_IDX: self.query(_lib.Q_NUM_ENTRIES0, brc),
            _lib.A_RAW_W: self.query(_lib.Q_NUM_ENTRIES0, brc),
            _lib.A_SEG_PTR: self.query(_lib.Q_NUM_SEGMENTS, brc) + 1, _lib.A_SEG_PTR0: self.query(_lib.Q_NUM_SEGMENTS, brc) + 1,
            _lib.A_SEG_OWN: self.query(_lib.Q_NUM_SEGMENTS, brc), _lib.A_SEG_REL: self.query(_lib.Q_NUM_SEGMENTS, brc),
            _lib.A_E_IDX: self.query(_lib.Q_NUM_ENTRIES, brc), _lib.A_E_W: self.query(_lib.Q_NUM_ENTRIES, brc),
            _lib.A_CHUNK_BEG: self.query(_lib.Q_NUM_CHUNKS, brc), _lib.A_CHUNK_END: self.query(_lib.Q_NUM_CHUNKS, brc),
            _lib.A_BAT_SEG0: self.query(_lib.Q_NUM_BATCHES, brc), _lib.A_BAT_INFO: self.query(_lib.Q_NUM_BATCHES, brc),
            _lib.A_E_OWN: self.query(_lib.Q_NUM_ENTRIES, brc), _lib.A_TILE_E0: self.query(_lib.Q_NUM_TILES, brc),
            _lib.A_TILE_INFO: self.query(_lib.Q_NUM_TILES, brc),
            # only a FWD_REL structure of its own (several ranges) carries the map onto FWD's chunk numbering
            _lib.A_CHUNK_OUT: (self.query(_lib.Q_NUM_CHUNKS, brc)
                               if brc == _lib.BRC_FWD_REL and self.query(_lib.Q_RANGE_NODES) < self.num_owned else 0),
        }[array]
        out = np.empty(n, dtype=_ARRAY_DTYPES.get(array, np.int32))
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = _lib.load().rgcn_graph_export(self._h, brc, array, out.ctypes.data, out.nbytes, stream)
        _lib.check(rc, 'rgcn_graph_export')
        return out

    def workspace_bytes(self, fin: int, fout: int, backward: bool) -> int:
        return int(_lib.load().rgcn_layer_workspace_bytes(self._h, fin, fout, 1 if backward else 0))


# (edge_index storage ptr, offsets, strides, E, edge_type ptr, N, R, device) -> graph.
# Entries hold the tensors weakly: when the Data object drops them the graph goes too.
_CACHE: dict = {}


def cached_graph(edge_index: torch.Tensor, edge_type: torch.Tensor, num_nodes: int, num_relations: int) -> RGCNGraph:
    key = (edge_index.data_ptr(), tuple(edge_index.stride()), tuple(edge_index.shape), edge_index._version,
           edge_type.data_ptr(), tuple(edge_type.stride()), edge_type._version,
           int(num_nodes), int(num_relations), str(edge_index.device))
    hit = _CACHE.get(key)
    if hit is not None:
        ref_ei, ref_et, g = hit
        if ref_ei() is not None and ref_et() is not None:
            return g
        del _CACHE[key]
    g = RGCNGraph(edge_index, edge_type, num_nodes, num_relations)
    # drop dead entries so data_ptr reuse cannot alias a stale graph
    for k in [k for k, (a, b, _) in _CACHE.items() if a() is None or b() is None]:
        del _CACHE[k]
    _CACHE[key] = (weakref.ref(edge_index), weakref.ref(edge_type), g)
    return g


def clear_cache() -> None:
    _CACHE.clear()
