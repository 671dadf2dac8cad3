"""Summary -> original embedding transfer (reference model/embeddingTricks.py:8-49) on the
engine: the string/dict resolution (graphs/graphProcessing.py:41-52 map dicts, graph.py:53
node enumeration) is done once on the host into int32 index vectors — exact integer work —
and the row traffic runs in the K5 kernel (rgcn_map_gather).

``sum_embeddings`` / ``concat_embeddings`` / ``stack_embeddings`` keep the reference's
signatures ``(graph, sum_graphs, emb_dim)`` and its ``torch.rand`` fallback for unmapped rows;
``fallbacks=`` lets a caller inject those rows (parity tests).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch
from torch import Tensor

from . import _lib

MODE_SUM, MODE_CONCAT, MODE_STACK = 0, 1, 2


def build_map_index(org_node_to_enum: dict, sum_node_to_enum: dict, org2sum: dict) -> Tensor:
    """idx[i] = summary row feeding original node i, -1 where the reference keeps the random row
    (node absent from the map, or mapped to a summary node the summary graph does not contain)."""
    idx = torch.full((len(org_node_to_enum),), -1, dtype=torch.int32)
    get_sum = org2sum.get
    get_row = sum_node_to_enum.get
    for node, i in org_node_to_enum.items():
        s = get_sum(node)
        if s is not None:
            row = get_row(s)
            if row is not None:
                idx[i] = row
    return idx


def map_gather(embs: Sequence[Tensor], idxs: Sequence[Tensor], fallbacks: Sequence[Tensor], mode: int) -> Tensor:
    """Device map-gather: all tensors CUDA; embs[s] [N_s,F] fp32, idxs[s] [N] int32, fallbacks[s] [N,F]."""
    lib = _lib.load()
    s_count = len(embs)
    if not (s_count == len(idxs) == len(fallbacks)) or s_count == 0:
        raise ValueError('map_gather: need one index and one fallback per summary embedding')
    dev = embs[0].device
    if dev.type != 'cuda':
        raise _lib.EngineError('map_gather: CUDA tensors required (no CPU path)')
    n, f = fallbacks[0].shape
    embs = [e.detach().to(torch.float32).contiguous() for e in embs]
    idxs = [i.to(device=dev, dtype=torch.int32).contiguous() for i in idxs]
    fallbacks = [t.to(device=dev, dtype=torch.float32).contiguous() for t in fallbacks]
    for e, i, t in zip(embs, idxs, fallbacks):
        if e.size(1) != f or i.numel() != n or tuple(t.shape) != (n, f):
            raise ValueError('map_gather: inconsistent shapes')
    shape = {MODE_SUM: (n, f), MODE_CONCAT: (n, s_count * f), MODE_STACK: (s_count, n, f)}[mode]
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    arr = C.c_void_p * s_count
    with torch.cuda.device(dev):
        rc = lib.rgcn_map_gather(arr(*[e.data_ptr() for e in embs]), arr(*[i.data_ptr() for i in idxs]),
                                 arr(*[t.data_ptr() for t in fallbacks]), s_count, n, f, mode, out.data_ptr(),
                                 torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, 'rgcn_map_gather')
    return out


def _transfer(graph, sum_graphs: list, emb_dim: int, mode: int, fallbacks: Optional[List[Tensor]], device) -> Tensor:
    device = torch.device(device if device is not None else 'cuda')
    embs, idxs, fbs = [], [], []
    for k, sg in enumerate(sum_graphs):
        idx = getattr(sg, '_rgcn_b200_map_index', None)
        if idx is None or idx.numel() != graph.num_nodes:
            idx = build_map_index(graph.node_to_enum, sg.node_to_enum, sg.orgNode2sumNode_dict)
            sg._rgcn_b200_map_index = idx
        fb = fallbacks[k] if fallbacks is not None else torch.rand(graph.num_nodes, emb_dim)
        embs.append(sg.embedding.detach().to(device))
        idxs.append(idx.to(device))
        fbs.append(fb.to(device))
    return map_gather(embs, idxs, fbs, mode).detach()


def sum_embeddings(graph, sum_graphs: list, emb_dim: int, fallbacks=None, device=None) -> Tensor:
    """[num_graph_nodes, emb_dim] (embeddingTricks.py:43-49)."""
    return _transfer(graph, sum_graphs, emb_dim, MODE_SUM, fallbacks, device)


def concat_embeddings(graph, sum_graphs: list, emb_dim: int, fallbacks=None, device=None) -> Tensor:
    """[num_graph_nodes, num_summaries * emb_dim] (embeddingTricks.py:35-41)."""
    return _transfer(graph, sum_graphs, emb_dim, MODE_CONCAT, fallbacks, device)


def stack_embeddings(graph, sum_graphs: list, emb_dim: int, fallbacks=None, device=None) -> Tensor:
    """[num_summaries, num_graph_nodes, emb_dim] (embeddingTricks.py:27-33)."""
    return _transfer(graph, sum_graphs, emb_dim, MODE_STACK, fallbacks, device)
