"""Synthetic graph generators for the benchmark configs.

``am_shape`` follows the stated recipe: seed 0; N nodes; T triples; P predicates with Zipf(1.0)
frequencies; subject uniform; object Zipf(0.9) over a random node permutation; duplicates
kept; emitted exactly like /root/reference/graphs/graph.py:62-63 — columns interleaved
(s, o, 2p), (o, s, 2p+1) — as strided views of one [E, 3] int64 buffer, R = 2P + 1.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

AM_NODES, AM_TRIPLES, AM_PREDICATES = 1_666_764, 5_988_321, 133


def _zipf_sampler(rng: np.random.Generator, n_items: int, alpha: float, size: int) -> np.ndarray:
    p = 1.0 / np.power(np.arange(1, n_items + 1, dtype=np.float64), alpha)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return np.searchsorted(cdf, rng.random(size), side='right').astype(np.int64)


def am_shape(scale: float = 1.0, seed: int = 0, num_nodes: int = None, num_triples: int = None,
             num_predicates: int = AM_PREDICATES) -> Tuple[torch.Tensor, torch.Tensor, int, int]:
    """-> (edge_index [2,E] int64 view, edge_type [E] int64 view, num_nodes, num_relations)."""
    n = int(num_nodes if num_nodes is not None else round(AM_NODES * scale))
    t = int(num_triples if num_triples is not None else round(AM_TRIPLES * scale))
    rng = np.random.default_rng(seed)
    subj = rng.integers(0, n, size=t, dtype=np.int64)
    perm = rng.permutation(n).astype(np.int64)
    obj = perm[np.minimum(_zipf_sampler(rng, n, 0.9, t), n - 1)]
    pred = np.minimum(_zipf_sampler(rng, num_predicates, 1.0, t), num_predicates - 1)
    buf = np.empty((2 * t, 3), dtype=np.int64)
    buf[0::2, 0], buf[0::2, 1], buf[0::2, 2] = subj, obj, 2 * pred
    buf[1::2, 0], buf[1::2, 1], buf[1::2, 2] = obj, subj, 2 * pred + 1
    edge = torch.from_numpy(buf).t()           # [3, E] view of the [E, 3] buffer, like graph.py:65
    return edge[:2], edge[2], n, 2 * num_predicates + 1


def labelled_split(num_nodes: int, num_classes: int, frac: float = 0.10, seed: int = 0):
    """10 % labelled nodes, one class each (CE path), 60/20/20 split -> (x_train, y_train one-hot)."""
    rng = np.random.default_rng(seed + 1)
    lab = rng.choice(num_nodes, size=max(1, int(num_nodes * frac)), replace=False)
    n_train = max(1, int(len(lab) * 0.6))
    x_train = torch.from_numpy(np.sort(lab[:n_train]).astype(np.int64))
    cls = torch.from_numpy(rng.integers(0, num_classes, size=n_train).astype(np.int64))
    y_train = torch.nn.functional.one_hot(cls, num_classes).to(torch.int64)
    return x_train, y_train


def random_multigraph(num_nodes: int, num_edges: int, num_relations: int, seed: int = 0, hub_frac: float = 0.0,
                      dup_frac: float = 0.0):
    """Small random multigraph for parity tests: optional hub destinations and duplicate edges."""
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, num_nodes, (num_edges,), generator=g)
    dst = torch.randint(0, num_nodes, (num_edges,), generator=g)
    rel = torch.randint(0, max(1, num_relations - 1), (num_edges,), generator=g)   # last slot stays empty
    if hub_frac > 0:
        k = int(num_edges * hub_frac)
        dst[:k] = 0
        rel[:k] = 0
    if dup_frac > 0 and num_edges > 1:
        k = int(num_edges * dup_frac)
        pick = torch.randint(0, num_edges, (k,), generator=g)
        src[:k], dst[:k], rel[:k] = src[pick].clone(), dst[pick].clone(), rel[pick].clone()
    buf = torch.stack([src, dst, rel], dim=1).contiguous()
    edge = buf.t()
    return edge[:2], edge[2]
