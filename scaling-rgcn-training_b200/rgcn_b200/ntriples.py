"""N-Triples -> node / relation ids -> edge tensors, vectorised.

The step in front of the hot path: the reference walks every triple twice in Python
(/root/reference/graphs/graph.py:24-69: sets of lower-cased tokens, ``sorted`` node enumeration, a dict
lookup per token, one Python list per edge) and once more per map file
(graphs/graphProcessing.py:41-52).  Here the file is tokenised once, tokens are factorised by hashing
(pandas), only the UNIQUE node strings are sorted, and the ids / inverse edges / map index are integer array
work.  The outputs equal the reference's exactly:

* ``nodes`` = ``sorted(subjects | objects)`` of the lower-cased tokens (str order = UTF-8 byte order);
* relation ids: the reference enumerates a ``set`` of strings (process-dependent order), so the
  caller either passes the reference's own ``relations`` dict (``relation_order=``) or gets the canonical
  sorted enumeration; ``rdf:type`` and ``<type>`` are dropped (graph.py:41-44);
* ``edge_index`` / ``edge_type``: for every triple in file order whose predicate is kept,
  ``[s, o, 2 r]`` then ``[o, s, 2 r + 1]``, as strided views of one ``[E, 3]`` int64 buffer (graph.py:55-69);
* the reference's quirks are kept: the last TWO characters of every line are cut before splitting
  (``triple[:-2]``), tokens split at the first two spaces only, empty lines skipped.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

TYPE_PREDICATES = ('<http://www.w3.org/1999/02/22-rdf-syntax-ns#type>', '<type>')


def read_lines(path: str) -> List[str]:
    """graphs/graphProcessing.py:7-10."""
    with open(path, 'r') as f:
        return f.read().splitlines()


def tokenize(lines: Sequence[str]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """-> object arrays (s, p, o) of the lower-cased tokens of every non-empty line, in file order."""
    import pandas as pd
    ser = pd.Series(list(lines), dtype=object)
    body = ser.str.slice(stop=-2)                       # triple[:-2]
    body = body[body != '']                             # triple_list != ['']
    parts = body.str.split(' ', n=2, expand=True)
    if parts.shape[1] < 3 or parts[2].isna().any():
        raise IndexError('N-Triples line with fewer than three space-separated fields (the reference raises here too)')
    low = [parts[i].str.lower().to_numpy(dtype=object) for i in range(3)]
    return low[0], low[1], low[2]


class ParsedGraph:
    """What Graph.init_graph leaves on the Graph object (graphs/graph.py:40-69)."""

    def __init__(self, nodes, node_to_enum, relations, edge_index, edge_type, num_edges):
        self.nodes: List[str] = nodes
        self.node_to_enum: Dict[str, int] = node_to_enum
        self.num_nodes = len(nodes)
        self.relations: Dict[str, int] = relations
        self.edge_index: torch.Tensor = edge_index
        self.edge_type: torch.Tensor = edge_type
        self.num_edges = num_edges           # distinct triple LINES (the reference's printed count)


def parse_graph(lines: Sequence[str], relation_order: Optional[Dict[str, int]] = None) -> ParsedGraph:
    """Vectorised Graph.init_graph.  relation_order: the reference's ``relations`` dict, to reproduce its
    process-dependent relation ids; default = predicates sorted."""
    import pandas as pd
    s, p, o = tokenize(lines)
    # nodes: factorise all subject / object tokens at once, sort only the uniques
    codes, uniques = pd.factorize(np.concatenate([s, o]), sort=False)
    order = np.argsort(np.asarray(uniques, dtype=object), kind='stable')        # Python str comparison = sorted()
    rank = np.empty(len(uniques), dtype=np.int64)
    rank[order] = np.arange(len(uniques), dtype=np.int64)
    nodes = [str(uniques[i]) for i in order]
    node_ids = rank[codes]
    src, dst = node_ids[:len(s)], node_ids[len(s):]
    # relations
    pcodes, puniq = pd.factorize(p, sort=False)
    kept = [str(u) for u in puniq if str(u) not in TYPE_PREDICATES]
    if relation_order is None:
        relations = {r: i for i, r in enumerate(sorted(kept))}
    else:
        if set(relation_order) != set(kept):
            raise ValueError('relation_order does not enumerate exactly the non-type predicates of this graph')
        relations = {str(k): int(v) for k, v in relation_order.items()}
    rel_of_code = np.array([relations.get(str(u), -1) for u in puniq], dtype=np.int64)
    rel = rel_of_code[pcodes]
    keep = rel >= 0
    src, dst, rel = src[keep], dst[keep], rel[keep]
    buf = np.empty((2 * src.size, 3), dtype=np.int64)
    buf[0::2, 0], buf[0::2, 1], buf[0::2, 2] = src, dst, 2 * rel
    buf[1::2, 0], buf[1::2, 1], buf[1::2, 2] = dst, src, 2 * rel + 1
    edge = torch.from_numpy(buf).t()
    return ParsedGraph(nodes, {n: i for i, n in enumerate(nodes)}, relations, edge[:2], edge[2], len(set(lines)))


def parse_graph_file(path: str, relation_order: Optional[Dict[str, int]] = None) -> ParsedGraph:
    return parse_graph(read_lines(path), relation_order)


def node_mappings(lines: Sequence[str]) -> Tuple[Dict[str, str], Dict[str, List[str]]]:
    """graphs/graphProcessing.py:41-52: (orgNode -> sumNode, sumNode -> [orgNodes]); a later map line overwrites an
    earlier one for the same original node; both dicts sorted by key."""
    s, _, o = tokenize(lines)
    org2sum = {}
    sum2org: Dict[str, List[str]] = {}
    for a, b in zip(s.tolist(), o.tolist()):
        sum2org.setdefault(a, []).append(b)
        org2sum[b] = a
    return dict(sorted(org2sum.items())), dict(sorted(sum2org.items()))


def map_index(map_lines: Sequence[str], org_nodes: Sequence[str], sum_node_to_enum: Dict[str, int]) -> torch.Tensor:
    """int32 [len(org_nodes)]: row of the summary embedding that feeds original node i, -1 where the reference keeps
    its torch.rand row (model/embeddingTricks.py:19-23) — the index rgcn_map_gather consumes, built without the
    O(N) Python dict walk: last occurrence per original node wins, exactly as the dict assignment does."""
    import pandas as pd
    s, _, o = tokenize(map_lines)
    df = pd.DataFrame({'org': o, 'sum': s}).drop_duplicates('org', keep='last')
    sum_row = df['sum'].map(sum_node_to_enum)                      # NaN: summary node absent from the summary graph
    org_pos = pd.Series(np.arange(len(org_nodes), dtype=np.int64), index=pd.Index(list(org_nodes), dtype=object))
    pos = df['org'].map(org_pos)
    ok = sum_row.notna() & pos.notna()
    idx = np.full(len(org_nodes), -1, dtype=np.int32)
    idx[pos[ok].to_numpy(dtype=np.int64)] = sum_row[ok].to_numpy(dtype=np.int64).astype(np.int32)
    return torch.from_numpy(idx)
