"""Attribute summariser (reference graphs/createAttributeSum.py:6-67), producer of
the on-disk summary / map files the transfer pipeline consumes — runnable where ``mmh3`` is absent.

A node's summary id is MurmurHash3-x64-128 (seed 0, unsigned, as ``mmh3.hash128`` returns it) of the
comma-joined SORTED set of its outgoing / incoming predicates (``rdf:type`` excluded, literals share one
incoming bucket), or the SUM of the two for the in+out summary.  Same function name, argument order and file
formats as the reference:

    <sub_hash> predicate <obj_hash> .                      (summary graph, one line per triple)
    <sum_hash> <isSummaryOf> original_node .               (map file, one line per original node, first-seen order)

The per-node predicate sets are built with one pandas group-by over the distinct (node, predicate) pairs instead
of a Python set update per triple; the hash runs once per DISTINCT predicate set.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

MASK64 = (1 << 64) - 1
RDF_TYPE = '<http://www.w3.org/1999/02/22-rdf-syntax-ns#type>'
LITERAL = 'http://example.org/literal'


def _rotl64(x: int, r: int) -> int:
    return ((x << r) | (x >> (64 - r))) & MASK64


def _fmix64(k: int) -> int:
    k ^= k >> 33
    k = (k * 0xff51afd7ed558ccd) & MASK64
    k ^= k >> 33
    k = (k * 0xc4ceb9fe1a85ec53) & MASK64
    k ^= k >> 33
    return k


def murmur3_x64_128(data: bytes, seed: int = 0) -> int:
    """MurmurHash3_x64_128 (Austin Appleby, public domain algorithm) as ``mmh3.hash128(data, seed)`` returns it:
    the 16 output bytes read as one little-endian unsigned integer (h2 << 64 | h1)."""
    c1, c2 = 0x87c37b91114253d5, 0x4cf5ad432745937f
    h1 = h2 = seed & MASK64
    n = len(data)
    nblocks = n // 16
    for i in range(nblocks):
        k1 = int.from_bytes(data[16 * i:16 * i + 8], 'little')
        k2 = int.from_bytes(data[16 * i + 8:16 * i + 16], 'little')
        k1 = (k1 * c1) & MASK64
        k1 = _rotl64(k1, 31)
        k1 = (k1 * c2) & MASK64
        h1 ^= k1
        h1 = _rotl64(h1, 27)
        h1 = (h1 + h2) & MASK64
        h1 = (h1 * 5 + 0x52dce729) & MASK64
        k2 = (k2 * c2) & MASK64
        k2 = _rotl64(k2, 33)
        k2 = (k2 * c1) & MASK64
        h2 ^= k2
        h2 = _rotl64(h2, 31)
        h2 = (h2 + h1) & MASK64
        h2 = (h2 * 5 + 0x38495ab5) & MASK64
    tail = data[16 * nblocks:]
    k1 = k2 = 0
    if len(tail) > 8:
        k2 = int.from_bytes(tail[8:], 'little')
        k2 = (k2 * c2) & MASK64
        k2 = _rotl64(k2, 33)
        k2 = (k2 * c1) & MASK64
        h2 ^= k2
    if len(tail) > 0:
        k1 = int.from_bytes(tail[:8], 'little')
        k1 = (k1 * c1) & MASK64
        k1 = _rotl64(k1, 31)
        k1 = (k1 * c2) & MASK64
        h1 ^= k1
    h1 ^= n
    h2 ^= n
    h1 = (h1 + h2) & MASK64
    h2 = (h2 + h1) & MASK64
    h1 = _fmix64(h1)
    h2 = _fmix64(h2)
    h1 = (h1 + h2) & MASK64
    h2 = (h2 + h1) & MASK64
    return (h2 << 64) | h1


def hash128(key, seed: int = 0) -> int:
    """Drop-in for ``mmh3.hash128`` (x64 variant, unsigned): str keys are UTF-8 encoded like mmh3 does."""
    return murmur3_x64_128(key.encode('utf8') if isinstance(key, str) else bytes(key), seed)


def property_hashes(lines: Sequence[str]):
    """-> (outgoing, incoming, combined) dicts node -> hash, in the reference's insertion orders
    (createAttributeSum.py:10-39)."""
    import pandas as pd
    from .ntriples import tokenize
    s, p, o = tokenize(lines)
    df = pd.DataFrame({'s': s, 'p': p, 'o': o})
    df = df[df['p'] != RDF_TYPE]
    lit = df['o'].str.startswith('"')
    out_pairs = df[['s', 'p']].drop_duplicates()
    in_pairs = pd.DataFrame({'n': df['o'].where(~lit, LITERAL), 'p': df['p']}).drop_duplicates()

    def hashed(pairs, key):
        # first-seen order of the nodes = insertion order of the reference's defaultdict
        keys = pairs[key].drop_duplicates().tolist()
        joined = pairs.sort_values('p', kind='stable').groupby(key, sort=False)['p'].agg(','.join)
        cache: Dict[str, int] = {}
        res = {}
        for k in keys:
            txt = joined[k]
            h = cache.get(txt)
            if h is None:
                h = cache[txt] = hash128(txt)
            res[k] = h
        return res

    outgoing = hashed(out_pairs, 's')
    incoming = hashed(in_pairs.rename(columns={'n': 's'}), 's')
    combined = {}
    for entity in set(incoming) | set(outgoing):         # (a set union in the reference too: only the VALUES matter,
        combined[entity] = incoming.get(entity, 0) + outgoing.get(entity, 0)   # the files follow the triple order)
    return outgoing, incoming, combined


def write_sum_map_files(prop: Dict[str, int], lines: Sequence[str], sum_path: str, map_path: str) -> None:
    """createAttributeSum.py:44-67: one summary line per triple (type triples included), one map line per original
    node in first-seen order, later triples overwriting the node's summary id."""
    from .ntriples import tokenize
    s, p, o = tokenize(lines)
    has_lit = LITERAL in prop
    mapping: Dict[str, object] = {}
    out: List[str] = []
    for a, b, c in zip(s.tolist(), p.tolist(), o.tolist()):
        if c.startswith('"') and has_lit:
            obj = prop[LITERAL]
        else:
            obj = prop[c] if c in prop else '0'
        sub = prop[a] if a in prop else '0'
        mapping[a] = sub
        mapping[c] = obj
        out.append(f'<{sub}> {b} <{obj}> .\n')
    with open(sum_path, 'w') as f:
        f.writelines(out)
    with open(map_path, 'w') as m:
        m.writelines(f'<{sn}> <isSummaryOf> {on} .\n' for on, sn in mapping.items())


def create_sum_map(path: str, sum_path: str, map_path: str, dataset: str) -> None:
    """Same signature and outputs as the reference's create_sum_map (main.py:39-42 calls it with -create_attr_sum)."""
    from .ntriples import read_lines
    lines = read_lines(path)
    outgoing, incoming, combined = property_hashes(lines)
    write_sum_map_files(outgoing, lines, f'{sum_path}{dataset}_sum_out.nt', f'{map_path}{dataset}_map_out.nt')
    write_sum_map_files(incoming, lines, f'{sum_path}{dataset}_sum_in.nt', f'{map_path}{dataset}_map_in.nt')
    write_sum_map_files(combined, lines, f'{sum_path}{dataset}_sum_in_out.nt', f'{map_path}{dataset}_map_in_out.nt')
