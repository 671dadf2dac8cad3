"""CPU ORACLE (numpy) for the engine's on-device graph build (kernel group K0) — test
infrastructure only.

The reference has no CSR: PyG's loop path masks ``edge_type == r`` per relation
 over the tensors ``Graph.init_graph`` emits
(/root/reference/graphs/graph.py:55-69).  The engine re-sorts those edges once; this file
restates that integer mapping with numpy (stable lexsort) so the device build can be
checked bit-exactly, and so that "the CSR is a pure permutation of the reference's
edges" is a testable property (``edges_from_brc``).

Blocked relational CSR (BRC), spec shared with scaling-rgcn-training_b200/csrc/graph_build.cu:
  entries   the E edges plus one self-loop per node (relation id R, weight 1)  -> root term
  key       (owner // NR) * (R+1) * NR + rel * NR + owner % NR ; self loops after all ranges, by owner;
            stable sort
  segment   maximal run of equal key = one (relation, owner) pair; cnt = multiplicity
  weight    1/cnt of the FORWARD segment of the entry (per-(relation, dst) mean normaliser)
  chunking  segments with cnt > T are replaced by ceil(cnt/CH) virtual entries N+chunk_id
            (weight 1) whose rows a pre-pass fills with the weighted partial sums
  group     maximal run of segments with equal (range, rel)
  batch     <= BS consecutive segments of one group
"""
from __future__ import annotations

import numpy as np

BS = 16
LAST_FLAG = np.uint32(0x80000000)


def edge_weights(dst, rel, n):
    """w[e] = 1 / multiplicity of e's (relation, dst) pair over the WHOLE graph (float32 division)."""
    key = np.asarray(rel, dtype=np.int64) * n + np.asarray(dst, dtype=np.int64)
    _, inv, counts = np.unique(key, return_inverse=True, return_counts=True)
    return (np.float32(1.0) / counts.astype(np.float32))[inv].astype(np.float32)


def build_brc(own, gat, rel, n, r, nr, t, ch, w_edge=None, lo=0, hi=None, push=False):
    """own/gat/rel: the graph's edges (global ids).  [lo,hi): owned node range (whole graph by
    default).  Entries = edges whose owner end is owned (input order), then the owned self loops;
    owner ids local, gather ids global.  push=True (rgcn_graph_create_push): the edges whose GATHER
    end is owned, gather ids local, owner ids global (owner space = the whole graph)."""
    own = np.asarray(own, dtype=np.int64)
    gat = np.asarray(gat, dtype=np.int64)
    rel = np.asarray(rel, dtype=np.int64)
    hi = n if hi is None else hi
    if w_edge is None:
        raise ValueError('w_edge required (edge_weights(dst, rel, n))')
    n_sel = hi - lo
    loops = np.arange(n_sel, dtype=np.int64)
    if push:
        sel = (gat >= lo) & (gat < hi)
        n_gat, n_own = n_sel, n
        own2 = np.concatenate([own[sel], loops + lo])
        gat2 = np.concatenate([gat[sel] - lo, loops])
    else:
        sel = (own >= lo) & (own < hi)
        n_gat, n_own = n, n_sel
        own2 = np.concatenate([own[sel] - lo, loops])
        gat2 = np.concatenate([gat[sel], loops + lo])
    nr = int(min(max(nr, 1), max(n_own, 1)))
    rel2 = np.concatenate([rel[sel], np.full(n_sel, r, dtype=np.int64)])
    w_entry = np.concatenate([np.asarray(w_edge, dtype=np.float32)[sel], np.ones(n_sel, dtype=np.float32)])
    e = int(sel.sum())
    n_self = n_sel
    n = n_own
    nranges = max((n + nr - 1) // nr, 1)
    self_base = nranges * (r + 1) * nr            # self loops (relation r) sort after every range, by owner
    key = np.where(rel2 == r, self_base + own2, (own2 // nr) * ((r + 1) * nr) + rel2 * nr + own2 % nr)
    perm = np.argsort(key, kind='stable')
    skey = key[perm]
    e2 = e + n_self
    head = np.ones(e2, dtype=bool)
    head[1:] = skey[1:] != skey[:-1]
    seg_of = np.cumsum(head) - 1
    seg_ptr0 = np.flatnonzero(head).astype(np.int64)
    s = seg_ptr0.size
    seg_ptr0 = np.concatenate([seg_ptr0, [e2]])
    cnt = np.diff(seg_ptr0)
    seg_key = skey[seg_ptr0[:-1]]
    is_self = seg_key >= self_base
    seg_rel = np.where(is_self, r, (seg_key // nr) % (r + 1))
    seg_own = np.where(is_self, seg_key - self_base, (seg_key // ((r + 1) * nr)) * nr + seg_key % nr)
    seg_range = seg_own // nr
    raw_idx = gat2[perm].astype(np.int32)
    raw_w = w_entry[perm].astype(np.float32)
    # chunking
    split = cnt > t
    nchunk = np.where(split, (cnt + ch - 1) // ch, 0)
    chunk_base = np.concatenate([[0], np.cumsum(nchunk)])
    nc = int(chunk_base[-1])
    out_cnt = np.where(split, nchunk, cnt)
    seg_ptr = np.concatenate([[0], np.cumsum(out_cnt)]).astype(np.int64)
    e3 = int(seg_ptr[-1])
    e_idx = np.zeros(e3, dtype=np.uint32)
    e_w = np.zeros(e3, dtype=np.float32)
    chunk_beg = np.zeros(nc, dtype=np.int32)
    chunk_end = np.zeros(nc, dtype=np.int32)
    for si in range(s):
        a, b = seg_ptr0[si], seg_ptr0[si + 1]
        o = seg_ptr[si]
        if split[si]:
            for j in range(int(nchunk[si])):
                c = chunk_base[si] + j
                chunk_beg[c] = a + j * ch
                chunk_end[c] = min(a + (j + 1) * ch, b)
                e_idx[o + j] = n_gat + c
                e_w[o + j] = 1.0
        else:
            e_idx[o:o + (b - a)] = raw_idx[a:b].astype(np.uint32)
            e_w[o:o + (b - a)] = raw_w[a:b]
    if s:
        e_idx[seg_ptr[1:] - 1] |= LAST_FLAG
    # groups and batches
    gkey = seg_range * (r + 1) + seg_rel
    ghead = np.ones(s, dtype=bool)
    ghead[1:] = gkey[1:] != gkey[:-1]
    grp_seg = np.concatenate([np.flatnonzero(ghead), [s]]).astype(np.int64)
    g = grp_seg.size - 1
    grp_rel = seg_rel[grp_seg[:-1]] if g else np.zeros(0, dtype=np.int64)
    grp_n = np.diff(grp_seg)
    grp_nb = (grp_n + BS - 1) // BS
    nb = int(grp_nb.sum())
    bat_seg0 = np.zeros(nb, dtype=np.int32)
    bat_info = np.zeros(nb, dtype=np.int32)
    b = 0
    for gi in range(g):
        for j in range(int(grp_nb[gi])):
            s0 = grp_seg[gi] + j * BS
            ns = min(BS, grp_seg[gi + 1] - s0)
            bat_seg0[b] = s0
            bat_info[b] = (int(grp_rel[gi]) << 8) | int(ns)
            b += 1
    # per-entry owner and entry tiles (<= 16 consecutive entries of one (range, relation) group)
    e_own = np.repeat(seg_own, out_cnt).astype(np.int32)
    tile_e0, tile_info = [], []
    for gi in range(g):
        a, bnd = int(seg_ptr[grp_seg[gi]]), int(seg_ptr[grp_seg[gi + 1]])
        for e0 in range(a, bnd, 16):
            tile_e0.append(e0)
            tile_info.append((int(grp_rel[gi]) << 8) | min(16, bnd - e0))
    num_tiles_noself = int(sum(1 for ti in tile_info if (ti >> 8) != r))
    return dict(num_tiles_noself=num_tiles_noself, e_own=e_own, tile_e0=np.asarray(tile_e0, dtype=np.int32), tile_info=np.asarray(tile_info, dtype=np.int32),
                num_tiles=len(tile_e0), perm=perm.astype(np.int32), num_seg=s, seg_ptr=seg_ptr.astype(np.int32),
                seg_own=seg_own.astype(np.int32), seg_rel=seg_rel.astype(np.int32), cnt=cnt.astype(np.int32),
                e_idx=e_idx, e_w=e_w, raw_idx=raw_idx, raw_w=raw_w, w_entry=w_entry,
                chunk_beg=chunk_beg, chunk_end=chunk_end, num_chunks=nc,
                num_groups=g, num_batches=nb, bat_seg0=bat_seg0, bat_info=bat_info)


def share_chunks(fwd, fwd_rel, n_gat):
    """FWD_REL shares FWD's chunk numbering (graph_build.cu share_chunks): both chunk the same
    (relation, dst) segments into the same pieces; FWD_REL's chunk c is the c-th chunk in
    (relation, dst, piece) order, and a stable sort of FWD's chunks by (relation, dst) gives FWD's id
    of it.  Returns a copy of fwd_rel with the chunk entries renumbered and ``chunk_out`` = that map."""
    nc = fwd['num_chunks']
    out = dict(fwd_rel)
    if nc == 0 or nc != fwd_rel['num_chunks']:
        out['chunk_out'] = np.zeros(0, dtype=np.int32)
        return out
    n_own = int(max(fwd['seg_own'].max(), fwd_rel['seg_own'].max())) + 1
    key = np.zeros(nc, dtype=np.int64)
    idx = fwd['e_idx'] & np.uint32(0x7fffffff)
    for s in range(fwd['num_seg']):
        a, b = int(fwd['seg_ptr'][s]), int(fwd['seg_ptr'][s + 1])
        if a < b and idx[a] >= n_gat:
            for o in range(a, b):
                key[int(idx[o]) - n_gat] = int(fwd['seg_rel'][s]) * n_own + int(fwd['seg_own'][s])
    order = np.argsort(key, kind='stable').astype(np.int32)
    e = fwd_rel['e_idx'].copy()
    ridx = e & np.uint32(0x7fffffff)
    is_chunk = ridx >= n_gat
    e[is_chunk] = (np.uint32(n_gat) + order[(ridx[is_chunk] - n_gat).astype(np.int64)].astype(np.uint32)) | (e[is_chunk] & LAST_FLAG)
    out['e_idx'] = e
    out['chunk_out'] = order
    return out


def build_graph(src, dst, rel, n, r, nr, t, ch, lo=0, hi=None, push=False):
    """Forward (owner = dst) and transposed (owner = src) BRCs as the engine builds them; push=True: the
    source-partitioned forward structure (the transposed one is the same in both modes)."""
    w = edge_weights(dst, rel, n)
    fwd = build_brc(dst, src, rel, n, r, nr, t, ch, w_edge=w, lo=lo, hi=hi, push=push)
    bwd = build_brc(src, dst, rel, n, r, nr, t, ch, w_edge=w, lo=lo, hi=hi)
    return fwd, bwd


def edges_from_brc(brc, n):
    """Recover the (owner, gather, rel) entry multiset the BRC encodes (self-loops carry
    rel == R; chunked segments are read from the raw sorted arrays).  Restricted to
    rel < R it must equal the reference's edge list as a multiset."""
    seg_ptr0 = np.concatenate([[0], np.cumsum(brc['cnt'])])
    seg_of = np.repeat(np.arange(brc['num_seg']), brc['cnt'])
    assert seg_of.size == seg_ptr0[-1]
    return np.stack([brc['seg_own'][seg_of], brc['raw_idx'], brc['seg_rel'][seg_of]], axis=1)
