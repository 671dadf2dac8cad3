"""Property tests (hypothesis, CPU) of the blocked relational CSR spec that the device build is
checked against bit for bit: for arbitrary small multigraphs and blocking parameters the structure
is a pure permutation of the reference's edges (graphs/graph.py:62-63 semantics: multi-edges kept),
weights are the per-(relation, dst) mean normalisers, and every derived table tiles its parent."""
import numpy as np
from hypothesis import given, settings, strategies as st

import csr_oracle


@st.composite
def graphs(draw):
    n = draw(st.integers(1, 40))
    r = draw(st.integers(1, 6))
    e = draw(st.integers(0, 120))
    src = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
    dst = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
    rel = draw(st.lists(st.integers(0, r - 1), min_size=e, max_size=e))
    nr = draw(st.sampled_from([1, 3, 8, 64]))
    t = draw(st.sampled_from([1, 2, 5, 1000]))
    ch = draw(st.sampled_from([1, 2, 7]))
    lo = draw(st.integers(0, n - 1))
    hi = draw(st.integers(lo + 1, n))
    return np.array(src), np.array(dst), np.array(rel), n, r, nr, t, ch, lo, hi


@settings(max_examples=120, deadline=None)
@given(graphs())
def test_brc_invariants(gr):
    src, dst, rel, n, r, nr, t, ch, lo, hi = gr
    w = csr_oracle.edge_weights(dst, rel, n)
    for own, gat in ((dst, src), (src, dst)):
        b = csr_oracle.build_brc(own, gat, rel, n, r, nr, t, ch, w_edge=w, lo=lo, hi=hi)
        sel = (own >= lo) & (own < hi)
        # permutation of the owned edges (+ one self loop per owned node)
        got = csr_oracle.edges_from_brc(b, n)
        edges = got[got[:, 2] < r]
        want = np.stack([own[sel] - lo, gat[sel], rel[sel]], 1) if sel.any() else np.zeros((0, 3), dtype=np.int64)
        assert sorted(map(tuple, edges.tolist())) == sorted(map(tuple, want.tolist()))
        loops = got[got[:, 2] == r]
        assert sorted(loops[:, 0].tolist()) == list(range(hi - lo)) and np.all(loops[:, 1] == loops[:, 0] + lo)
        # weights: 1 / multiplicity of the edge's (relation, dst) pair in the WHOLE graph; self loops 1
        mult = {}
        for d_, r_ in zip(dst.tolist(), rel.tolist()):
            mult[(r_, d_)] = mult.get((r_, d_), 0) + 1
        ent_w = b['w_entry']
        k = 0
        for i in np.flatnonzero(sel):
            assert ent_w[k] == np.float32(1.0) / np.float32(mult[(int(rel[i]), int(dst[i]))])
            k += 1
        assert np.all(ent_w[k:] == 1.0)
        # chunks tile exactly the segments longer than t
        cnt = b['cnt']
        assert int((b['chunk_end'] - b['chunk_beg']).sum()) == int(cnt[cnt > t].sum())
        # compacted entries: one per short-segment entry or per chunk; last-flag once per segment
        assert int((b['e_idx'] >> 31).sum()) == b['num_seg']
        assert b['e_idx'].size == int(np.where(cnt > t, (cnt + ch - 1) // ch, cnt).sum()) == b['e_own'].size
        # entry tiles: cover every entry once, one relation each, self-loop tiles last
        tn = b['tile_info'] & 0xff
        assert int(tn.sum()) == b['e_idx'].size and (tn.min() >= 1 if tn.size else True) and tn.max(initial=0) <= 16
        assert np.array_equal(b['tile_e0'], np.concatenate([[0], np.cumsum(tn)[:-1]]).astype(np.int32)[:tn.size])
        trel = b['tile_info'] >> 8
        assert np.all(trel[:b['num_tiles_noself']] < r) and np.all(trel[b['num_tiles_noself']:] == r)
        seg_of_entry = np.repeat(np.arange(b['num_seg']), np.diff(b['seg_ptr']))
        for i in range(tn.size):
            sl = slice(int(b['tile_e0'][i]), int(b['tile_e0'][i]) + int(tn[i]))
            assert np.all(b['seg_rel'][seg_of_entry[sl]] == trel[i])
            assert np.array_equal(b['e_own'][sl], b['seg_own'][seg_of_entry[sl]])


@settings(max_examples=80, deadline=None)
@given(graphs())
def test_shared_chunk_numbering(gr):
    """FWD_REL (relation-major) renumbered onto FWD's chunk ids: chunk c of FWD_REL and chunk
    chunk_out[c] of FWD are the same piece of the same (relation, dst) segment — same gathered rows
    and weights in the same order — so a chunk-row matrix computed for one serves the other."""
    src, dst, rel, n, r, nr, t, ch, lo, hi = gr
    w = csr_oracle.edge_weights(dst, rel, n)
    fwd = csr_oracle.build_brc(dst, src, rel, n, r, nr, t, ch, w_edge=w, lo=lo, hi=hi)
    own = csr_oracle.build_brc(dst, src, rel, n, r, hi - lo, t, ch, w_edge=w, lo=lo, hi=hi)   # one range
    shared = csr_oracle.share_chunks(fwd, own, n)
    nc = fwd['num_chunks']
    assert own['num_chunks'] == nc
    if nc == 0:
        return
    order = shared['chunk_out']
    assert sorted(order.tolist()) == list(range(nc))
    for c in range(nc):
        a0, a1 = int(own['chunk_beg'][c]), int(own['chunk_end'][c])
        b0, b1 = int(fwd['chunk_beg'][order[c]]), int(fwd['chunk_end'][order[c]])
        assert np.array_equal(own['raw_idx'][a0:a1], fwd['raw_idx'][b0:b1])
        assert np.array_equal(own['raw_w'][a0:a1], fwd['raw_w'][b0:b1])
    # the renumbered entries point at FWD ids; everything else is untouched
    idx_old = own['e_idx'] & np.uint32(0x7fffffff)
    idx_new = shared['e_idx'] & np.uint32(0x7fffffff)
    is_chunk = idx_old >= n
    assert np.array_equal(idx_new[~is_chunk], idx_old[~is_chunk])
    assert np.array_equal(idx_new[is_chunk].astype(np.int64) - n, order[(idx_old[is_chunk] - n).astype(np.int64)])
    assert np.array_equal(shared['e_idx'] >> 31, own['e_idx'] >> 31)
