"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: range planning, padded row
all-gather with global indexing, flat gradient all-reduce, and the owner-computes algebra of the
partitioned layer (local compute supplied by the ORACLE here — the engine itself has no CPU path)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ORACLE, PKG, REPO, golden_graph


def test_plan_ranges():
    from rgcn_b200.partition import plan_ranges
    chunk, r = plan_ranges(10, 3)
    assert chunk == 4 and r == [(0, 4), (4, 8), (8, 10)]
    chunk, r = plan_ranges(1666764, 8)
    assert r[0][0] == 0 and r[-1][1] == 1666764 and all(b - a <= chunk for a, b in r)
    with pytest.raises(ValueError):
        plan_ranges(5, 4)          # ceil(5/4)=2 -> 4th range empty
    with pytest.raises(ValueError):
        plan_ranges(3, 8)


def test_plan_ranges_balanced():
    import numpy as np
    from rgcn_b200.partition import plan_ranges_balanced
    rng = np.random.default_rng(0)
    w = np.minimum(rng.zipf(2.0, size=10007), 500).astype(np.float64)          # hubs
    for world in (2, 3, 8):
        r = plan_ranges_balanced(w, world)
        assert r[0][0] == 0 and r[-1][1] == w.size and all(a < b for a, b in r)
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        assert all(b % 4 == 0 for _, b in r[:-1])
        loads = np.array([w[a:b].sum() for a, b in r])
        equal = np.array([w[a:b].sum() for a, b in [(i * w.size // world, (i + 1) * w.size // world) for i in range(world)]])
        assert loads.max() <= loads.mean() + w.max() + 4 * w.mean()
        assert loads.max() <= equal.max() + w.max()
    assert plan_ranges_balanced(np.ones(5), 5) == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5)]
    with pytest.raises(ValueError):
        plan_ranges_balanced(np.ones(3), 4)


def _worker(rank, world, port, ret):
    for p in (REPO, PKG, ORACLE):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import numpy as np
        import rgcn_oracle
        from rgcn_b200.partition import RowComm
        sys.path.insert(0, os.path.join(REPO, 'tests'))
        ei, et, n, r = golden_graph('MUTAG_bisim_k1')          # N=66 -> ranges of 33
        n_odd = n - 1                                          # uneven last range exercises the padding
        comm = RowComm(n_odd)
        lo, hi = comm.lo, comm.hi
        torch.manual_seed(0)
        keep = (ei[0] < n_odd) & (ei[1] < n_odd)
        ei, et = ei[:, keep], et[keep]
        x = torch.randn(n_odd, 5)
        w = torch.randn(r, 5, 3) * 0.2
        root = torch.randn(5, 3) * 0.2
        bias = torch.randn(3) * 0.1
        gout = torch.randn(n_odd, 3)
        # all-gather: rows land at their global index
        x_all = comm.all_gather_rows(x[lo:hi].clone())
        assert x_all.shape == (comm.world * comm.chunk, 5) and torch.equal(x_all[:n_odd], x)
        # owner-computes forward: every rank needs only the edges whose dst it owns
        ref = rgcn_oracle.rgcn_forward(x, ei, et, w, root, bias)
        mine_dst = (ei[1] >= lo) & (ei[1] < hi)
        out_full = rgcn_oracle.rgcn_forward(x_all[:n_odd], ei[:, mine_dst], et[mine_dst], w, root, bias)
        assert torch.allclose(out_full[lo:hi], ref[lo:hi], atol=1e-6)
        # parameter gradients: partial sums over owned dst, one flat all-reduce
        leaves = [t.clone().requires_grad_() for t in (x, w, root, bias)]
        rgcn_oracle.rgcn_forward(*[leaves[0], ei, et] + leaves[1:]).backward(gout)
        part = [t.clone().requires_grad_() for t in (w, root, bias)]
        o = rgcn_oracle.rgcn_forward(x, ei[:, mine_dst], et[mine_dst], *part)
        o[lo:hi].backward(gout[lo:hi])
        grads = [p.grad.clone() for p in part]
        comm.all_reduce_sum_(grads)
        for a, b in zip(grads, leaves[1:]):
            assert torch.allclose(a, b.grad, atol=1e-5), 'all-reduced parameter gradients'
        # source-partitioned ("push") forward: every rank runs the layer over the edges whose SRC it owns from
        # its own x rows (root + bias on its own rows only) and the partials are reduce-scattered
        mine_src = (ei[0] >= lo) & (ei[0] < hi)
        # mean normalisers are those of the WHOLE graph: scale each edge's message by cnt_local / cnt_global
        pad_n = comm.world * comm.chunk
        partial = torch.zeros(pad_n, 3)
        key_all = et * n_odd + ei[1]
        cnt_all = torch.bincount(key_all, minlength=r * n_odd).float()
        for e_id in torch.nonzero(mine_src).flatten().tolist():
            s_, d_, t_ = int(ei[0, e_id]), int(ei[1, e_id]), int(et[e_id])
            partial[d_] += (x[s_] @ w[t_]) / cnt_all[t_ * n_odd + d_]
        partial[lo:hi] += x[lo:hi] @ root + bias
        mine = comm.reduce_scatter_rows(partial)
        assert mine.shape == (comm.chunk, 3) and comm.bytes_reduced == partial.numel() * 4
        assert torch.allclose(mine[:hi - lo], ref[lo:hi], atol=1e-5), 'reduce-scattered partial outputs'
        # sparse exchange plan (host logic of NvlComm.set_touched): which ranks reach an owned row / which remote
        # rows this rank reaches, and that exchanging ONLY those reproduces the dense exchanges
        from rgcn_b200.partition import sparse_plan
        touched = torch.zeros(n_odd, dtype=torch.bool)
        touched[ei[1][mine_src]] = True
        touched[lo:hi] = True
        plan = sparse_plan(touched, comm.ranges, rank)
        all_touched = [torch.zeros(n_odd, dtype=torch.bool) for _ in range(world)]
        for p_, (a, b) in enumerate(comm.ranges):
            m_ = (ei[0] >= a) & (ei[0] < b)
            all_touched[p_][ei[1][m_]] = True
            all_touched[p_][a:b] = True
        want_mask = sum(all_touched[p_][lo:hi].to(torch.int32) << p_ for p_ in range(world))
        assert torch.equal(plan['peer_mask'], want_mask) and plan['peer_mask'].dtype == torch.int32
        rows = plan['row_list'].long()
        assert torch.equal(rows, torch.nonzero(touched & ~torch.isin(torch.arange(n_odd), torch.arange(lo, hi))).flatten())
        offs = plan['offsets']
        assert len(offs) == world + 1 and offs[0] == 0 and offs[-1] == rows.numel() and offs[rank] == offs[rank + 1]
        for p_, (a, b) in enumerate(comm.ranges):
            seg = rows[offs[p_]:offs[p_ + 1]]
            assert bool(((seg >= a) & (seg < b)).all())
        # every non-zero row of this rank's partial is a touched row; masked sum == dense sum
        assert not bool(partial[:n_odd][~touched].any())
        parts = [torch.empty_like(partial) for _ in range(world)]
        dist.all_gather(parts, partial)
        sparse_sum = torch.zeros(hi - lo, 3)
        for p_ in range(world):
            sel = ((plan['peer_mask'] >> p_) & 1).bool()
            sparse_sum[sel] += parts[p_][lo:hi][sel]
        assert torch.equal(sparse_sum, sum(q[lo:hi] for q in parts))
        assert 0.0 <= plan['remote_saving_mean'] <= 1.0 and 0.0 < plan['touched_frac_mean'] <= 1.0
        ret[rank] = 'ok'
    finally:
        dist.destroy_process_group()


def test_row_comm_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29611, ret), nprocs=world, join=True)
    assert dict(ret) == {0: 'ok', 1: 'ok'}
