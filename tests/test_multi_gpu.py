"""Real multi-GPU check of the destination-partitioned path (needs >= 2 CUDA devices; run with
`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`): two NCCL ranks, each owning
half of the nodes, must reproduce the single-GPU forward and every gradient."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ORACLE, PKG, REPO, golden_graph

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    for p in (REPO, PKG, os.path.join(REPO, 'tests')):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from rgcn_b200 import RGCNGraph, rgcn_layer
        from rgcn_b200.partition import RowComm
        ei, et, n, r = golden_graph('AIFB_bisim_k3')
        ei, et = ei.to(dev), et.to(dev)
        torch.manual_seed(0)
        x = torch.randn(n, 63, device=dev)
        w = (torch.rand(r, 63, 16, device=dev) - 0.5) * 0.3
        root = (torch.rand(63, 16, device=dev) - 0.5) * 0.3
        bias = torch.rand(16, device=dev)
        gout = torch.randn(n, 16, device=dev)
        full = [t.clone().requires_grad_() for t in (x, w, root, bias)]
        ref = rgcn_layer(*full, RGCNGraph(ei, et, n, r))
        ref.backward(gout)
        comm = RowComm(n)

        def rel(a, b):
            return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
        errs = []
        for push in (False, True):      # destination-partitioned (all-gather) and source-partitioned (reduce-scatter)
            g = RGCNGraph(ei, et, n, r, own_range=(comm.lo, comm.hi), push=push)
            part = [x[comm.lo:comm.hi].clone().requires_grad_()] + [t.clone().requires_grad_() for t in (w, root, bias)]
            out = rgcn_layer(*part, g, comm=comm)
            out.backward(gout[comm.lo:comm.hi])
            errs += [rel(out, ref[comm.lo:comm.hi]), rel(part[0].grad, full[0].grad[comm.lo:comm.hi])]
            errs += [rel(a.grad, b.grad) for a, b in zip(part[1:], full[1:])]
        ret[rank] = max(errs)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_rank_partitioned_layer_matches_single_gpu():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29641, ret), nprocs=2, join=True)
    assert set(ret.keys()) == {0, 1}
    assert max(ret.values()) < 1e-5, dict(ret)
