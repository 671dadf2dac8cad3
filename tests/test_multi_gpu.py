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
        from rgcn_b200.partition import NvlComm
        nvl = NvlComm(n)                # the engine's own exchanges over NVLink multicast / peer memory
        # destination-partitioned (all-gather), source-partitioned (reduce-scatter) over NCCL, and source-
        # partitioned over the engine's exchange kernels
        for push, c in ((False, comm), (True, comm), (True, nvl)):
            g = RGCNGraph(ei, et, n, r, own_range=(comm.lo, comm.hi), push=push)
            part = [x[comm.lo:comm.hi].clone().requires_grad_()] + [t.clone().requires_grad_() for t in (w, root, bias)]
            out = rgcn_layer(*part, g, comm=c)
            out.backward(gout[comm.lo:comm.hi])
            errs += [rel(out, ref[comm.lo:comm.hi]), rel(part[0].grad, full[0].grad[comm.lo:comm.hi])]
            errs += [rel(a.grad, b.grad) for a, b in zip(part[1:], full[1:])]
        # two layers with the fused ReLU: the mask rides on the exchange of the upper layer's gx (handoff)
        w2 = (torch.rand(r, 16, 11, device=dev) - 0.5) * 0.5
        f2 = [t.clone().requires_grad_() for t in (x, w, root, bias, w2)]
        g1 = RGCNGraph(ei, et, n, r)
        ref2 = rgcn_layer(rgcn_layer(f2[0], f2[1], f2[2], f2[3], g1), f2[4], None, None, g1, relu_in=True)
        gout2 = torch.randn(n, 11, device=dev)
        ref2.backward(gout2)
        g = RGCNGraph(ei, et, n, r, own_range=(comm.lo, comm.hi), push=True)
        p2 = [x[comm.lo:comm.hi].clone().requires_grad_()] + [t.clone().requires_grad_() for t in (w, root, bias, w2)]
        out2 = rgcn_layer(rgcn_layer(p2[0], p2[1], p2[2], p2[3], g, comm=nvl, comm_key=1), p2[4], None, None, g,
                          relu_in=True, comm=nvl, comm_key=2)
        out2.backward(gout2[comm.lo:comm.hi])
        errs += [rel(out2, ref2[comm.lo:comm.hi]), rel(p2[0].grad, f2[0].grad[comm.lo:comm.hi])]
        errs += [rel(a.grad, b.grad) for a, b in zip(p2[1:], f2[1:])]
        # the same two layers with the SPARSE exchanges planned from the graph (only the rows a rank reaches cross
        # the links): bit-identical to the dense peer-pointer exchanges, hence the same errors
        os.environ['RGCN_B200_NVL_SPARSE'] = '1'
        assert nvl.set_touched(g.touched) and nvl.sparse is not None
        assert nvl.sparse[1] == int(g.touched.sum()) - (comm.hi - comm.lo)
        p3 = [x[comm.lo:comm.hi].clone().requires_grad_()] + [t.clone().requires_grad_() for t in (w, root, bias, w2)]
        out3 = rgcn_layer(rgcn_layer(p3[0], p3[1], p3[2], p3[3], g, comm=nvl, comm_key=1), p3[4], None, None, g,
                          relu_in=True, comm=nvl, comm_key=2)
        out3.backward(gout2[comm.lo:comm.hi])
        errs += [rel(out3, ref2[comm.lo:comm.hi]), rel(p3[0].grad, f2[0].grad[comm.lo:comm.hi])]
        errs += [rel(a.grad, b.grad) for a, b in zip(p3[1:], f2[1:])]
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
