"""Probe (GPU box, torchrun): torch symmetric memory over NVLink — peer buffers, multicast, barrier cost."""
import os, sys, time, json
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
out = {}
try:
    n = 16 * 1024 * 1024
    t = sm.empty((n,), dtype=torch.float32, device=dev)
    h = sm.rendezvous(t, dist.group.WORLD)
except Exception as e:
    out['rendezvous_err'] = repr(e)[:300]
    h = None
if h is not None:
    try:
        from torch._C._autograd import DeviceType
        out['multicast'] = bool(sm._SymmetricMemory.has_multicast_support(DeviceType.CUDA, local))
    except Exception as e:
        out['multicast_err'] = repr(e)[:200]
    try:
        out['multicast_ptr'] = int(h.multicast_ptr)
    except Exception as e:
        out['multicast_ptr_err'] = repr(e)[:200]
    out['buffer_ptrs'] = [hex(p) for p in h.buffer_ptrs]
    t.fill_(rank + 1)
    torch.cuda.synchronize(); dist.barrier()
    peer = h.get_buffer((rank + 1) % world, (n,), torch.float32, 0)
    out['peer_val'] = float(peer[0].item())
    # barrier cost
    for _ in range(5): h.barrier(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): h.barrier(0)
    e1.record(); torch.cuda.synchronize()
    out['barrier_us'] = e0.elapsed_time(e1) * 10
    # push copy bandwidth: local -> peer buffer
    src = torch.randn(n, device=dev)
    for _ in range(3): peer.copy_(src)
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for _ in range(20): peer.copy_(src)
    e1.record(); torch.cuda.synchronize()
    out['p2p_push_gbs'] = 20 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    # graph capture of barrier + copy
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            h.barrier(0); peer.copy_(src)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize(); dist.barrier()
        with torch.cuda.graph(g):
            h.barrier(0); peer.copy_(src); h.barrier(0)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        out['graph_capture'] = 'ok'
    except Exception as e:
        out['graph_capture'] = repr(e)[:300]
    # NCCL all_gather / reduce_scatter bandwidth at bench sizes for comparison
    nrow = 1666764 // world + 1
    for f in (16, 12):
        a = torch.randn(nrow, f, device=dev); b = torch.empty(world * nrow, f, device=dev)
        for _ in range(3): dist.all_gather_into_tensor(b, a)
        torch.cuda.synchronize(); dist.barrier()
        e0.record()
        for _ in range(10): dist.all_gather_into_tensor(b, a)
        e1.record(); torch.cuda.synchronize()
        out[f'nccl_ag_{f}_ms'] = e0.elapsed_time(e1) / 10
        for _ in range(3): dist.reduce_scatter_tensor(a, b)
        torch.cuda.synchronize(); dist.barrier()
        e0.record()
        for _ in range(10): dist.reduce_scatter_tensor(a, b)
        e1.record(); torch.cuda.synchronize()
        out[f'nccl_rs_{f}_ms'] = e0.elapsed_time(e1) / 10
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
