"""Probe (GPU box, torchrun): bandwidth of the engine's NVLink exchange kernels in their three modes."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'scaling-rgcn-training_b200')]
import torch, torch.distributed as dist
import __graft_entry__; __graft_entry__.build()
from rgcn_b200 import _lib
from rgcn_b200.partition import NvlComm

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
n = 1666764
comm = NvlComm(n, steady_state=True)
out = {'world': world, 'multicast': comm.multicast}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for mode in (1, 0, 2):
    _lib.set_option(_lib.OPT_NVL_MODE, mode)
    for f in (16, 12):
        t = torch.randn(comm.hi - comm.lo, f, device=dev)
        for _ in range(3): comm.all_gather_rows(t, key=('p', f))
        torch.cuda.synchronize(); dist.barrier()
        e0.record()
        for _ in range(10): comm.all_gather_rows(t, key=('p', f))
        e1.record(); torch.cuda.synchronize()
        out[f'mode{mode}_ag{f}_ms'] = e0.elapsed_time(e1) / 10
        pb = comm.partial_buffer(f, ('p', f), dev)
        pb.normal_()
        for _ in range(3): comm.reduce_scatter_rows(pb, ('p', f))
        torch.cuda.synchronize(); dist.barrier()
        e0.record()
        for _ in range(10): comm.reduce_scatter_rows(pb, ('p', f))
        e1.record(); torch.cuda.synchronize()
        out[f'mode{mode}_rs{f}_ms'] = e0.elapsed_time(e1) / 10
# correctness of the reduce in every mode
for mode in (1, 0, 2):
    _lib.set_option(_lib.OPT_NVL_MODE, mode)
    pb = comm.partial_buffer(16, ('p', 16), dev)
    pb.fill_(float(rank + 1))
    torch.cuda.synchronize(); dist.barrier()
    r = comm.reduce_scatter_rows(pb, ('p', 16))
    out[f'mode{mode}_sum_ok'] = bool((r == world * (world + 1) / 2).all().item())
for _ in range(20): comm.barrier()
torch.cuda.synchronize()
e0.record()
for _ in range(100): comm.barrier()
e1.record(); torch.cuda.synchronize()
out['barrier_us'] = e0.elapsed_time(e1) * 10
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
