import os, sys, time, torch
REPO='/root/repo'
sys.path[:0]=[REPO, os.path.join(REPO,'scaling-rgcn-training_b200')]
import __graft_entry__
from rgcn_b200 import RGCNGraph
from rgcn_b200.synthetic import am_shape
ei, et, n, r = am_shape(1.0)
dev=torch.device('cuda:0')
torch.cuda.synchronize()
for rep in range(3):
    t0=time.perf_counter(); ei_d, et_d = ei.to(dev), et.to(dev); torch.cuda.synchronize(); t1=time.perf_counter()
    g=RGCNGraph(ei_d, et_d, n, r); torch.cuda.synchronize(); t2=time.perf_counter()
    print('rep',rep,'h2d_ms',(t1-t0)*1e3,'create_ms',(t2-t1)*1e3, flush=True)
    del g
    torch.cuda.synchronize()
# device-time of the create call
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    g=RGCNGraph(ei_d, et_d, n, r); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=60))
print(prof.key_averages().table(sort_by='cpu_time_total', row_limit=12, max_name_column_width=60))
