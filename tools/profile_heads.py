"""Kernel-time breakdown of one training step of the mlp / attention experiment models at AM shape
(torch.profiler, CUDA activities only).  usage: python tools/profile_heads.py [mlp|attention]"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'scaling-rgcn-training_b200')]
import __graft_entry__  # noqa: E402,F401
from rgcn_b200 import Data, Emb_ATT_Layers, Emb_MLP_Layers  # noqa: E402
from rgcn_b200.synthetic import am_shape  # noqa: E402
from rgcn_b200.trainer import ce_loss, identity, make_optimizer  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'mlp'
dev = torch.device('cuda:0')
ei, et, n, r = am_shape(1.0, seed=0)
d = Data(edge_index=ei.to(dev))
d.edge_type = et.to(dev)
S, EMB = 3, 63
if which == 'mlp':
    m = Emb_MLP_Layers(r, 16, 11, n, EMB, S)
    m.load_embedding(torch.randn(n, S * EMB), freeze=True)
else:
    m = Emb_ATT_Layers(r, 16, 11, n, EMB, S)
    m.load_embedding(torch.randn(S, n, EMB), freeze=True)
m = m.to(dev)
opt = make_optimizer(m)
xt = torch.randint(0, n, (1000,), device=dev)
yt = torch.zeros(1000, 11, device=dev)
yt[:, 0] = 1


def step():
    opt.zero_grad()
    loss = ce_loss(m(d, identity)[xt], yt)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40, max_name_column_width=70))
# timeline of the last step: every device activity in start order, with the idle gap in front of it
import json
import tempfile
path = os.path.join(tempfile.gettempdir(), 'heads_trace.json')
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset') and 'dur' in e]
ev.sort(key=lambda e: e['ts'])
ev = ev[len(ev) * 2 // 3:]
t0, end, busy = ev[0]['ts'], ev[0]['ts'], 0.0
print('--- last step: start_us  dur_us  gap_before_us  stream  name')
for e in ev:
    gap = e['ts'] - end
    print('%9.1f %8.1f %8.1f  %s  %s' % (e['ts'] - t0, e['dur'], max(gap, 0.0), e.get('args', {}).get('stream', '?'), e['name'][:90]))
    busy += e['dur']
    end = max(end, e['ts'] + e['dur'])
print('span_us', end - t0, 'sum_of_durations_us', busy)
