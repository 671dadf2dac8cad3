"""BASELINE.json configs[4]: synthetic 10x AM-shape graph, hidden 64, basis-decomposed weights
(``num_bases=30``, PyG semantics — the reference itself never sets it, /root/reference/model/layers.py:15-16),
transferred R-GCN weights frozen (``-w_grad False``, main.py:87) and an MLP transfer head
(model/layers.py:90-112) as the only trainable part.  ``bench.py --workload am10x_h64_bases30_frozenW``.

What one step computes
  value leg   W_r = comp @ bases (both layers), forward 63 -> 64 -> 11 with the fused ReLU, backward down to
              dL/dx0 (the only gradient anyone upstream needs when the weights are frozen)
  e2e leg     x0 = lin2(tanh(lin1(E_cat))) from the frozen concatenated summary embeddings [N, 3*63], the two
              layers, CE loss on the labelled batch (copied from pinned host memory every step), backward
              through the layers into the head, Adam on the head's parameters, loss.item()
Algorithmic bytes: BASELINE.md section 2 row 4 (127.5 GB per step at 10x).
"""
from __future__ import annotations

import json
import os
import time

import torch
import torch.distributed as dist
from torch import nn

WORKLOAD = 'am10x_h64_bases30_frozenW'
EMB, HIDDEN, CLASSES, BASES, SUMS = 63, 64, 11, 30, 3


def _step_bytes(n: int, e: int) -> int:
    def layer(fin, fout):
        fwd = e * (4 * fin + 8) + n * 4 * (fin + fout)
        bwd = n * 4 * fout + e * (4 * fout + 8) + n * 4 * fin        # dL/dx only
        return fwd + bwd
    return layer(EMB, HIDDEN) + layer(HIDDEN, CLASSES)


class FrozenBasisRGCN(nn.Module):
    """Two basis-decomposed R-GCN layers with frozen parameters over a (possibly partitioned) graph."""

    def __init__(self, graph, num_relations: int, comm=None, seed: int = 0) -> None:
        super().__init__()
        from .conv import RGCNConv
        self.graph, self.comm = graph, comm
        if comm is not None and getattr(graph, 'touched', None) is not None and hasattr(comm, 'set_touched'):
            comm.set_touched(graph.touched)
        torch.manual_seed(seed)                               # same replicated weights on every rank
        self.rgcn1 = RGCNConv(EMB, HIDDEN, num_relations, num_bases=BASES)
        self.rgcn2 = RGCNConv(HIDDEN, CLASSES, num_relations, num_bases=BASES)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, x0: torch.Tensor) -> torch.Tensor:
        from .conv import rgcn_layer
        c1, c2 = self.rgcn1, self.rgcn2
        # basis expansion as RGCNConv.forward does it on every call (a [R,B] x [B, Fin*Fout] product)
        w1 = (c1.comp @ c1.weight.view(BASES, -1)).view(c1.num_relations, EMB, HIDDEN)
        w2 = (c2.comp @ c2.weight.view(BASES, -1)).view(c2.num_relations, HIDDEN, CLASSES)
        h = rgcn_layer(x0, w1, c1.root, c1.bias, self.graph, comm=self.comm, comm_key=1)
        return rgcn_layer(h, w2, c2.root, c2.bias, self.graph, relu_in=True, comm=self.comm, comm_key=2)


class MLPHead(nn.Module):
    """lin2(tanh(lin1(E_cat))) with the reference's widths and init (model/layers.py:91-103)."""

    def __init__(self, seed: int = 1) -> None:
        super().__init__()
        torch.manual_seed(seed)
        width_in = SUMS * EMB
        width_mid = round((width_in * (2 / 3)) + CLASSES)
        self.lin1 = nn.Linear(width_in, width_mid)
        self.lin2 = nn.Linear(width_mid, EMB)
        for lin in (self.lin1, self.lin2):
            nn.init.kaiming_uniform_(lin.weight, mode='fan_in')

    def forward(self, e_cat: torch.Tensor) -> torch.Tensor:
        if e_cat.is_cuda and os.environ.get('RGCN_B200_ENGINE_HEAD', '1') != '0':
            from .heads import mlp_head                       # the engine's tcgen05 contraction (csrc/gemm_tc.cu)
            return mlp_head(e_cat, self.lin1, self.lin2)
        return self.lin2(torch.tanh(self.lin1(e_cat)))


def run(args, rank: int, world: int, device, metric: str, unit: str) -> None:
    import bench as B
    from . import _lib
    from .graph import RGCNGraph
    from .partition import NvlComm, RowComm, balanced_ranges_for, comm_backend, make_comm
    from .synthetic import am_shape, labelled_split
    from .trainer import make_optimizer
    scale = 10.0 * args.scale
    ei, et, n, r = am_shape(scale=scale)
    e = et.numel()
    comm = None
    if world > 1:
        nvl = comm_backend() == 'nvl'
        comm = make_comm(n, steady_state=True, ranges=balanced_ranges_for(ei, n, world) if nvl else None)
    lo, hi = (comm.lo, comm.hi) if comm is not None else (0, n)
    t0 = time.perf_counter()
    ei_d, et_d = ei.to(device), et.to(device)
    graph = RGCNGraph(ei_d, et_d, n, r, own_range=(lo, hi) if comm is not None else None, push=comm is not None)
    del ei_d, et_d
    torch.cuda.synchronize(device)
    setup_ms = (time.perf_counter() - t0) * 1e3
    model = FrozenBasisRGCN(graph, r, comm).to(device)
    gen = torch.Generator().manual_seed(2)
    x0 = torch.randn(hi - lo, EMB, generator=gen).to(device).requires_grad_()
    gout = torch.randn(hi - lo, CLASSES, generator=gen).to(device)

    def step():
        x0.grad = None
        model(x0).backward(gout)

    steps, warmup = args.steps, max(args.warmup, 3)
    peak, peak_src = B.load_peaks()
    sampler = B.ClockSampler(device.index)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize(device)
    launches0 = _lib.launch_count()
    step()
    launches = _lib.launch_count() - launches0
    _lib.set_option(_lib.OPT_OVERLAP, 0)
    _lib.profile_enable(True)
    _lib.profile_collect()
    step()
    torch.cuda.synchronize(device)
    passes = {}
    for name, dims, pms in _lib.profile_collect():
        k = f'{name}_{dims[0]}x{dims[1]}'
        passes[k] = passes.get(k, 0.0) + pms
    _lib.profile_enable(False)
    _lib.set_option(_lib.OPT_OVERLAP, 1)
    run_step = step
    graph_obj = None
    if world > 1 and os.environ.get('RGCN_B200_PART_GRAPH', '1') != '0':
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        dist.barrier()
        graph_obj = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_obj):
            step()
        run_step = graph_obj.replay
    if rank == 0:
        sampler.start()
    ms = B.time_steps(run_step, steps, warmup, world, device) / steps
    clocks = sampler.stop() if rank == 0 else None
    value = e / (ms * 1e-3)
    step_bytes = _step_bytes(n, e)

    # end to end with the trainable MLP head
    e2e = None
    if not args.no_e2e:
        head = MLPHead().to(device)
        from .heads import rows16
        e_cat = rows16(torch.randn(hi - lo, SUMS * EMB, generator=torch.Generator().manual_seed(3)).to(device))   # frozen: rows padded once
        x_all, y_all = labelled_split(n, CLASSES)
        mine = (x_all >= lo) & (x_all < hi)
        x_h, y_h = (x_all[mine] - lo).contiguous().pin_memory(), y_all[mine].contiguous().pin_memory()
        m_total = float(x_all.numel())
        opt = make_optimizer(head)
        hp = list(head.parameters())

        def e2e_step():
            xs, ys = x_h.to(device, non_blocking=True), y_h.to(device, non_blocking=True)
            opt.zero_grad()
            out = model(head(e_cat))[xs]
            picked = out.gather(1, ys.to(torch.float32).argmax(-1, keepdim=True)).squeeze(1)
            local = (torch.logsumexp(out, dim=1) - picked).sum() / m_total
            local.backward()
            if comm is not None:
                comm.all_reduce_sum_([p.grad for p in hp])
            opt.step()
            tot = local.detach().clone()
            if world > 1:
                dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            return tot.item()

        e2e_ms = B.time_steps(e2e_step, steps, warmup, world, device) / steps
        h2d = torch.tensor([x_h.numel() * 8 + y_h.numel() * 8], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(h2d, op=dist.ReduceOp.SUM)
        e2e = {'value': e / (e2e_ms * 1e-3), 'unit': unit, 'h2d_bytes_per_step': int(h2d.item()),
               'd2h_bytes_per_step': 4 * world, 'ms_per_step': e2e_ms,
               'what': 'x0 = lin2(tanh(lin1(E_cat))) (MLP transfer head, the only trainable part; forward and dL/dh on the engine tcgen05 kernel, parameter-gradient reductions in torch), '
                       'the two frozen basis layers, CE loss on the labelled batch from pinned host memory, backward '
                       'into the head, FusedAdam on the head, loss.item()'}
    if rank == 0:
        line = {
            'metric': metric, 'value': value, 'unit': unit, 'n_gpus': world, 'steps': steps, 'warmup': warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'scale': scale, 'nodes': n, 'directed_edges': e, 'relations': r,
                       'emb': EMB, 'hidden': HIDDEN, 'classes': CLASSES, 'num_bases': BASES,
                       'frozen': 'weight/comp/root/bias of both layers (requires_grad False); dL/dx0 only',
                       'partition': (f'source-partitioned x{world}' if world > 1 else 'one GPU'),
                       'comm_backend': ('nvl' if isinstance(comm, NvlComm) else 'nccl') if comm is not None else None,
                       'exchange_sparse': getattr(comm, 'sparse', None) is not None,
                       'rows_reached_per_rank_frac_mean': getattr(comm, 'touched_frac', None),
                       'step_algorithmic_bytes': step_bytes,
                       'step_roofline_frac': step_bytes / (ms * 1e-3) / 1e9 / peak / max(world, 1),
                       'peak_source': peak_src, 'graph_build_ms_once': setup_ms,
                       'passes_rank0_serialised_ms': passes,
                       'l2_policy': 'inputs larger than L2; no flush'},
            'clocks': clocks, 'gpu_launches': int(launches * steps), 'e2e': e2e,
            'roofline': {'bound': 'hbm', 'kernel': 'whole step (BASELINE.md section 2 row 4)', 'achieved': step_bytes / (ms * 1e-3) / 1e9 / max(world, 1),
                         'peak': peak, 'unit': 'GB/s', 'frac': step_bytes / (ms * 1e-3) / 1e9 / peak / max(world, 1),
                         'traffic': None, 'peak_source': peak_src},
            'cpu_baseline': None,
        }
        print(json.dumps(line), flush=True)
    del graph_obj, run_step
    if world > 1:
        torch.cuda.synchronize(device)
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[2]: multi-summary pre-training, one summary graph per GPU
# ----------------------------------------------------------------------------------------------------------------
MULTI_SUMMARY = 'multi_summary_pretraining_8_aifb_summaries'


def _golden_summary_graphs(count: int, classes: int):
    """`count` AIFB-sized summary graphs: the real AIFB attr / bisim summaries shipped as fixtures (tests/golden/,
    made by the reference's own Graph.init_graph), cycled, each with its own seeded fractional labels."""
    import types
    import numpy as np
    from .data import Data
    here = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    names = ['AIFB_sum_in', 'AIFB_sum_in_out', 'AIFB_bisim_k3']
    out = []
    for k in range(count):
        g = dict(np.load(os.path.join(here, 'tests', 'golden', f'graph_{names[k % len(names)]}.npz'), allow_pickle=False))
        buf = torch.from_numpy(np.concatenate([g['edge_index'], g['edge_type'][None]], 0).astype(np.int64).T.copy())
        edge = buf.t()
        n, r = int(g['num_nodes']), int(g['num_relations'])
        sg = types.SimpleNamespace(num_nodes=n, relations={f'p{i}': i for i in range((r - 1) // 2)}, name=names[k % len(names)])
        td = Data(edge_index=edge[:2])
        td.edge_type = edge[2]
        gen = torch.Generator().manual_seed(100 + k)
        td.x_train = torch.randperm(n, generator=gen)[:max(2, n // 2)]
        td.y_train = torch.rand(td.x_train.numel(), classes, generator=gen)     # summary labels are fractions (BCE)
        sg.training_data = td
        out.append(sg)
    return out


def run_multi_summary(args, rank: int, world: int, device, metric: str, unit: str) -> None:
    import bench as B
    from . import _lib
    from .trainer import PARALLEL_SEMANTICS, train_summaries_parallel
    graphs = _golden_summary_graphs(8, 26)
    epochs = max(args.steps, 1)
    edges = sum(g.training_data.edge_type.numel() for g in graphs)
    # warm-up: CSR builds, lazy module loads (the graphs are cached by tensor identity inside RGCNConv)
    for g in graphs:
        g.training_data.to(device)
    train_summaries_parallel(graphs, 26, 16, 2, 63, 0.01, 5e-5, device, seed=0)
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    sampler = B.ClockSampler(device.index)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    res = train_summaries_parallel(graphs, 26, 16, epochs, 63, 0.01, 5e-5, device, seed=0)
    t1.record()
    torch.cuda.synchronize(device)
    ms = torch.tensor([t0.elapsed_time(t1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(ms.item())
    value = edges * epochs / (total_ms * 1e-3)
    first, last = res['losses'][0][0], res['losses'][-1][-1]
    if rank == 0:
        line = {
            'metric': metric, 'value': value, 'unit': unit, 'n_gpus': world, 'steps': epochs, 'warmup': 2,
            'ms_per_step': total_ms / epochs, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'real AIFB summary graphs (tests/golden fixtures), seeded labels',
            'config': {'workload': MULTI_SUMMARY, 'summaries': [g.name for g in graphs], 'directed_edges_all_summaries': edges,
                       'epochs': epochs, 'rounds': res['rounds'], 'emb': 63, 'hidden': 16, 'classes': 26,
                       'step': 'Trainer.train iteration body per summary (fwd, BCE on sigmoid, bwd, FusedAdam, loss.item()); '
                               'value = edges of all summaries x epochs / wall time of the whole pre-training',
                       'semantics': PARALLEL_SEMANTICS if world > 1 else
                       'sequential with carried weights and a fresh Adam per summary = the reference (modelTrainer.py:76-82)',
                       'loss_first_last_rank0': [first, last],
                       'l2_policy': 'L2-resident working set (52 MB per summary step): reported, not graded against HBM'},
            'clocks': clocks, 'gpu_launches': int(_lib.launch_count() - l0), 'e2e': {
                'value': value, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 4 * min(world, 8),
                'ms_per_step': total_ms / epochs, 'what': 'same run: the loss of every summary is read back every epoch'},
            'roofline': None, 'cpu_baseline': None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize(device)
        dist.barrier()
        dist.destroy_process_group()
