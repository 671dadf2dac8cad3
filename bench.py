#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 R-GCN engine (contract: see DESIGN.md section 6).

Metric (BASELINE.json): R-GCN fwd+bwd directed edges/s on the synthetic AM-shape graph
(N=1 666 764, 5 988 321 triples -> E=11 976 642 directed edges, 133 predicates -> R=267,
63 -> 16 -> 11, fp32, all gradients).  One "step" = both layers forward + backward over the
whole graph (inter-layer ReLU, root and bias included; loss/optimizer excluded).

  python bench.py --gpus N --steps K --warmup W            # engine arm
  python bench.py --impl reference ...                     # reference CPU path (oracle port) on host cores

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, 'scaling-rgcn-training_b200')
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'rgcn_fwd_bwd_directed_edges_per_s_am_shape'
UNIT = 'edges/s'
EMB, HIDDEN, CLASSES = 63, 16, 11


def algorithmic_bytes(n, e, fin, fout, need_w=True, need_x=True):
    """SURVEY.md section 8(d) / BASELINE.md section 2 formula, per layer."""
    fwd = e * (4 * fin + 8) + n * 4 * (fin + fout)
    bwd = n * 4 * fout
    if need_w:
        bwd += e * (4 * fin + 8) + n * 4 * fin
    if need_x:
        bwd += e * (4 * fout + 8) + n * 4 * fin
    return fwd, bwd


def pass_bytes(n, e, fin, fout):
    """Per-kernel split of the same formula: the dL/dW pass carries the read
    of gout (N*4*Fout) because it runs first; the dL/dx pass does not count it again."""
    return {
        ('tile_fwd', fin, fout): e * (4 * fin + 8) + n * 4 * (fin + fout),
        ('wgrad', fin, fout): n * 4 * fout + e * (4 * fin + 8) + n * 4 * fin,
        ('tile_dx', fout, fin): e * (4 * fout + 8) + n * 4 * fin,
    }


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), or None."""
    path = os.path.join(REPO, 'profiles', 'ncu_traffic.json')
    try:
        return json.load(open(path)).get(kernel)
    except Exception:
        return None


def load_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def dist_setup(n_gpus: int):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        # NCCL prints its version banner on stdout when the first communicator is created; stdout
        # carries exactly one JSON line, so fd 1 points at stderr until that has happened
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
            warm = torch.zeros(1, device=torch.device('cuda', local))
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    return rank, world, local


def time_steps(step_fn, steps: int, warmup: int, world: int, device):
    """W untimed steps, then exactly K steps between barrier+synchronize, CUDA events, max over ranks."""
    import torch.distributed as dist
    for _ in range(warmup):
        step_fn()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step_fn()
    t1.record()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    ms = torch.tensor([t0.elapsed_time(t1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's PyG loop path on host cores
# ----------------------------------------------------------------------------------------------
def cpu_reference(steps: int, warmup: int, scale: float, threads: int):
    sys.path.insert(0, os.path.join(REPO, 'oracle'))
    import rgcn_oracle                                 # bench.py's cpu_baseline leg may execute oracle/
    from rgcn_b200.synthetic import am_shape
    torch.set_num_threads(threads)
    ei, et, n, r = am_shape(scale=scale)
    e = et.numel()
    torch.manual_seed(0)
    emb = torch.randn(n, EMB, requires_grad=True)
    c1 = rgcn_oracle.RGCNConv(EMB, HIDDEN, r)
    c2 = rgcn_oracle.RGCNConv(HIDDEN, CLASSES, r)
    gout = torch.randn(n, CLASSES)
    params = [emb] + list(c1.parameters()) + list(c2.parameters())

    def step():
        for p in params:
            p.grad = None
        out = c2(torch.relu(c1(emb, ei, et)), ei, et)
        out.backward(gout)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    sample = (f'AM-shape at scale {scale:g} (N={n}, E={e}, R={r}), 63->16->11 fwd+bwd all grads, fp32, '
              f'{steps} timed steps after {warmup} warm-up; the loop path keeps an N x Fin tensor per relation, '
              f'full scale does not fit host RAM (SURVEY.md F11)')
    return e / dt, dt * 1e3, n, e, sample


def oracle_parity_check(device, scale: float = 1 / 64):
    """The engine's two layers (63 -> 16 -> 11, fused ReLU), output and ALL gradients, against the CPU oracle
    (oracle/rgcn_oracle.py: the reference's PyG loop path restated) on the AM-shape generator at `scale`."""
    sys.path.insert(0, os.path.join(REPO, 'oracle'))
    import rgcn_oracle                                 # the checker, not the thing measured
    from rgcn_b200 import RGCNGraph, rgcn_layer
    from rgcn_b200.synthetic import am_shape
    ei, et, n, r = am_shape(scale=scale)
    torch.manual_seed(4)
    p = [torch.randn(n, EMB), (torch.rand(r, EMB, HIDDEN) - 0.5) * 0.3, (torch.rand(EMB, HIDDEN) - 0.5) * 0.3,
         torch.rand(HIDDEN) - 0.5, (torch.rand(r, HIDDEN, CLASSES) - 0.5) * 0.5, (torch.rand(HIDDEN, CLASSES) - 0.5) * 0.5,
         torch.rand(CLASSES) - 0.5]
    gout = torch.randn(n, CLASSES)
    ref = [t.clone().requires_grad_() for t in p]
    want = rgcn_oracle.rgcn_forward(torch.relu(rgcn_oracle.rgcn_forward(ref[0], ei, et, ref[1], ref[2], ref[3])), ei, et,
                                    ref[4], ref[5], ref[6])
    want.backward(gout)
    g = RGCNGraph(ei.to(device), et.to(device), n, r)
    dl = [t.clone().to(device).requires_grad_() for t in p]
    out = rgcn_layer(rgcn_layer(dl[0], dl[1], dl[2], dl[3], g), dl[4], dl[5], dl[6], g, relu_in=True)
    out.backward(gout.to(device))

    def rel(a, b):
        return float((a.detach().cpu().double() - b.detach().double()).abs().max() / b.detach().abs().max().clamp_min(1e-30))
    errs = {'out': rel(out, want)}
    for name, a, b in zip(('gx', 'gW1', 'groot1', 'gbias1', 'gW2', 'groot2', 'gbias2'), dl, ref):
        errs[name] = rel(a.grad, b.grad)
    return {'against': f'CPU oracle (PyG 2.3.1 loop path restated), AM-shape at scale {scale:g} (N={n}, E={et.numel()})',
            'max_rel_err': errs, 'worst': max(errs.values()),
            'argmax_equal': bool(torch.equal(out.argmax(1).cpu(), want.argmax(1)))}


def bench_heads(n, r, data, x_train_h, y_train_h, device, world, args, peak):
    from rgcn_b200 import Emb_ATT_Layers, Emb_Layers, Emb_MLP_Layers
    from rgcn_b200 import heads as H
    from rgcn_b200.trainer import ce_loss, identity, make_optimizer
    S = 3
    out = {}
    gen = torch.Generator().manual_seed(5)
    e_cat = torch.randn(n, S * EMB, generator=gen).to(device)
    mlp = Emb_MLP_Layers(r, HIDDEN, CLASSES, n, EMB, S)
    mlp.load_embedding(e_cat, freeze=True)
    mlp = mlp.to(device)
    a16 = H.rows16(mlp.embedding.weight.detach())
    peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(REPO, 'MEASURED_PEAKS.json')) else {}
    bf16 = float(peaks.get('bf16_tflops_sustained', 1400.9))

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / reps

    mid = mlp.lin1.out_features
    h = H.gemm(a16, mlp.lin1.weight, mlp.lin1.bias, act='tanh')
    for name, fn, k_, n_ in (('lin1_tanh', lambda: H.gemm(a16, mlp.lin1.weight, mlp.lin1.bias, act='tanh'), S * EMB, mid),
                             ('lin2_into_mirror', lambda: H.gemm(h, mlp.lin2.weight, mlp.lin2.bias, out_ld=64), mid, EMB)):
        ms = timed(fn)
        useful = 2.0 * n * k_ * n_
        issued = 3 * 2.0 * n * ((k_ + 31) // 32 * 32) * ((n_ + 15) // 16 * 16)       # three tf32 MMAs per product, padded tile
        bytes_ = n * 4 * (k_ + n_)
        out[name] = {'ms': ms, 'm': n, 'k': k_, 'n': n_, 'useful_tflops': useful / ms / 1e9, 'issued_tf32_tflops': issued / ms / 1e9,
                     'tensor_frac_of_bf16_sustained_over_2': issued / ms / 1e9 / (bf16 / 2),
                     'tensor_frac_of_bf16_sustained_over_4': issued / ms / 1e9 / (bf16 / 4),
                     'hbm_frac': bytes_ / (ms * 1e-3) / 1e9 / peak}
    torch_ms = timed(lambda: mlp.lin2(torch.tanh(mlp.lin1(mlp.embedding.weight))))
    out['torch_head_fwd_ms'] = torch_ms
    out['engine_head_fwd_ms'] = out['lin1_tanh']['ms'] + out['lin2_into_mirror']['ms']
    # Trainer step per experiment model (labelled batch from pinned host memory, loss.item())
    steps = {}
    e_stack = e_cat.view(n, S, EMB).permute(1, 0, 2).contiguous()
    models = {'summation': None, 'mlp': mlp, 'mlp_torch_head': mlp, 'attention': None, 'attention_torch_head': None}
    for name in models:
        if name == 'summation':
            m = Emb_Layers(r, HIDDEN, CLASSES, n, EMB, S)
            m.load_embedding(e_cat[:, :EMB].contiguous(), freeze=True)
        elif name.startswith('attention'):
            m = Emb_ATT_Layers(r, HIDDEN, CLASSES, n, EMB, S)
            m.load_embedding(e_stack, freeze=True)
            m.engine_head = name == 'attention'
        else:
            m = mlp
            m.engine_head = name == 'mlp'
        m = m.to(device)
        opt = make_optimizer(m)

        def step():
            data.x_train = x_train_h.to(device, non_blocking=True)
            data.y_train = y_train_h.to(device, non_blocking=True)
            m.train()
            opt.zero_grad()
            loss = ce_loss(m(data, identity)[data.x_train], data.y_train.to(torch.float32))
            loss.backward()
            opt.step()
            return loss.item()
        steps[name] = time_steps(step, min(args.steps, 10), 3, world, device) / min(args.steps, 10)
        del opt
        if name in ('mlp', 'attention'):
            # the same step captured once in a CUDA graph (the eager step spends ~8 ms per step on the host side)
            from rgcn_b200.trainer import GraphedTrainStep
            data.x_train, data.y_train = x_train_h.to(device), y_train_h.to(device)
            graphed = GraphedTrainStep(m, data, make_optimizer(m, capturable=True), ce_loss, identity)
            graphed.prefetch(x_train_h, y_train_h)

            def graphed_step():
                loss = graphed()
                graphed.prefetch(x_train_h, y_train_h)
                return loss.item()
            steps[name + '_graphed'] = time_steps(graphed_step, min(args.steps, 10), 3, world, device) / min(args.steps, 10)
            del graphed
        if name == 'summation' or name.startswith('attention'):
            del m
            torch.cuda.empty_cache()
    out['trainer_step_ms'] = steps
    out['what'] = ('frozen transferred summary embeddings (S = 3), -e_freeze True; mlp = engine tcgen05 head, mlp_torch_head = '
                   'nn.Linear calls, attention = engine head (tcgen05 projections + one softmax kernel, query position 0 only), '
                   'attention_torch_head = nn.MultiheadAttention; *_graphed = the engine-head step captured in one CUDA graph (GraphedTrainStep)')
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    val, ms, n, e, sample = cpu_reference(args.steps, args.warmup, args.ref_scale, threads)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'am_shape_full_graph_rgcn_63_16_11_all_grads', 'reference_sample_scale': args.ref_scale,
                   'sample_nodes': n, 'sample_edges': e},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# engine arm
# ----------------------------------------------------------------------------------------------
def run_engine(args):
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device — the engine has no CPU path (use --impl reference for the CPU baseline)')
    import __graft_entry__
    __graft_entry__.build()
    from rgcn_b200 import Data, Emb_Layers, RGCNGraph, _lib, rgcn_layer
    from rgcn_b200.synthetic import am_shape, labelled_split
    from rgcn_b200.trainer import ce_loss, identity, make_optimizer
    rank, world, local = dist_setup(args.gpus)
    device = torch.device('cuda', local)
    torch.cuda.set_device(device)
    if args.workload == 'am10x_h64_bases30_frozenW':        # BASELINE.json configs[4]
        from rgcn_b200 import workloads
        return workloads.run(args, rank, world, device, METRIC, UNIT)
    if args.workload == 'multi_summary':                    # BASELINE.json configs[2]
        from rgcn_b200 import workloads
        return workloads.run_multi_summary(args, rank, world, device, METRIC, UNIT)
    if world > 1:
        from rgcn_b200 import partition
        return partition.run_partitioned_bench(args, rank, world, device, METRIC, UNIT)

    ei, et, n, r = am_shape(scale=args.scale)
    e = et.numel()
    # CUDA context + lazy module load (seconds on a fresh box) are not part of the graph build time
    _wi, _wt = torch.zeros((2, 4), dtype=torch.int64, device=device), torch.zeros(4, dtype=torch.int64, device=device)
    RGCNGraph(_wi, _wt, 2, 3)
    torch.cuda.synchronize(device)
    t_setup = time.perf_counter()
    data = Data(edge_index=ei)
    data.edge_type = et
    data = data.to(device)
    graph = RGCNGraph(data.edge_index, data.edge_type, n, r)
    torch.cuda.synchronize(device)
    setup_ms = (time.perf_counter() - t_setup) * 1e3

    torch.manual_seed(0)
    model = Emb_Layers(r, HIDDEN, CLASSES, n, EMB, 1).to(device)
    gout = torch.randn(n, CLASSES, device=device)
    params = list(model.parameters())
    c1, c2 = model.rgcn1, model.rgcn2

    def layer_step():
        for p in params:
            p.grad = None
        h = rgcn_layer(model.embedding.weight, c1.weight, c1.root, c1.bias, graph)
        out = rgcn_layer(h, c2.weight, c2.root, c2.bias, graph, relu_in=True)   # F.relu fused into the load
        out.backward(gout)

    sampler = ClockSampler(local)
    launches0 = None

    def timed():
        nonlocal launches0
        layer_step()

    for _ in range(args.warmup):
        layer_step()
    torch.cuda.synchronize(device)
    # the headline: K steps as the library runs them (independent passes of a layer call overlap on the
    # engine's side stream), CUDA events around the whole region
    launches0 = _lib.launch_count()
    sampler.start()
    total_ms = time_steps(layer_step, args.steps, 0, world, device)
    clocks = sampler.stop()
    launches = _lib.launch_count() - launches0
    ms_per_step = total_ms / args.steps
    value = e / (ms_per_step * 1e-3)
    # per-pass device times: the same K steps once more with the passes SERIALISED (overlap off) and an event
    # pair around every launch, so that a pass's time is its own; the roofline of the dominant kernel uses these
    _lib.set_option(_lib.OPT_OVERLAP, 0)
    _lib.profile_enable(True)
    _lib.profile_collect()
    serial_ms = time_steps(layer_step, args.steps, 1, world, device) / args.steps
    recs = _lib.profile_collect()
    _lib.profile_enable(False)
    _lib.set_option(_lib.OPT_OVERLAP, 1)
    recs = recs[len(recs) - (len(recs) // (args.steps + 1)) * args.steps:]      # drop the warm-up step's records

    # dominant kernel and its roofline
    peak, peak_src = load_peaks()
    agg = {}
    for name, dims, ms in recs:
        agg.setdefault((name, dims[0], dims[1]), []).append(ms)
    # the self-loop (root + bias) launch belongs to the pass whose bytes include the N x (Fin + Fout) term
    self_ms = {}
    for (name, d0, d1), v in list(agg.items()):
        if name == 'self_loop':
            for host in ('tile_fwd', 'tile_dx'):
                if (host, d0, d1) in agg and len(agg[(host, d0, d1)]) == len(v):
                    agg[(host, d0, d1)] = [x + y for x, y in zip(agg[(host, d0, d1)], v)]
                    self_ms[(host, d0, d1)] = statistics.mean(v)
                    del agg[(name, d0, d1)]
                    break
    tot = {k: sum(v) for k, v in agg.items()}
    pb = {}
    for (fin, fout) in ((EMB, HIDDEN), (HIDDEN, CLASSES)):
        pb.update(pass_bytes(n, e, fin, fout))
    graded = {k: v for k, v in tot.items() if k in pb}
    dom = max(graded, key=graded.get) if graded else None
    roofline = None
    if dom is not None:
        avg_ms = statistics.mean(agg[dom])
        ach = pb[dom] / (avg_ms * 1e-3) / 1e9
        kname = f'{dom[0]}_{dom[1]}x{dom[2]}'
        roofline = {'bound': 'hbm', 'kernel': kname, 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
                    'frac': ach / peak, 'traffic': ncu_traffic(kname) if args.scale == 1.0 else None, 'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': pb[dom], 'avg_launch_ms': avg_ms,
                    'share_of_step': tot[dom] / args.steps / serial_ms,
                    'timing': 'CUDA-event pair around every launch, passes serialised (RGCN_OPT_OVERLAP 0), '
                              f'{args.steps} steps; the step then takes {serial_ms:.3f} ms'}
    f1, b1 = algorithmic_bytes(n, e, EMB, HIDDEN)
    f2, b2 = algorithmic_bytes(n, e, HIDDEN, CLASSES)
    step_bytes = f1 + b1 + f2 + b2
    passes = {f'{k[0]}_{k[1]}x{k[2]}': {'avg_ms': statistics.mean(v), 'launches': len(v),
                                        'gbps_algorithmic': (pb[k] / (statistics.mean(v) * 1e-3) / 1e9) if k in pb else None,
                                        **({'of_which_self_loop_ms': self_ms[k]} if k in self_ms else {})}
              for k, v in sorted(agg.items())}

    # frozen variants of BASELINE.md section 2 (-e_freeze True / -w_grad False), same graph and kernels
    variants = {}
    if not args.no_e2e:
        emb_w = model.embedding.weight

        def frozen_x0_step():
            for p in params:
                p.grad = None
            h = rgcn_layer(emb_w.detach(), c1.weight, c1.root, c1.bias, graph)
            rgcn_layer(h, c2.weight, c2.root, c2.bias, graph, relu_in=True).backward(gout)

        wd = [t.detach() for t in (c1.weight, c1.root, c1.bias, c2.weight, c2.root, c2.bias)]

        def frozen_x0_w_step():
            with torch.no_grad():
                h = rgcn_layer(emb_w, wd[0], wd[1], wd[2], graph)
            h.requires_grad_(True)          # something upstream of layer 2 still wants dL/dh1 (a transfer head)
            rgcn_layer(h, wd[3], wd[4], wd[5], graph, relu_in=True).backward(gout)

        for name, fn, flags in (('frozen_x0', frozen_x0_step, (True, False, True, True)),
                                ('frozen_x0_frozen_w', frozen_x0_w_step, (False, False, False, True))):
            v_ms = time_steps(fn, args.steps, args.warmup, world, device) / args.steps
            fa, ba = algorithmic_bytes(n, e, EMB, HIDDEN, need_w=flags[0], need_x=flags[1])
            fb, bb = algorithmic_bytes(n, e, HIDDEN, CLASSES, need_w=flags[2], need_x=flags[3])
            vb = fa + ba + fb + bb
            variants[name] = {'ms_per_step': v_ms, 'edges_per_s': e / (v_ms * 1e-3), 'algorithmic_bytes': vb,
                              'roofline_frac': vb / (v_ms * 1e-3) / 1e9 / peak}

    # parity leg: the same two layers, all gradients, against the CPU oracle on the AM-shape generator at 1/64
    parity = None
    if not args.no_check:
        parity = oracle_parity_check(device)

    # end to end: the reference's Trainer.train iteration body through the public (drop-in) API,
    # labelled batch copied from pinned host memory every step, loss read back every step
    x_train_h, y_train_h = labelled_split(n, CLASSES)
    x_train_h, y_train_h = x_train_h.pin_memory(), y_train_h.pin_memory()
    h2d = x_train_h.numel() * x_train_h.element_size() + y_train_h.numel() * y_train_h.element_size()

    def make_e2e_step(opt):
        def e2e_step():
            data.x_train = x_train_h.to(device, non_blocking=True)
            data.y_train = y_train_h.to(device, non_blocking=True)
            model.train()
            opt.zero_grad()
            out = model(data, identity)
            loss = ce_loss(out[data.x_train], data.y_train.to(torch.float32))
            loss.backward()
            opt.step()
            return loss.item()
        return e2e_step

    e2e_ms = torch_adam_ms = eager_ms = float('nan')
    if not args.no_e2e:
        # the package's Trainer step: engine layers + the engine's one-pass Adam (rgcn_b200.trainer.make_optimizer)
        eager_ms = time_steps(make_e2e_step(make_optimizer(model)), args.steps, args.warmup, world, device) / args.steps
        # the same step captured once in a CUDA graph (rgcn_b200.trainer.GraphedTrainStep) and replayed:
        # the labelled batch is still copied from pinned host memory and the loss read back every step
        from rgcn_b200.trainer import GraphedTrainStep
        data.x_train, data.y_train = x_train_h.to(device), y_train_h.to(device)
        graphed = GraphedTrainStep(model, data, make_optimizer(model, capturable=True), ce_loss, identity)

        graphed.prefetch(x_train_h, y_train_h)

        def graphed_step():
            loss = graphed()                              # consumes the batch prefetched during the previous step
            graphed.prefetch(x_train_h, y_train_h)        # next step's H2D overlaps this step's kernels
            return loss.item()
        e2e_ms = time_steps(graphed_step, args.steps, args.warmup, world, device) / args.steps
        # same step with torch.optim.Adam, i.e. what the unmodified reference Trainer runs on the drop-in layers
        torch_adam_ms = time_steps(make_e2e_step(make_optimizer(model, fused=False)), args.steps, args.warmup, world,
                                   device) / args.steps
    # one run of the reference = 51 epochs (main.py -epochs default) after ONE graph build (edge tensors host ->
    # device + K0) and ONE upload of the [N,63] embedding: the amortised epoch time carries both
    _emb_host = torch.empty(n, EMB).pin_memory()           # (page-locking the host buffer is not part of the copy)
    torch.cuda.synchronize(device)
    t_up = time.perf_counter()
    _emb_up = _emb_host.to(device, non_blocking=True)
    torch.cuda.synchronize(device)
    upload_ms = (time.perf_counter() - t_up) * 1e3
    del _emb_up, _emb_host
    e2e = {'value': e / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': 4,
           'ms_per_step': e2e_ms, 'epoch_ms_51': (setup_ms + upload_ms + 51 * e2e_ms) / 51,
           'epoch_ms_51_what': f'(graph build {setup_ms:.1f} ms + embedding upload {upload_ms:.1f} ms from pinned memory + 51 x step) / 51',
           'h2d_bytes_once': int(2 * e * 8 + e * 8 + n * EMB * 4), 'ms_per_step_eager': eager_ms, 'ms_per_step_eager_torch_adam': torch_adam_ms,
           'what': 'Trainer.train iteration body (modelTrainer.py:61-69) via Emb_Layers on the drop-in RGCNConv, captured '
                   'in a CUDA graph (GraphedTrainStep): pinned H2D of x_train/y_train every step (copy stream, overlapping the previous step), fwd, CE loss, bwd, Adam step '
                   '(engine FusedAdam, same update rule as torch.optim.Adam(lr, weight_decay); incl. the [N,63] '
                   'embedding), loss.item(); ms_per_step_eager = the same step launched eagerly, '
                   '..._torch_adam = eagerly with torch.optim.Adam'}

    # K5 map gather at this graph's size (3 summaries, sum mode): achieved GB/s, reported beside the layer
    map_gather = None
    if not args.no_e2e:
        from rgcn_b200 import map_gather as mg
        gen = torch.Generator().manual_seed(3)
        n_sum = 50_000
        embs = [torch.randn(n_sum, EMB, generator=gen).to(device) for _ in range(3)]
        idxs = [torch.randint(-1, n_sum, (n,), generator=gen, dtype=torch.int32).to(device) for _ in range(3)]
        fbs = [torch.rand(n, EMB, generator=gen).to(device) for _ in range(3)]
        for _ in range(3):
            mg(embs, idxs, fbs, 0)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            mg(embs, idxs, fbs, 0)
        ev1.record()
        torch.cuda.synchronize(device)
        mg_ms = ev0.elapsed_time(ev1) / 10
        mg_bytes = 3 * n * (4 * EMB + 4) + n * 4 * EMB        # S x (row + index) read, one row written
        # what has to cross the HBM pins: the output, the index vectors, the (L2-resident) tables once
        mg_dram = n * 4 * EMB + 3 * n * 4 + 3 * n_sum * EMB * 4
        map_gather = {'ms': mg_ms, 'algorithmic_bytes': mg_bytes, 'gbps': mg_bytes / (mg_ms * 1e-3) / 1e9,
                      'frac_of_peak': mg_bytes / (mg_ms * 1e-3) / 1e9 / peak, 'dram_bytes': mg_dram,
                      'frac_of_peak_dram_bytes': mg_dram / (mg_ms * 1e-3) / 1e9 / peak, 'mode': 'sum', 'summaries': 3,
                      'summary_rows': n_sum}
        del embs, idxs, fbs

    # transfer heads at this graph's size: the MLP head's two contractions on the engine's tcgen05
    # kernel, and the Trainer step of the three experiment models (summation / mlp / attention) beside each other
    heads = None
    if not args.no_e2e:
        heads = bench_heads(n, r, data, x_train_h, y_train_h, device, world, args, peak)

    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cval, cms, cn, ce_, sample = cpu_reference(min(args.steps, 20), min(args.warmup, 5), args.ref_scale, threads)
        cpu = {'value': cval, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample, 'ms_per_step': cms}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'am_shape_full_graph_rgcn_63_16_11_all_grads', 'scale': args.scale, 'nodes': n,
                   'directed_edges': e, 'triples_per_s': value / 2, 'relations': r, 'emb': EMB, 'hidden': HIDDEN,
                   'classes': CLASSES, 'l2_policy': 'inputs larger than L2 (features 420 MB + CSR > 126 MB L2); no flush',
                   'graph_build_ms_once': setup_ms, 'range_nodes': graph.query(_lib.Q_RANGE_NODES),
                   'ms_per_step_passes_serialised': serial_ms, 'variants': variants,
                   'step_algorithmic_bytes': step_bytes, 'step_roofline_frac': step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                   'passes': passes, 'map_gather': map_gather, 'heads': heads},
        'parity': parity,
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', choices=['engine', 'reference'], default='engine')
    ap.add_argument('--workload', choices=['am_63_16_11', 'am10x_h64_bases30_frozenW', 'multi_summary'], default='am_63_16_11',
                    help='am_63_16_11 = BASELINE.json configs[3] (the headline); am10x_... = configs[4]')
    ap.add_argument('--scale', type=float, default=1.0, help='AM-shape scale (1.0 = the BASELINE.json config)')
    ap.add_argument('--ref-scale', type=float, default=1 / 32, help='bounded sample for the CPU reference arm')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='kernel-level timing only (tuning runs; not a bench line)')
    ap.add_argument('--no-check', action='store_true', help='skip the parity leg (reduced-scale check printed in the line)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_engine(args)


if __name__ == '__main__':
    main()
