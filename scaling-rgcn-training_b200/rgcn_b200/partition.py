"""Destination-partitioned full-graph R-GCN over the GPUs of one node (one process per GPU,
``torch.distributed``; BASELINE.json configs 4/5).

The reference is single-process (model/modelTrainer.py:16); this is the engine's multi-GPU mode
for the same layer arithmetic.  Owner-computes in BOTH directions, so the only data-path
collectives are all-gathers of the gather-side rows and one all-reduce of parameter gradients:

  rank p owns nodes [p*chunk, (p+1)*chunk)  (chunk = ceil(N / P): equal ranges make the
  all-gathered tensor directly indexable by global node id)
  forward   x_all  = all_gather(x_owned)        ->  out_owned  (edges whose dst is owned)
  backward  g_all  = all_gather(gout_owned)     ->  gx_owned   (edges whose src is owned)
            dW, droot, dbias partial over owned dst  ->  all_reduce(sum)
The embedding rows, activations and their gradients stay sharded; weights are replicated.
"""
from __future__ import annotations

import json
import os
import statistics
import time
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor, nn


def plan_ranges(num_nodes: int, world: int) -> Tuple[int, List[Tuple[int, int]]]:
    """Equal contiguous node ranges; trailing ranks may own fewer (never zero) nodes."""
    if world < 1 or num_nodes < world:
        raise ValueError('plan_ranges: need 1 <= world <= num_nodes')
    chunk = -(-num_nodes // world)
    ranges = [(min(p * chunk, num_nodes), min((p + 1) * chunk, num_nodes)) for p in range(world)]
    if any(lo >= hi for lo, hi in ranges):
        raise ValueError(f'plan_ranges: {world} ranks leave an empty range for {num_nodes} nodes')
    return chunk, ranges


def plan_ranges_balanced(node_weight, world: int) -> List[Tuple[int, int]]:
    """Contiguous node ranges with (nearly) equal total weight — e.g. out-degree + 1, the work a node brings to
    the rank that owns it in the source-partitioned layers (balance by edges, not by
    nodes: the hubs make equal node ranges ~9 % uneven at 8 ranks).  Range ends are rounded to multiples of 4
    nodes (16-byte row blocks) and every range keeps at least one node."""
    import numpy as np
    w = np.asarray(node_weight, dtype=np.float64)
    n = int(w.size)
    if world < 1 or n < world:
        raise ValueError('plan_ranges_balanced: need 1 <= world <= num_nodes')
    csum = np.cumsum(w)
    cuts = [0]
    for p in range(1, world):
        c = int(np.searchsorted(csum, csum[-1] * p / world, side='left')) + 1
        c = (c + 3) // 4 * 4
        c = max(c, cuts[-1] + 1)
        c = min(c, n - (world - p))
        cuts.append(c)
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


class RowComm:
    """The two collectives of the partitioned layer, over any torch.distributed backend
    (nccl on the GPUs; gloo in the CPU tests of this host logic)."""

    def __init__(self, num_nodes: int, group=None, ranges=None) -> None:
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.num_nodes = num_nodes
        self.chunk, self.ranges = plan_ranges(num_nodes, self.world)
        if ranges is not None:
            if type(self) is RowComm and list(ranges) != self.ranges:
                raise ValueError('RowComm (NCCL collectives) needs equal node ranges; NvlComm takes any')
            self.ranges = [(int(a), int(b)) for a, b in ranges]
        self.lo, self.hi = self.ranges[self.rank]
        self.bytes_gathered = 0
        self.bytes_reduced = 0
        self.events = None        # list of (start, stop) CUDA-event pairs when timing is on (bench)

    def all_gather_rows_async(self, t: Tensor):
        """Start the all-gather on the communicator's stream; returns (out, work). The caller runs
        independent kernels (dL/dW needs only the owned gout rows) and calls work.wait() before the
        first consumer of `out` — the transfer then overlaps that compute."""
        n_own, f = t.shape
        if n_own != self.chunk:
            pad = t.new_zeros((self.chunk, f))
            pad[:n_own] = t
            t = pad
        out = t.new_empty((self.world * self.chunk, f))
        work = dist.all_gather_into_tensor(out, t.contiguous(), group=self.group, async_op=True)
        self.bytes_gathered += out.numel() * out.element_size()
        return out, work

    def all_gather_rows(self, t: Tensor, key=None) -> Tensor:
        """[n_owned, F] on every rank -> [world*chunk, F]; row i is node i for i < num_nodes."""
        n_own, f = t.shape
        if n_own != self.hi - self.lo:
            raise ValueError('all_gather_rows: wrong number of owned rows')
        if n_own != self.chunk:        # last rank(s): pad to the common chunk
            pad = t.new_zeros((self.chunk, f))
            pad[:n_own] = t
            t = pad
        out = t.new_empty((self.world * self.chunk, f))
        ev = self._tic(t)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        self._toc(ev)
        self.bytes_gathered += out.numel() * out.element_size()
        return out

    def partial_buffer(self, cols: int, key=None, device=None) -> Tensor:
        """[world*chunk, cols] buffer a source-partitioned layer accumulates its partial output into (rows >=
        num_nodes are zero); reduce_scatter_rows takes it back."""
        t = torch.empty((self.world * self.chunk, cols), dtype=torch.float32, device=device)
        if t.size(0) > self.num_nodes:
            t[self.num_nodes:].zero_()
        return t

    def reduce_scatter_rows(self, t: Tensor, key=None) -> Tensor:
        """[world*chunk, F] partial sums on every rank -> [chunk, F]: this rank's rows of the total."""
        if t.size(0) != self.world * self.chunk:
            raise ValueError('reduce_scatter_rows: expected world * chunk rows')
        out = t.new_empty((self.chunk, t.size(1)))
        ev = self._tic(t)
        dist.reduce_scatter_tensor(out, t.contiguous(), op=dist.ReduceOp.SUM, group=self.group)
        self._toc(ev)
        self.bytes_reduced += t.numel() * t.element_size()
        return out

    def _tic(self, t: Tensor):
        if self.events is None or not t.is_cuda:
            return None
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        return ev

    def _toc(self, ev) -> None:
        if ev is not None:
            ev[1].record()
            self.events.append(ev)

    def comm_ms(self) -> float:
        """Sum of the timed synchronous collectives (call after a device synchronize)."""
        return sum(a.elapsed_time(b) for a, b in (self.events or []))

    def all_reduce_sum_(self, tensors: List[Tensor]) -> None:
        """One flat all-reduce for the (small) parameter gradients."""
        tensors = [t for t in tensors if t is not None]
        if not tensors or self.world == 1:
            return
        flat = torch.cat([t.reshape(-1) for t in tensors])
        ev = self._tic(flat)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        self._toc(ev)
        off = 0
        for t in tensors:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()


class NvlComm(RowComm):
    """The same exchanges as the engine's own kernels over NVLink / NVSwitch peer memory (csrc/nvl_comm.cu)
    instead of NCCL calls.  Buffers are symmetric allocations (torch.distributed._symmetric_memory: same size on
    every rank, peer-mapped, with an NVSwitch multicast address):

      all-gather      one multimem.st per 16 bytes of the rank's own rows -> the switch replicates them into every
                      rank's copy; the ReLU backward of the inter-layer activation can ride on the same kernel
      reduce-scatter  the layer kernels accumulate the rank's partial output straight into its copy
                      (partial_buffer); after a barrier each rank reads ITS rows with multimem.ld_reduce.add.f32,
                      the switch summing the P copies in flight
      all-reduce      the same load-reduce over the flat parameter-gradient buffer

    Ordering: every exchange is write -> device barrier over all ranks -> read (the barrier is the signal-pad
    kernel of the symmetric-memory handle; everything is stream-ordered and CUDA-graph capturable).  In the
    training loop one barrier per exchange is enough — a buffer is rewritten only after a LATER exchange's
    barrier of the same step, which no rank passes before every rank has finished reading (`steady_state=True`,
    what PartitionedRGCN's step satisfies: distinct buffers per call site, >= 1 other exchange between two uses
    of a buffer).  `steady_state=False` adds a second barrier on the other side of every exchange.
    Source-partitioned ("push") graphs only: nothing the layers save for backward may alias these buffers."""

    def __init__(self, num_nodes: int, group=None, steady_state: bool = False, ranges=None) -> None:
        super().__init__(num_nodes, group, ranges)
        self.rows = (num_nodes + 3) // 4 * 4      # rows of every exchange buffer: ranges need not be equal here
        import ctypes as C
        import torch.distributed._symmetric_memory as sm
        from . import _lib
        self._C, self._sm, self._lib = C, sm, _lib.load()
        self._check = _lib.check
        self.steady_state = steady_state
        self.device = torch.device('cuda', torch.cuda.current_device())
        self._bufs = {}
        self._grp = self.group if self.group is not None else dist.group.WORLD
        self._ctl = sm.empty((1024,), dtype=torch.float32, device=self.device)
        self._ctl_h = sm.rendezvous(self._ctl, self._grp)
        try:
            self.multicast = int(self._ctl_h.multicast_ptr) != 0
        except Exception:
            self.multicast = False
        # multicast (multimem.*) or plain peer pointers (RGCN_B200_NVL_MODE: 1 multicast, 2 peers).  Measured
        # (tools/probe_nvl.py, 16-wide rows of the AM-shape graph, ms per exchange incl. the barrier):
        #   ranks   all-gather mc / peers    reduce-scatter mc / peers
        #     2        0.174 / 0.095            0.177 / 0.102
        #     4        0.175 / 0.147            0.172 / 0.161
        #     8        0.169 / 0.163            0.165 / 0.183
        # a multicast exchange costs the same at every rank count (every copy of the buffer, the local one
        # included, is written through the switch: 107 MB in 0.17 ms = 630 GB/s of link ingress); peer pointers
        # keep the local share off the links, which wins up to 4 ranks
        mode = os.environ.get('RGCN_B200_NVL_MODE')
        self.mode = int(mode) if mode else (1 if self.multicast and self.world > 4 else 2)
        _lib.set_option(_lib.OPT_NVL_MODE, self.mode)
        self.handoff = None          # (data_ptr of an exchanged tensor, its gathered rows): see all_gather_rows
        self.sparse = None           # set_touched(): (peer_mask [n_own] int32, number of remote rows this rank reaches)
        self.touched_frac = None

    def set_touched(self, touched: Tensor, min_saving: float = 0.15) -> bool:
        """Plan the sparse exchanges from the graph (collective: every rank calls it).  touched [num_nodes] bool = the
        rows this rank's passes reach (RGCNGraph.touched): its partial output is zero elsewhere and it gathers gout
        from nowhere else.  The reduce-scatter then reads an owned row only from the ranks that reach it and the
        all-gather stores an owned row only into those ranks' copies — both over peer pointers.  Kept only when the remote rows drop
        by at least min_saving on average (RGCN_B200_NVL_SPARSE=0 / 1 forces the choice); returns whether it is on."""
        plan = sparse_plan(touched, self.ranges, self.rank, self.group)
        self.touched_frac = plan['touched_frac_mean']
        env = os.environ.get('RGCN_B200_NVL_SPARSE')
        on = plan['remote_saving_mean'] >= min_saving if env is None else env != '0'
        if on:
            self.sparse = (plan['peer_mask'].to(self.device), int(plan['row_list'].numel()))
        else:
            self.sparse = None
        return on

    def barrier(self) -> None:
        self._ctl_h.barrier(0)

    def side_stream(self):
        """Stream for compute that may run under an exchange (dL/dW of the upper layer while its gx is stored)."""
        if getattr(self, '_side', None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def _symm(self, key, rows: int, cols: int):
        ent = self._bufs.get(key)
        if ent is None or tuple(ent[0].shape) != (rows, cols):
            t = self._sm.empty((rows, cols), dtype=torch.float32, device=self.device)
            h = self._sm.rendezvous(t, self._grp)          # collective: every rank asks for the same keys in order
            peers = (self._C.c_void_p * self.world)(*[int(p) for p in h.buffer_ptrs])
            mc = int(h.multicast_ptr) if self.multicast else 0
            t.zero_()
            ent = (t, h, mc, peers)
            self._bufs[key] = ent
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
        return ent

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def all_gather_rows(self, t: Tensor, key=None, relu_pre: Optional[Tensor] = None) -> Tensor:
        """[n_owned, F] -> [world*chunk, F] (a view of the symmetric buffer, rows 16-byte addressable: the row
        stride is F rounded up to 4).  relu_pre: zero t where relu_pre <= 0 first, in the same kernel."""
        n_own, f = t.shape
        if n_own != self.hi - self.lo or t.stride(1) != 1:
            raise ValueError('all_gather_rows: wrong number of owned rows / non-unit column stride')
        ld = (f + 3) // 4 * 4
        buf, _, mc, peers = self._symm(('ag', key if key is not None else f, ld), self.rows, ld)
        if not self.steady_state:
            self.barrier()
        if self.sparse is not None:
            # every owned row goes only into the copies of the ranks that reach it (ReLU mask fused)
            rc = self._lib.rgcn_nvl_store_rows_sparse(t.data_ptr(), t.stride(0), f,
                                                      relu_pre.data_ptr() if relu_pre is not None else None,
                                                      relu_pre.stride(0) if relu_pre is not None else 0,
                                                      peers, self.world, self.sparse[0].data_ptr(), ld, self.lo, n_own,
                                                      self._stream())
            self._check(rc, 'rgcn_nvl_store_rows_sparse')
            self.barrier()
            self.bytes_gathered += (self.sparse[1] + n_own) * ld * 4
            return buf[:, :f] if ld != f else buf
        rc = self._lib.rgcn_nvl_store_rows(t.data_ptr(), t.stride(0), f,
                                           relu_pre.data_ptr() if relu_pre is not None else None,
                                           relu_pre.stride(0) if relu_pre is not None else 0,
                                           mc if mc else None, peers, self.world, ld, self.lo, n_own, self._stream())
        self._check(rc, 'rgcn_nvl_store_rows')
        self.barrier()
        self.bytes_gathered += buf.numel() * 4
        return buf[:, :f] if ld != f else buf

    all_gather_rows_async = None          # (no NCCL work handle: the exchange is stream-ordered)

    def partial_buffer(self, cols: int, key=None, device=None) -> Tensor:
        if cols % 4:
            raise ValueError('partial_buffer: width must be a multiple of 4')
        return self._symm(('rs', key if key is not None else cols, cols), self.rows, cols)[0]

    def reduce_scatter_rows(self, t: Tensor, key=None) -> Tensor:
        cols = t.size(1)
        buf, _, mc, peers = self._symm(('rs', key if key is not None else cols, cols), self.rows, cols)
        if t.data_ptr() != buf.data_ptr():
            raise ValueError('reduce_scatter_rows: pass the tensor partial_buffer() returned')
        self.barrier()
        n_own = self.hi - self.lo
        out = torch.empty((n_own, cols), dtype=torch.float32, device=self.device)
        if self.sparse is not None:
            rc = self._lib.rgcn_nvl_reduce_rows_sparse(peers, self.world, self.sparse[0].data_ptr(), cols, self.lo, n_own,
                                                       out.data_ptr(), cols, cols, self._stream())
            self._check(rc, 'rgcn_nvl_reduce_rows_sparse')
            self.bytes_reduced += int(self.touched_frac * buf.numel() * 4)
        else:
            rc = self._lib.rgcn_nvl_reduce_rows(mc if mc else None, peers, self.world, cols, self.lo, n_own, out.data_ptr(),
                                                cols, cols, self._stream())
            self._check(rc, 'rgcn_nvl_reduce_rows')
            self.bytes_reduced += buf.numel() * 4
        if not self.steady_state:
            self.barrier()
        return out

    def all_reduce_sum_(self, tensors: List[Tensor]) -> None:
        tensors = [t for t in tensors if t is not None]
        if not tensors or self.world == 1:
            return
        n = sum(t.numel() for t in tensors)
        width = 256
        rows = (n + width - 1) // width
        buf, _, mc, peers = self._symm(('ar', rows), rows, width)
        flat = buf.view(-1)
        off = 0
        for t in tensors:
            flat[off:off + t.numel()].copy_(t.reshape(-1))
            off += t.numel()
        self.barrier()
        out = torch.empty((rows, width), dtype=torch.float32, device=self.device)
        rc = self._lib.rgcn_nvl_reduce_rows(mc if mc else None, peers, self.world, width, 0, rows, out.data_ptr(), width,
                                            width, self._stream())
        self._check(rc, 'rgcn_nvl_reduce_rows')
        if not self.steady_state:
            self.barrier()
        res = out.view(-1)
        off = 0
        for t in tensors:
            t.copy_(res[off:off + t.numel()].view_as(t))
            off += t.numel()


def sparse_plan(touched: Tensor, ranges, rank: int, group=None) -> dict:
    """Host logic of NvlComm.set_touched (any backend; gloo in the CPU tests).  touched [N] bool per rank ->
    peer_mask [n_own] int32 (bit p: rank p reaches owned row i), row_list int32 (the remote rows this rank reaches,
    ascending = grouped by owner), offsets [world + 1] into row_list by owner, and the traffic it saves."""
    world = dist.get_world_size(group)
    n = touched.numel()
    lo, hi = ranges[rank]
    mine = touched.to(torch.uint8).contiguous()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    mask = torch.zeros(hi - lo, dtype=torch.int32, device=touched.device)
    for p, t in enumerate(parts):
        mask |= t[lo:hi].to(torch.int32) << p
    mask |= 1 << rank
    remote = touched.clone()
    remote[lo:hi] = False
    rows = torch.nonzero(remote).flatten().to(torch.int32)
    bounds = torch.tensor([a for a, _ in ranges] + [ranges[-1][1]], dtype=torch.int64)
    offsets = torch.searchsorted(rows.to(torch.int64).cpu(), bounds).tolist()
    offsets[0], offsets[-1] = 0, rows.numel()
    fracs = torch.stack([t.to(torch.float32).mean() for t in parts]).cpu()
    dense_remote = torch.tensor([1.0 - (b - a) / max(n, 1) for a, b in ranges])
    sparse_remote = torch.clamp(fracs - (1.0 - dense_remote), min=0.0)
    saving = 1.0 - float((sparse_remote / torch.clamp(dense_remote, min=1e-9)).mean())
    return {'peer_mask': mask, 'row_list': rows, 'offsets': [int(v) for v in offsets],
            'touched_frac_mean': float(fracs.mean()), 'remote_saving_mean': saving}


def comm_backend(backend: Optional[str] = None) -> str:
    backend = backend or os.environ.get('RGCN_B200_COMM', 'nvl' if torch.cuda.is_available() else 'nccl')
    if backend not in ('nvl', 'nccl'):
        raise ValueError('RGCN_B200_COMM must be nvl or nccl')
    return backend


def make_comm(num_nodes: int, backend: Optional[str] = None, steady_state: bool = False, ranges=None) -> RowComm:
    """RGCN_B200_COMM=nvl (default on CUDA: engine kernels over NVLink multicast) | nccl (torch.distributed calls)."""
    if comm_backend(backend) == 'nvl':
        return NvlComm(num_nodes, steady_state=steady_state, ranges=ranges)
    return RowComm(num_nodes)


def balanced_ranges_for(edge_index: Tensor, num_nodes: int, world: int) -> List[Tuple[int, int]]:
    """Node ranges balanced by the work of the source-partitioned layers: out-degree (every pass walks the edges
    whose src is owned) + 1 for the node's own row."""
    import numpy as np
    deg = np.bincount(edge_index[0].numpy(), minlength=num_nodes).astype(np.float64)
    return plan_ranges_balanced(deg + 1.0, world)


class PartitionedRGCN(nn.Module):
    """2-layer R-GCN (Emb_Layers arithmetic, reference model/layers.py:20-25) on a partitioned
    graph: sharded embedding rows, replicated layer weights, fused inter-layer ReLU.  The graph decides the
    exchange: destination-partitioned ("pull": all-gather of the input rows forward and of gout backward) or
    source-partitioned ("push": reduce-scatter of partial outputs forward, all-gather of gout backward — the
    63-wide layer-1 input never crosses the links)."""

    def __init__(self, graph, comm: RowComm, num_relations: int, hidden_l: int, num_labels: int, emb_dim: int,
                 seed: int = 0) -> None:
        super().__init__()
        from .conv import RGCNConv
        self.graph, self.comm = graph, comm
        if getattr(graph, 'touched', None) is not None and hasattr(comm, 'set_touched'):
            comm.set_touched(graph.touched)
        gen = torch.Generator().manual_seed(seed)          # same replicated weights on every rank
        full = torch.randn(comm.num_nodes, emb_dim, generator=gen)
        self.embedding = nn.Parameter(full[comm.lo:comm.hi].clone())
        torch.manual_seed(seed)
        self.rgcn1 = RGCNConv(emb_dim, hidden_l, num_relations)
        self.rgcn2 = RGCNConv(hidden_l, num_labels, num_relations)
        for conv in (self.rgcn1, self.rgcn2):
            nn.init.kaiming_uniform_(conv.weight, mode='fan_in')

    def forward(self) -> Tensor:
        from .conv import rgcn_layer
        c1, c2 = self.rgcn1, self.rgcn2
        h = rgcn_layer(self.embedding, c1.weight, c1.root, c1.bias, self.graph, comm=self.comm, comm_key=1)
        return rgcn_layer(h, c2.weight, c2.root, c2.bias, self.graph, relu_in=True, comm=self.comm, comm_key=2)


def partition_mode() -> str:
    m = os.environ.get('RGCN_B200_PARTITION', 'push')
    if m not in ('push', 'pull'):
        raise ValueError('RGCN_B200_PARTITION must be push or pull')
    return m


def partitioned_parity_check(rank: int, world: int, device, mode: str, hidden: int, classes: int, emb: int,
                             scale: float = 1 / 16, backend: Optional[str] = None) -> Optional[dict]:
    """bench.py --gpus N, N > 1: the partitioned fwd+bwd on the AM-shape graph at `scale` against the SAME step on
    one GPU (rank 0, unpartitioned graph): max-norm relative errors of the output, the embedding gradient and all
    parameter gradients.  Proves on the measuring box that the N-rank path computes what the 1-rank path does."""
    from .conv import rgcn_layer
    from .graph import RGCNGraph
    from .synthetic import am_shape
    ei, et, n, r = am_shape(scale=scale)
    nvl = mode == 'push' and comm_backend(backend) == 'nvl'
    comm = make_comm(n, backend, steady_state=False, ranges=balanced_ranges_for(ei, n, world) if nvl else None) \
        if mode == 'push' else RowComm(n)
    ei_d, et_d = ei.to(device), et.to(device)
    graph = RGCNGraph(ei_d, et_d, n, r, own_range=(comm.lo, comm.hi), push=(mode == 'push'))
    model = PartitionedRGCN(graph, comm, r, hidden, classes, emb, seed=7).to(device)
    gen = torch.Generator().manual_seed(8)
    gout_full = torch.randn(n, classes, generator=gen)
    out = model()
    out.backward(gout_full[comm.lo:comm.hi].to(device))
    def gather_all(t):          # (the comparison itself gathers over NCCL; ranges may be unequal)
        parts = [torch.empty((b - a, t.size(1)), dtype=t.dtype, device=t.device) for a, b in comm.ranges]
        dist.all_gather(parts, t.contiguous())
        return torch.cat(parts)
    out_all = gather_all(out.detach())
    gx_all = gather_all(model.embedding.grad)
    res = None
    if rank == 0:
        g1 = RGCNGraph(ei_d, et_d, n, r)
        full = torch.randn(n, emb, generator=torch.Generator().manual_seed(7)).to(device).requires_grad_()
        ps = [p.detach().clone().requires_grad_() for p in (model.rgcn1.weight, model.rgcn1.root, model.rgcn1.bias,
                                                            model.rgcn2.weight, model.rgcn2.root, model.rgcn2.bias)]
        h = rgcn_layer(full, ps[0], ps[1], ps[2], g1)
        ref = rgcn_layer(h, ps[3], ps[4], ps[5], g1, relu_in=True)
        ref.backward(gout_full.to(device))

        def rel(a, b):
            return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
        got = [model.rgcn1.weight.grad, model.rgcn1.root.grad, model.rgcn1.bias.grad, model.rgcn2.weight.grad,
               model.rgcn2.root.grad, model.rgcn2.bias.grad]
        errs = {'out': rel(out_all, ref.detach()), 'gx': rel(gx_all, full.grad)}
        for name, a, b in zip(('gW1', 'groot1', 'gbias1', 'gW2', 'groot2', 'gbias2'), got, ps):
            errs[name] = rel(a, b.grad)
        res = {'against': f'the same step on 1 GPU (unpartitioned), AM-shape at scale {scale:g} (N={n}, E={et.numel()})',
               'max_rel_err': errs, 'worst': max(errs.values()), 'argmax_equal': bool(torch.equal(out_all.argmax(1), ref.argmax(1)))}
        del g1
    del graph, model
    torch.cuda.synchronize(device)
    dist.barrier()
    return res


def run_partitioned_bench(args, rank: int, world: int, device, metric: str, unit: str) -> None:
    """bench.py --gpus N (N > 1): strong scaling of the AM-shape fwd+bwd step."""
    from . import _lib
    from .graph import RGCNGraph
    from .synthetic import am_shape
    import bench as B                                        # helpers of the single-GPU arm
    mode = partition_mode()
    parity = None
    if not getattr(args, 'no_check', False):
        parity = partitioned_parity_check(rank, world, device, mode, B.HIDDEN, B.CLASSES, B.EMB)
    ei, et, n, r = am_shape(scale=args.scale)                # every rank derives the same graph (seed 0)
    e = et.numel()
    nvl = mode == 'push' and comm_backend() == 'nvl'
    comm = make_comm(n, steady_state=True, ranges=balanced_ranges_for(ei, n, world) if nvl else None) \
        if mode == 'push' else RowComm(n)
    backend = 'nvl' if isinstance(comm, NvlComm) else 'nccl'
    t0 = time.perf_counter()
    ei_d, et_d = ei.to(device), et.to(device)
    graph = RGCNGraph(ei_d, et_d, n, r, own_range=(comm.lo, comm.hi), push=(mode == 'push'))
    del ei_d, et_d
    torch.cuda.synchronize(device)
    setup_ms = (time.perf_counter() - t0) * 1e3
    model = PartitionedRGCN(graph, comm, r, B.HIDDEN, B.CLASSES, B.EMB).to(device)
    gen = torch.Generator().manual_seed(1)
    gout = torch.randn(n, B.CLASSES, generator=gen)[comm.lo:comm.hi].to(device)
    params = list(model.parameters())

    def step():
        for p in params:
            p.grad = None
        model().backward(gout)

    # the step is captured once in a CUDA graph (collectives included) and replayed: at 8 GPUs a rank's
    # kernels are ~50 us each and eager launches would show; RGCN_B200_PART_GRAPH=0 launches eagerly
    use_graph = os.environ.get('RGCN_B200_PART_GRAPH', '1') != '0'
    sampler = B.ClockSampler(device.index)
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(device)
    comm.bytes_gathered = comm.bytes_reduced = 0
    step()
    gathered, reduced = comm.bytes_gathered, comm.bytes_reduced
    launches0 = _lib.launch_count()
    step()
    launches_per_step = _lib.launch_count() - launches0
    # synchronous-collective time, measured on an eager step (events cannot be recorded inside a capture)
    comm.events = []
    for _ in range(3):
        step()
    torch.cuda.synchronize(device)
    sync_ms = comm.comm_ms() / 3
    comm.events = None
    # per-pass device times of one eager step with the concurrent passes serialised (rank 0's view)
    _lib.set_option(_lib.OPT_OVERLAP, 0)
    _lib.profile_enable(True)
    _lib.profile_collect()
    step()
    torch.cuda.synchronize(device)
    prof = {}
    for name, dims, pms in _lib.profile_collect():
        k = f'{name}_{dims[0]}x{dims[1]}'
        prof[k] = prof.get(k, 0.0) + pms
    _lib.profile_enable(False)
    _lib.set_option(_lib.OPT_OVERLAP, 1)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        step()
    ev[1].record()
    torch.cuda.synchronize(device)
    eager_ms = ev[0].elapsed_time(ev[1]) / 5
    cuda_graph = None
    run = step
    if use_graph:
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        dist.barrier()
        cuda_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cuda_graph):
            step()
        run = cuda_graph.replay
        torch.cuda.synchronize(device)
    if rank == 0:
        sampler.start()
    total_ms = B.time_steps(run, args.steps, max(args.warmup, 3), world, device)
    clocks = sampler.stop() if rank == 0 else None
    ms = total_ms / args.steps
    value = e / (ms * 1e-3)
    # end to end: the Trainer.train iteration body on the partitioned model — this rank's share of the
    # labelled batch copied from pinned host memory every step, forward, CE loss (global mean: local
    # sum / total count), backward (weight gradients all-reduced inside the layer), Adam on the local
    # embedding shard and the replicated weights, loss all-reduced and read back
    e2e = None
    e2e_graph = None
    if not getattr(args, 'no_e2e', False):
        from .synthetic import labelled_split
        from .trainer import make_optimizer
        x_all, y_all = labelled_split(n, B.CLASSES)
        mine = (x_all >= comm.lo) & (x_all < comm.hi)
        x_h = (x_all[mine] - comm.lo).contiguous().pin_memory()
        y_h = y_all[mine].contiguous().pin_memory()
        m_total = float(x_all.numel())
        opt = make_optimizer(model, capturable=use_graph)
        xs, ys = torch.empty_like(x_h, device=device), torch.empty_like(y_h, device=device)
        tot = torch.zeros((), device=device)

        def body():
            opt.zero_grad(set_to_none=True)
            out = model()[xs]
            picked = out.gather(1, ys.to(torch.float32).argmax(-1, keepdim=True)).squeeze(1)
            local = (torch.logsumexp(out, dim=1) - picked).sum() / m_total
            local.backward()
            opt.step()
            tot.copy_(local.detach())
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)

        if use_graph:
            snapshot = [p.detach().clone() for p in params]
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                xs.copy_(x_h, non_blocking=True)
                ys.copy_(y_h, non_blocking=True)
                body()
                with torch.no_grad():
                    for p, s_ in zip(params, snapshot):
                        p.copy_(s_)
                opt.reset_state()
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            dist.barrier()
            e2e_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(e2e_graph):
                body()

        def e2e_step():
            xs.copy_(x_h, non_blocking=True)
            ys.copy_(y_h, non_blocking=True)
            if e2e_graph is not None:
                e2e_graph.replay()
            else:
                body()
            return tot.item()

        e2e_ms = B.time_steps(e2e_step, args.steps, max(args.warmup, 3), world, device) / args.steps
        h2d = torch.tensor([x_h.numel() * x_h.element_size() + y_h.numel() * y_h.element_size()], device=device,
                           dtype=torch.float64)
        dist.all_reduce(h2d, op=dist.ReduceOp.SUM)
        e2e = {'value': e / (e2e_ms * 1e-3), 'unit': unit, 'h2d_bytes_per_step': int(h2d.item()),
               'd2h_bytes_per_step': 4 * world, 'ms_per_step': e2e_ms,
               'what': 'Trainer.train iteration body on the partitioned model (' + ('one CUDA-graph replay per step' if use_graph else 'eager') +
                       '): per-rank share of x_train/y_train '
                       'from pinned host memory, fwd, CE loss (global mean), bwd, FusedAdam on the embedding shard and '
                       'the replicated weights, all-reduced loss .item(); bytes summed over ranks'}
    own_edges = torch.tensor([graph.query(_lib.Q_NUM_ENTRIES0, _lib.BRC_BWD) - graph.num_owned], device=device,
                             dtype=torch.float64)
    mx = own_edges.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    if rank == 0:
        if mode == 'push':
            coll = ('per layer: reduce_scatter(partial out) fwd, all_gather(gout) bwd; one all_reduce of dW/droot/dbias; '
                    'the 63-wide layer-1 input is never exchanged; ' +
                    (('engine kernels over NVLink peer memory, SPARSE: only the rows a rank\'s own edges reach cross the '
                      'links (reduce-scatter reads an owned row from the ranks that reach it; all-gather stores it only into '
                      'their copies), one device barrier each'
                      if getattr(comm, 'sparse', None) is not None else
                      'engine kernels over NVSwitch multicast (multimem.st / multimem.ld_reduce), one device barrier each'
                      if getattr(comm, 'mode', 2) != 2 else
                      'engine kernels over NVLink peer pointers, one device barrier each')
                     if backend == 'nvl' else 'NCCL calls'))
            limiting = 'all_gather(gout of layer 1, 16 wide) + reduce_scatter(partial h1, 16 wide)'
        else:
            coll = 'per layer: all_gather(x) fwd, all_gather(gout) bwd; one all_reduce of dW/droot/dbias'
            limiting = 'all_gather(x0 mirror, 64 wide)'
        line = {
            'metric': metric, 'value': value, 'unit': unit, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': 'am_shape_full_graph_rgcn_63_16_11_all_grads', 'scale': args.scale, 'nodes': n,
                       'directed_edges': e, 'relations': r,
                       'partition': f'{"source" if mode == "push" else "dst"}-partitioned x{world}, ' +
                                    ('node ranges balanced by out-degree' if backend == 'nvl' else 'equal node ranges'),
                       'collectives': coll, 'comm_backend': backend, 'limiting_collective': limiting,
                       'exchange_sparse': getattr(comm, 'sparse', None) is not None,
                       'rows_reached_per_rank_frac_mean': getattr(comm, 'touched_frac', None),
                       'all_gather_bytes_per_step_per_rank': gathered,
                       'reduce_scatter_bytes_per_step_per_rank': reduced,
                       'max_rank_edges': int(mx.item()), 'mean_rank_edges': e / world,
                       'sync_collectives_ms_per_step_rank0_eager': sync_ms,
                       'step_launch': 'cuda graph replay (collectives captured)' if use_graph else 'eager',
                       'ms_per_step_eager': eager_ms, 'passes_rank0_serialised_ms': prof,
                       'graph_build_ms_once': setup_ms,
                       'l2_policy': 'inputs larger than L2; no flush'},
            'parity': parity,
            'clocks': clocks, 'gpu_launches': int(launches_per_step * args.steps),
            'e2e': e2e, 'roofline': None, 'cpu_baseline': None,
        }
        print(json.dumps(line), flush=True)
    # graphs hold references to the communicator's work: drop them before the process group goes away
    del cuda_graph, e2e_graph, run
    torch.cuda.synchronize(device)
    dist.barrier()
    dist.destroy_process_group()
