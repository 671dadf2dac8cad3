"""RGCNConv — the drop-in for ``torch_geometric.nn.RGCNConv`` (PyG 2.3.1) at the reference's
call sites: constructed at /root/reference/model/layers.py:15-16,54-55,98-99, its ``weight``
re-initialised in place at :17-18, called at :21,23,62,64,108,110, its parameters cloned at
model/modelTrainer.py:28-35 and replaced by attribute assignment at model/layers.py:34-46.

Same constructor, same parameter names/shapes, same forward signature.  The arithmetic runs
only in librgcn_b200.so (hand-written sm_100a kernels); CPU tensors are rejected — there is no
fallback path.  Parameters are read at call time, never cached, so in-place init and
re-assignment keep working.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
from torch import Tensor, nn

from . import _lib
from .graph import RGCNGraph, cached_graph


def _glorot_(t: Tensor) -> None:   # PyG inits.glorot
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)


_PAD_WIDE = os.environ.get('RGCN_B200_PAD', '1') != '0'
_KEEP_CHUNKS = os.environ.get('RGCN_B200_KEEP_CHUNKS', '1') != '0'


def _ptr(t: Optional[Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _RGCNLayerFn(torch.autograd.Function):
    """One R-GCN layer. forward -> rgcn_layer_fwd, backward -> rgcn_layer_bwd."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, root: Optional[Tensor], bias: Optional[Tensor],
                graph: RGCNGraph, flags: int, comm=None, comm_key=None) -> Tensor:
        # comm (partitioned graphs only): object with all_gather_rows(t) -> [N, F] and
        # all_reduce_sum_(list of tensors); x then holds the OWNED rows and is all-gathered here.
        lib = _lib.load()
        if not x.is_cuda:
            raise _lib.EngineError('RGCNConv: input is not a CUDA tensor; the B200 engine has no CPU path')
        if x.dtype != torch.float32 or weight.dtype != torch.float32:
            raise TypeError('RGCNConv: fp32 features and parameters expected')
        fin = x.size(1) if x.dim() == 2 else -1

        def pad_rows(t: Tensor) -> Tensor:
            # odd-width wide rows (the reference's emb = 63): gather from a zero-padded 16-byte
            # addressable mirror (one streaming copy) so the kernels can use 128-bit row loads;
            # the mirror is what backward re-gathers, the caller's tensor is not kept
            if not (_PAD_WIDE and fin > 32 and (t.stride(0) % 4 != 0 or t.data_ptr() % 16 != 0)):
                return t
            ldp = (fin + 3) // 4 * 4
            tp = torch.empty((t.size(0), ldp), dtype=torch.float32, device=t.device)
            with torch.cuda.device(t.device):
                _lib.check(lib.rgcn_pad_rows(t.data_ptr(), t.stride(0), fin, tp.data_ptr(), ldp, t.size(0),
                                             _stream(t.device)), 'rgcn_pad_rows')
            return tp

        push = comm is not None and graph.push
        if comm is not None:
            if x.dim() != 2 or x.size(0) != graph.num_owned:
                raise ValueError(f'RGCNConv (partitioned): x must hold the {graph.num_owned} owned rows')
            x = x.contiguous()
            if not push:                                     # pull: every rank needs the rows of all sources
                x = comm.all_gather_rows(pad_rows(x))        # (the owned shard made 16-byte addressable first)
        if not push and (x.dim() != 2 or x.size(0) < graph.num_nodes or (comm is None and x.size(0) != graph.num_nodes)):
            raise ValueError(f'RGCNConv: x must be [num_nodes={graph.num_nodes}, in_channels]')
        x = x if x.stride(1) == 1 or x.size(1) == 1 else x.contiguous()
        weight = weight.contiguous()
        root_c = root.contiguous() if root is not None else None
        bias_c = bias.contiguous() if bias is not None else None
        n = x.size(0)
        r, fin_w, fout = weight.shape
        if fin_w != fin or r != graph.num_relations:
            raise ValueError('RGCNConv: weight shape does not match input / num_relations')
        # unpartitioned: the engine writes the mirror itself (fused with the root pass where it can)
        mirror = None
        if (comm is None or push) and _PAD_WIDE and 32 < fin <= 64 and (x.stride(0) % 4 != 0 or x.data_ptr() % 16 != 0):
            mirror = torch.empty((x.size(0), (fin + 3) // 4 * 4), dtype=torch.float32, device=x.device)
        ldo = (fout + 3) // 4 * 4
        if push:
            # partial output over ALL nodes (padded to world * chunk rows for the reduce-scatter)
            out_buf = comm.partial_buffer(ldo, comm_key, x.device)
        else:
            out_buf = torch.empty((graph.num_owned, ldo), dtype=torch.float32, device=x.device)
        out = out_buf[:, :fout] if ldo != fout else out_buf
        ws_bytes = graph.workspace_bytes(fin, fout, False)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        # the pre-reduced rows of long (relation, dst) segments are kept for dL/dW when a weight
        # gradient will be asked for (a few MB; saves recomputing them in backward)
        keep = _KEEP_CHUNKS and any(ctx.needs_input_grad[1:4]) and not (flags & _lib.F_FORCE_SIMPLE)
        xc_bytes = lib.rgcn_layer_chunk_rows_bytes(graph.handle, fin) if keep else 0
        xc = torch.empty(xc_bytes, dtype=torch.uint8, device=x.device) if xc_bytes > 0 else None
        with torch.cuda.device(x.device):
            rc = lib.rgcn_layer_fwd_keep(graph.handle, x.data_ptr(), x.stride(0), fin, weight.data_ptr(), _ptr(root_c),
                                         _ptr(bias_c), out_buf.data_ptr(), ldo, fout, flags, ws.data_ptr(), ws_bytes,
                                         _ptr(xc), _ptr(mirror), mirror.stride(0) if mirror is not None else 0,
                                         _stream(x.device))
        _lib.check(rc, 'rgcn_layer_fwd')
        if push:   # sum of the ranks' partials, each rank keeping its own rows
            out_buf = comm.reduce_scatter_rows(out_buf, comm_key)[:graph.num_owned]
            out = out_buf[:, :fout] if ldo != fout else out_buf
        # an odd-width input that arrives in zero-padded 16-byte addressable rows (the engine's transfer heads write x0
        # that way) gets its gradient in the same layout: the head's backward reads it without a padding copy
        ctx.gx_ld = x.stride(0) if (comm is None and mirror is None and fin % 4 != 0
                                    and x.stride(0) == (fin + 3) // 4 * 4 and x.data_ptr() % 16 == 0) else 0
        if mirror is not None:
            x = mirror              # what backward re-gathers; the caller's tensor is not kept
        ctx.graph, ctx.flags, ctx.comm, ctx.fin, ctx.comm_key = graph, flags, comm, fin, comm_key
        ctx.x_chunk_rows = xc
        ctx.has_root, ctx.has_bias = root is not None, bias is not None
        ctx.save_for_backward(x, weight, root_c if root_c is not None else x.new_empty(0))
        return out

    @staticmethod
    def backward(ctx, gout: Tensor):
        lib = _lib.load()
        x, weight, root = ctx.saved_tensors
        root = root if ctx.has_root else None
        graph = ctx.graph
        need_x, need_w, need_root, need_bias = ctx.needs_input_grad[:4]
        need_root = need_root and ctx.has_root
        need_bias = need_bias and ctx.has_bias
        gout = gout.contiguous()
        fin = ctx.fin              # x may be the zero-padded mirror (wider than fin)
        fout = weight.size(2)
        dev = x.device
        comm = ctx.comm
        # partitioned graph: the all-gather of gout (needed only by dL/dx) is started first and
        # overlaps the dL/dW pass, which reads the owned rows only
        gout_all, work = None, None
        fuse_mask = False
        if comm is not None and graph.push:
            # source-partitioned: dL/dW and dL/dx both read the rows of all destinations
            if need_x or need_w or need_root or need_bias:
                ho = getattr(comm, 'handoff', None)
                if ho is not None and ho[0] == gout.data_ptr() and ho[1] == tuple(gout.shape):
                    gout_all = ho[2]          # the layer above exchanged its gx while it wrote it
                    comm.handoff = None
                else:
                    gout_all = comm.all_gather_rows(gout, ('g', ctx.comm_key))
            # ReLU backward of the fused inter-layer activation rides on the exchange of gx (engine comm only)
            fuse_mask = need_x and bool(ctx.flags & _lib.F_RELU_IN) and hasattr(comm, 'handoff')
        elif comm is not None and need_x:
            if getattr(comm, 'all_gather_rows_async', None) is not None:
                gout_all, work = comm.all_gather_rows_async(gout)
            else:
                gout_all = comm.all_gather_rows(gout)
        gx = None
        if need_x:
            # 64 x 64 column passes run on the tcgen05 kernel, which adds whole 256-byte rows: an odd-width
            # gradient (emb = 63 next to hidden 64) is accumulated in 64-wide rows and returned as a view
            wide = 32 < fin <= 64 and 32 < fout <= 64 and fin % 4 != 0 and os.environ.get('RGCN_B200_TC', '1') != '0'
            if ctx.gx_ld:
                gx = torch.empty((graph.num_owned, ctx.gx_ld), dtype=torch.float32, device=dev)[:, :fin]
            elif wide:
                gx = torch.empty((graph.num_owned, 64), dtype=torch.float32, device=dev)[:, :fin]
            else:
                gx = torch.empty((graph.num_owned, fin), dtype=torch.float32, device=dev)
        gw = torch.empty_like(weight) if need_w else None
        groot = torch.empty((fin, fout), dtype=torch.float32, device=dev) if need_root else None
        gbias = torch.empty((fout,), dtype=torch.float32, device=dev) if need_bias else None

        def call(gx_t, gw_t, groot_t, gbias_t):
            ws_bytes = graph.workspace_bytes(fin, fout, True)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                rc = lib.rgcn_layer_bwd_reuse(graph.handle, x.data_ptr(), x.stride(0), fin, weight.data_ptr(), _ptr(root),
                                              gout.data_ptr(), gout.stride(0), _ptr(gout_all),
                                              gout_all.stride(0) if gout_all is not None else 0, fout, _ptr(gx_t),
                                              gx_t.stride(0) if gx_t is not None else fin,
                                              _ptr(gw_t), _ptr(groot_t), _ptr(gbias_t),
                                              ctx.flags | (_lib.F_NO_RELU_MASK if fuse_mask else 0), ws.data_ptr(),
                                              ws_bytes, _ptr(ctx.x_chunk_rows), _stream(dev))
            _lib.check(rc, 'rgcn_layer_bwd')

        if fuse_mask:
            # dL/dx first; its rows are masked and stored into every rank's copy by ONE kernel, and the exchange
            # is in flight while dL/dW runs; the layer below picks the gathered rows up (comm.handoff)
            call(gx, None, None, None)
            cur = torch.cuda.current_stream(dev)
            side = None
            if need_w or need_root or need_bias:          # dL/dW on a side stream, under the exchange
                side = comm.side_stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    call(None, gw, groot, gbias)
            gathered = comm.all_gather_rows(gx, ('g', 'below', ctx.comm_key), relu_pre=x)
            comm.handoff = (gx.data_ptr(), tuple(gx.shape), gathered)
            if side is not None:
                cur.wait_stream(side)
        elif work is not None and (need_w or need_root or need_bias):
            call(None, gw, groot, gbias)      # runs while the all-gather is in flight
            work.wait()
            call(gx, None, None, None)
        elif need_x or need_w or need_root or need_bias:
            if work is not None:
                work.wait()
            call(gx, gw, groot, gbias)
        if comm is not None:
            comm.all_reduce_sum_([t for t in (gw, groot, gbias) if t is not None])
        return gx, gw, groot, gbias, None, None, None, None


def rgcn_layer(x: Tensor, weight: Tensor, root: Optional[Tensor], bias: Optional[Tensor], graph: RGCNGraph,
               relu_in: bool = False, force_simple: bool = False, comm=None, comm_key=None) -> Tensor:
    flags = (_lib.F_RELU_IN if relu_in else 0) | (_lib.F_FORCE_SIMPLE if force_simple else 0)
    if comm is None and graph.num_owned != graph.num_nodes:
        raise ValueError('rgcn_layer: a partitioned graph needs comm= (see rgcn_b200.partition)')
    return _RGCNLayerFn.apply(x, weight, root, bias, graph, flags, comm, comm_key)


class RGCNConv(nn.Module):
    """``RGCNConv(in_channels, out_channels, num_relations, num_bases=None, ...)`` with parameters
    ``weight [R, in, out]`` (or ``[B, in, out]`` + ``comp [R, B]``), ``root [in, out]``, ``bias [out]``."""

    def __init__(self, in_channels: int, out_channels: int, num_relations: int, num_bases: Optional[int] = None,
                 num_blocks: Optional[int] = None, aggr: str = 'mean', root_weight: bool = True,
                 is_sorted: bool = False, bias: bool = True, **kwargs) -> None:
        super().__init__()
        if num_blocks is not None:
            raise NotImplementedError('RGCNConv(num_blocks=...) is outside the reference path (never used there)')
        if aggr != 'mean':
            raise NotImplementedError("RGCNConv: only aggr='mean' (PyG default, what the reference runs)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_relations, self.num_bases, self.num_blocks = num_relations, num_bases, None
        self.relu_in = False        # engine extension: apply ReLU to the input on load (fused F.relu)
        self.force_simple = False   # engine extension: generic scalar kernels (cross-check)
        w_slots = num_bases if num_bases is not None else num_relations
        self.weight = nn.Parameter(torch.empty(w_slots, in_channels, out_channels))
        if num_bases is not None:
            self.comp = nn.Parameter(torch.empty(num_relations, num_bases))
        else:
            self.register_parameter('comp', None)
        if root_weight:
            self.root = nn.Parameter(torch.empty(in_channels, out_channels))
        else:
            self.register_parameter('root', None)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        _glorot_(self.weight)
        if self.comp is not None:
            _glorot_(self.comp)
        if self.root is not None:
            _glorot_(self.root)
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()

    def forward(self, x: Tensor, edge_index, edge_type: Optional[Tensor] = None) -> Tensor:
        if isinstance(edge_index, RGCNGraph):
            graph = edge_index
        else:
            if edge_type is None:
                raise ValueError('RGCNConv.forward: edge_type is required')
            graph = cached_graph(edge_index, edge_type, x.size(0), self.num_relations)
        weight = self.weight
        if self.num_bases is not None:   # basis decomposition: tiny dense product, autograd carries comp/bases
            weight = (self.comp @ weight.view(self.num_bases, -1)).view(
                self.num_relations, self.in_channels, self.out_channels)
        return rgcn_layer(x, weight, self.root, self.bias, graph, self.relu_in, self.force_simple)

    def __repr__(self) -> str:
        return f'{self.__class__.__name__}({self.in_channels}, {self.out_channels}, num_relations={self.num_relations})'
