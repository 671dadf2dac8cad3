"""Transfer heads on the engine (reference model/layers.py:49-66, 90-112).

MLP head: ``x0 = lin2(tanh(lin1(E_cat)))`` — two dense contractions per step at the size of the whole
graph (AM-shape: 1.67 M x 189 x 137 and x 137 x 63).  Forward runs on the engine's tcgen05 kernel
(csrc/gemm_tc.cu: TMA operand loads, ``tcgen05.mma kind::tf32`` with the error-compensated 3xTF32 split,
accumulators in tensor memory, bias / tanh in the epilogue); ``lin2``'s result is written straight into
zero-padded 64-wide rows, i.e. the 16-byte addressable layout the first R-GCN layer gathers from, so the
layer needs no padding copy of x0.  Backward: ``dL/dh`` through the same kernel (W2 as the transposed
operand); the parameter gradients are reductions over all nodes with tiny outputs (137 x 189): the engine's
split-K tcgen05 kernel (csrc/gram_tc.cu, ``gram``; RGCN_B200_GRAM=0 restores the torch matmuls).

Attention head: ``x0 = MHA(E, E, E)[0]`` with S heads over the S stacked summary embeddings.  Only query position 0 is
kept by the reference, so the engine projects Q from summary 0 alone and K | V from all S summaries (two tcgen05
contractions), runs the per-(node, head) S-way softmax in one kernel (csrc/attn_head.cu) and writes ``out_proj``'s
result into the same 64-wide mirror rows.  Training-mode dropout draws the reference's own mask (a dropout over a
ones tensor of the [N * heads, S, S] attention-weight shape, the call nn.MultiheadAttention makes) and keeps row 0.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib


def _pad(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def prepack(weight: Tensor, transpose: bool = False) -> Tuple[Tensor, Tensor, int, int]:
    """nn.Linear weight [n, k] (or, transpose=True, the same tensor used as [k, n] -> n = weight.size(1)) split
    into tf32 hi / lo parts, zero-padded to [n_pad, k_pad]."""
    lib = _lib.load()
    w = weight.detach()
    if not w.is_cuda or w.dtype != torch.float32:
        raise _lib.EngineError('gemm: CUDA fp32 weights only (no CPU path)')
    w = w if w.stride(1) == 1 else w.contiguous()
    n, k = (w.size(1), w.size(0)) if transpose else (w.size(0), w.size(1))
    n_pad, k_pad = _pad(n, 16), _pad(k, 32)
    if n_pad > 256:
        raise _lib.EngineError('gemm: output width > 256 is outside the heads this kernel serves')
    hi = torch.empty((n_pad, k_pad), dtype=torch.float32, device=w.device)
    lo = torch.empty_like(hi)
    with torch.cuda.device(w.device):
        _lib.check(lib.rgcn_gemm_prepack(w.data_ptr(), w.stride(0), n, k, n_pad, k_pad, 1 if transpose else 0,
                                         hi.data_ptr(), lo.data_ptr(), _stream(w.device)), 'rgcn_gemm_prepack')
    return hi, lo, n_pad, k_pad


def rows16(a: Tensor) -> Tensor:
    """a [m, k] with 16-byte addressable rows (row stride a multiple of 4 floats, aligned base): a itself or a
    zero-padded copy."""
    if a.stride(1) == 1 and a.stride(0) % 4 == 0 and a.data_ptr() % 16 == 0:
        return a
    ld = _pad(a.size(1), 4)
    buf = torch.zeros((a.size(0), ld), dtype=torch.float32, device=a.device)
    buf[:, :a.size(1)] = a
    return buf[:, :a.size(1)]


def gemm(a: Tensor, weight: Tensor, bias: Optional[Tensor] = None, act: str = 'none', out_ld: Optional[int] = None,
         transpose_w: bool = False, aux: Optional[Tensor] = None) -> Tensor:
    """act(a @ W^T + bias) on the tcgen05 kernel.  a [m, k]; weight [n, k] (transpose_w: weight is [k, n]).
    act: 'none', 'tanh', or 'dtanh' = multiply by (1 - aux^2) with aux [m, n] the saved tanh output.
    Returns a [m, n] VIEW of an [m, out_ld] buffer whose extra columns are zero (out_ld default = n rounded up to 4)."""
    lib = _lib.load()
    if not a.is_cuda or a.dtype != torch.float32:
        raise _lib.EngineError('gemm: CUDA fp32 input only (no CPU path)')
    a = rows16(a)
    m, k = a.shape
    hi, lo, n_pad, k_pad = prepack(weight, transpose_w)
    n = weight.size(1) if transpose_w else weight.size(0)
    ld = out_ld if out_ld is not None else _pad(n, 4)
    if ld % 4 or ld < n or ld > n_pad:
        raise ValueError('gemm: out_ld must be a multiple of 4 in [n, n_pad]')
    bp = None
    if bias is not None:
        bp = torch.zeros(n_pad, dtype=torch.float32, device=a.device)
        bp[:n] = bias.detach()
    out = torch.empty((m, ld), dtype=torch.float32, device=a.device)
    code = {'none': 0, 'tanh': 1, 'dtanh': 2}[act]
    if code == 2:
        if aux is None or aux.shape != (m, n) or not aux.is_cuda:
            raise ValueError("gemm: act='dtanh' needs aux [m, n]")
        aux = rows16(aux)
    with torch.cuda.device(a.device):
        rc = lib.rgcn_gemm3x_tf32(a.data_ptr(), a.stride(0), m, k, hi.data_ptr(), lo.data_ptr(), n_pad, k_pad,
                                  bp.data_ptr() if bp is not None else None, code,
                                  aux.data_ptr() if code == 2 else None, aux.stride(0) if code == 2 else 0,
                                  out.data_ptr(), ld, ld, _stream(a.device))
    _lib.check(rc, 'rgcn_gemm3x_tf32')
    return out[:, :n] if ld != n else out


def padded_like(rows: int, cols: int, device) -> Tensor:
    """An uninitialised [rows, cols] view of a [rows, ceil4(cols)] buffer (16-byte addressable rows)."""
    return torch.empty((rows, _pad(cols, 4)), dtype=torch.float32, device=device)[:, :cols]


def gram(a: Tensor, b: Tensor, colsums: bool = False):
    """a^T @ b for a [M, n1], b [M, n2] with M = all nodes: the parameter-gradient reduction of a head, on the
    engine's split-K tcgen05 kernel (csrc/gram_tc.cu).  Operands that are not 16-byte addressable are copied.
    colsums=True: returns (a^T b, a.sum(0), b.sum(0)) — the bias gradients, accumulated inside the same kernel."""
    lib = _lib.load()
    if not a.is_cuda or not b.is_cuda or a.dtype != torch.float32 or b.dtype != torch.float32 or a.size(0) != b.size(0):
        raise _lib.EngineError('gram: two CUDA fp32 matrices with the same number of rows (no CPU path)')
    n1, n2 = a.size(1), b.size(1)

    def cost(m_side, n_side):                    # tensor-core work: 128-column blocks of the M side x the padded N side
        return float('inf') if n_side > 192 else ((m_side + 127) // 128) * (128 + _pad(n_side, 32))
    swap = cost(n2, n1) < cost(n1, n2)
    if swap:
        a, b, n1, n2 = b, a, n2, n1
    if n2 > 192:
        raise _lib.EngineError('gram: one side must be at most 192 wide')
    a, b = rows16(a), rows16(b)
    c = torch.empty((n1, n2), dtype=torch.float32, device=a.device)
    sa = torch.empty(n1, dtype=torch.float32, device=a.device) if colsums else None
    sb = torch.empty(n2, dtype=torch.float32, device=a.device) if colsums else None
    with torch.cuda.device(a.device):
        rc = lib.rgcn_gram3x_tf32(a.data_ptr(), a.stride(0), n1, b.data_ptr(), b.stride(0), n2, a.size(0), c.data_ptr(), n2,
                                  sa.data_ptr() if colsums else None, sb.data_ptr() if colsums else None, _stream(a.device))
    _lib.check(rc, 'rgcn_gram3x_tf32')
    c = c.t().contiguous() if swap else c
    if not colsums:
        return c
    return (c, sb, sa) if swap else (c, sa, sb)


class _MLPHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e_cat: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor) -> Tensor:
        a = rows16(e_cat)
        h = gemm(a, w1, b1, act='tanh')                       # [N, mid], rows padded to a multiple of 4
        x0 = gemm(h, w2, b2, out_ld=_pad(w2.size(0), 4))      # [N, emb] view of [N, ceil4(emb)]: the layer's mirror layout
        ctx.save_for_backward(a, h, w1, w2)
        return x0

    @staticmethod
    def backward(ctx, g: Tensor):
        a, h, w1, w2 = ctx.saved_tensors
        need_e, need_w1, need_b1, need_w2, need_b2 = ctx.needs_input_grad
        g = rows16(g)                                         # (one padding copy when the layer hands over packed rows)
        use_gram = os.environ.get('RGCN_B200_GRAM', '1') != '0'

        def red(p, q):          # (p^T q, p.sum(0)): [out, in] weight gradient + bias gradient, reductions over all nodes
            if use_gram:
                c, sp, _ = gram(p, q, colsums=True)
                return c, sp
            return p.t() @ q, p.sum(0)
        gw2 = gb2 = None
        if need_w2 or need_b2:
            gw2, gb2 = red(g, h)                              # [emb, mid], [emb]
        ge = gw1 = gb1 = None
        if need_e or need_w1 or need_b1:
            # (g @ W2) * (1 - h^2) -> [N, mid]: the tanh backward rides in the epilogue of the tensor-core kernel
            dpre = gemm(g, w2, transpose_w=True, act='dtanh', aux=h)
            if need_w1 or need_b1:
                gw1, gb1 = red(dpre, a)
            if need_e:
                ge = gemm(dpre, w1, transpose_w=True)
        return ge, gw1, gb1, gw2, gb2


def mlp_head(e_cat: Tensor, lin1: torch.nn.Linear, lin2: torch.nn.Linear) -> Tensor:
    """lin2(tanh(lin1(e_cat))) (reference model/layers.py:105-107) on the engine."""
    return _MLPHeadFn.apply(e_cat, lin1.weight, lin1.bias, lin2.weight, lin2.bias)


def _flat_rows16(e: Tensor) -> Tensor:
    """e [S, N, emb] as [S * N, emb] rows with 16-byte addressable rows: a view when e already is one padded block
    (Emb_ATT_Layers caches frozen embeddings that way), else a padded copy."""
    s, n, emb = e.shape
    if e.stride(2) == 1 and e.stride(1) % 4 == 0 and e.stride(0) == n * e.stride(1) and e.data_ptr() % 16 == 0:
        return e.as_strided((s * n, emb), (e.stride(1), 1))
    return rows16(e.reshape(s * n, emb))


def stack_rows16(e: Tensor) -> Tensor:
    """[S, N, emb] -> the same values as a view of one zero-padded [S, N, ceil4(emb)] block."""
    s, n, emb = e.shape
    buf = torch.zeros((s, n, _pad(emb, 4)), dtype=torch.float32, device=e.device)
    buf[:, :, :emb] = e
    return buf[:, :, :emb]


def _kv_weights(w_in: Tensor, b_in: Optional[Tensor], emb: int) -> Tuple[Tensor, Optional[Tensor], int]:
    """in_proj rows of K and V as one operand whose two halves start at multiples of 4 columns (zero rows between):
    the projected K | V rows are then 16-byte addressable on both sides of the split."""
    ep = _pad(emb, 4)
    w = torch.zeros((2 * ep, emb), dtype=torch.float32, device=w_in.device)
    w[:emb] = w_in[emb:2 * emb].detach()
    w[ep:ep + emb] = w_in[2 * emb:].detach()
    b = None
    if b_in is not None:
        b = torch.zeros(2 * ep, dtype=torch.float32, device=w_in.device)
        b[:emb] = b_in[emb:2 * emb].detach()
        b[ep:ep + emb] = b_in[2 * emb:].detach()
    return w, b, ep


class _AttentionHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e: Tensor, w_in: Tensor, b_in: Tensor, w_out: Tensor, b_out: Tensor, keep: Optional[Tensor],
                heads: int) -> Tensor:
        lib = _lib.load()
        s, n, emb = e.shape
        d = emb // heads
        e_flat = _flat_rows16(e)
        w_kv, b_kv, ep = _kv_weights(w_in, b_in, emb)
        q = gemm(e_flat[:n], w_in[:emb], b_in[:emb] if b_in is not None else None)
        kv = gemm(e_flat, w_kv, b_kv)                         # [S * N, 2 ep]: keys at column 0, values at column ep
        probs = torch.empty((n, heads, s), dtype=torch.float32, device=e.device)
        o = torch.empty((n, ep), dtype=torch.float32, device=e.device)       # pad columns are never read (TMA bounds)
        with torch.cuda.device(e.device):
            rc = lib.rgcn_attn_head_fwd(q.data_ptr(), q.stride(0), kv.data_ptr(), kv.stride(0), ep, s, n, heads, d,
                                        keep.data_ptr() if keep is not None else None, probs.data_ptr(), o.data_ptr(), ep,
                                        _stream(e.device))
        _lib.check(rc, 'rgcn_attn_head_fwd')
        o = o[:, :emb]
        x0 = gemm(o, w_out, b_out, out_ld=ep)
        ctx.save_for_backward(e_flat, q, kv, probs, o, w_in, w_out, keep, w_kv)
        ctx.dims = (s, n, emb, heads, d, b_in is not None, ep)
        return x0

    @staticmethod
    def backward(ctx, g: Tensor):
        lib = _lib.load()
        e_flat, q, kv, probs, o, w_in, w_out, keep, w_kv = ctx.saved_tensors
        s, n, emb, heads, d, has_b_in, ep = ctx.dims
        need_e, need_w_in, need_b_in, need_w_out, need_b_out = ctx.needs_input_grad[:5]
        g = rows16(g)
        use_gram = os.environ.get('RGCN_B200_GRAM', '1') != '0'

        def red(p, q):          # (p^T q, p.sum(0))
            if use_gram:
                c, sp, _ = gram(p, q, colsums=True)
                return c, sp
            return p.t() @ q, p.sum(0)
        gw_out = gb_out = None
        if need_w_out or need_b_out:
            gw_out, gb_out = red(g, o)
        ge = gw_in = gb_in = None
        if need_e or need_w_in or need_b_in:
            go = gemm(g, w_out, transpose_w=True)                                       # dL/do [N, emb]
            gq = torch.empty((n, ep), dtype=torch.float32, device=g.device)
            gkv = torch.empty((s * n, 2 * ep), dtype=torch.float32, device=g.device)
            with torch.cuda.device(g.device):
                rc = lib.rgcn_attn_head_bwd(q.data_ptr(), q.stride(0), kv.data_ptr(), kv.stride(0), ep, s, n, heads, d,
                                            keep.data_ptr() if keep is not None else None, probs.data_ptr(),
                                            go.data_ptr(), go.stride(0), gq.data_ptr(), gq.stride(0), gkv.data_ptr(),
                                            gkv.stride(0), _stream(g.device))
            _lib.check(rc, 'rgcn_attn_head_bwd')
            gq = gq[:, :emb]
            if need_w_in or (need_b_in and has_b_in):
                gkv_w, gkv_b = red(gkv, e_flat)                                         # [2 ep, emb], [2 ep]; pad rows dropped
                gq_w, gq_b = red(gq, e_flat[:n])
                gw_in = torch.cat([gq_w, gkv_w[:emb], gkv_w[ep:ep + emb]], 0)
                if has_b_in:
                    gb_in = torch.cat([gq_b, gkv_b[:emb], gkv_b[ep:ep + emb]], 0)
            if need_e:
                if ep != emb:                                 # the pad columns meet zero weight rows: they must be finite
                    gkv[:, emb:ep].zero_()
                    gkv[:, ep + emb:].zero_()
                ge = gemm(gkv, w_kv, transpose_w=True).contiguous()
                ge[:n] += gemm(gq, w_in[:emb], transpose_w=True)
                ge = ge.view(s, n, emb)
        return ge, gw_in, gb_in, gw_out, gb_out, None, None


def attention_keep_mask(num_nodes: int, heads: int, num_sums: int, p: float, device) -> Tensor:
    """The dropout keep mask (scaled by 1 / (1 - p)) of query position 0, drawn the way nn.MultiheadAttention draws it:
    one dropout over the [N * heads, L = S, S] attention weights (torch/nn/functional.py multi_head_attention_forward),
    so the same generator state gives the same mask as the reference module."""
    ones = torch.ones((num_nodes * heads, num_sums, num_sums), dtype=torch.float32, device=device)
    return torch.nn.functional.dropout(ones, p=p, training=True)[:, 0, :].contiguous()


def attention_head(e: Tensor, att: torch.nn.MultiheadAttention) -> Tensor:
    """att(e, e, e)[0][0] (reference model/layers.py:59-61; e [S, N, emb], S heads) on the engine."""
    s, n, emb = e.shape
    if att.in_proj_weight is None or att.batch_first or att.bias_k is not None or att.add_zero_attn or emb % att.num_heads:
        raise _lib.EngineError('attention_head: serves the reference configuration (packed in_proj, sequence-first, '
                               'no bias_k / zero_attn)')
    if s > 8 or emb // att.num_heads > 64:
        raise _lib.EngineError('attention_head: at most 8 summaries and head_dim <= 64')
    keep = None
    if att.training and att.dropout > 0.0:
        keep = attention_keep_mask(n, att.num_heads, s, att.dropout, e.device)
    return _AttentionHeadFn.apply(e, att.in_proj_weight, att.in_proj_bias, att.out_proj.weight, att.out_proj.bias, keep,
                                  att.num_heads)
