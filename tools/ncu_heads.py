"""The head kernels at AM shape, one launch each after a warm-up (for ncu):
   ncu --set full -k regex:'k_gemm3x|k_gram3x|k_attn' --launch-skip 6 --launch-count 6 python tools/ncu_heads.py"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'scaling-rgcn-training_b200')]
import __graft_entry__  # noqa: E402,F401
from rgcn_b200 import heads as H  # noqa: E402

dev = torch.device('cuda:0')
n = 1666764
torch.manual_seed(0)
a = H.rows16(torch.randn(n, 189, device=dev))
w1 = torch.randn(137, 189, device=dev) * 0.1
b1 = torch.randn(137, device=dev)
w2 = torch.randn(63, 137, device=dev) * 0.1
g = H.rows16(torch.randn(n, 63, device=dev))
att = torch.nn.MultiheadAttention(63, 3, dropout=0.0).to(dev)
e = H.stack_rows16(torch.randn(3, n, 63, device=dev)).requires_grad_(False)
for rep in range(2):            # second round = the profiled one
    h = H.gemm(a, w1, b1, act='tanh')                         # k_gemm3x  189 -> 137, tanh
    x0 = H.gemm(h, w2, None, out_ld=64)                       # k_gemm3x  137 -> 63
    dpre = H.gemm(g, w2, transpose_w=True, act='dtanh', aux=h)   # k_gemm3x  63 -> 137, tanh backward
    gw2 = H.gram(g, h)                                        # k_gram3x  63 x 137
    gw1 = H.gram(dpre, a)                                     # k_gram3x  137 x 189 (swapped)
    with torch.enable_grad():
        out = H.attention_head(e, att)                        # 2 x k_gemm3x, k_attn_fwd_st, k_gemm3x
        out.sum().backward()                                  # k_gemm3x, k_attn_bwd_st, 3 x k_gram3x
    att.zero_grad()
    torch.cuda.synchronize()
print('ok')
