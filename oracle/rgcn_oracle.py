"""CPU ORACLE — test infrastructure only, never a product path.

Restatement of the R-GCN layer the reference calls: ``torch_geometric.nn.RGCNConv``
of the pinned ``torch_geometric==2.3.1`` (reference ``requirements.txt:7``), per-relation
*loop path* (no ``pyg_lib`` in ``requirements.txt`` -> ``WITH_PYG_LIB`` is false).
PyG is a third-party dependency that is NOT vendored under /root/reference and is not
installable here (no network), so this file restates its published algorithm:

  torch_geometric/nn/conv/rgcn_conv.py   RGCNConv.__init__/reset_parameters/forward
  torch_geometric/utils/scatter.py       reduce='mean' = scatter_add sum / scatter_add count.clamp(min=1)
  torch_geometric/nn/inits.py            glorot, zeros
  torch_geometric/data/data.py           Data (attribute bag, .to())

Reference call sites this must serve (all under /root/reference):
  construct  model/layers.py:15-16,54-55,98-99
  init       model/layers.py:17-18 (kaiming_uniform_ on .weight in place)
  forward    model/layers.py:21,23,62,64,108,110
  params     model/layers.py:34-46, model/modelTrainer.py:28-35

PARITY UNPINNED: the reference ships no tests, golden vectors or seeds for this path
 and real PyG cannot be run here.  The oracle is cross-checked
instead by (i) an independent dense-adjacency formulation (``dense_rgcn_forward``),
(ii) fp64 ``torch.autograd.gradcheck``, (iii) fixtures produced by running the reference's
own unmodified callers (graphs/graph.py, model/layers.py, model/embeddingTricks.py) on top
of this oracle (tests/golden/, generator: tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
from __future__ import annotations

import copy
import math
from typing import Optional

import torch
from torch import Tensor, nn


def _glorot_(t: Tensor) -> None:
    # torch_geometric/nn/inits.py glorot: U(-a, a), a = sqrt(6 / (size(-2) + size(-1)))
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)


def scatter_mean_rows(src_rows: Tensor, index: Tensor, dim_size: int) -> Tensor:
    """torch_geometric.utils.scatter(..., reduce='mean') for dim=0, as in PyG 2.3.1."""
    count = src_rows.new_zeros(dim_size)
    count.scatter_add_(0, index, src_rows.new_ones(src_rows.size(0)))
    count = count.clamp(min=1)
    idx = index.view(-1, 1).expand_as(src_rows)
    out = src_rows.new_zeros((dim_size, src_rows.size(1))).scatter_add_(0, idx, src_rows)
    return out / count.view(-1, 1)


def rgcn_forward(x: Tensor, edge_index: Tensor, edge_type: Tensor, weight: Tensor,
                 root: Optional[Tensor], bias: Optional[Tensor],
                 comp: Optional[Tensor] = None) -> Tensor:
    """RGCNConv.forward, loop path.

    out = sum_r mean_{e: type=r, dst=i} x[src_e] @ W_r  +  x @ root  +  bias
    Relations iterate ascending and accumulate sequentially, exactly like PyG.
    """
    n = x.size(0)
    if comp is not None:  # basis decomposition
        nb, fin, fout = weight.shape
        w = (comp @ weight.view(nb, -1)).view(comp.size(0), fin, fout)
    else:
        w = weight
    num_relations = w.size(0)
    out = torch.zeros(n, w.size(2), dtype=x.dtype, device=x.device)
    for r in range(num_relations):
        m = edge_type == r
        src = edge_index[0, m]
        dst = edge_index[1, m]
        h = scatter_mean_rows(x.index_select(0, src), dst, n)   # flow source_to_target
        out = out + h @ w[r]
    if root is not None:
        out = out + x @ root
    if bias is not None:
        out = out + bias
    return out


def dense_rgcn_forward(x, edge_index, edge_type, weight, root, bias):
    """Independent formulation used to cross-check ``rgcn_forward`` on small graphs:
    one dense row-normalised adjacency matrix per relation (multi-edges counted)."""
    n = x.size(0)
    out = x @ root + bias
    for r in range(weight.size(0)):
        a = torch.zeros(n, n, dtype=x.dtype)
        m = edge_type == r
        for s, d in zip(edge_index[0, m].tolist(), edge_index[1, m].tolist()):
            a[d, s] += 1.0
        deg = a.sum(1, keepdim=True).clamp(min=1)
        out = out + (a / deg) @ x @ weight[r]
    return out


class RGCNConv(nn.Module):
    """Constructor/attribute-compatible stand-in for torch_geometric.nn.RGCNConv 2.3.1
    (the subset the reference uses: num_blocks unsupported, aggr='mean', root_weight, bias)."""

    def __init__(self, in_channels: int, out_channels: int, num_relations: int,
                 num_bases: Optional[int] = None, num_blocks: Optional[int] = None,
                 aggr: str = 'mean', root_weight: bool = True, is_sorted: bool = False,
                 bias: bool = True, **kwargs) -> None:
        super().__init__()
        if num_blocks is not None:
            raise NotImplementedError('num_blocks is never used by the reference')
        if aggr != 'mean':
            raise NotImplementedError("only aggr='mean' (the PyG default) is restated")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_relations = num_relations
        self.num_bases = num_bases
        self.num_blocks = None
        if num_bases is not None:
            self.weight = nn.Parameter(torch.empty(num_bases, in_channels, out_channels))
            self.comp = nn.Parameter(torch.empty(num_relations, num_bases))
        else:
            self.weight = nn.Parameter(torch.empty(num_relations, in_channels, out_channels))
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

    def forward(self, x: Tensor, edge_index: Tensor, edge_type: Tensor) -> Tensor:
        return rgcn_forward(x, edge_index, edge_type, self.weight, self.root, self.bias, self.comp)

    def __repr__(self) -> str:
        return (f'{self.__class__.__name__}({self.in_channels}, {self.out_channels}, '
                f'num_relations={self.num_relations})')


class Data:
    """torch_geometric.data.Data restated as the attribute bag the reference needs
    (graphs/graph.py:68-69, graphs/dataset.py:30-35; .to() in place, returns self)."""

    def __init__(self, **kwargs) -> None:
        for k, v in kwargs.items():
            setattr(self, k, v)

    def to(self, device, *args, **kwargs) -> 'Data':
        for k, v in list(self.__dict__.items()):
            if isinstance(v, Tensor):
                setattr(self, k, v.to(device, *args, **kwargs))
        return self

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith('_')]

    def __deepcopy__(self, memo):
        out = Data()
        for k, v in self.__dict__.items():
            setattr(out, k, copy.deepcopy(v, memo))
        return out


# ---------------------------------------------------------------------------------------
# Map-gather oracle (reference model/embeddingTricks.py:8-49) on integer index form.
# ---------------------------------------------------------------------------------------

def build_map_index(org_node_to_enum: dict, sum_node_to_enum: dict, org2sum: dict):
    """Integer restatement of the dict walk in get_tensor_list (embeddingTricks.py:19-23):
    idx[i] = row of the summary node org node i maps to, or -1 (keeps the fallback row)."""
    idx = torch.full((len(org_node_to_enum),), -1, dtype=torch.int64)
    for org_node, i in org_node_to_enum.items():
        if org_node in org2sum:
            s = org2sum[org_node]
            if s in sum_node_to_enum:
                idx[i] = sum_node_to_enum[s]
    return idx


def map_gather(emb_list, idx_list, fallback_list, mode: str) -> Tensor:
    """get_tensor_list + sum/concat/stack (embeddingTricks.py:27-49) with the torch.rand
    fallback rows supplied by the caller so that both sides see the same values."""
    tensors = []
    for emb, idx, fb in zip(emb_list, idx_list, fallback_list):
        t = fb.clone()
        m = idx >= 0
        t[m] = emb[idx[m]].detach()
        tensors.append(t)
    if mode == 'sum':
        return sum(tensors).detach()          # python sum: starts from int 0, s = 0..S-1
    if mode == 'concat':
        return torch.concat(tensors, dim=-1).detach()
    if mode == 'stack':
        return torch.stack(tensors).detach()
    raise ValueError(mode)
