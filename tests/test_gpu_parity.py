"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs, against the committed golden fixtures, and through
size-independent properties at larger sizes.  Tolerance (BASELINE.json north_star): integer
work bit-exact; fp32 outputs and gradients within 1e-5 relative (max-norm), identical argmax."""
import numpy as np
import pytest
import torch

from conftest import golden_graph, load_golden, rel_err

import csr_oracle
import rgcn_oracle
from rgcn_b200 import (Data, Emb_ATT_Layers, Emb_Layers, Emb_MLP_Layers, RGCNConv, RGCNGraph, _lib, map_gather,
                       rgcn_layer)
from rgcn_b200.synthetic import am_shape, random_multigraph
from rgcn_b200.trainer import bce_loss, ce_loss

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = 'cuda:0'

GRAPHS = ['TEST_complete', 'TEST_sum_in_out', 'AIFB_sum_in', 'AIFB_sum_in_out', 'MUTAG_bisim_k1', 'AIFB_bisim_k3']


# ------------------------------------------------------------------ K0: integer, bit-exact
def _check_brc(g, which, want):
    S = g.query(_lib.Q_NUM_SEGMENTS, which)
    assert S == want['num_seg']
    assert g.query(_lib.Q_NUM_CHUNKS, which) == want['num_chunks']
    assert g.query(_lib.Q_NUM_BATCHES, which) == want['num_batches']
    assert g.query(_lib.Q_NUM_GROUPS, which) == want['num_groups']
    assert g.query(_lib.Q_NUM_TILES, which) == want['num_tiles']
    assert g.query(_lib.Q_NUM_TILES_NOSELF, which) == want['num_tiles_noself']
    pairs = [(_lib.A_PERM, 'perm'), (_lib.A_SEG_PTR, 'seg_ptr'), (_lib.A_SEG_OWN, 'seg_own'), (_lib.A_SEG_REL, 'seg_rel'),
             (_lib.A_E_IDX, 'e_idx'), (_lib.A_E_W, 'e_w'), (_lib.A_RAW_IDX, 'raw_idx'), (_lib.A_RAW_W, 'raw_w'),
             (_lib.A_CHUNK_BEG, 'chunk_beg'), (_lib.A_CHUNK_END, 'chunk_end'), (_lib.A_BAT_SEG0, 'bat_seg0'),
             (_lib.A_BAT_INFO, 'bat_info'), (_lib.A_E_OWN, 'e_own'), (_lib.A_TILE_E0, 'tile_e0'),
             (_lib.A_TILE_INFO, 'tile_info')]
    for aid, name in pairs:
        got = g.export(aid, which)
        assert got.dtype == want[name].dtype or got.view(want[name].dtype).dtype == want[name].dtype
        assert np.array_equal(got.view(want[name].dtype), want[name]), name
    assert np.array_equal(np.diff(g.export(_lib.A_SEG_PTR0, which)), want['cnt'])


@pytest.mark.parametrize('name', GRAPHS)
@pytest.mark.parametrize('nr,t,ch', [(0, 0, 0), (64, 8, 4), (5, 2, 2)])
def test_graph_build_bit_exact(name, nr, t, ch):
    ei, et, n, r = golden_graph(name)
    g = RGCNGraph(ei.to(DEV), et.to(DEV), n, r, range_nodes=nr, split_threshold=t, chunk_size=ch)
    nr_eff, t_eff, ch_eff = (nr or 16384), (t or 16), (ch or 64)
    src, dst, rel = ei[0].numpy(), ei[1].numpy(), et.numpy()
    fwd, bwd = csr_oracle.build_graph(src, dst, rel, n, r, nr_eff, t_eff, ch_eff)
    _check_brc(g, _lib.BRC_FWD, fwd)
    _check_brc(g, _lib.BRC_BWD, bwd)
    fwd_rel = csr_oracle.build_brc(dst, src, rel, n, r, n, t_eff, ch_eff, w_edge=csr_oracle.edge_weights(dst, rel, n))
    if nr_eff < n:      # several ranges: FWD_REL is its own structure and shares FWD's chunk numbering
        fwd_rel = csr_oracle.share_chunks(fwd, fwd_rel, n)
        assert np.array_equal(g.export(_lib.A_CHUNK_OUT, _lib.BRC_FWD_REL), fwd_rel['chunk_out'])
    _check_brc(g, _lib.BRC_FWD_REL, fwd_rel)


def test_graph_build_rejects_bad_indices_and_cpu_tensors():
    ei = torch.tensor([[0, 1], [1, 5]], device=DEV)
    with pytest.raises(_lib.EngineError, match='outside'):
        RGCNGraph(ei, torch.tensor([0, 1], device=DEV), 3, 2)
    with pytest.raises(_lib.EngineError, match='outside'):
        RGCNGraph(torch.tensor([[0, 1], [1, 2]], device=DEV), torch.tensor([0, 2], device=DEV), 3, 2)
    with pytest.raises(_lib.EngineError):
        RGCNGraph(torch.tensor([[0, 1], [1, 2]]), torch.tensor([0, 1]), 3, 2)


def test_graph_build_empty_edge_list():
    ei = torch.zeros((2, 0), dtype=torch.long, device=DEV)
    et = torch.zeros((0,), dtype=torch.long, device=DEV)
    conv = RGCNConv(5, 3, 4).to(DEV)
    x = torch.randn(6, 5, device=DEV)
    out = conv(x, ei, et)
    ref = x.cpu() @ conv.root.detach().cpu() + conv.bias.detach().cpu()
    assert rel_err(out, ref) < TOL


# ------------------------------------------------------------------ K1/K3/K4 vs oracle
def _layer_case(ei, et, n, r, fin, fout, seed, relu_in=False, force_simple=False, graph_kw=None, need=(True,) * 4):
    torch.manual_seed(seed)
    x = torch.randn(n, fin)
    w = (torch.rand(r, fin, fout) - 0.5) * 0.4
    root = (torch.rand(fin, fout) - 0.5) * 0.4
    bias = (torch.rand(fout) - 0.5) * 0.2
    gout = torch.randn(n, fout)
    leaves = [t.clone().requires_grad_(nd) for t, nd in zip((x, w, root, bias), need)]
    xin = torch.relu(leaves[0]) if relu_in else leaves[0]
    ref = rgcn_oracle.rgcn_forward(xin, ei, et, leaves[1], leaves[2], leaves[3])
    if any(need):
        ref.backward(gout)
    # fp64 truth for an accuracy bound that does not depend on the oracle's own fp32 round-off
    x64 = torch.relu(x.double()) if relu_in else x.double()
    ref64 = rgcn_oracle.rgcn_forward(x64, ei, et, w.double(), root.double(), bias.double())
    g = RGCNGraph(ei.to(DEV), et.to(DEV), n, r, **(graph_kw or {}))
    dl = [t.clone().to(DEV).requires_grad_(nd) for t, nd in zip((x, w, root, bias), need)]
    out = rgcn_layer(dl[0], dl[1], dl[2], dl[3], g, relu_in=relu_in, force_simple=force_simple)
    if any(need):
        out.backward(gout.to(DEV))
    torch.cuda.synchronize()
    assert rel_err(out, ref) < TOL, ('out', rel_err(out, ref))
    assert rel_err(out, ref64) < TOL
    for name, a, b, nd in zip(('gx', 'gw', 'groot', 'gbias'), dl, leaves, need):
        if nd:
            assert rel_err(a.grad, b.grad) < TOL, (name, rel_err(a.grad, b.grad))
        else:
            assert a.grad is None
    return out


@pytest.mark.parametrize('name', GRAPHS)
@pytest.mark.parametrize('fin,fout', [(63, 16), (16, 11), (64, 26)])
def test_layer_fwd_bwd_on_reference_graphs(name, fin, fout):
    ei, et, n, r = golden_graph(name)
    _layer_case(ei, et, n, r, fin, fout, seed=fin + fout)


@pytest.mark.parametrize('fin,fout', [(1, 1), (3, 2), (16, 16), (17, 33), (20, 7), (32, 64), (60, 60), (64, 64), (63, 11),
                                      (63, 64), (64, 63), (36, 40)])
def test_layer_shapes_on_random_multigraph(fin, fout):
    ei, et = random_multigraph(257, 3000, 7, seed=fin * 100 + fout, hub_frac=0.25, dup_frac=0.2)
    _layer_case(ei, et, 257, 7, fin, fout, seed=1)


@pytest.mark.parametrize('n', [1, 15, 16, 17, 1003])
@pytest.mark.parametrize('fin', [57, 61, 62, 63, 64])
def test_dx_bulk_scatter_paths(n, fin):
    """dL/dx with 64 padded columns is scattered by the bulk-copy engine: whole rows for fin = 64, the
    shifted 16-byte windows of tightly packed odd-width rows otherwise (every row phase 63 i mod 4, the
    scalar fallback of the last rows, graphs smaller than one tile)."""
    ei, et = random_multigraph(n, 40 * n, 9, seed=n + fin, hub_frac=0.3, dup_frac=0.2)
    _layer_case(ei, et, n, 9, fin, 16, seed=3)


@pytest.mark.parametrize('fin,fout', [(100, 70), (65, 3), (5, 130)])
def test_layer_wide_shapes_use_generic_kernels(fin, fout):
    ei, et = random_multigraph(90, 800, 5, seed=7, hub_frac=0.2, dup_frac=0.1)
    _layer_case(ei, et, 90, 5, fin, fout, seed=2)


@pytest.mark.parametrize('kw', [dict(range_nodes=16, split_threshold=4, chunk_size=3), dict(range_nodes=1, split_threshold=1, chunk_size=1),
                                dict(range_nodes=100000, split_threshold=100000)])
def test_layer_invariant_to_blocking_and_chunking(kw):
    ei, et = random_multigraph(120, 2500, 6, seed=11, hub_frac=0.5, dup_frac=0.3)
    _layer_case(ei, et, 120, 6, 63, 16, seed=3, graph_kw=kw)
    _layer_case(ei, et, 120, 6, 16, 11, seed=4, graph_kw=kw)


def test_layer_simple_and_tensor_paths_agree_with_oracle():
    ei, et, n, r = golden_graph('AIFB_sum_in_out')
    a = _layer_case(ei, et, n, r, 63, 16, seed=5, force_simple=True)
    b = _layer_case(ei, et, n, r, 63, 16, seed=5, force_simple=False)
    assert rel_err(a, b.cpu()) < TOL


@pytest.mark.parametrize('need', [(False, True, True, True), (True, False, False, False), (False, False, False, True),
                                  (False, True, False, False), (True, False, True, False)])
def test_layer_respects_requires_grad_flags(need):
    """-e_freeze True / -w_grad False (main.py:84-87) switch whole backward passes off."""
    ei, et = random_multigraph(64, 700, 5, seed=13, hub_frac=0.2)
    _layer_case(ei, et, 64, 5, 63, 16, seed=6, need=need)


@pytest.mark.parametrize('fin,fout', [(16, 11), (63, 16)])
def test_layer_fused_relu_input(fin, fout):
    ei, et = random_multigraph(150, 2000, 5, seed=17, hub_frac=0.3, dup_frac=0.1)
    _layer_case(ei, et, 150, 5, fin, fout, seed=7, relu_in=True, graph_kw=dict(split_threshold=8, chunk_size=4))


def test_isolated_nodes_unused_relation_slot_and_duplicates():
    # node 5 isolated; relation 3 (the 2|rel| slot, modelTrainer.py:78) never used; duplicate edges
    ei = torch.tensor([[0, 0, 0, 1, 2, 3, 3], [1, 1, 1, 2, 0, 4, 4]])
    et = torch.tensor([0, 0, 0, 1, 2, 0, 0])
    out = _layer_case(ei, et, 6, 4, 63, 16, seed=8)
    assert torch.isfinite(out).all()


def test_basis_decomposition_module():
    ei, et = random_multigraph(80, 900, 6, seed=19)
    torch.manual_seed(0)
    ref = rgcn_oracle.RGCNConv(20, 12, 6, num_bases=3)
    mine = RGCNConv(20, 12, 6, num_bases=3)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(DEV)
    x = torch.randn(80, 20)
    a = ref(x, ei, et)
    b = mine(x.to(DEV), ei.to(DEV), et.to(DEV))
    a.sum().backward()
    b.sum().backward()
    assert rel_err(b, a) < TOL
    assert rel_err(mine.comp.grad, ref.comp.grad) < TOL and rel_err(mine.weight.grad, ref.weight.grad) < TOL


def test_strided_views_and_graph_cache():
    ei, et, n, r = golden_graph('MUTAG_bisim_k1')
    assert not ei.is_contiguous()
    d = Data(edge_index=ei)
    d.edge_type = et
    d = d.to(DEV)
    from rgcn_b200.graph import _CACHE, clear_cache
    clear_cache()
    c1, c2 = RGCNConv(8, 8, r).to(DEV), RGCNConv(8, 4, r).to(DEV)
    x = torch.randn(n, 8, device=DEV)
    c2(c1(x, d.edge_index, d.edge_type), d.edge_index, d.edge_type)
    assert len(_CACHE) == 1          # both layers share one build (model/layers.py:21,23)


# ------------------------------------------------------------------ full models vs reference-caller goldens
@pytest.mark.parametrize('name', ['TEST_complete', 'AIFB_sum_in', 'AIFB_sum_in_out', 'MUTAG_bisim_k1', 'AIFB_bisim_k3'])
@pytest.mark.parametrize('fused', [False, True])
def test_emb_layers_matches_reference_golden(name, fused):
    ei, et, n, r = golden_graph(name)
    g = load_golden(f'layers_{name}.npz')
    emb, hid, c = g['emb'].shape[1], g['w1'].shape[2], g['w2'].shape[2]
    model = Emb_Layers(r, hid, c, n, emb, 1)
    with torch.no_grad():
        model.embedding.weight.copy_(torch.from_numpy(g['emb']))
        for conv, sfx in ((model.rgcn1, '1'), (model.rgcn2, '2')):
            conv.weight.copy_(torch.from_numpy(g['w' + sfx]))
            conv.root.copy_(torch.from_numpy(g['r' + sfx]))
            conv.bias.copy_(torch.from_numpy(g['b' + sfx]))
    model.fused = fused
    model = model.to(DEV)
    d = Data(edge_index=ei)
    d.edge_type = et
    d = d.to(DEV)
    bce = str(g['loss_kind']) == 'bce'
    out = model(d, torch.sigmoid if bce else (lambda t: t))
    xt, yt = torch.from_numpy(g['x_train']).to(DEV), torch.from_numpy(g['y_train']).to(DEV)
    loss = (bce_loss if bce else ce_loss)(out[xt], yt)
    loss.backward()
    want_out = torch.from_numpy(g['out'])
    assert rel_err(out, want_out) < TOL
    assert torch.equal(out.argmax(1).cpu(), want_out.argmax(1)) or \
        float((out.argmax(1).cpu() != want_out.argmax(1)).float().mean()) == 0.0
    assert abs(loss.item() - float(g['loss'])) < TOL * max(1.0, abs(float(g['loss'])))
    for pname, key in (('embedding.weight', 'g_emb'), ('rgcn1.weight', 'g_w1'), ('rgcn1.root', 'g_r1'), ('rgcn1.bias', 'g_b1'),
                       ('rgcn2.weight', 'g_w2'), ('rgcn2.root', 'g_r2'), ('rgcn2.bias', 'g_b2')):
        got = dict(model.named_parameters())[pname].grad
        assert rel_err(got, torch.from_numpy(g[key])) < TOL, (pname, rel_err(got, torch.from_numpy(g[key])))


def test_transfer_heads_match_reference_golden():
    ei, et, n, r = golden_graph('TEST_complete')
    g = load_golden('heads_TEST.npz')
    S, emb, hid, C = int(g['S']), int(g['emb']), int(g['hidden']), int(g['C'])
    d = Data(edge_index=ei)
    d.edge_type = et
    d = d.to(DEV)
    mlp = Emb_MLP_Layers(r, hid, C, n, emb, S)
    mlp.load_embedding(torch.from_numpy(g['e_cat']), freeze=True)
    mlp.load_state_dict({k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('mlp.')})
    out = mlp.to(DEV)(d, torch.sigmoid)
    assert rel_err(out, torch.from_numpy(g['out_mlp'])) < TOL
    att = Emb_ATT_Layers(r, hid, C, n, emb, S)
    att.load_embedding(torch.from_numpy(g['e_stack']), freeze=True)
    att.load_state_dict({k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('att.')})
    att.eval()
    out = att.to(DEV)(d, torch.sigmoid)
    assert rel_err(out, torch.from_numpy(g['out_att'])) < TOL


# ------------------------------------------------------------------ K5 map gather: bit-exact
@pytest.mark.parametrize('tag', ['TEST_attr', 'AIFB_attr', 'AIFB_bisim'])
def test_map_gather_bit_exact(tag):
    g = load_golden(f'mapgather_{tag}.npz')
    s = int(g['num_sums'])
    embs = [torch.from_numpy(g[f'emb{i}']).to(DEV) for i in range(s)]
    idxs = [torch.from_numpy(g[f'idx{i}']).to(DEV) for i in range(s)]
    fbs = [torch.from_numpy(g[f'fallback{i}']).to(DEV) for i in range(s)]
    for mode, key in ((0, 'out_sum'), (1, 'out_concat'), (2, 'out_stack')):
        got = map_gather(embs, idxs, fbs, mode).cpu()
        assert torch.equal(got, torch.from_numpy(g[key])), key


# ------------------------------------------------------------------ larger sizes: properties
@pytest.fixture(scope='module')
def am16():
    ei, et, n, r = am_shape(scale=1 / 16)
    return ei.to(DEV), et.to(DEV), n, r, RGCNGraph(ei.to(DEV), et.to(DEV), n, r)


def test_am_shape_sixteenth_matches_oracle_forward(am16):
    ei, et, n, r, g = am16
    torch.manual_seed(0)
    x = torch.randn(n, 63)
    w = (torch.rand(r, 63, 16) - 0.5) * 0.3
    root = (torch.rand(63, 16) - 0.5) * 0.3
    bias = torch.rand(16)
    with torch.no_grad():
        ref = rgcn_oracle.rgcn_forward(x, ei.cpu(), et.cpu(), w, root, bias)
        out = rgcn_layer(x.to(DEV), w.to(DEV), root.to(DEV), bias.to(DEV), g)
    assert rel_err(out, ref) < TOL


def test_am_shape_known_answer_constant_features(am16):
    """x = ones  =>  every segment mean is ones  =>  out[i] = sum_{r present at i} colsum(W_r) + colsum(root) + bias."""
    ei, et, n, r, g = am16
    torch.manual_seed(1)
    w = (torch.rand(r, 63, 16, device=DEV) - 0.5)
    root = torch.rand(63, 16, device=DEV)
    bias = torch.rand(16, device=DEV)
    out = rgcn_layer(torch.ones(n, 63, device=DEV), w, root, bias, g)
    pairs = torch.unique(ei[1] * r + et)                       # distinct (dst, rel)
    want = torch.zeros(n, 16, device=DEV, dtype=torch.float64)
    want.index_add_(0, pairs // r, w.sum(1).double()[pairs % r])
    want += root.sum(0).double() + bias.double()
    assert rel_err(out, want) < TOL


def test_am_shape_linearity_and_backward_adjointness(am16):
    """<gout, F(x)> == <F^T(gout), x> for the bias-free linear map x -> out (fwd vs dL/dx kernels), and
    dL/dW is the matching adjoint in W — ties K1, K3 and K4 together at a size the oracle cannot reach."""
    ei, et, n, r, g = am16
    torch.manual_seed(2)
    x = torch.randn(n, 63, device=DEV, requires_grad=True)
    w = ((torch.rand(r, 63, 16, device=DEV) - 0.5) * 0.3).requires_grad_()
    root = ((torch.rand(63, 16, device=DEV) - 0.5) * 0.3).requires_grad_()
    gout = torch.randn(n, 16, device=DEV)
    out = rgcn_layer(x, w, root, None, g)
    out.backward(gout)
    lhs = float((gout.double() * out.detach().double()).sum())
    rhs_x = float((x.grad.double() * x.detach().double()).sum())
    rhs_w = float((w.grad.double() * w.detach().double()).sum() + (root.grad.double() * root.detach().double()).sum())
    assert abs(lhs - rhs_x) < 1e-5 * abs(lhs) + 1e-3
    assert abs(lhs - rhs_w) < 1e-5 * abs(lhs) + 1e-3
    out2 = rgcn_layer(2.0 * x.detach(), w.detach(), root.detach(), None, g)
    assert rel_err(out2, 2.0 * out.detach().cpu()) < TOL


def test_am_shape_tensor_and_generic_kernels_agree(am16):
    ei, et, n, r, g = am16
    torch.manual_seed(3)
    x = torch.randn(n, 16, device=DEV, requires_grad=True)
    w = ((torch.rand(r, 16, 11, device=DEV) - 0.5) * 0.3).requires_grad_()
    root = torch.rand(16, 11, device=DEV, requires_grad=True)
    bias = torch.rand(11, device=DEV, requires_grad=True)
    gout = torch.randn(n, 11, device=DEV)
    res = []
    for simple in (False, True):
        for t in (x, w, root, bias):
            t.grad = None
        out = rgcn_layer(x, w, root, bias, g, force_simple=simple)
        out.backward(gout)
        res.append([out.detach().cpu()] + [t.grad.cpu() for t in (x, w, root, bias)])
    # out and dL/dx at the parity tolerance; the parameter gradients of the GENERIC path are sums of
    # ~1e5-1e6 fp32 atomic adds into single floats (round-off ~1e-5 of the result), so that side is
    # the loose one here — the tensor path itself is checked against the oracle elsewhere
    for i, (a, b) in enumerate(zip(*res)):
        assert rel_err(a, b) < (TOL if i < 2 else 1e-4), i


# ------------------------------------------------------------------ destination-partitioned graphs (1 GPU emulation)
class _LocalComm:
    """Stands in for rgcn_b200.partition.RowComm on one GPU: the 'other ranks' rows come from a
    table filled by running the ranks one after another (no concurrent waiting kernels)."""

    def __init__(self, chunk, world):
        self.chunk, self.world, self.table, self.reduced = chunk, world, {}, []

    def set_full(self, key_rows, full):
        self.table[key_rows] = full

    def all_gather_rows(self, t, key=None):
        # the layer gathers the zero-padded 16-byte aligned mirror of odd-width rows (63 -> 64)
        full = self.table.get(t.size(1), self.table.get(t.size(1) - 1))
        out = full.new_zeros((self.world * self.chunk, t.size(1)))
        out[:full.size(0), :full.size(1)] = full
        return out

    def all_reduce_sum_(self, tensors):
        self.reduced.append([t.clone() for t in tensors])

    # source-partitioned ("push") layers: the rank's partial over all nodes is kept for the test to sum; what
    # comes back are this rank's rows of the true total (supplied by the test), so autograd continues from
    # exactly what a real reduce-scatter would deliver
    def partial_buffer(self, cols, key=None, device=None):
        return torch.zeros((self.world * self.chunk, cols), dtype=torch.float32, device=device)

    def reduce_scatter_rows(self, t, key=None):
        self.partials.append(t.detach().clone())
        lo = self.rank * self.chunk
        out = t.new_zeros((self.chunk, t.size(1)))
        rows = self.total[lo:lo + self.chunk]
        out[:rows.size(0), :rows.size(1)] = rows
        return out


@pytest.mark.parametrize('world', [2, 3])
def test_partitioned_graph_matches_unpartitioned(world):
    from rgcn_b200.partition import plan_ranges
    ei, et, n, r = golden_graph('AIFB_bisim_k3')
    eid, etd = ei.to(DEV), et.to(DEV)
    chunk, ranges = plan_ranges(n, world)
    torch.manual_seed(0)
    x = torch.randn(n, 63, device=DEV)
    w = ((torch.rand(r, 63, 16, device=DEV) - 0.5) * 0.3)
    root = ((torch.rand(63, 16, device=DEV) - 0.5) * 0.3)
    bias = torch.rand(16, device=DEV)
    gout = torch.randn(n, 16, device=DEV)
    g_full = RGCNGraph(eid, etd, n, r)
    leaves = [t.clone().requires_grad_() for t in (x, w, root, bias)]
    ref = rgcn_layer(*leaves, g_full)
    ref.backward(gout)
    outs, gxs, gws, groots, gbiases = [], [], 0, 0, 0
    for lo, hi in ranges:
        # bit-exact structure of this rank's partition
        g = RGCNGraph(eid, etd, n, r, own_range=(lo, hi), range_nodes=64, split_threshold=8, chunk_size=4)
        fwd, bwd = csr_oracle.build_graph(ei[0].numpy(), ei[1].numpy(), et.numpy(), n, r, 64, 8, 4, lo=lo, hi=hi)
        _check_brc(g, _lib.BRC_FWD, fwd)
        _check_brc(g, _lib.BRC_BWD, bwd)
        comm = _LocalComm(chunk, world)
        comm.set_full(63, x)
        comm.set_full(16, gout)
        lv = [x[lo:hi].clone().requires_grad_()] + [t.clone().requires_grad_() for t in (w, root, bias)]
        out = rgcn_layer(*lv, g, comm=comm)
        out.backward(gout[lo:hi])
        outs.append(out.detach())
        gxs.append(lv[0].grad)
        gw, gr, gb = comm.reduced[0]
        gws, groots, gbiases = gws + gw, groots + gr, gbiases + gb
    assert rel_err(torch.cat(outs), ref.detach().cpu()) < TOL
    assert rel_err(torch.cat(gxs), leaves[0].grad.cpu()) < TOL
    assert rel_err(gws, leaves[1].grad.cpu()) < TOL
    assert rel_err(groots, leaves[2].grad.cpu()) < TOL
    assert rel_err(gbiases, leaves[3].grad.cpu()) < TOL


@pytest.mark.parametrize('world', [2, 3])
@pytest.mark.parametrize('fin,fout,relu', [(63, 16, False), (16, 11, True)])
def test_source_partitioned_graph_matches_unpartitioned(world, fin, fout, relu):
    """rgcn_graph_create_push: per-rank structures bit-exact against the oracle; the ranks' PARTIAL outputs sum
    to the unpartitioned layer's output (what the reduce-scatter computes); gx rows and the summed parameter
    gradients match.  x never leaves its rank."""
    from rgcn_b200.partition import plan_ranges
    ei, et, n, r = golden_graph('AIFB_bisim_k3')
    eid, etd = ei.to(DEV), et.to(DEV)
    chunk, ranges = plan_ranges(n, world)
    torch.manual_seed(1)
    x = torch.randn(n, fin, device=DEV)
    w = ((torch.rand(r, fin, fout, device=DEV) - 0.5) * 0.3)
    root = ((torch.rand(fin, fout, device=DEV) - 0.5) * 0.3)
    bias = torch.rand(fout, device=DEV)
    gout = torch.randn(n, fout, device=DEV)
    g_full = RGCNGraph(eid, etd, n, r)
    leaves = [t.clone().requires_grad_() for t in (x, w, root, bias)]
    ref = rgcn_layer(*leaves, g_full, relu_in=relu)
    ref.backward(gout)
    total = torch.zeros(world * chunk, fout, device=DEV)
    gxs, gws, groots, gbiases = [], 0, 0, 0
    for rank, (lo, hi) in enumerate(ranges):
        g = RGCNGraph(eid, etd, n, r, own_range=(lo, hi), range_nodes=64, split_threshold=8, chunk_size=4, push=True)
        fwd, bwd = csr_oracle.build_graph(ei[0].numpy(), ei[1].numpy(), et.numpy(), n, r, 64, 8, 4, lo=lo, hi=hi, push=True)
        _check_brc(g, _lib.BRC_FWD, fwd)
        _check_brc(g, _lib.BRC_BWD, bwd)
        comm = _LocalComm(chunk, world)
        comm.rank, comm.partials, comm.total = rank, [], ref.detach()
        comm.set_full(fout, gout)
        lv = [x[lo:hi].clone().requires_grad_()] + [t.clone().requires_grad_() for t in (w, root, bias)]
        out = rgcn_layer(*lv, g, relu_in=relu, comm=comm)
        assert tuple(out.shape) == (hi - lo, fout)
        out.backward(gout[lo:hi])
        part = comm.partials[0]
        total[:, :] += part[:, :fout]
        gxs.append(lv[0].grad)
        gw, gr, gb = comm.reduced[0]
        gws, groots, gbiases = gws + gw, groots + gr, gbiases + gb
    assert rel_err(total[:n], ref.detach().cpu()) < TOL
    assert float(total[n:].abs().max()) == 0.0 if total.size(0) > n else True
    assert rel_err(torch.cat(gxs), leaves[0].grad.cpu()) < TOL
    assert rel_err(gws, leaves[1].grad.cpu()) < TOL
    assert rel_err(groots, leaves[2].grad.cpu()) < TOL
    assert rel_err(gbiases, leaves[3].grad.cpu()) < TOL


# ------------------------------------------------------------------ device-side evaluate
@pytest.mark.parametrize('sigmoid_path', [False, True])
def test_eval_counts_match_sklearn(sigmoid_path):
    """The reference's evaluate (model/evaluation.py:14-31) restated on CPU with sklearn vs the device counts."""
    from sklearn.metrics import accuracy_score, f1_score
    from rgcn_b200.evaluation import confusion_counts, metrics_from_counts
    torch.manual_seed(5)
    n, c, m = 500, 7, 180
    logits = torch.randn(n, c)
    logits[3] = 0.5                      # ties: argmax takes the first index; round(0.5) -> 0 (half to even)
    x = torch.randperm(n)[:m]
    if sigmoid_path:
        out = torch.sigmoid(logits)
        out[5] = 0.5
        y = (torch.rand(m, c) < 0.3).long()
        pred_ref = torch.round(out).type(torch.int64)
    else:
        out = logits
        y = torch.nn.functional.one_hot(torch.randint(0, c, (m,)), c)
        y[::9, 0] = 1                    # multi-hot rows occur in the reference's labels
        a = torch.softmax(out, 1).argmax(1)
        pred_ref = torch.zeros(out.shape).scatter(1, a.unsqueeze(1), 1.0)
    want = (accuracy_score(y, pred_ref[x]), f1_score(y, pred_ref[x], average='weighted', zero_division=0),
            f1_score(y, pred_ref[x], average='macro', zero_division=0))
    counts = confusion_counts(out.to(DEV), x.to(DEV), y.to(DEV), sigmoid_path)
    got = metrics_from_counts(counts, m)
    for g_, w_ in zip(got, want):
        assert abs(g_ - w_) < 1e-12


# ------------------------------------------------------------------ fused Adam
def test_fused_adam_matches_torch_adam():
    from rgcn_b200 import FusedAdam
    torch.manual_seed(9)
    shapes = [(1000, 63), (89, 63, 16), (16,), (7,)]
    pa = [torch.randn(s, device=DEV).requires_grad_() for s in shapes]
    pb = [p.detach().clone().requires_grad_() for p in pa]
    oa = torch.optim.Adam(pa, lr=0.01, weight_decay=5e-5)
    ob = FusedAdam(pb, lr=0.01, weight_decay=5e-5)
    for step in range(25):
        gs = [torch.randn_like(p) * (0.1 + step) for p in pa]
        for p, q, g_ in zip(pa, pb, gs):
            p.grad = g_.clone()
            q.grad = g_.clone()
        oa.step()
        ob.step()
    for p, q in zip(pa, pb):
        assert rel_err(q, p.detach().cpu()) < 1e-6
    cpu_param = torch.zeros(3, requires_grad=True)       # no CPU path: fails loudly
    o = FusedAdam([cpu_param])
    cpu_param.grad = torch.ones(3)
    with pytest.raises(_lib.EngineError):
        o.step()


# ------------------------------------------------------------------ tcgen05 entry-tile pass (hidden 64)
@pytest.mark.parametrize('name', ['AIFB_sum_in', 'AIFB_bisim_k3'])
@pytest.mark.parametrize('fin,fout', [(63, 64), (64, 64)])
def test_hidden64_layer_on_reference_graphs(name, fin, fout):
    """64 x 64 column passes (forward and dL/dx) run on the tcgen05 kernel: hubs (chunk rows), duplicates, tiles of
    every fill from 1 to 128 entries, relations with a single entry."""
    ei, et, n, r = golden_graph(name)
    _layer_case(ei, et, n, r, fin, fout, seed=fin * 3 + fout)


def test_hidden64_am_shape_two_layers_match_mma_sync_path(am16):
    """config-5 shape (63 -> 64 -> 11, fused ReLU, frozen weights: dL/dx0 only) on the AM-shape graph at 1/16:
    the tcgen05 passes against the same step with RGCN_B200_TC=0 semantics (generic kernels as the cross-check)."""
    ei, et, n, r, g = am16
    torch.manual_seed(9)
    x = torch.randn(n, 63, device=DEV, requires_grad=True)
    w1 = (torch.rand(r, 63, 64, device=DEV) - 0.5) * 0.3
    r1 = (torch.rand(63, 64, device=DEV) - 0.5) * 0.3
    b1 = torch.rand(64, device=DEV) - 0.5
    w2 = (torch.rand(r, 64, 11, device=DEV) - 0.5) * 0.3
    r2 = (torch.rand(64, 11, device=DEV) - 0.5) * 0.3
    gout = torch.randn(n, 11, device=DEV)
    res = []
    for simple in (False, True):
        x.grad = None
        h = rgcn_layer(x, w1, r1, b1, g, force_simple=simple)
        out = rgcn_layer(h, w2, r2, None, g, relu_in=True, force_simple=simple)
        out.backward(gout)
        res.append((out.detach().cpu(), x.grad.detach().cpu().clone()))
    assert rel_err(res[0][0], res[1][0]) < TOL
    assert rel_err(res[0][1], res[1][1]) < TOL
