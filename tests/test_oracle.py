"""CPU tests of the oracle itself (no GPU): independent dense formulation, fp64 gradcheck,
golden fixtures produced by the reference's own callers, CSR oracle properties."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ORACLE, REFERENCE, REPO, golden_graph, load_golden, rel_err

import csr_oracle
import rgcn_oracle
from rgcn_b200.synthetic import am_shape, random_multigraph


def _rand_params(fin, fout, R, dtype=torch.float32, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(R, fin, fout, generator=g, dtype=dtype) * 0.2, torch.randn(fin, fout, generator=g, dtype=dtype) * 0.2,
            torch.randn(fout, generator=g, dtype=dtype) * 0.1)


@pytest.mark.parametrize('n,e,r', [(7, 30, 5), (12, 14, 5), (20, 200, 9), (5, 0, 3)])
def test_oracle_matches_dense_formulation(n, e, r):
    ei, et = random_multigraph(n, e, r, seed=n + e, hub_frac=0.3, dup_frac=0.2)
    x = torch.randn(n, 6, dtype=torch.float64)
    w, root, b = _rand_params(6, 4, r, torch.float64)
    a = rgcn_oracle.rgcn_forward(x, ei, et, w, root, b)
    d = rgcn_oracle.dense_rgcn_forward(x, ei, et, w, root, b)
    assert torch.allclose(a, d, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('name', ['AIFB_bisim_k3', 'AIFB_sum_in_out', 'MUTAG_bisim_k1'])
def test_oracle_equals_normalised_adjacency_formulation_on_real_graphs(name):
    """An independent statement of the layer at the size of the reference's real summary graphs (the dense check
    above stops at 20 nodes): the R-GCN paper's  h_i = W_0 x_i + b + sum_r sum_{j in N_r(i)} (1 / c_{i,r}) W_r x_j
    written as one row-normalised sparse adjacency per relation (multi-edges counted with their multiplicity, as a
    mean over edges does), in fp64, forward and every gradient — no index_select / scatter in common with the
    oracle's loop path."""
    ei, et, n, r = golden_graph(name)
    torch.manual_seed(5)
    fin, fout = 7, 5
    x = torch.randn(n, fin, dtype=torch.float64)
    w, root, b = _rand_params(fin, fout, r, torch.float64, seed=9)
    gout = torch.randn(n, fout, dtype=torch.float64)
    leaves_a = [t.clone().requires_grad_() for t in (x, w, root, b)]
    leaves_b = [t.clone().requires_grad_() for t in (x, w, root, b)]
    out_a = rgcn_oracle.rgcn_forward(leaves_a[0], ei, et, *leaves_a[1:])
    xb, wb, rootb, bb = leaves_b
    out_b = xb @ rootb + bb
    for rel in range(r):
        m = et == rel
        if not bool(m.any()):
            continue
        src, dst = ei[0][m], ei[1][m]
        deg = torch.bincount(dst, minlength=n).to(torch.float64)            # c_{i,r}: edges of relation r into i
        adj = torch.sparse_coo_tensor(torch.stack([dst, src]), 1.0 / deg[dst], (n, n)).coalesce()   # duplicates add up
        out_b = out_b + torch.sparse.mm(adj, xb) @ wb[rel]
    assert torch.allclose(out_a, out_b, rtol=1e-11, atol=1e-11)
    out_a.backward(gout)
    out_b.backward(gout)
    for a, bb_ in zip(leaves_a, leaves_b):
        assert torch.allclose(a.grad, bb_.grad, rtol=1e-10, atol=1e-10)


def test_oracle_gradcheck_fp64():
    ei, et = random_multigraph(6, 25, 4, seed=3, hub_frac=0.4, dup_frac=0.3)
    x = torch.randn(6, 5, dtype=torch.float64, requires_grad=True)
    w, root, b = [t.requires_grad_() for t in _rand_params(5, 3, 4, torch.float64)]
    assert torch.autograd.gradcheck(lambda *a: rgcn_oracle.rgcn_forward(a[0], ei, et, a[1], a[2], a[3]), (x, w, root, b))


def test_oracle_semantics_edge_cases():
    # isolated node gets x.root + bias; empty relation contributes 0; duplicates weight the mean
    ei = torch.tensor([[0, 0, 1], [2, 2, 2]])
    et = torch.tensor([0, 0, 0])
    x = torch.tensor([[1.0], [4.0], [10.0], [7.0]])
    w = torch.ones(2, 1, 1)
    out = rgcn_oracle.rgcn_forward(x, ei, et, w, torch.full((1, 1), 2.0), torch.tensor([0.5]))
    assert torch.allclose(out[:, 0], torch.tensor([2.5, 8.5, 20.5 + (1 + 1 + 4) / 3, 14.5]))


def test_basis_decomposition_equals_expanded():
    ei, et = random_multigraph(9, 40, 5, seed=5)
    conv = rgcn_oracle.RGCNConv(4, 3, 5, num_bases=2)
    x = torch.randn(9, 4)
    w_full = (conv.comp @ conv.weight.view(2, -1)).view(5, 4, 3)
    assert torch.allclose(conv(x, ei, et), rgcn_oracle.rgcn_forward(x, ei, et, w_full, conv.root, conv.bias))


@pytest.mark.parametrize('name', ['TEST_complete', 'AIFB_sum_in', 'AIFB_sum_in_out', 'MUTAG_bisim_k1', 'AIFB_bisim_k3'])
def test_oracle_reproduces_reference_caller_goldens(name):
    """layers_*.npz were produced by the reference's own Emb_Layers / losses on the oracle shim."""
    ei, et, n, r = golden_graph(name)
    g = load_golden(f'layers_{name}.npz')
    emb = torch.from_numpy(g['emb']).requires_grad_()
    p = {k: torch.from_numpy(g[k]).requires_grad_() for k in ('w1', 'r1', 'b1', 'w2', 'r2', 'b2')}
    h1 = rgcn_oracle.rgcn_forward(emb, ei, et, p['w1'], p['r1'], p['b1'])
    assert rel_err(h1, torch.from_numpy(g['h1'])) < 1e-6
    out = rgcn_oracle.rgcn_forward(torch.relu(h1), ei, et, p['w2'], p['r2'], p['b2'])
    xt, yt = torch.from_numpy(g['x_train']), torch.from_numpy(g['y_train'])
    if str(g['loss_kind']) == 'bce':
        out = torch.sigmoid(out)
        loss = torch.nn.functional.binary_cross_entropy(out[xt], yt)
    else:
        loss = torch.nn.functional.cross_entropy(out[xt], yt.argmax(-1))
    assert rel_err(out, torch.from_numpy(g['out'])) < 1e-6
    assert abs(loss.item() - float(g['loss'])) < 1e-6 * max(1.0, abs(float(g['loss'])))
    loss.backward()
    assert rel_err(emb.grad, torch.from_numpy(g['g_emb'])) < 1e-5
    assert rel_err(p['w1'].grad, torch.from_numpy(g['g_w1'])) < 1e-5
    assert rel_err(p['r2'].grad, torch.from_numpy(g['g_r2'])) < 1e-5


@pytest.mark.parametrize('tag', ['TEST_attr', 'AIFB_attr', 'AIFB_bisim'])
def test_map_gather_oracle_vs_reference_golden(tag):
    g = load_golden(f'mapgather_{tag}.npz')
    s = int(g['num_sums'])
    embs = [torch.from_numpy(g[f'emb{i}']) for i in range(s)]
    idxs = [torch.from_numpy(g[f'idx{i}']).long() for i in range(s)]
    fbs = [torch.from_numpy(g[f'fallback{i}']) for i in range(s)]
    for mode in ('sum', 'concat', 'stack'):
        got = rgcn_oracle.map_gather(embs, idxs, fbs, mode)
        assert torch.equal(got, torch.from_numpy(g[f'out_{mode}'])), mode
    if tag == 'AIFB_attr':
        assert any(int((i < 0).sum()) > 0 for i in idxs)   # fallback rows are exercised


@pytest.mark.parametrize('nr,t,ch', [(10 ** 9, 10 ** 9, 256), (64, 8, 4), (7, 3, 2), (1, 1, 1)])
def test_csr_oracle_is_a_permutation_of_the_reference_edges(nr, t, ch):
    ei, et, n, r = golden_graph('AIFB_sum_in_out')
    src, dst, rel = ei[0].numpy(), ei[1].numpy(), et.numpy()
    fwd, bwd = csr_oracle.build_graph(src, dst, rel, n, r, nr, t, ch)
    want = np.stack([dst, src, rel], 1)
    got = csr_oracle.edges_from_brc(fwd, n)
    got = got[got[:, 2] < r]
    assert np.array_equal(want[np.lexsort(want.T[::-1])], got[np.lexsort(got.T[::-1])])
    wantb = np.stack([src, dst, rel], 1)
    gotb = csr_oracle.edges_from_brc(bwd, n)
    gotb = gotb[gotb[:, 2] < r]
    assert np.array_equal(wantb[np.lexsort(wantb.T[::-1])], gotb[np.lexsort(gotb.T[::-1])])
    # stable: within a segment entries keep the reference's edge order
    seg_of = np.repeat(np.arange(fwd['num_seg']), fwd['cnt'])
    for s in np.flatnonzero(fwd['cnt'] > 1)[:50]:
        ids = fwd['perm'][seg_of == s]
        assert np.all(np.diff(ids) > 0)
    # weights are 1/multiplicity of the forward (relation, dst) pair, shared by both directions
    assert np.array_equal(fwd['raw_w'], (np.float32(1) / fwd['cnt'].astype(np.float32))[seg_of])
    assert np.array_equal(np.sort(fwd['w_entry']), np.sort(bwd['raw_w']))
    # chunk rows tile the long segments exactly
    if fwd['num_chunks']:
        assert np.all(fwd['chunk_end'] - fwd['chunk_beg'] <= ch) and np.all(fwd['chunk_end'] > fwd['chunk_beg'])
        assert int((fwd['chunk_end'] - fwd['chunk_beg']).sum()) == int(fwd['cnt'][fwd['cnt'] > t].sum())
    # batches: <= 16 segments of one relation, covering every segment once
    ns = fwd['bat_info'] & 0xff
    assert ns.min() >= 1 and ns.max() <= 16 and int(ns.sum()) == fwd['num_seg']
    for b in range(0, fwd['num_batches'], max(1, fwd['num_batches'] // 40)):
        s0 = fwd['bat_seg0'][b]
        assert np.all(fwd['seg_rel'][s0:s0 + ns[b]] == (fwd['bat_info'][b] >> 8))


def test_synthetic_am_shape_layout():
    ei, et, n, r = am_shape(scale=0.002)
    assert r == 267 and ei.shape[0] == 2 and ei.stride() == (1, 3) and et.stride() == (3,)
    assert int(et.max()) < r - 1                       # slot 2|rel| is never used (modelTrainer.py:78)
    assert torch.equal(ei[0, 0::2], ei[1, 1::2]) and torch.equal(et[0::2] + 1, et[1::2])   # inverse edges


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='/root/reference only exists in the build container')
def test_unmodified_reference_runs_end_to_end_on_the_oracle_shim(tmp_path):
    """python main.py -dataset TEST -sum attr -exp summation, reference code untouched."""
    work = tmp_path / 'work'
    work.mkdir()
    os.symlink(os.path.join(REFERENCE, 'graphs'), work / 'graphs')
    (work / 'results').mkdir()
    base = work / 'baselines' / 'TEST_baseline'
    base.mkdir(parents=True)
    import json
    curve = [[0.0] * 51, [0.0] * 51, [0.0] * 51]
    (base / 'run_results_baseline_i=5.json').write_text(json.dumps(
        {'baseline': {m: curve for m in ('accuracy', 'loss', 'f1 weighted', 'f1 macro')}}))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ORACLE, 'shim'), ORACLE, REFERENCE]))
    res = subprocess.run([sys.executable, os.path.join(REFERENCE, 'main.py'), '-dataset', 'TEST', '-sum', 'attr',
                          '-exp', 'summation', '-epochs', '3'], cwd=work, env=env, capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert 'ACC ON TEST SET' in res.stdout
