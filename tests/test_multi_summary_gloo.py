"""World-size-2 gloo test (CPU) of the one-summary-graph-per-rank pre-training (BASELINE.json configs[2],
rgcn_b200.trainer.train_summaries_parallel): host logic only — the layer arithmetic is supplied by the ORACLE
conv here because the engine has no CPU path.  Checks that (a) the shared R-GCN weights stay identical on all ranks,
(b) they equal a one-process emulation that averages the two graphs' gradients by hand, (c) each rank keeps its
own embedding, (d) an odd number of graphs (idle rank in the last round) is handled."""
import os
import sys
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from conftest import ORACLE, PKG, REPO, golden_graph


def _graphs(n_graphs, classes):
    from rgcn_b200 import Data
    out = []
    for k, name in enumerate(['AIFB_sum_in', 'MUTAG_bisim_k1', 'TEST_complete'][:n_graphs]):
        ei, et, n, r = golden_graph(name)
        r = 9 if name != 'AIFB_sum_in' else r
        g = types.SimpleNamespace(num_nodes=n, relations={f'p{i}': i for i in range((89 - 1) // 2)})
        keep = et < 88
        td = Data(edge_index=ei[:, keep])
        td.edge_type = et[keep]
        gen = torch.Generator().manual_seed(k)
        td.x_train = torch.randperm(n, generator=gen)[:max(2, n // 2)]
        td.y_train = torch.rand(td.x_train.numel(), classes, generator=gen)
        g.training_data = td
        out.append(g)
    return out


def _oracle_layers():
    sys.path.insert(0, ORACLE)
    import rgcn_oracle
    from rgcn_b200 import layers as engine_layers

    class OracleLayers(engine_layers.Emb_Layers):
        def _build_convs(self, num_relations, hidden_l, num_labels, emb_dim):
            self.rgcn1 = rgcn_oracle.RGCNConv(emb_dim, hidden_l, num_relations)
            self.rgcn2 = rgcn_oracle.RGCNConv(hidden_l, num_labels, num_relations)
            for c in (self.rgcn1, self.rgcn2):
                nn.init.kaiming_uniform_(c.weight, mode='fan_in')
            self.fused = False
    return OracleLayers


def _worker(rank, world, port, n_graphs, ret):
    for p in (REPO, PKG, ORACLE, os.path.join(REPO, 'tests')):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from rgcn_b200.trainer import train_summaries_parallel
        res = train_summaries_parallel(_graphs(n_graphs, 4), 4, 8, 3, 12, 0.01, 5e-5, 'cpu', layers_cls=_oracle_layers(),
                                       fused_adam=False, seed=3)
        m = res['model']
        ret[rank] = {'w': [p.detach().clone() for c in (m.rgcn1, m.rgcn2) for p in (c.weight, c.root, c.bias)],
                     'emb': {k: v.clone() for k, v in res['embeddings'].items()}, 'losses': res['losses'],
                     'rounds': res['rounds']}
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_graphs', [2, 3])
def test_parallel_summary_pretraining_world2_gloo(n_graphs):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29631 + n_graphs, n_graphs, ret), nprocs=world, join=True)
    a, b = ret[0], ret[1]
    assert a['rounds'] == (n_graphs + 1) // 2
    for x, y in zip(a['w'], b['w']):
        assert torch.equal(x, y)                                   # same updates on every rank
    assert set(a['emb']) | set(b['emb']) == set(range(n_graphs)) and not (set(a['emb']) & set(b['emb']))
    # one-process emulation of round 0: two models with tied weights, gradients averaged by hand
    from rgcn_b200.trainer import bce_loss
    cls = _oracle_layers()
    graphs = _graphs(n_graphs, 4)
    torch.manual_seed(3)
    base = cls(2 * len(graphs[0].relations) + 1, 8, 4, graphs[0].num_nodes, 12, n_graphs)
    import copy
    models = []
    for gi in range(2):
        m = copy.deepcopy(base)
        torch.manual_seed(3 + 1000 + gi)
        m.reset_embedding(graphs[gi].num_nodes, 12)
        models.append(m)
    opts = [torch.optim.Adam(m.parameters(), lr=0.01, weight_decay=5e-5) for m in models]
    shared = [[p for c in (m.rgcn1, m.rgcn2) for p in (c.weight, c.root, c.bias)] for m in models]
    for _ in range(3):
        for m, o, g in zip(models, opts, graphs):
            o.zero_grad()
            td = g.training_data
            bce_loss(m(td, torch.sigmoid)[td.x_train], td.y_train.float()).backward()
        for p0, p1 in zip(*shared):
            avg = (p0.grad + p1.grad) / 2
            p0.grad.copy_(avg)
            p1.grad.copy_(avg)
        for o in opts:
            o.step()
    if n_graphs == 2:
        for x, y in zip(a['w'], shared[0]):
            assert torch.allclose(x, y.detach(), atol=1e-6), 'weights after the round of averaged steps'
        assert torch.allclose(a['emb'][0], models[0].embedding.weight.detach(), atol=1e-6)
        assert torch.allclose(b['emb'][1], models[1].embedding.weight.detach(), atol=1e-6)
