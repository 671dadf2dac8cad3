"""Generate tests/golden/*.npz by running the reference's OWN, unmodified callers
(/root/reference/graphs/graph.py, graphs/graphProcessing.py, model/layers.py,
model/embeddingTricks.py, model/evaluation.py) on top of the CPU oracle shim
(oracle/shim/torch_geometric -> oracle/rgcn_oracle.py).

Run in the build container only (needs /root/reference):
    PYTHONHASHSEED=0 python tests/golden/make_golden.py
The fixtures travel to the GPU box; /root/reference does not.

What each fixture pins
  graph_<name>.npz    edge_index / edge_type exactly as Graph.init_graph (graphs/graph.py:55-69)
                      emits them (strided views of one [E,3] buffer, inverse edges 2r/2r+1,
                      multi-edges kept), num_nodes, num_relations = 2*|rel|+1 (modelTrainer.py:78)
  layers_<name>.npz   Emb_Layers (model/layers.py:11-25) forward, BCE loss (evaluation.py:39-42)
                      and all parameter gradients for seeded parameters on that graph
  mapgather_*.npz     get_tensor_list + sum/concat/stack (model/embeddingTricks.py:8-49) with the
                      torch.rand fallback captured, plus the integer index the dict walk implies
"""
import os
import sys

if os.environ.get('PYTHONHASHSEED') != '0':
    os.environ['PYTHONHASHSEED'] = '0'
    os.execv(sys.executable, [sys.executable] + sys.argv)

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, os.path.join(REPO, 'oracle', 'shim'))
sys.path.insert(0, os.path.join(REPO, 'oracle'))
sys.path.insert(0, REF)

from graphs.graph import Graph                                    # noqa: E402  (reference)
from graphs.graphProcessing import parse_graph_nt, get_node_mappings_dict  # noqa: E402
from model.layers import Emb_Layers, Emb_MLP_Layers, Emb_ATT_Layers        # noqa: E402
from model import embeddingTricks                                  # noqa: E402
from model.evaluation import bce_loss, ce_loss                     # noqa: E402

GRAPHS = {
    'TEST_complete': 'graphs/TEST/TEST_complete.nt',
    'TEST_sum_in_out': 'graphs/TEST/attr/sum/TEST_sum_in_out.nt',
    'AIFB_sum_in': 'graphs/AIFB/attr/sum/AIFB_sum_in.nt',
    'AIFB_sum_in_out': 'graphs/AIFB/attr/sum/AIFB_sum_in_out.nt',
    'AIFB_bisim_k3': 'graphs/AIFB/bisim/sum/AIFB_bisim_k3.nt',
    'MUTAG_bisim_k1': 'graphs/MUTAG/bisim/sum/MUTAG_bisim_k1.nt',
}


def load_graph(rel_path):
    g = Graph(os.path.basename(rel_path), {})
    g.init_graph(parse_graph_nt(os.path.join(REF, rel_path)))
    return g


def save_graph(name, g):
    ei = g.training_data.edge_index
    et = g.training_data.edge_type
    assert not ei.is_contiguous() or ei.size(1) <= 1   # the reference hands out strided views
    np.savez_compressed(
        os.path.join(HERE, f'graph_{name}.npz'),
        edge_index=ei.numpy().astype(np.int32), edge_type=et.numpy().astype(np.int32),
        num_nodes=np.int64(g.num_nodes), num_relations=np.int64(2 * len(g.relations) + 1),
        ei_stride=np.array(ei.stride(), dtype=np.int64), et_stride=np.array(et.stride(), dtype=np.int64))


def layers_golden(name, g, emb, hidden, classes, seed, loss_kind):
    torch.manual_seed(seed)
    R = 2 * len(g.relations) + 1
    model = Emb_Layers(R, hidden, classes, g.num_nodes, emb, 1)
    with torch.no_grad():                       # non-trivial bias (PyG zero-inits it)
        model.rgcn1.bias.uniform_(-0.1, 0.1)
        model.rgcn2.bias.uniform_(-0.1, 0.1)
    n_lab = max(1, g.num_nodes // 2)
    x_train = torch.randperm(g.num_nodes)[:n_lab]
    if loss_kind == 'bce':
        y = torch.rand(n_lab, classes)
        act, loss_f = torch.sigmoid, bce_loss
    else:
        y = torch.nn.functional.one_hot(torch.randint(0, classes, (n_lab,)), classes).float()
        act, loss_f = (lambda t: t), ce_loss
    td = g.training_data
    out = model(td, act)
    loss = loss_f(out[x_train], y)
    loss.backward()
    # first-layer pre-activation, for per-layer checks
    h1 = model.rgcn1(model.embedding.weight, td.edge_index, td.edge_type)
    np.savez_compressed(
        os.path.join(HERE, f'layers_{name}.npz'),
        loss_kind=loss_kind, emb=model.embedding.weight.detach().numpy(),
        w1=model.rgcn1.weight.detach().numpy(), r1=model.rgcn1.root.detach().numpy(),
        b1=model.rgcn1.bias.detach().numpy(),
        w2=model.rgcn2.weight.detach().numpy(), r2=model.rgcn2.root.detach().numpy(),
        b2=model.rgcn2.bias.detach().numpy(),
        x_train=x_train.numpy(), y_train=y.numpy(),
        h1=h1.detach().numpy(), out=out.detach().numpy(), loss=np.float64(loss.item()),
        g_emb=model.embedding.weight.grad.numpy(),
        g_w1=model.rgcn1.weight.grad.numpy(), g_r1=model.rgcn1.root.grad.numpy(),
        g_b1=model.rgcn1.bias.grad.numpy(),
        g_w2=model.rgcn2.weight.grad.numpy(), g_r2=model.rgcn2.root.grad.numpy(),
        g_b2=model.rgcn2.bias.grad.numpy())


class _FakeOrg:
    """Stands in for the absent original graph: node ids are the sorted
    node set of the map file, exactly how Graph.init_graph enumerates nodes (graph.py:47,53)."""

    def __init__(self, nodes):
        self.nodes = sorted(nodes)
        self.num_nodes = len(self.nodes)
        self.node_to_enum = {n: i for i, n in enumerate(self.nodes)}


def mapgather_golden(tag, sum_paths, map_paths, emb_dim, seed, drop_every=0):
    sum_graphs, org_nodes = [], set()
    for sp, mp in zip(sum_paths, map_paths):
        sg = load_graph(sp)
        sg.orgNode2sumNode_dict, sg.sumNode2orgNode_dict = get_node_mappings_dict(
            parse_graph_nt(os.path.join(REF, mp)))
        org_nodes |= set(sg.orgNode2sumNode_dict.keys())
        sum_graphs.append(sg)
    org = _FakeOrg(org_nodes)
    if drop_every:   # make some org nodes unmapped in summary 0 -> exercises the fallback rows
        for k in list(sum_graphs[0].orgNode2sumNode_dict.keys())[::drop_every]:
            del sum_graphs[0].orgNode2sumNode_dict[k]
    torch.manual_seed(seed)
    for sg in sum_graphs:
        sg.embedding = torch.randn(sg.num_nodes, emb_dim)
    out = {}
    for mode, fn in (('sum', embeddingTricks.sum_embeddings), ('concat', embeddingTricks.concat_embeddings),
                     ('stack', embeddingTricks.stack_embeddings)):
        torch.manual_seed(seed + 1)             # same torch.rand fallback stream for every mode
        out[mode] = fn(org, sum_graphs, emb_dim).numpy()
    torch.manual_seed(seed + 1)
    fallbacks = [torch.rand(org.num_nodes, emb_dim).numpy() for _ in sum_graphs]
    from rgcn_oracle import build_map_index
    idx = [build_map_index(org.node_to_enum, sg.node_to_enum, sg.orgNode2sumNode_dict).numpy().astype(np.int32)
           for sg in sum_graphs]
    payload = {f'emb{i}': sg.embedding.numpy() for i, sg in enumerate(sum_graphs)}
    payload.update({f'idx{i}': a for i, a in enumerate(idx)})
    payload.update({f'fallback{i}': a for i, a in enumerate(fallbacks)})
    payload.update({f'out_{k}': v for k, v in out.items()})
    np.savez_compressed(os.path.join(HERE, f'mapgather_{tag}.npz'), num_sums=np.int64(len(sum_graphs)), **payload)


def heads_golden(seed):
    """Transfer heads (model/layers.py:49-66, 90-112) on the TEST graph: MLP head and the
    attention head (MHA dropout disabled via eval())."""
    g = load_graph(GRAPHS['TEST_complete'])
    R = 2 * len(g.relations) + 1
    S, emb, hidden, C = 3, 12, 16, 4
    torch.manual_seed(seed)
    td = g.training_data
    mlp = Emb_MLP_Layers(R, hidden, C, g.num_nodes, emb, S)
    e_cat = torch.randn(g.num_nodes, S * emb)
    mlp.load_embedding(e_cat, freeze=True)
    out_mlp = mlp(td, torch.sigmoid)
    att = Emb_ATT_Layers(R, hidden, C, g.num_nodes, emb, S)
    e_stack = torch.randn(S, g.num_nodes, emb)
    att.load_embedding(e_stack, freeze=True)
    att.eval()
    out_att = att(td, torch.sigmoid)
    pay = {f'mlp.{k}': v.detach().numpy() for k, v in mlp.state_dict().items()}
    pay.update({f'att.{k}': v.detach().numpy() for k, v in att.state_dict().items()})
    np.savez_compressed(os.path.join(HERE, 'heads_TEST.npz'), e_cat=e_cat.numpy(), e_stack=e_stack.numpy(),
                        att_embedding=att.embedding.detach().numpy(),
                        out_mlp=out_mlp.detach().numpy(), out_att=out_att.detach().numpy(),
                        S=np.int64(S), emb=np.int64(emb), hidden=np.int64(hidden), C=np.int64(C), **pay)


def heads_aifb_golden(seed):
    """Transfer heads at the shape of BASELINE config 2 (S = 3 summaries, emb 63, hidden 16): the real AIFB
    attr summary graphs and their real map files (8243 original nodes) feed the reference's own
    concat_embeddings / stack_embeddings; the original graph itself is absent, so its edges are
    a seeded synthetic multigraph over the real node set with AIFB's counts (24 919 triples, 44 predicates).
    Emb_MLP_Layers / Emb_ATT_Layers (model/layers.py:49-66, 90-112, unmodified) run forward, CE loss and
    backward with the embedding trainable; every parameter gradient is stored (embedding gradient: every
    16th row)."""
    from torch_geometric.data import Data
    tags = ('in', 'in_out', 'out')
    sum_graphs, org_nodes = [], set()
    for t in tags:
        sg = load_graph(f'graphs/AIFB/attr/sum/AIFB_sum_{t}.nt')
        sg.orgNode2sumNode_dict, sg.sumNode2orgNode_dict = get_node_mappings_dict(
            parse_graph_nt(os.path.join(REF, f'graphs/AIFB/attr/map/AIFB_map_{t}.nt')))
        org_nodes |= set(sg.orgNode2sumNode_dict.keys())
        sum_graphs.append(sg)
    org = _FakeOrg(org_nodes)
    for k in list(sum_graphs[1].orgNode2sumNode_dict.keys())[::97]:     # some fallback rows
        del sum_graphs[1].orgNode2sumNode_dict[k]
    n, preds, triples = org.num_nodes, 44, 24919
    rng = np.random.default_rng(seed)
    s_, o_, p_ = rng.integers(0, n, triples), rng.integers(0, n, triples), rng.integers(0, preds, triples)
    buf = np.empty((2 * triples, 3), dtype=np.int64)
    buf[0::2] = np.stack([s_, o_, 2 * p_], 1)
    buf[1::2] = np.stack([o_, s_, 2 * p_ + 1], 1)
    edge = torch.from_numpy(buf).t()
    td = Data(edge_index=edge[:2])
    td.edge_type = edge[2]
    R, S, emb, hidden, C = 2 * preds + 1, 3, 63, 16, 11
    torch.manual_seed(seed)
    for sg in sum_graphs:
        sg.embedding = torch.randn(sg.num_nodes, emb)
    torch.manual_seed(seed + 1)
    e_cat = embeddingTricks.concat_embeddings(org, sum_graphs, emb)
    torch.manual_seed(seed + 1)
    e_stack = embeddingTricks.stack_embeddings(org, sum_graphs, emb)
    torch.manual_seed(seed + 1)
    fallbacks = [torch.rand(n, emb) for _ in sum_graphs]
    from rgcn_oracle import build_map_index
    idx = [build_map_index(org.node_to_enum, sg.node_to_enum, sg.orgNode2sumNode_dict).numpy().astype(np.int32)
           for sg in sum_graphs]
    torch.manual_seed(seed + 2)
    x_train = torch.randperm(n)[:n // 2]
    y = torch.nn.functional.one_hot(torch.randint(0, C, (n // 2,)), C).float()
    pay = {'S': np.int64(S), 'emb': np.int64(emb), 'hidden': np.int64(hidden), 'C': np.int64(C),
           'edge_index': edge[:2].numpy().astype(np.int32), 'edge_type': edge[2].numpy().astype(np.int32),
           'num_nodes': np.int64(n), 'num_relations': np.int64(R), 'x_train': x_train.numpy(), 'y_train': y.numpy()}
    for i, sg in enumerate(sum_graphs):
        pay[f'emb{i}'] = sg.embedding.numpy()
        pay[f'idx{i}'] = idx[i]
        rows = np.nonzero(idx[i] < 0)[0]
        pay[f'fb_rows{i}'] = rows.astype(np.int32)
        pay[f'fb_vals{i}'] = fallbacks[i].numpy()[rows]
    mlp = Emb_MLP_Layers(R, hidden, C, n, emb, S)
    mlp.load_embedding(e_cat.clone(), freeze=False)
    out = mlp(td, lambda t: t)
    loss = ce_loss(out[x_train], y)
    loss.backward()
    pay.update({f'mlp.{k}': v.detach().numpy() for k, v in mlp.state_dict().items() if k != 'embedding.weight'})
    pay.update({f'mlp.grad.{k}': v.grad.numpy() for k, v in mlp.named_parameters() if k != 'embedding.weight'})
    pay['mlp.grad.embedding_rows16'] = mlp.embedding.weight.grad.numpy()[::16]
    pay['mlp.out'], pay['mlp.loss'] = out.detach().numpy(), np.float64(loss.item())
    att = Emb_ATT_Layers(R, hidden, C, n, emb, S)
    att.load_embedding(e_stack.clone(), freeze=False)
    att.eval()                                  # MHA dropout off; gradients still flow
    out = att(td, lambda t: t)
    loss = ce_loss(out[x_train], y)
    loss.backward()
    pay.update({f'att.{k}': v.detach().numpy() for k, v in att.state_dict().items() if k != 'embedding'})
    pay.update({f'att.grad.{k}': v.grad.numpy() for k, v in att.named_parameters() if k != 'embedding'})
    pay['att.grad.embedding_rows16'] = att.embedding.grad.numpy()[:, ::16]
    pay['att.out'], pay['att.loss'] = out.detach().numpy(), np.float64(loss.item())
    # spot rows of the gathered inputs, so the test's own map-gather is tied to the reference's
    pay['e_cat_rows16'] = e_cat.numpy()[::16]
    np.savez_compressed(os.path.join(HERE, 'heads_AIFB_attr.npz'), **pay)


def main():
    graphs = {k: load_graph(v) for k, v in GRAPHS.items()}
    for k, g in graphs.items():
        save_graph(k, g)
        print(k, 'N', g.num_nodes, 'E', g.training_data.edge_type.numel(), 'R', 2 * len(g.relations) + 1)
    layers_golden('TEST_complete', graphs['TEST_complete'], emb=63, hidden=16, classes=3, seed=1, loss_kind='bce')
    layers_golden('AIFB_sum_in', graphs['AIFB_sum_in'], emb=63, hidden=16, classes=7, seed=2, loss_kind='bce')
    layers_golden('AIFB_sum_in_out', graphs['AIFB_sum_in_out'], emb=63, hidden=16, classes=26, seed=3, loss_kind='bce')
    layers_golden('MUTAG_bisim_k1', graphs['MUTAG_bisim_k1'], emb=64, hidden=16, classes=2, seed=4, loss_kind='ce')
    layers_golden('AIFB_bisim_k3', graphs['AIFB_bisim_k3'], emb=20, hidden=16, classes=11, seed=5, loss_kind='ce')
    mapgather_golden('TEST_attr',
                     [f'graphs/TEST/attr/sum/TEST_sum_{s}.nt' for s in ('in', 'in_out', 'out')],
                     [f'graphs/TEST/attr/map/TEST_map_{s}.nt' for s in ('in', 'in_out', 'out')], emb_dim=63, seed=11)
    mapgather_golden('AIFB_attr',
                     [f'graphs/AIFB/attr/sum/AIFB_sum_{s}.nt' for s in ('in', 'in_out', 'out')],
                     [f'graphs/AIFB/attr/map/AIFB_map_{s}.nt' for s in ('in', 'in_out', 'out')], emb_dim=9, seed=12,
                     drop_every=97)
    mapgather_golden('AIFB_bisim',
                     [f'graphs/AIFB/bisim/sum/AIFB_bisim_k{k}.nt' for k in (1, 2, 3)],
                     [f'graphs/AIFB/bisim/map/AIFB_bisim_map_k{k}.nt' for k in (1, 2, 3)], emb_dim=6, seed=13)
    heads_golden(21)
    heads_aifb_golden(31)
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()
