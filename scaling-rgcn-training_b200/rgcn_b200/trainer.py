"""One training step of the reference's ``Trainer.train`` loop body
(/root/reference/model/modelTrainer.py:61-69) and its two losses (model/evaluation.py:33-42),
mirrored so bench.py and the tests can drive the engine the way the reference does."""
from __future__ import annotations

import copy
from typing import Callable, Tuple

import torch
from torch import Tensor, nn


def ce_loss(pred: Tensor, targets: Tensor) -> Tensor:
    """``nn.CrossEntropyLoss()(pred, targets.argmax(-1))`` (model/evaluation.py:33-36) written as
    mean(logsumexp - picked logit): the same value, but torch's nll_loss forward/backward reduce
    kernels run in a single block (0.17 ms per step at 100 k labelled rows) and these do not."""
    picked = pred.gather(1, targets.argmax(-1, keepdim=True)).squeeze(1)
    return (torch.logsumexp(pred, dim=1) - picked).mean()


def bce_loss(pred: Tensor, targets: Tensor) -> Tensor:
    return nn.functional.binary_cross_entropy(pred, targets)


def identity(x: Tensor) -> Tensor:
    return x


def get_losst(dataset: str, sumModel: bool = False) -> Tuple[Callable, Callable]:
    """(loss, activation) selector, evaluation.py:44-48."""
    if sumModel or dataset == 'AIFB':
        return bce_loss, torch.sigmoid
    return ce_loss, identity


def make_optimizer(model: nn.Module, lr: float = 0.01, weight_decay: float = 5e-5, fused: bool = True,
                   capturable: bool = False):
    """Adam as modelTrainer.py:44 builds it (weight_d = 5e-5 from main.py:52): the engine's one-pass
    FusedAdam (same update rule) by default, torch.optim.Adam with fused=False."""
    if fused:
        from .optim import FusedAdam
        return FusedAdam(model.parameters(), lr=lr, weight_decay=weight_decay, capturable=capturable)
    return torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay)


def train_step(model: nn.Module, training_data, optimizer, loss_f: Callable, activation: Callable) -> float:
    model.train()
    optimizer.zero_grad()
    out = model(training_data, activation)
    targets = training_data.y_train.to(torch.float32)
    loss = loss_f(out[training_data.x_train], targets)
    loss.backward()
    optimizer.step()
    return loss.item()


class GraphedTrainStep:
    """The same iteration body (modelTrainer.py:61-69) captured ONCE in a CUDA graph and replayed:
    one graph launch per step instead of ~60 kernel launches plus the Python between them.

    ``x_train`` / ``y_train`` live in static device buffers that ``__call__`` refills (from pinned
    host memory or device tensors) before each replay; the loss comes back as a device scalar.
    Capture needs one eager step (lazy initialisation, CSR build, optimiser state); the parameters
    and optimiser state are restored afterwards, so step k of the graphed loop equals step k of the
    eager loop.  The optimiser must be a ``FusedAdam(capturable=True)``
    (``make_optimizer(..., capturable=True)``)."""

    def __init__(self, model: nn.Module, training_data, optimizer, loss_f: Callable, activation: Callable) -> None:
        if not getattr(optimizer, 'capturable', False):
            raise ValueError('GraphedTrainStep needs make_optimizer(..., capturable=True)')
        self.model, self.td, self.opt = model, training_data, optimizer
        self.loss_f, self.activation = loss_f, activation
        dev = training_data.x_train.device
        if dev.type != 'cuda':
            raise ValueError('GraphedTrainStep: training data must be on the GPU')
        self._copy_stream, self._next_slot, self._pending = None, 0, []
        self.x_static = training_data.x_train.clone()
        self.y_static = training_data.y_train.clone()
        self._td = copy.copy(training_data)     # shallow: same edge tensors (the CSR cache is keyed on them)
        self._td.x_train, self._td.y_train = self.x_static, self.y_static
        params = [p for p in model.parameters()]
        snapshot = [p.detach().clone() for p in params]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        rng = torch.cuda.get_rng_state(dev)                # dropout (attention head) must see the same stream
        with torch.cuda.stream(side):
            self._body()                                   # eager warm-up on the capture stream
            with torch.no_grad():
                for p, s in zip(params, snapshot):
                    p.copy_(s)
            optimizer.reset_state()
        torch.cuda.set_rng_state(rng, dev)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss_static = self._body()

    def _body(self) -> Tensor:
        self.model.train()
        self.opt.zero_grad(set_to_none=True)
        out = self.model(self._td, self.activation)
        targets = self.y_static.to(torch.float32)
        loss = self.loss_f(out[self.x_static], targets)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def prefetch(self, x_train: Tensor, y_train: Tensor) -> None:
        """Starts the host->device copy of the NEXT step's labelled batch on a copy stream (double-
        buffered staging), so it overlaps the step that is running; the next ``__call__()`` without
        arguments consumes it."""
        dev = self.x_static.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staging = [(torch.empty_like(self.x_static), torch.empty_like(self.y_static)) for _ in range(2)]
            self._events = [torch.cuda.Event(), torch.cuda.Event()]
            self._consumed = [None, None]
        k = self._next_slot
        self._next_slot ^= 1
        if self._consumed[k] is not None:                  # the D2D that last read this slot must be over
            self._copy_stream.wait_event(self._consumed[k])
        with torch.cuda.stream(self._copy_stream):
            self._staging[k][0].copy_(x_train, non_blocking=True)
            self._staging[k][1].copy_(y_train, non_blocking=True)
            self._events[k].record(self._copy_stream)
        self._pending.append(k)

    def __call__(self, x_train: Tensor = None, y_train: Tensor = None) -> Tensor:
        """Replays one step; returns the loss as a device scalar (``.item()`` to read it).  With
        arguments the batch is copied on the current stream first; without, a prefetched batch (if
        any) is used, else the batch already in the static buffers."""
        cur = torch.cuda.current_stream(self.x_static.device)
        if x_train is not None or y_train is not None:
            if x_train is not None:
                self.x_static.copy_(x_train, non_blocking=True)
            if y_train is not None:
                self.y_static.copy_(y_train, non_blocking=True)
        elif self._pending:
            k = self._pending.pop(0)
            cur.wait_event(self._events[k])
            self.x_static.copy_(self._staging[k][0], non_blocking=True)
            self.y_static.copy_(self._staging[k][1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cur)
            self._consumed[k] = ev
        self.graph.replay()
        return self.loss_static


class Trainer:
    """Mirror of the reference's ``Trainer`` (/root/reference/model/modelTrainer.py:14-116) on the
    engine: sequential summary pre-training with carried R-GCN weights (:76-82),
    weight transfer (:26-39), embedding transfer (:94-96) and full-graph training with per-epoch
    validation (:41-74), everything on the device.  ``data`` needs ``sumGraphs``, ``orgGraph`` and
    ``num_classes`` exactly like the reference's ``Dataset``; a graph needs ``training_data``,
    ``num_nodes``, ``relations``, ``node_to_enum`` (and ``orgNode2sumNode_dict`` for summaries).
    """

    def __init__(self, data, hidden_l: int, epochs: int, emb_dim: int, lr: float, weight_d: float,
                 device='cuda:0', fused_adam: bool = True) -> None:
        self.data, self.hidden_l, self.epochs, self.emb_dim = data, hidden_l, epochs, emb_dim
        self.lr, self.weight_d = lr, weight_d
        self.device = torch.device(device)
        self.fused_adam = fused_adam
        self.sumModel = None

    def transfer_weights(self, orgModel: nn.Module, grad: bool) -> None:
        s = self.sumModel
        orgModel.override_params(s.rgcn1.weight.clone(), s.rgcn1.bias.clone(), s.rgcn1.root.clone(),
                                 s.rgcn2.weight.clone(), s.rgcn2.bias.clone(), s.rgcn2.root.clone(), grad)

    def train(self, model: nn.Module, graph, loss_f: Callable, activation: Callable, sum_graph: bool = True):
        from .evaluation import evaluate
        model = model.to(self.device)
        td = graph.training_data.to(self.device)
        opt = make_optimizer(model, self.lr, self.weight_d, fused=self.fused_adam)
        accuracies, losses, f1_ws, f1_ms = [], [], [], []
        for _ in range(self.epochs):
            if not sum_graph:
                model.eval()
                acc, f1_w, f1_m = evaluate(model, activation, td, td.x_val, td.y_val)
                accuracies.append(acc)
                f1_ws.append(f1_w)
                f1_ms.append(f1_m)
            losses.append(train_step(model, td, opt, loss_f, activation))
        return accuracies, losses, f1_ws, f1_ms

    def train_summaries(self, dataset: str = 'AIFB') -> None:
        from .layers import Emb_Layers
        loss_f, activation = get_losst(dataset, sumModel=True)
        first = self.data.sumGraphs[0]
        self.sumModel = Emb_Layers(2 * len(first.relations) + 1, self.hidden_l, self.data.num_classes, first.num_nodes,
                                   self.emb_dim, len(self.data.sumGraphs))
        for sg in self.data.sumGraphs:
            self.sumModel.reset_embedding(sg.num_nodes, self.emb_dim)
            self.train(self.sumModel, sg, loss_f, activation, sum_graph=True)
            sg.embedding = self.sumModel.embedding.weight.clone()

    def train_original(self, org_layers, embedding_trick, configs: dict, exp: str):
        from .evaluation import evaluate
        org = self.data.orgGraph
        model = org_layers(2 * len(org.relations) + 1, self.hidden_l, self.data.num_classes, org.num_nodes,
                           self.emb_dim, configs['num_sums'])
        if exp != 'baseline' and configs['e_trans']:
            emb = embedding_trick(org, self.data.sumGraphs, self.emb_dim, device=self.device)
            model.load_embedding(emb, freeze=configs['e_freeze'])
        if exp != 'baseline' and configs['w_trans']:
            self.transfer_weights(model, configs['w_grad'])
        loss_f, activation = get_losst(configs['dataset'], sumModel=False)
        acc, loss, f1_w, f1_m = self.train(model, org, loss_f, activation, sum_graph=False)
        td = org.training_data
        test = evaluate(model, activation, td, td.x_test, td.y_test)
        return acc, loss, f1_w, f1_m, test[0], test[1], test[2], model


PARALLEL_SEMANTICS = ('parallel multi-summary pre-training: one summary graph per rank and round, every step the 6 R-GCN '
                      'weight gradients are averaged over the ranks (one flat all-reduce), embeddings stay local; the '
                      'reference trains the summaries ONE AFTER ANOTHER with carried weights and a fresh Adam each '
                      '(model/modelTrainer.py:76-82) — same per-graph step arithmetic, different weight trajectory')


def train_summaries_parallel(sum_graphs: list, num_classes: int, hidden_l: int, epochs: int, emb_dim: int, lr: float,
                             weight_d: float, device, layers_cls=None, group=None, fused_adam: bool = True,
                             seed: int = 0) -> dict:
    """BASELINE.json configs[2]: multi-summary pre-training with one summary graph per GPU.

    Rank r of P trains summary graphs r, r + P, ... (one per ROUND); within a round every rank runs the reference's
    iteration body (modelTrainer.py:61-69: BCE on sigmoid outputs for summary models) on ITS graph and embedding,
    and the gradients of the shared R-GCN parameters (rgcn1/rgcn2 weight, root, bias) are averaged over the ranks that
    hold a graph in this round before the optimiser step, so every rank applies the same update to the same weights.
    Returns {'model', 'losses' (this rank's, per round), 'embeddings' {graph index: tensor}, 'semantics'}; each graph
    also gets ``.embedding`` like modelTrainer.py:82.  Works on any torch.distributed backend (gloo in the CPU test)."""
    import torch.distributed as dist
    if layers_cls is None:
        from .layers import Emb_Layers as layers_cls
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    first = sum_graphs[0]
    torch.manual_seed(seed)                                   # identical replicated weights on every rank
    model = layers_cls(2 * len(first.relations) + 1, hidden_l, num_classes, first.num_nodes, emb_dim, len(sum_graphs))
    shared = [p for conv in (model.rgcn1, model.rgcn2) for p in (conv.weight, conv.root, conv.bias)]
    loss_f, activation = get_losst('', sumModel=True)
    rounds = (len(sum_graphs) + world - 1) // world
    losses, embeddings = [], {}
    for rnd in range(rounds):
        gi = rnd * world + rank
        mine = sum_graphs[gi] if gi < len(sum_graphs) else None
        active = float(min(world, len(sum_graphs) - rnd * world))
        if mine is not None:
            torch.manual_seed(seed + 1000 + gi)               # the fresh N(0,1) embedding of this graph
            model.reset_embedding(mine.num_nodes, emb_dim)
        model = model.to(device)
        td = mine.training_data.to(device) if mine is not None else None
        use_fused = fused_adam and torch.device(device).type == 'cuda'
        opt = make_optimizer(model, lr, weight_d, fused=use_fused)
        curve = []
        for _ in range(epochs):
            model.train()
            opt.zero_grad()
            if mine is not None:
                out = model(td, activation)
                loss = loss_f(out[td.x_train], td.y_train.to(torch.float32))
                loss.backward()
                curve.append(float(loss.item()))
            if world > 1:
                flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in shared])
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
                flat /= active
                off = 0
                for p in shared:
                    g = flat[off:off + p.numel()].view_as(p)
                    off += p.numel()
                    if p.grad is None:
                        p.grad = g.clone()
                    else:
                        p.grad.copy_(g)
            if mine is None:                                  # an idle rank still follows the shared weights
                for p in model.parameters():
                    if not any(p is q for q in shared):
                        p.grad = None
            opt.step()
        losses.append(curve)
        if mine is not None:
            mine.embedding = model.embedding.weight.detach().clone()
            embeddings[gi] = mine.embedding
    return {'model': model, 'losses': losses, 'embeddings': embeddings, 'semantics': PARALLEL_SEMANTICS,
            'rounds': rounds, 'world': world}
