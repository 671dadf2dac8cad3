import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, 'scaling-rgcn-training_b200')
ORACLE = os.path.join(REPO, 'oracle')
GOLDEN = os.path.join(REPO, 'tests', 'golden')
REFERENCE = '/root/reference'
for p in (REPO, PKG, ORACLE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    # the library is git-ignored: build it (nvcc cross-compiles sm_100a without a GPU); stale objects are
    # rebuilt, so the GPU suite never runs against a binary older than csrc/
    import __graft_entry__
    __graft_entry__.build()          # compiles only what is missing or older than its sources


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def golden_graph(name):
    """-> (edge_index [2,E] int64 STRIDED view, edge_type [E] int64 strided view, N, R) laid out like
    the reference's Graph.init_graph output (views of one [E,3] buffer)."""
    g = load_golden(f'graph_{name}.npz')
    ei, et = g['edge_index'].astype(np.int64), g['edge_type'].astype(np.int64)
    buf = torch.from_numpy(np.concatenate([ei, et[None, :]], axis=0).T.copy())   # [E,3]
    edge = buf.t()
    return edge[:2], edge[2], int(g['num_nodes']), int(g['num_relations'])


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = max(float(b.abs().max()), 1e-30)
    return float((a - b).abs().max()) / denom


@pytest.fixture(scope='session')
def cuda_device():
    return torch.device('cuda:0')
