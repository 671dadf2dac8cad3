"""CPU tests of the boundary: the C-ABI library loads and exports every symbol the header
declares; the product path fails loudly (no CPU fallback, no oracle import) without a GPU."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import PKG, REPO

from rgcn_b200 import _lib


def _header_functions():
    text = open(os.path.join(REPO, 'include', 'rgcn_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(rgcn_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _header_functions()
    assert set(declared) == set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.rgcn_abi_version() == 6


def test_library_has_sm100a_code_only():
    out = os.popen(f'cuobjdump -lelf {_lib.LIB_PATH} 2>/dev/null').read()
    if not out.strip():
        pytest.skip('cuobjdump unavailable')
    assert 'sm_100a' in out and 'sm_90' not in out and 'sm_80' not in out


def test_product_never_imports_the_oracle():
    for root, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                for line in open(os.path.join(root, f)):
                    code = line.split('//')[0].split('#include')[-1] if f.endswith(('.cu', '.cuh', '.h')) else line
                    if re.match(r'\s*(import|from)\s', line) or '#include' in line or 'sys.path' in line or 'exec' in line:
                        assert 'oracle' not in code, (f, line)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_fails_loudly_without_a_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.rgcn_graph_create(None, 1, None, 1, None, 1, 0, 4, 2, 0, 0, 0, None, C.byref(h))
    assert rc != 0 and h.value is None
    assert b'no CUDA device' in lib.rgcn_last_error() or b'CPU fallback' in lib.rgcn_last_error()
    from rgcn_b200 import RGCNConv
    conv = RGCNConv(4, 3, 5)
    ei = torch.tensor([[0, 1], [1, 2]])
    with pytest.raises(_lib.EngineError):
        conv(torch.randn(3, 4), ei, torch.tensor([0, 1]))


def test_invalid_arguments_are_rejected():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.rgcn_graph_create(None, 1, None, 1, None, 1, -1, 4, 2, 0, 0, 0, None, C.byref(h)) == -1
    assert lib.rgcn_graph_create(None, 1, None, 1, None, 1, 5, 4, 2, 0, 0, 0, None, C.byref(h)) == -1   # null tensors
    assert lib.rgcn_graph_create(None, 1, None, 1, None, 1, 0, 0, 2, 0, 0, 0, None, C.byref(h)) == -1
    assert b'rgcn_graph_create' in lib.rgcn_last_error()
    assert lib.rgcn_layer_fwd(None, None, 0, 0, None, None, None, None, 0, 0, 0, None, 0, None) == -1
    assert lib.rgcn_map_gather(None, None, None, 0, 0, 0, 0, None, None) == -1


def test_constructor_and_parameter_contract():
    """What model/layers.py:15-18,34-46 and modelTrainer.py:28-35 rely on."""
    from rgcn_b200 import RGCNConv, Emb_Layers
    c = RGCNConv(in_channels=63, out_channels=16, num_relations=89, num_bases=None)
    assert tuple(c.weight.shape) == (89, 63, 16) and tuple(c.root.shape) == (63, 16) and tuple(c.bias.shape) == (16,)
    assert float(c.bias.detach().abs().sum()) == 0 and c.comp is None
    torch.nn.init.kaiming_uniform_(c.weight, mode='fan_in')
    assert abs(float(c.weight.detach().abs().max()) - (6 / (63 * 16)) ** 0.5) < 2e-3
    b = RGCNConv(8, 4, 5, num_bases=3)
    assert tuple(b.weight.shape) == (3, 8, 4) and tuple(b.comp.shape) == (5, 3)
    m = Emb_Layers(5, 16, 3, 12, 63, 1)
    names = sorted(n for n, _ in m.named_parameters())
    assert names == ['embedding.weight', 'rgcn1.bias', 'rgcn1.root', 'rgcn1.weight', 'rgcn2.bias', 'rgcn2.root', 'rgcn2.weight']
    new = [p.detach().clone() for p in (m.rgcn1.weight, m.rgcn1.bias, m.rgcn1.root, m.rgcn2.weight, m.rgcn2.bias, m.rgcn2.root)]
    m.override_params(*new, grad=False)
    assert not any(p.requires_grad for n, p in m.named_parameters() if n.startswith('rgcn'))
    m.load_embedding(torch.zeros(12, 63), freeze=True)
    assert not m.embedding.weight.requires_grad
    import copy
    from rgcn_b200 import Data
    d = Data(edge_index=torch.zeros(2, 3, dtype=torch.long))
    d.edge_type = torch.zeros(3, dtype=torch.long)
    d2 = copy.deepcopy(d)
    assert d2.edge_type is not d.edge_type and d.to('cpu') is d
