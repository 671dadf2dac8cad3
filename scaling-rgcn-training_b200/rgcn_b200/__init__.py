"""rgcn_b200 — host side of the B200-native R-GCN message-passing engine.

Everything numeric runs in ``lib/librgcn_b200.so`` (hand-written sm_100a CUDA, C ABI in
``include/rgcn_b200.h``); this package is the thin Python/PyTorch mirror of the reference
interface for that path.  No CPU fallback, no Triton, no backend dispatch.
"""
from . import _lib
from ._lib import EngineError
from .conv import RGCNConv, rgcn_layer
from .data import Data
from .graph import RGCNGraph, cached_graph, clear_cache
from .layers import Emb_ATT_Layers, Emb_Layers, Emb_MLP_Layers
from .evaluation import evaluate
from .optim import FusedAdam
from .embedding_tricks import (build_map_index, concat_embeddings, map_gather, stack_embeddings, sum_embeddings)

__all__ = ['EngineError', 'RGCNConv', 'rgcn_layer', 'Data', 'RGCNGraph', 'cached_graph', 'clear_cache',
           'Emb_Layers', 'Emb_MLP_Layers', 'Emb_ATT_Layers', 'build_map_index', 'map_gather',
           'sum_embeddings', 'concat_embeddings', 'stack_embeddings', 'evaluate', 'FusedAdam']
