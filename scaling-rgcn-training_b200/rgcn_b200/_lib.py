"""ctypes binding of librgcn_b200.so (C ABI: include/rgcn_b200.h).

The library is the product; there is no Python or CPU fallback.  Importing this module
without the built library raises immediately (build it with ``python -c "import
__graft_entry__ as g; g.build()"``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'lib', 'librgcn_b200.so')

# every symbol include/rgcn_b200.h declares
EXPORTS = (
    'rgcn_last_error', 'rgcn_abi_version', 'rgcn_graph_create', 'rgcn_graph_create_part', 'rgcn_graph_destroy', 'rgcn_graph_query',
    'rgcn_graph_export', 'rgcn_layer_workspace_bytes', 'rgcn_layer_fwd', 'rgcn_layer_bwd', 'rgcn_map_gather',
    'rgcn_kernel_launch_count', 'rgcn_profile_enable', 'rgcn_profile_collect', 'rgcn_pad_rows',
    'rgcn_eval_counts', 'rgcn_adam_step', 'rgcn_adam_step_dev', 'rgcn_layer_chunk_rows_bytes', 'rgcn_layer_fwd_keep',
    'rgcn_layer_bwd_reuse', 'rgcn_set_option', 'rgcn_graph_create_push', 'rgcn_nvl_store_rows', 'rgcn_nvl_reduce_rows', 'rgcn_gemm_prepack', 'rgcn_gemm3x_tf32', 'rgcn_attn_head_fwd', 'rgcn_attn_head_bwd',
    'rgcn_nvl_reduce_rows_sparse', 'rgcn_nvl_store_rows_sparse', 'rgcn_gram3x_tf32',
)

BRC_FWD, BRC_BWD, BRC_FWD_REL = 0, 1, 2
Q_NUM_NODES, Q_NUM_EDGES, Q_NUM_RELATIONS, Q_NUM_SEGMENTS, Q_NUM_ENTRIES, Q_NUM_CHUNKS, Q_NUM_GROUPS, \
    Q_NUM_BATCHES, Q_RANGE_NODES, Q_DEVICE_BYTES, Q_NUM_OWNED, Q_OWN_LO, Q_NUM_ENTRIES0, Q_NUM_TILES, \
    Q_NUM_TILES_NOSELF, Q_PUSH = range(16)
A_PERM, A_SEG_PTR, A_SEG_OWN, A_SEG_REL, A_SEG_PTR0, A_E_IDX, A_E_W, A_RAW_IDX, A_RAW_W, A_CHUNK_BEG, \
    A_CHUNK_END, A_BAT_SEG0, A_BAT_INFO, A_E_OWN, A_TILE_E0, A_TILE_INFO, A_CHUNK_OUT = range(17)
F_RELU_IN, F_FORCE_SIMPLE, F_NO_RELU_MASK = 1, 2, 4
OPT_OVERLAP, OPT_OVERLAP_WGRAD_CTAS, OPT_OVERLAP_DX_CTAS, OPT_NVL_MODE = 0, 1, 2, 3

_lib = None


class EngineError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f'{LIB_PATH} is missing: the CUDA engine is not built. There is no CPU fallback; '
            'run __graft_entry__.build() (nvcc, sm_100a) first.')
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32, u32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32
    lib.rgcn_last_error.restype = C.c_char_p
    lib.rgcn_last_error.argtypes = []
    lib.rgcn_abi_version.restype = C.c_int
    lib.rgcn_graph_create.restype = C.c_int
    lib.rgcn_graph_create.argtypes = [vp, i64, vp, i64, vp, i64, i64, i64, i32, i32, i32, i32, vp, C.POINTER(vp)]
    lib.rgcn_graph_create_part.restype = C.c_int
    lib.rgcn_graph_create_part.argtypes = [vp, i64, vp, i64, vp, i64, i64, i64, i32, i64, i64, i32, i32, i32, vp,
                                           C.POINTER(vp)]
    lib.rgcn_graph_create_push.restype = C.c_int
    lib.rgcn_graph_create_push.argtypes = lib.rgcn_graph_create_part.argtypes
    lib.rgcn_graph_destroy.restype = None
    lib.rgcn_graph_destroy.argtypes = [vp]
    lib.rgcn_graph_query.restype = C.c_int
    lib.rgcn_graph_query.argtypes = [vp, i32, i32, C.POINTER(i64)]
    lib.rgcn_graph_export.restype = C.c_int
    lib.rgcn_graph_export.argtypes = [vp, i32, i32, vp, i64, vp]
    lib.rgcn_layer_workspace_bytes.restype = i64
    lib.rgcn_layer_workspace_bytes.argtypes = [vp, i32, i32, i32]
    lib.rgcn_layer_fwd.restype = C.c_int
    lib.rgcn_layer_fwd.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp, i64, i32, u32, vp, i64, vp]
    lib.rgcn_layer_bwd.restype = C.c_int
    lib.rgcn_layer_bwd.argtypes = [vp, vp, i64, i32, vp, vp, vp, i64, vp, i64, i32, vp, i64, vp, vp, vp, u32, vp, i64, vp]
    lib.rgcn_layer_chunk_rows_bytes.restype = i64
    lib.rgcn_layer_chunk_rows_bytes.argtypes = [vp, i32]
    lib.rgcn_layer_fwd_keep.restype = C.c_int
    lib.rgcn_layer_fwd_keep.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp, i64, i32, u32, vp, i64, vp, vp, i64, vp]
    lib.rgcn_layer_bwd_reuse.restype = C.c_int
    lib.rgcn_layer_bwd_reuse.argtypes = [vp, vp, i64, i32, vp, vp, vp, i64, vp, i64, i32, vp, i64, vp, vp, vp, u32, vp, i64,
                                         vp, vp]
    lib.rgcn_map_gather.restype = C.c_int
    lib.rgcn_map_gather.argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), i32, i64, i32, i32, vp, vp]
    lib.rgcn_eval_counts.restype = C.c_int
    lib.rgcn_eval_counts.argtypes = [vp, i64, i32, vp, i64, vp, i32, vp, vp]
    lib.rgcn_adam_step.restype = C.c_int
    lib.rgcn_adam_step.argtypes = [vp, vp, vp, vp, i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, i64, vp]
    lib.rgcn_adam_step_dev.restype = C.c_int
    lib.rgcn_adam_step_dev.argtypes = [vp, vp, vp, vp, i64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp]
    lib.rgcn_pad_rows.restype = C.c_int
    lib.rgcn_pad_rows.argtypes = [vp, i64, i32, vp, i64, i64, vp]
    lib.rgcn_nvl_store_rows.restype = C.c_int
    lib.rgcn_nvl_store_rows.argtypes = [vp, i64, i32, vp, i64, vp, vp, i32, i64, i64, i64, vp]
    lib.rgcn_nvl_reduce_rows_sparse.restype = C.c_int
    lib.rgcn_nvl_reduce_rows_sparse.argtypes = [vp, i32, vp, i64, i64, i64, vp, i64, i32, vp]
    lib.rgcn_nvl_store_rows_sparse.restype = C.c_int
    lib.rgcn_nvl_store_rows_sparse.argtypes = [vp, i64, i32, vp, i64, vp, i32, vp, i64, i64, i64, vp]
    lib.rgcn_nvl_reduce_rows.restype = C.c_int
    lib.rgcn_nvl_reduce_rows.argtypes = [vp, vp, i32, i64, i64, i64, vp, i64, i32, vp]
    lib.rgcn_gemm_prepack.restype = C.c_int
    lib.rgcn_gemm_prepack.argtypes = [vp, i64, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.rgcn_gemm3x_tf32.restype = C.c_int
    lib.rgcn_gemm3x_tf32.argtypes = [vp, i64, i64, i32, vp, vp, i32, i32, vp, i32, vp, i64, vp, i64, i32, vp]
    lib.rgcn_gram3x_tf32.restype = C.c_int
    lib.rgcn_gram3x_tf32.argtypes = [vp, i64, i32, vp, i64, i32, i64, vp, i64, vp, vp, vp]
    lib.rgcn_attn_head_fwd.restype = C.c_int
    lib.rgcn_attn_head_fwd.argtypes = [vp, i64, vp, i64, i64, i32, i64, i32, i32, vp, vp, vp, i64, vp]
    lib.rgcn_attn_head_bwd.restype = C.c_int
    lib.rgcn_attn_head_bwd.argtypes = [vp, i64, vp, i64, i64, i32, i64, i32, i32, vp, vp, vp, i64, vp, i64, vp, i64, vp]
    lib.rgcn_set_option.restype = C.c_int
    lib.rgcn_set_option.argtypes = [i32, i64]
    lib.rgcn_kernel_launch_count.restype = i64
    lib.rgcn_kernel_launch_count.argtypes = []
    lib.rgcn_profile_enable.restype = C.c_int
    lib.rgcn_profile_enable.argtypes = [i32]
    lib.rgcn_profile_collect.restype = C.c_int
    lib.rgcn_profile_collect.argtypes = [vp, vp, vp, i32, C.POINTER(i32)]
    _lib = lib
    return lib


def check(rc: int, what: str = '') -> None:
    if rc != 0:
        msg = load().rgcn_last_error()
        raise EngineError(f'{what} failed (code {rc}): {msg.decode() if msg else "?"}')


PASS_NAMES = {1: 'wprep', 2: 'chunk_prepass', 3: 'tile_fwd', 4: 'tile_dx', 5: 'wgrad', 6: 'copy_cols', 7: 'relu_mask',
              8: 'generic', 9: 'map_gather', 10: 'self_loop', 11: 'nvl_store', 12: 'nvl_reduce', 13: 'gemm3x'}


def set_option(option: int, value: int) -> None:
    check(load().rgcn_set_option(option, value), 'rgcn_set_option')


def launch_count() -> int:
    return int(load().rgcn_kernel_launch_count())


def profile_enable(on: bool) -> None:
    load().rgcn_profile_enable(1 if on else 0)


def profile_collect(max_records: int = 65536):
    """-> list of (pass name, (gathered width, output width), milliseconds)."""
    import numpy as np
    tags = np.zeros(max_records, dtype=np.int32)
    dims = np.zeros(2 * max_records, dtype=np.int32)
    ms = np.zeros(max_records, dtype=np.float32)
    n = C.c_int32()
    check(load().rgcn_profile_collect(tags.ctypes.data, dims.ctypes.data, ms.ctypes.data, max_records, C.byref(n)),
          'rgcn_profile_collect')
    return [(PASS_NAMES.get(int(tags[i]), str(int(tags[i]))), (int(dims[2 * i]), int(dims[2 * i + 1])), float(ms[i]))
            for i in range(n.value)]
