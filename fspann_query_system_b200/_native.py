"""ctypes binding of libfspann_gpu.so (include/fspann_gpu.h).  There is no fallback: if the library is missing
or no GPU is present, loading / context creation raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libfspann_gpu.so")
DEBUG_LIB_PATH = os.path.join(_HERE, "csrc", "libfspann_gpu_debug.so")

OK, E_ARG, E_STATE, E_CUDA, E_NOMEM = 0, -1, -2, -3, -4
V_OK, V_NOT_FOUND, V_NO_KEY, V_TAG_FAIL, V_NON_FINITE = 0, 1, 2, 3, 4
V_OTHER_SHARD = 0xFD
COUNTERS = 6

EXPORTS = [
    "fspann_ctx_create", "fspann_ctx_destroy", "fspann_last_error", "fspann_ctx_stream", "fspann_ctx_sync",
    "fspann_ctx_launch_count", "fspann_routing_upload", "fspann_gfunctions_upload", "fspann_deleted_set",
    "fspann_store_upload", "fspann_store_upload_shard", "fspann_store_update", "fspann_keys_set", "fspann_keys_retire", "fspann_tokengen_batch", "fspann_tokengen_batch_dev",
    "fspann_route_batch", "fspann_refine_batch", "fspann_refine_batch_ex", "fspann_search_batch", "fspann_search_batch_dev", "fspann_search_tokens", "fspann_search_tokens_dev", "fspann_touched_fetch",
    "fspann_last_stage_ms", "fspann_debug_decrypt", "fspann_set_option", "fspann_get_info", "fspann_migrate", "fspann_encrypt_batch", "fspann_routing_build", "fspann_groundtruth", "fspann_recall_batch", "fspann_route_batch_dev", "fspann_refine_batch_dev", "fspann_merge_topk_dev",
    "fspann_comm_unique_id", "fspann_comm_init", "fspann_comm_destroy", "fspann_sharded_search_batch", "fspann_sharded_search_batch_dev",
    "fspann_sharded_last_stage_ms", "fspann_routing_build_begin", "fspann_routing_build_add", "fspann_routing_build_add_dev",
    "fspann_routing_build_finish", "fspann_store_alloc_shard", "fspann_store_encrypt_dev",
]

_libs = {}


def load(debug: bool = False):
    path = DEBUG_LIB_PATH if debug else LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  There is no CPU fallback for the FSPANN hot path.")
    lib = C.CDLL(path)
    lib.fspann_last_error.restype = C.c_char_p
    lib.fspann_ctx_stream.restype = C.c_void_p
    lib.fspann_ctx_launch_count.restype = C.c_int64
    lib.fspann_last_stage_ms.restype = C.c_int64
    lib.fspann_get_info.restype = C.c_int64
    lib.fspann_sharded_last_stage_ms.restype = C.c_int64
    _libs[path] = lib
    return lib


def ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(int(a))  # raw (device or pinned host) address


class FspannError(RuntimeError):
    """Base class; subclasses mirror the Java exception classes the reference throws on this path."""


class IllegalArgumentError(FspannError, ValueError):
    pass


class IllegalStateError(FspannError):
    pass


class CudaError(FspannError):
    pass


def check(lib, ctx, rc):
    if rc == OK:
        return
    msg = lib.fspann_last_error(ctx).decode() if ctx else "context creation failed (no CUDA device? there is no CPU fallback)"
    if rc == E_ARG:
        raise IllegalArgumentError(msg)
    if rc == E_STATE:
        raise IllegalStateError(msg)
    raise CudaError(f"rc={rc}: {msg}")
