"""ctypes binding of liblater_b200.so (the C ABI in include/later_b200.h).

There is no fallback: if the shared library is missing or does not load, importing this module
raises.  PyTorch is used by callers only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "liblater_b200.so"

# every symbol include/later_b200.h declares: name -> (restype, argtypes)
_c_ctx = C.c_void_p
SYMBOLS = {
    "later_b200_create": (C.c_int, [C.POINTER(_c_ctx), C.c_int, C.c_void_p]),
    "later_b200_destroy": (C.c_int, [_c_ctx]),
    "later_b200_last_error": (C.c_char_p, [_c_ctx]),
    "later_b200_set_graph": (C.c_int, [_c_ctx, C.c_int]),
    "later_b200_workspace_bytes": (C.c_size_t, [_c_ctx, C.c_int, C.c_int]),
    "later_b200_rgsqrf": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "later_b200_rgsqrf_reorth": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "later_b200_qdwh_polar": (C.c_int, [_c_ctx, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float,
                                        C.c_int, C.POINTER(C.c_int)]),
    "later_b200_rgsqrf_host": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "later_b200_oc_qr": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "later_b200_rgsqrf_stream_in": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                              C.c_void_p, C.c_int]),
    "later_b200_panel_qr": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "later_b200_panel32_qr": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "later_b200_tsqr_apply": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "later_b200_rhouqr": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                    C.c_int, C.c_int]),
    "later_b200_ormqr": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "later_b200_ormqr2": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "later_b200_gemm_gram": (C.c_int, [_c_ctx, C.c_void_p, C.c_int, C.c_int, C.c_long, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_int]),
    "later_b200_gemm_update": (C.c_int, [_c_ctx, C.c_void_p, C.c_int, C.c_int, C.c_long, C.c_int, C.c_int,
                                         C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_long, C.c_void_p,
                                         C.c_long, C.c_int]),
    "later_b200_peer_export": (C.c_int, [_c_ctx, C.c_size_t, C.c_void_p]),
    "later_b200_peer_import": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p]),
    "later_b200_peer_init_all": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_size_t]),
    "later_b200_comm_unique_id": (C.c_int, [C.c_void_p]),
    "later_b200_comm_init": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p]),
    "later_b200_comm_init_all": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "later_b200_rgsqrf_dist": (C.c_int, [_c_ctx, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "later_b200_rgsqrf_mgpu": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int,
                                         C.POINTER(C.c_void_p), C.c_int]),
    "later_b200_mgpu_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int)]),
    "later_b200_mgpu_destroy": (C.c_int, [C.c_void_p]),
    "later_b200_tsqr_mgpu": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int,
                                       C.POINTER(C.c_void_p), C.c_int]),
    "later_b200_mgpu_sync": (C.c_int, [C.c_void_p]),
    "later_b200_mgpu_last_error": (C.c_char_p, [C.c_void_p]),
    "later_b200_last_launch_count": (C.c_long, [_c_ctx]),
    "later_b200_last_info": (C.c_int, [_c_ctx, C.POINTER(C.c_int)]),
    "later_b200_graph_stats": (C.c_int, [_c_ctx, C.POINTER(C.c_long), C.POINTER(C.c_long)]),
}


def load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m later_b200.build` "
            "(nvcc, sm_100a).  later_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()
