"""ctypes binding of the drop-in shared library (include/coexpression_b200.h).

The library is loaded exactly the way the reference loads its own
(`ctypes.cdll.LoadLibrary(<dir>/coexpression_v2.so)`, reference 1-make-network.py:66-67).
There is no fallback: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "csrc", "coexpression_v2.so")

c_ll = ctypes.c_longlong
c_int_p = ctypes.POINTER(ctypes.c_int)


class CoexParams(ctypes.Structure):
    """struct coex_params of include/coexpression_b200.h"""
    _fields_ = [
        ("measure", ctypes.c_int),
        ("anticorrelation", ctypes.c_int),
        ("noiter", ctypes.c_int),
        ("useaverage", ctypes.c_int),
        ("use_specific_p", ctypes.c_int),
        ("reserved", ctypes.c_int),
        ("pval_to_join", ctypes.c_double),
        ("selectionthreshold", ctypes.c_double),
        ("soft_separation", ctypes.c_longlong),
    ]


# name -> (restype, argtypes); every symbol include/coexpression_b200.h declares
SIGNATURES = {
    "getPearson": (None, [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_int]),
    "coex_last_error": (ctypes.c_char_p, []),
    "coex_abi_version": (ctypes.c_int, []),
    "coex_device_count": (ctypes.c_int, []),
    "coex_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "coex_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "coex_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "coex_stream": (ctypes.c_void_p, [ctypes.c_void_p]),
    "coex_set_expression": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_ll, ctypes.c_int, ctypes.c_int]),
    "coex_prepare": (ctypes.c_int, [ctypes.c_void_p]),
    "coex_get_ranks": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "coex_set_features": (ctypes.c_int, [ctypes.c_void_p] * 5 + [c_ll]),
    "coex_set_feature_order": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_ll]),
    "coex_set_global_list": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_ll]),
    "coex_permute_map": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, c_ll, c_ll,
                                        ctypes.c_void_p, ctypes.c_void_p, c_ll, ctypes.c_int, ctypes.c_void_p]),
    "coex_pairs": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, c_ll,
                                  ctypes.c_void_p, ctypes.c_int]),
    "coex_allpairs_histogram": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_ll, c_ll, ctypes.c_int,
                                               ctypes.c_void_p, ctypes.c_void_p]),
    "coex_allpairs_accumulate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_ll, c_ll, ctypes.c_int, ctypes.c_int]),
    "coex_allpairs_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "coex_corr_block": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_ll, c_ll, ctypes.c_void_p]),
    "coex_set_pearson": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "coex_network_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(CoexParams), ctypes.c_int] +
                           [ctypes.c_void_p] * 10 + [ctypes.c_int]),
    "coex_shard_table_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "coex_shard_table_open": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "coex_shard_prepare_rows": (ctypes.c_int, [ctypes.c_void_p]),
    "coex_shard_prepare_finish": (ctypes.c_int, [ctypes.c_void_p]),
    "coex_shard_scores_alloc": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_ll, ctypes.c_void_p]),
    "coex_shard_scores_open": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "coex_shard_count": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(CoexParams), ctypes.c_int] + [ctypes.c_void_p] * 4),
    "coex_shard_finish": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(CoexParams), ctypes.c_int] + [ctypes.c_void_p] * 9 +
                          [ctypes.c_int]),
    "coex_read_scores": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "coex_set_score_readback": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "coex_collate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_ll, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "coex_launch_count": (c_ll, [ctypes.c_void_p]),
    "coex_stage_ms": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                     ctypes.POINTER(c_ll)]),
    "coex_set_gemm_mode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "coex_set_count_mode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "coex_set_workspace_mb": (ctypes.c_int, [ctypes.c_void_p, c_ll]),
    "coex_selftest_gemm": (ctypes.c_int, [ctypes.c_void_p, c_ll, ctypes.POINTER(c_ll), ctypes.POINTER(c_ll)]),
}

_lib = None


def build(force=False):
    """Compile csrc/ into csrc/coexpression_v2.so with nvcc for sm_100a (in-tree)."""
    if force and os.path.exists(SO_PATH):
        os.remove(SO_PATH)
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc")])
    return SO_PATH


def load():
    """-> ctypes library with typed signatures.  Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f"{SO_PATH} is missing: build it with `make -C coexpression_b200/csrc` "
            "(there is no CPU fallback for the coexpression hot path)")
    lib = ctypes.cdll.LoadLibrary(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class CoexError(RuntimeError):
    pass


def check(status, what=""):
    if status != 0:
        msg = load().coex_last_error()
        raise CoexError(f"{what}: {msg.decode() if msg else 'unknown error'}")
