"""ctypes binding of coex_hostio.so (include/coexpression_b200_io.h): the two large result objects of a permutation,
`individualscores` and `indjoined` (reference 1-make-network.py:523-540), written straight from the score matrix with
the text json.dump would produce -- byte for byte (tests/test_hostio.py) -- at ~25 ms instead of ~0.7 s per permutation
of 650 hit regions.  Host code only; the library is built next to coexpression_v2.so by the same Makefile."""
from __future__ import annotations

import ctypes
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "csrc", "coex_hostio.so")

c_ll = ctypes.c_longlong

# name -> (restype, argtypes); every symbol include/coexpression_b200_io.h declares
SIGNATURES = {
    "coex_io_last_error": (ctypes.c_char_p, []),
    "coex_io_write_json_matrix": (c_ll, [ctypes.c_char_p, ctypes.c_void_p, c_ll, ctypes.c_int, ctypes.c_char_p,
                                         ctypes.c_void_p, ctypes.c_int]),
    "coex_io_format_doubles": (c_ll, [ctypes.c_void_p, c_ll, ctypes.c_void_p, c_ll]),
}

_lib = None


def build():
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "coex_hostio.so"])
    return SO_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        try:                                  # host code only (g++): built on first use if the tree was never built
            build()
        except (OSError, subprocess.CalledProcessError):
            pass
    if not os.path.exists(SO_PATH):
        raise RuntimeError(f"{SO_PATH} is missing: build it with `make -C coexpression_b200/csrc`")
    lib = ctypes.cdll.LoadLibrary(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _fail(what):
    msg = load().coex_io_last_error()
    raise OSError(f"{what}: {msg.decode() if msg else 'unknown error'}")


def format_doubles(values):
    """-> [repr(v) for v in values], formatted by the library (the number format of write_json_matrix)."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    buf = ctypes.create_string_buffer(34 * max(1, v.size))
    n = load().coex_io_format_doubles(v.ctypes.data, v.size, buf, len(buf))
    if n < 0:
        _fail("coex_io_format_doubles")
    return buf.raw[:n].decode("ascii").split("\n")[:-1]


def write_json_matrix(path, S, labels, skip_diag=True):
    """json.dump({labels[i]: {labels[j]: float(S[i, j]) for j != i} for i}, open(path, "w")), same text.  S: R x R."""
    S = np.asarray(S, dtype=np.float64)
    R = len(labels)
    if S.shape != (R, R):
        raise ValueError("write_json_matrix: S must be len(labels) x len(labels)")
    if len(set(labels)) != R:
        raise ValueError("write_json_matrix: duplicate labels (a dict would keep the last one only)")
    if S.strides[1] != 8 or S.strides[0] % 8 or S.strides[0] < 8 * R:
        S = np.ascontiguousarray(S)
    keys = [json.dumps(nm).encode("ascii") for nm in labels]
    off = np.zeros(R + 1, dtype=np.int64)
    np.cumsum([len(k) for k in keys], out=off[1:])
    n = load().coex_io_write_json_matrix(os.fsencode(path), S.ctypes.data if R else None, S.strides[0] // 8 if R else 0, R,
                                         b"".join(keys), off.ctypes.data, 1 if skip_diag else 0)
    if n < 0:
        _fail("coex_io_write_json_matrix")
    return int(n)
