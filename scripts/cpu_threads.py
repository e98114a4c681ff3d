#!/usr/bin/env python3
"""Pairs/s of the reference's own getPearson (oracle/_ref/coexpression_v2.so: the unmodified C file) at 1 thread (the default of
`0-prepare-coex.py -t 1`), 4 threads (the library's default, coexpression_v2.c:29-33) and every host core -- SURVEY.md section 8(d)
asks for the three figures as context for `cpu_baseline`.  One child process per setting: libgomp reads OMP_NUM_THREADS when loaded."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import ctypes, sys, time, numpy as np
sys.path.insert(0, %r)
from coexpression_b200 import synth
rows, samples, pairs = %d, %d, %d
t = synth.make_table(rows, samples, seed=1234)
lib = ctypes.cdll.LoadLibrary(%r)
rng = np.random.default_rng(1)
a = rng.integers(0, rows, pairs).astype(np.int32); b = rng.integers(0, rows, pairs).astype(np.int32)
X = np.ascontiguousarray(t.X, dtype=np.float64); out = np.zeros(pairs)
best = 1e30
for _ in range(3):
    t0 = time.perf_counter()
    lib.getPearson(ctypes.c_void_p(X.ctypes.data), ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(a.ctypes.data),
                   ctypes.c_void_p(b.ctypes.data), ctypes.c_int(pairs), ctypes.c_int(samples))
    best = min(best, time.perf_counter() - t0)
print(pairs / best)
"""

if __name__ == "__main__":
    rows, samples, pairs = (int(x) for x in (sys.argv[1:4] + ["20000", "256", "1100000"][len(sys.argv) - 1:]))
    so = os.path.join(ROOT, "oracle", "_ref", "coexpression_v2.so")
    res = {}
    for th in sorted({1, 4, os.cpu_count() or 1}):
        env = dict(os.environ, OMP_NUM_THREADS=str(th))
        out = subprocess.check_output([sys.executable, "-c", CHILD % (ROOT, rows, samples, pairs, so)], env=env, text=True)
        res[str(th)] = float(out.strip().split("\n")[-1])
    print(json.dumps({"table": [rows, samples], "pairs": pairs, "pairs_per_sec_by_threads": res}))
