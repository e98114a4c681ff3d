#!/usr/bin/env python
"""Development aid: all-pairs histogram (C3) on one GPU at the FANTOM5 shape, stage times."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coexpression_b200 import synth
from coexpression_b200.api import CoexContext
N, n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000, 1829
t = synth.make_table(N, n, seed=1333)
with CoexContext(0) as ctx:
    ctx.set_workspace_mb(2048)
    ctx.set_expression(t.X); ctx.prepare()
    NB = int(os.environ.get("NB", str(1 << 20)))
    for ranked in (True, True, False):
        t0 = time.perf_counter(); g = c = 0.0
        for i, r0 in enumerate(range(0, N, 4096)):
            ctx.allpairs_accumulate(ranked, r0, min(N, r0 + 4096), NB, reset=(i == 0))
            st = ctx.stage_ms(); g += st["corr_gemm"][0]; c += st["count_score"][0]
        h, nan = ctx.allpairs_read(NB)
        dt = time.perf_counter() - t0
        chk = int((np.arange(NB, dtype=np.uint64) * h).sum() % (1 << 61))
        print("spearman" if ranked else "pearson", round(dt, 4), "s  gemm", round(g, 1), "ms hist", round(c, 1), "ms  total", int(h.sum()) + nan, "checksum", chk, flush=True)
