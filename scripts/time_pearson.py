#!/usr/bin/env python
"""Times one C2-sized batch with `-cm Pearson` (Pearson null + Spearman statistic, quirk Q2) against `-cm Spearman`."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coexpression_b200 import synth
from coexpression_b200.api import CoexContext, Options, STAGES

N, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (20000, 256)
perms = int(sys.argv[3]) if len(sys.argv) > 3 else 100
t = synth.make_table(N, n, seed=1234)
hits, _, _ = synth.make_permutation_hits(t, 500, perms, window=10000, seed=1251)
ga, gb = synth.global_pair_indices(t.names, 200000, seed=1239)
with CoexContext() as ctx:
    ctx.set_expression(t.X); ctx.prepare()
    ctx.set_features(t.chrom, t.start, t.end, t.names)
    for cm in ("Spearman", "Pearson"):
        g = ctx.pairs(ga, gb, ranked=(cm == "Spearman"))
        ctx.set_global_list(np.sort(np.where(np.isnan(g), -99.0, g)))
        opts = Options(correlationmeasure=cm)
        for rep in range(3):
            t0 = time.perf_counter()
            res = ctx.network_batch(hits, opts)
            dt = time.perf_counter() - t0
        st = ctx.stage_ms()
        print(cm, "batch wall ms %.2f" % (1e3 * dt), {s: round(st[s][0], 2) for s in STAGES}, "densities", float(np.sum(res.group_density)))
    if len(sys.argv) > 4:                      # cross-check against the exact in-order strip + per-query kernel
        ctx.set_count_mode("per_query")
        t0 = time.perf_counter(); ref = ctx.network_batch(hits, Options(correlationmeasure="Pearson")); dt = time.perf_counter() - t0
        st = ctx.stage_ms()
        print("Pearson exact strip (count mode per_query) wall ms %.2f" % (1e3 * dt), {s: round(st[s][0], 2) for s in STAGES},
              "identical densities:", bool(np.array_equal(ref.group_density, res.group_density)))
