#!/usr/bin/env python
"""Development aid: -cm Pearson node-specific null at the FANTOM5 shape (tensor-core filter + exact refinements), K hit sets."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=200000); ap.add_argument("--samples", type=int, default=1829)
ap.add_argument("--perms", type=int, default=25)
a = ap.parse_args()
from coexpression_b200 import synth
from coexpression_b200.api import CoexContext, Options
t = synth.make_table(a.rows, a.samples, seed=1333)
hits, _, _ = synth.make_permutation_hits(t, 500, a.perms, window=10000, seed=1350)
ga, gb = synth.global_pair_indices(t.names, 200000, seed=1338)
with CoexContext(0) as ctx:
    ctx.set_expression(t.X); ctx.prepare()
    ctx.set_features(t.chrom, t.start, t.end, t.names)
    g = ctx.pairs(ga, gb, ranked=False)
    ctx.set_global_list(np.sort(np.where(np.isnan(g), -99.0, g)))
    for cm in ("Spearman", "Pearson", "Pearson"):
        t0 = time.perf_counter()
        res = ctx.network_batch(hits, Options(correlationmeasure=cm))
        dt = time.perf_counter() - t0
        print(cm, len(hits), "sets", round(dt * 1e3, 1), "ms", {k: round(v[0], 1) for k, v in ctx.stage_ms().items() if v[0] > 0}, flush=True)
