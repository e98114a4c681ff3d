#!/usr/bin/env python
"""Launches every HBM-bound kernel of the path once on a FANTOM5-row-length table, for an ncu launch list
(north_star: achieved HBM GB/s of the rank, normalise and gather kernels):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/r02_kernels.csv python scripts/bench_kernels.py --rows 60000 --samples 1829
    python scripts/bench_kernels.py --summarise gpurun_out/r02_kernels.csv --rows 60000 --samples 1829

Algorithmic bytes per kernel follow SURVEY.md section 8(d) / DESIGN.md section 4."""
import argparse
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=60000)
ap.add_argument("--samples", type=int, default=1829)
ap.add_argument("--pairs", type=int, default=1100000)
ap.add_argument("--summarise", default="")
a = ap.parse_args()
N, n, NP = a.rows, a.samples, a.pairs
n_pad = (n + 127) // 128 * 128

if a.summarise:
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    alg = {   # kernel name fragment -> (algorithmic bytes per launch, what)
        "rank_pack_kernel": (N * n * 12.0, "8 B read + 2 B 2*rank + 2 B two int8 digits per element"),
        "rowstats_kernel": (N * n * 16.0, "two in-order passes over the f64 row"),
        "pearson_pack_kernel": (N * n * (8.0 + 4.0), "8 B read + four int8 digit planes per element"),
        "pairs_rank_kernel": (NP * (2.0 * n * 2 + 16), "two rank rows (2 B / element) + indices + result per pair"),
        "pairs_exact_kernel": (NP * (2.0 * n * 8 + 16), "two f64 rows + indices + result per pair"),
        "gather_rows_kernel": (2.0 * 2.0 * 4096 * n_pad, "two int8 digit planes of 4096 gathered rows, read + written"),
    }
    rows = list(csv.reader(open(a.summarise)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mn, mv, mu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    idc = h.index("ID")
    per = {}
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        d = per.setdefault((r[idc], r[kn]), {})
        val = float(r[mv].replace(",", ""))
        unit = r[mu]
        scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r[mn]] = val * scale
    print("| kernel | time [ms] | algorithmic GB | achieved GB/s | fraction of HBM peak (%.0f GB/s) | DRAM read + write GB (ncu) | bytes |" % hbm)
    print("|---|---|---|---|---|---|---|")
    for (kid, name), d in per.items():
        for frag, (b, what) in alg.items():
            if frag in name:
                t = d.get("gpu__time_duration.sum", 0.0)
                tr = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
                print("| `%s` | %.3f | %.3f | %.0f | %.3f | %.3f | %s |" % (frag, t * 1e3, b / 1e9, b / t / 1e9, b / t / 1e9 / hbm, tr / 1e9, what))
    sys.exit(0)

from coexpression_b200 import synth
from coexpression_b200.api import CoexContext

t = synth.make_table(N, n, seed=5)
rng = np.random.default_rng(1)
ia = rng.integers(0, N, NP).astype(np.int32)
ib = rng.integers(0, N, NP).astype(np.int32)
with CoexContext(0) as ctx:
    ctx.set_expression(t.X)
    ctx.prepare()                                         # rank_pack_kernel, rowstats_kernel
    ctx.pairs(ia, ib, ranked=True)                        # pairs_rank_kernel
    ctx.pairs(ia, ib, ranked=False)                       # pairs_exact_kernel
    ctx.corr_block("pearson", 0, 128)                     # pearson_pack_kernel (+ the digit GEMM of one row block)
    ctx.allpairs_histogram(True, 0, 4096, 1 << 16)        # gather_rows_kernel of 4096 rows (+ GEMM + histogram)
print("done")
