#!/usr/bin/env python
"""Config C3 (all unordered row pairs reduced to a histogram of r): pairs/s of the tensor-core
strips, Spearman (exact) and Pearson (4 or 3 int8 digits), plus the in-order double Pearson
kernel for reference.  Device time by CUDA events inside the library (coex_stage_ms).

    python scripts/bench_allpairs.py --rows 50000 --samples 1829 [--row-block 4096]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from coexpression_b200 import synth
from coexpression_b200.api import CoexContext


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=50000)
    ap.add_argument("--samples", type=int, default=1829)
    ap.add_argument("--row-block", type=int, default=0, help="rows [0, row_block) against all later rows (0 = all rows)")
    ap.add_argument("--nbins", type=int, default=1 << 20)
    ap.add_argument("--exact-rows", type=int, default=512, help="row block for the in-order double Pearson kernel")
    a = ap.parse_args()
    t = synth.make_table(a.rows, a.samples, seed=77)
    N, n = t.N, t.n
    rb = a.row_block or N
    pairs = sum(N - 1 - r for r in range(rb))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops", 1590.0)
    out = {"rows": N, "samples": n, "row_block": rb, "pairs": pairs, "nbins": a.nbins, "results": {}}
    with CoexContext() as ctx:
        ctx.set_workspace_mb(2048)
        ctx.set_expression(t.X)
        ctx.prepare()
        for name, ranked, digits, mode, rows in (("spearman_tensor", True, 4, "tensor", rb),
                                                 ("pearson_tensor_4digit", False, 4, "tensor", rb),
                                                 ("pearson_tensor_3digit", False, 3, "tensor", rb),
                                                 ("pearson_inorder_f64", False, 4, "exact", min(rb, a.exact_rows))):
            ctx.set_pearson(digits, mode)
            best = None
            for rep in range(2):
                t0 = time.perf_counter()
                hist, nan = ctx.allpairs_histogram(ranked, 0, rows, a.nbins)
                wall = time.perf_counter() - t0
                st = ctx.stage_ms()
                if best is None or wall < best[0]:
                    best = (wall, st)
            p = sum(N - 1 - r for r in range(rows))
            gemm_ms = best[1]["corr_gemm"][0]
            # the GEMM computes full tiles: executed pairs = rows x N (upper tiles only on the Pearson tensor path)
            out["results"][name] = {
                "rows": rows, "pairs": p, "wall_s": best[0], "pairs_per_s": p / best[0],
                "gemm_ms": gemm_ms, "hist_ms": best[1]["count_score"][0],
                "gemm_algorithmic_tflops": 2.0 * n * p / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
                "frac_of_bf16_peak": (2.0 * n * p / (gemm_ms * 1e-3) / 1e12) / peak if gemm_ms else None,
                "hist_total": int(hist.sum()), "nan_pairs": nan,
            }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
