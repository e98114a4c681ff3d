#!/usr/bin/env python
"""Config C3 of BASELINE.json on N GPUs of one box: ALL unordered row pairs of the table reduced to one histogram of r.
Row blocks are dealt to the ranks (coexpression_b200.dist.shard_row_blocks, no data-path collective); every rank keeps
its own resident copy of the packed table; the ONE exchange is the all-reduce of the per-rank histograms (NCCL).

    python scripts/bench_allpairs_multi.py --rows 200000 --samples 1829 --measure spearman            # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/bench_allpairs_multi.py --rows 200000 --samples 1829 --measure pearson

Prints one JSON line on rank 0: pairs/s = N(N-1)/2 / (max over ranks of the device time of the rank's blocks + reduce).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=50000)
    ap.add_argument("--samples", type=int, default=1829)
    ap.add_argument("--measure", default="spearman", choices=["spearman", "pearson"])
    ap.add_argument("--digits", type=int, default=4)
    ap.add_argument("--block", type=int, default=4096)
    ap.add_argument("--nbins", type=int, default=1 << 20)
    ap.add_argument("--max-blocks", type=int, default=0, help="per rank, 0 = all (a bounded sample for quick runs)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from coexpression_b200 import synth
    from coexpression_b200.api import CoexContext
    from coexpression_b200.dist import pairs_of_blocks, shard_row_blocks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        saved = os.dup(1); os.dup2(2, 1)                      # NCCL's banner must not reach stdout
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev)); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    t = synth.make_table(a.rows, a.samples, seed=77)
    N, n = t.N, t.n
    mine = shard_row_blocks(N, world, a.block)[rank]
    if a.max_blocks:
        mine = mine[:a.max_blocks]
    ranked = a.measure == "spearman"
    with CoexContext(local) as ctx:
        ctx.set_workspace_mb(2048)
        ctx.set_expression(t.X)
        ctx.prepare()
        ctx.set_pearson(a.digits, "tensor")
        ctx.allpairs_histogram(ranked, mine[0][0], mine[0][1], a.nbins)        # warm-up (packing, tensor maps)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dev_ms = 0.0
        acc = torch.zeros(a.nbins + 1, dtype=torch.int64, device=dev)          # nbins counters + the NaN-pair count
        t0 = time.perf_counter()
        for i, (r0, r1) in enumerate(mine):
            ctx.allpairs_accumulate(ranked, r0, r1, a.nbins, reset=(i == 0))   # histogram stays on the device
            st = ctx.stage_ms()
            dev_ms += st["corr_gemm"][0] + st["count_score"][0]
        ctx.allpairs_read(a.nbins, device_ptr=acc.data_ptr(), host=False)
        if world > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)                         # the one collective of C3
        out = acc.cpu().numpy()
        tot, nan_tot = out[:-1].view(np.uint64), int(out[-1])
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    pairs_mine = pairs_of_blocks(N, mine)
    tm = torch.tensor([wall, dev_ms * 1e-3, float(pairs_mine)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tm.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tm.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = sm = tm
    if rank == 0:
        pairs = int(sm[2].item())
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
        except Exception:
            pass
        line = {"metric": "correlation_pairs_per_sec", "value": pairs / mx[0].item(), "unit": "pairs/s", "n_gpus": world,
                "measure": a.measure, "rows": N, "samples": n, "row_block": a.block, "pairs": pairs,
                "all_pairs": N * (N - 1) // 2, "complete": pairs == N * (N - 1) // 2,
                "wall_s_max_over_ranks": mx[0].item(), "device_s_max_over_ranks": mx[1].item(),
                "pairs_per_s_device": pairs / mx[1].item(),
                "algorithmic_tflops_device": 2.0 * n * pairs / mx[1].item() / 1e12,
                "frac_of_bf16_peak": 2.0 * n * pairs / mx[1].item() / 1e12 / (world * peaks.get("bf16_tflops", 1590.0)),
                "hist_total": int(tot.sum()), "nan_pairs": nan_tot, "hist_plus_nan_equals_pairs": int(tot.sum()) + nan_tot == pairs,
                "scaling": "strong", "data": "synthetic"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
