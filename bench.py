#!/usr/bin/env python
"""bench.py -- network-density hot path, config C2 of BASELINE.json by default:
"GWAS-hit set of ~500 loci, -n 100 circular permutation, Spearman, minimal FANTOM5 expression
shape, 1xB200".  One STEP = one whole job over one batch of synthetic input: rank + row
statistics + packing of the table, then for every permutation the correlation of its hit
regions against the whole table, the node-specific empirical p / edge scores, grouping (-j),
winners, network density and iterative removal (reference 1-make-network.py:100-482 run once
per permutation), then -- on N > 1 GPUs -- one NCCL gather of the density vectors.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path

Prints ONE JSON line (see the contract in the task description).  Multi-GPU is weak scaling:
every rank reduces its own 101 hit sets (independent permutations, no data-path collective);
`value` = permutations all ranks processed / max-over-ranks device time.
"""
import os
import sys

# libgomp reads OMP_NUM_THREADS when it is loaded (the reference sets it before LoadLibrary,
# 0-prepare-coex.py:148 then :223), so this must precede every import that may load it.
os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
if ("--impl" in sys.argv and "reference" in sys.argv) or "--impl=reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and must get every host core
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import argparse
import json
import statistics
import subprocess
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="minimal", choices=["tiny", "small", "minimal", "full"])
    ap.add_argument("--rows", type=int, default=0, help="override table rows (with --samples)")
    ap.add_argument("--samples", type=int, default=0)
    ap.add_argument("--loci", type=int, default=500)
    ap.add_argument("--perms", type=int, default=100, help="-n: permutations besides the real data")
    ap.add_argument("--window", type=int, default=10000, help="-w")
    ap.add_argument("--anticorrelation", action="store_true")
    ap.add_argument("--pval-to-join", type=float, default=0.1)
    ap.add_argument("--gemm", default="tcgen05", choices=["tcgen05", "simt", "tcgen05_bn64", "tcgen05_bn128", "tcgen05_persistent", "tcgen05_persistent128"])
    ap.add_argument("--cpu-sample-perms", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=1234)
    return ap.parse_args()


def workload(args, rank):
    """Synthetic C2 inputs: table, global correlation pair indices, hit sets of this rank."""
    from coexpression_b200 import synth
    N, n = synth.SHAPES[args.shape]
    if args.rows and args.samples:
        N, n = args.rows, args.samples
    t = synth.make_table(N, n, seed=args.seed)
    hits, _, _ = synth.make_permutation_hits(t, args.loci, args.perms, window=args.window,
                                             seed=args.seed + 17 + 1000 * rank)
    ga, gb = synth.global_pair_indices(t.names, 200000, seed=args.seed + 5)
    return t, hits, ga, gb


def config_dict(args, t, hits, extra=None):
    P = [len(h) for h in hits]
    cfg = {
        "workload": "C2: GWAS-hit set, circular permutation, Spearman, minimal FANTOM5 shape (assumed "
                    f"{t.N}x{t.n}: the reference's minimal table is an external file of unrecorded shape)",
        "table_rows": t.N, "samples": t.n, "loci": args.loci, "numperms": args.perms,
        "hit_sets_per_gpu": len(hits), "hits_real": P[0], "hits_perm_mean": float(np.mean(P[1:])) if len(P) > 1 else 0.0,
        "window": args.window, "correlationmeasure": "Spearman", "permutationmode": "circular",
        "anticorrelation": bool(args.anticorrelation), "pval_to_join": args.pval_to_join,
        "l2": "L2 flushed (256 MiB write) between timed steps; per-step working set (int32 correlation strips) "
              "also exceeds the 126 MB L2",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except Exception:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:           # region shorter than one sample: take the nearest samples
            for ts, line in self.rows[-3:]:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1])); smax.append(float(f[2]))
                except Exception:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own path (unmodified C file through ctypes + the Python-3
# restatement of 1-make-network.py), one process, all host threads
# ------------------------------------------------------------------------------------------------
def cpu_permutations(t, hit_sets, glist, args):
    """Runs the oracle per permutation exactly as a 1-make-network.py process would: the rank
    transform of the whole table is redone for every permutation (1-make-network.py:108-111)."""
    from oracle import network_oracle as no
    t0 = time.perf_counter()
    pairs = 0
    for h in hit_sets:
        R = no.rank_rows(t.X)
        no.make_network(t.names, t.X, [t.names[r] for r in h], t.addresses(), glist,
                        anticorrelation=args.anticorrelation, pval_to_join=args.pval_to_join, Xrank=R)
        pairs += len(h) * (t.N - 1)
    dt = time.perf_counter() - t0
    return dt, pairs


def cpu_global_list(t, ga, gb):
    from oracle import network_oracle as no
    R = no.rank_rows(t.X)
    g = no.get_pearson(R, ga, gb)
    return np.sort(np.where(np.isnan(g), -99.0, g)).tolist()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import network_oracle as no
    kind = no.default_kind()
    cores = no.num_threads()
    t, hits, ga, gb = workload(args, 0)
    glist = cpu_global_list(t, ga, gb)
    sample = hits[1:1 + max(1, args.cpu_sample_perms // 2)] or hits[:1]
    times, pairs = [], 0
    for i in range(args.warmup + args.steps):
        if i == min(args.warmup, 1) and args.warmup > 1:
            pass
        dt, pr = cpu_permutations(t, sample, glist, args)
        if i >= args.warmup:
            times.append(dt); pairs += pr
    tot = sum(times)
    value = len(sample) * len(times) / tot
    line = {
        "impl": "reference", "metric": "permutations_per_sec", "value": value, "unit": "permutations/s",
        "pairs_per_sec": pairs / tot, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, t, hits, {"step": f"{len(sample)} permuted hit set(s) of the workload per step"}),
        "cpu_baseline": {"value": value, "unit": "permutations/s", "cores": cores, "kind": kind,
                         "sample": f"{len(sample)} permutation(s) per step (hits {[len(h) for h in sample]}), rank transform "
                                   "redone per permutation as each reference process does; hit-vs-hit statistic from "
                                   "the C library (consistent oracle), which is FASTER than the reference's P^2 scipy calls"},
        "e2e": {"value": value, "unit": "permutations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from coexpression_b200.api import CoexContext, Options, STAGES

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on fd 1 when the communicator is created; stdout must
        # carry ONE JSON line, so fd 1 points at stderr while NCCL initialises
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    t, hits, ga, gb = workload(args, rank)
    K = len(hits)
    opts = Options(anticorrelation=args.anticorrelation, pval_to_join=args.pval_to_join)
    perm_off = np.zeros(K + 1, dtype=np.int64)
    perm_off[1:] = np.cumsum([len(h) for h in hits])
    H = int(perm_off[-1])
    hit_rows = np.concatenate(hits).astype(np.int32)
    pairs_per_step = int(sum(len(h) * (t.N - 1) for h in hits))

    ctx = CoexContext(local)
    ctx.set_gemm_mode(args.gemm)
    ctx.set_expression(t.X)
    ctx.prepare()
    ctx.set_features(t.chrom, t.start, t.end, t.names)
    g = ctx.pairs(ga, gb, ranked=True)                       # global list via the resident-table pair kernel
    glist = np.sort(np.where(np.isnan(g), -99.0, g))
    ctx.set_global_list(glist)

    # ---- device-resident inputs / outputs for `value` -------------------------------------------
    X_dev = torch.from_numpy(t.X).to(dev)
    hits_dev = torch.from_numpy(hit_rows).to(dev)
    o_ng = torch.zeros(K, dtype=torch.int32, device=dev)
    o_grp = torch.zeros(H, dtype=torch.int32, device=dev)
    o_lab = torch.zeros(H, dtype=torch.int32, device=dev)
    o_win = torch.zeros(H, dtype=torch.int32, device=dev)
    # the gathered density block must have one size on every rank: pad to the largest hit count
    hpad = torch.tensor([H], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(hpad, op=dist.ReduceOp.MAX)
    Hpad = int(hpad.item())
    o_den = torch.zeros(Hpad, dtype=torch.float64, device=dev)
    o_aps = torch.zeros(H, dtype=torch.float64, device=dev)
    ptrs = dict(hit_rows=hits_dev.data_ptr(), n_groups=o_ng.data_ptr(), group_of_hit=o_grp.data_ptr(),
                group_label=o_lab.data_ptr(), group_winner=o_win.data_ptr(), group_density=o_den.data_ptr(),
                hit_score=o_aps.data_ptr())
    stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gather_buf = torch.zeros(world * Hpad, dtype=torch.float64, device=dev) if world > 1 else None

    def step_device():
        ctx.set_expression_device(X_dev.data_ptr(), t.N, t.n, keepalive=X_dev)
        ctx.prepare()
        ctx.network_batch_device(opts, K, perm_off, ptrs)
        if world > 1:
            # the one collective of the path: density vectors of every rank (2-collate pools them).
            # coex_network_batch has drained the library stream when it returns, so the gather can
            # be queued on torch's current stream.
            dist.all_gather_into_tensor(gather_buf, o_den)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stage_tot = {s: 0.0 for s in STAGES}
    stage_launch = {s: 0 for s in STAGES}
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()                                                # every rank enters the timed steps together
    launches0 = ctx.launch_count()
    t_wall0 = time.time()
    ms = []
    for _ in range(args.steps):
        flush.fill_(1)                                       # L2 flush, outside the events
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        h0 = time.perf_counter()
        ctx.set_expression_device(X_dev.data_ptr(), t.N, t.n, keepalive=X_dev)
        ctx.prepare()
        h1 = time.perf_counter()
        h2 = h1                                              # (prepare does not synchronise: the host plans the batch meanwhile)
        ctx.network_batch_device(opts, K, perm_off, ptrs)
        h3 = time.perf_counter()
        if os.environ.get("COEX_BENCH_DEBUG"):
            print(f"[rank {rank}] host ms: prepare {1e3*(h1-h0):.2f} stage_ms {1e3*(h2-h1):.2f} network_batch {1e3*(h3-h2):.2f}",
                  file=sys.stderr)
        if world > 1:
            dist.all_gather_into_tensor(gather_buf, o_den)
            e1.record()                                      # torch's current stream (after the gather)
        else:
            e1.record(stream)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        st = ctx.stage_ms()
        for s in STAGES:
            stage_tot[s] += st[s][0]; stage_launch[s] += st[s][1]
    t_wall1 = time.time()
    launches = ctx.launch_count() - launches0
    barrier()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    step_ms = float(np.mean(ms))
    tmax = torch.tensor([step_ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    step_ms_max = float(tmax.item())

    # ---- e2e: host buffers through the public API, copies inside the timed region ----------------
    X_pin = torch.from_numpy(t.X).pin_memory()
    Xh = X_pin.numpy()
    e2e_ms = []
    h2d = Xh.nbytes + hit_rows.nbytes + perm_off.nbytes
    d2h = 4 * K + 3 * 4 * H + 2 * 8 * H
    for i in range(2 + max(2, args.steps // 2)):
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        ctx.set_expression(Xh)
        ctx.prepare()
        res = ctx.network_batch(hits, opts)
        dens = res.group_density
        if world > 1:
            o_den[:H].copy_(torch.from_numpy(dens))
            dist.all_gather_into_tensor(gather_buf, o_den)
            all_dens = gather_buf.cpu()
        w1 = time.perf_counter()
        if i >= 2:
            e2e_ms.append(1e3 * (w1 - w0))
    e2e_t = torch.tensor([float(np.mean(e2e_ms))], device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_step_ms = float(e2e_t.item())

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        tc_peak = peaks.get("bf16_tflops", 1590.0)
        peak_src = "measured (MEASURED_PEAKS.json, burst)" if peaks else "fallback (B200_PROFILING.md)"
        per = {s: (stage_tot[s] / args.steps, stage_launch[s] / args.steps) for s in STAGES}
        # algorithmic work per step
        # the correlation strips are formed once per UNIQUE hit row (queries of several permutations share it)
        U = int(np.unique(hit_rows).size)
        gemm_flops = 2.0 * t.n * U * t.N
        count_bytes = 4.0 * U * t.N + 8.0 * float(np.sum(np.diff(perm_off) ** 2))
        traffic = {}
        is_default = (args.shape == "minimal" and not args.rows and args.loci == 500 and args.perms == 100 and
                      args.window == 10000 and args.seed == 1234 and world == 1)
        try:
            if is_default:
                with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
                    traffic = json.load(f)
        except Exception:
            traffic = {}
        rank_bytes = float(t.N) * t.n * (8 + 2 + 2)
        roof = {}
        if per["corr_gemm"][0] > 0:
            a = gemm_flops / (per["corr_gemm"][0] * 1e-3) / 1e12
            roof["corr_gemm"] = {"bound": "tensor", "achieved": a, "peak": tc_peak, "unit": "TFLOP/s", "frac": a / tc_peak,
                                 "traffic": traffic.get("corr_gemm"), "ms_per_step": per["corr_gemm"][0], "launches_per_step": per["corr_gemm"][1]}
        if per["count_score"][0] > 0:
            a = count_bytes / (per["count_score"][0] * 1e-3) / 1e9
            roof["count_score"] = {"bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak,
                                   "traffic": traffic.get("count_score"),
                                   "algorithmic_bytes_per_launch": count_bytes / max(per["count_score"][1], 1), "ms_per_step": per["count_score"][0],
                                   "launches_per_step": per["count_score"][1]}
        if per["rank_pack"][0] > 0:
            a = rank_bytes / (per["rank_pack"][0] * 1e-3) / 1e9
            roof["rank_pack"] = {"bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak,
                                 "traffic": traffic.get("rank_pack"), "ms_per_step": per["rank_pack"][0], "launches_per_step": per["rank_pack"][1]}
        dominant = max(("corr_gemm", "count_score", "rank_pack", "network"), key=lambda s: per[s][0])
        main = dict(roof.get(dominant) or roof.get("corr_gemm") or {})
        main["kernel"] = dominant if dominant in roof else "corr_gemm"
        main["kernel_symbol"] = {"count_score": "count_dedup_kernel<0> (row-rank regime: count_dedup_kernel<2> + rank_score_kernel)",
                                 "corr_gemm": "gemm_i8_tc_persistent_kernel / gemm_i8_tc_persistent128_kernel (rows of >= 1024 samples)",
                                 "rank_pack": "rank_pack_kernel"}.get(main["kernel"], main["kernel"])
        main["peak_source"] = peak_src
        main["note"] = ("roofline of the kernel with the largest share of the step (CUDA events per launch on the library "
                        "stream); traffic = ncu dram read+write per launch (profiles/r01_end_summary.md), null off the default config")
        total_perms = K * world
        line = {
            "metric": "permutations_per_sec", "value": total_perms / (step_ms_max * 1e-3), "unit": "permutations/s",
            "pairs_per_sec": pairs_per_step * world / (step_ms_max * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms_max,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s8 x s8 -> s32 (exact), f64 epilogue",
            "data": "synthetic",
            "config": config_dict(args, t, hits, {"gemm": args.gemm, "parallelism": f"permutations sharded over {world} GPU(s)"}),
            "roofline": main,
            "roofline_all": roof,
            "stage_ms_per_step": {s: per[s][0] for s in STAGES},
            "e2e": {"value": total_perms / (e2e_step_ms * 1e-3), "unit": "permutations/s", "ms_per_step": e2e_step_ms,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "CoexContext.set_expression + prepare + network_batch (pinned host table, host hit lists)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            from oracle import network_oracle as no
            sample = hits[1:1 + args.cpu_sample_perms] or hits[:1]
            dt, pr = cpu_permutations(t, sample, glist.tolist(), args)
            line["cpu_baseline"] = {
                "value": len(sample) / dt, "unit": "permutations/s", "cores": no.num_threads(), "kind": no.default_kind(),
                "pairs_per_sec": pr / dt,
                "sample": f"{len(sample)} permuted hit sets of this workload (hits {[len(h) for h in sample]}), "
                          "rank transform redone per permutation as every reference process does; consistent oracle "
                          "(statistic from the C library, faster than the reference's P^2 scipy calls)"}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(run_reference(a) if a.impl == "reference" else run_ours(a))
