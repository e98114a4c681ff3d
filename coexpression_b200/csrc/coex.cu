// coex.cu -- the C ABI of include/coexpression_b200.h on top of the sm_100a kernels.
// Built into coexpression_v2.so (the file name the reference loads through ctypes,
// reference 1-make-network.py:66-67).  No CPU fallback anywhere in this file.
#include "common.cuh"
#include "prepare.cuh"
#include "pairs.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "count.cuh"
#include "count_dedup.cuh"
#include "count_fine.cuh"
#include "plan.cuh"
#include "network.cuh"
#include "allpairs.cuh"
#include "permute.cuh"
#include "collate.cuh"

#include <algorithm>
#include <ctime>
#include <mutex>

namespace coex {
thread_local std::string g_last_error;

// `bedtools window` keeps the B file in the UCSC binning scheme (bedtools2 bedFile.h: _binFirstShift 14, _binNextShift 3, 7 levels;
// getBin = the smallest bin that contains [start, end - 1]) and BedFile::allHits walks, for one A record, the levels from the finest
// up, the bins of a level in ascending order and the records of a bin in B-FILE order.  For the B records overlapping one A record
// that is the ascending order of (level, bin, line number in the B file): a global order of the features, whose rank per table
// row is what permute_map_kernel sorts the promoters of a locus by (the order decides which null column quirk Q1 skips).
static int upload_window_order(coex_ctx *c) {
    const long long N = c->N;
    std::vector<unsigned long long> k1((size_t)N);
    for (long long i = 0; i < N; ++i) {
        long long s = c->h_feat_start[(size_t)i] >> 14, e = (std::max(c->h_feat_end[(size_t)i], c->h_feat_start[(size_t)i] + 1) - 1) >> 14;
        int level = 0;
        while (level < 6 && s != e) { s >>= 3; e >>= 3; ++level; }
        k1[(size_t)i] = ((unsigned long long)c->h_chrom_id[(size_t)i] << 40) | ((unsigned long long)level << 36) | (unsigned long long)s;
    }
    std::vector<int> order((size_t)N), rank((size_t)N);
    for (long long i = 0; i < N; ++i) order[(size_t)i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int x, int y) {
        if (k1[(size_t)x] != k1[(size_t)y]) return k1[(size_t)x] < k1[(size_t)y];
        if (c->h_bed_index[(size_t)x] != c->h_bed_index[(size_t)y]) return c->h_bed_index[(size_t)x] < c->h_bed_index[(size_t)y];
        return x < y;
    });
    for (long long r = 0; r < N; ++r) rank[(size_t)order[(size_t)r]] = (int)r;
    COEX_TRY(c->feat_wrank.reserve(N)); COEX_TRY(c->feat_wrow.reserve(N));
    COEX_CUDA(cudaMemcpyAsync(c->feat_wrank.p, rank.data(), N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    COEX_CUDA(cudaMemcpyAsync(c->feat_wrow.p, order.data(), N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    COEX_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

static int check_ctx(coex_ctx *c) {
    if (!c) COEX_FAIL("null context");
    COEX_CUDA(cudaSetDevice(c->device));
    return 0;
}

// The packed table: sizes, buffers (stable addresses as long as the shape does not grow: their CUDA IPC handles are exchanged once
// by a job shared by several GPUs), zero padding.
static int prepare_reserve(coex_ctx *c) {
    if (!c->raw) COEX_FAIL("coex_prepare: call coex_set_expression first");
    const long long N = c->N;
    const int n = c->n;
    if (n > 8192) COEX_FAIL("coex_prepare: more than 8192 samples per row is not supported");
    c->n_pad = (int)round_up(n, 128);
    c->N_pad = round_up(N, 256);
    COEX_TRY(c->raw_sum.reserve(N)); COEX_TRY(c->raw_mean.reserve(N)); COEX_TRY(c->raw_xx.reserve(N));
    COEX_TRY(c->rawsum_pos.reserve(N));
    COEX_TRY(c->rank2.reserve((size_t)N * n));
    COEX_TRY(c->u_hi.reserve((size_t)c->N_pad * c->n_pad));
    COEX_TRY(c->u_lo.reserve((size_t)c->N_pad * c->n_pad));
    COEX_TRY(c->suu.reserve(c->N_pad)); COEX_TRY(c->sx.reserve(c->N_pad));
    COEX_TRY(c->icol.reserve(c->N_pad)); COEX_TRY(c->d_ndeg.reserve(1)); COEX_TRY(c->sxg.reserve(c->N_pad));
    return 0;
}

// rows [row0, row1) of the table: rank transform + packing, in-order row statistics; stored to every destination of `ro` / `so`
static int prepare_rows(coex_ctx *c, long long row0, long long row1, const RankOut &ro, const StatOut &so) {
    const long long N = c->N;
    const int n = c->n;
    stage_reset(c);
    c->z_digits = 0;
    c->prepared = false;
    // padding rows must be exact zeros (they enter the GEMM as B rows); every rank zeroes its own copy
    const size_t tail = (size_t)(c->N_pad - N) * c->n_pad;
    if (tail) {
        COEX_CUDA(cudaMemsetAsync(c->u_hi.p + (size_t)N * c->n_pad, 0, tail, c->stream));
        COEX_CUDA(cudaMemsetAsync(c->u_lo.p + (size_t)N * c->n_pad, 0, tail, c->stream));
    }
    if (c->N_pad > N) {
        COEX_CUDA(cudaMemsetAsync(c->suu.p + N, 0, (c->N_pad - N) * sizeof(long long), c->stream));
        COEX_CUDA(cudaMemsetAsync(c->sx.p + N, 0, (c->N_pad - N) * sizeof(double), c->stream));
    }
    COEX_TRY(stage_begin(c, ST_RANK));
    if (row1 > row0) {
        const int P2 = std::max(32, next_pow2(n));         // at least one warp-wide block of the in-register sort stages
        const size_t smem = (size_t)P2 * 10 + (size_t)c->n_pad * 2;
        COEX_CUDA(cudaFuncSetAttribute(rank_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (c->up_pending) {
            // the table is still arriving (coex_set_expression, on_device = 2): rank every chunk of rows as it lands
            for (int k = 0; k < c->up_pending; ++k) {
                const long long a0 = std::max(row0, (long long)k * c->up_rows), a1 = std::min(row1, (long long)(k + 1) * c->up_rows);
                COEX_CUDA(cudaStreamWaitEvent(c->stream, c->up_event[k], 0));
                if (a1 <= a0) continue;
                rank_pack_kernel<<<(unsigned)(a1 - a0), kRankThreads, smem, c->stream>>>(c->raw, N, n, c->n_pad, P2, a0, ro);
                COEX_CUDA(cudaGetLastError());
            }
            c->up_pending = 0;                             // the compute stream is behind every chunk from here on
        } else {
            rank_pack_kernel<<<(unsigned)(row1 - row0), kRankThreads, smem, c->stream>>>(c->raw, N, n, c->n_pad, P2, row0, ro);
            COEX_CUDA(cudaGetLastError());
        }
    }
    COEX_TRY(stage_end(c, 1));
    COEX_TRY(stage_begin(c, ST_ROWSTATS));
    if (row1 > row0) {
        rowstats_kernel<<<(unsigned)((row1 - row0 + kStatRows - 1) / kStatRows), 256, 0, c->stream>>>(c->raw, row1, n, row0, so);
        COEX_CUDA(cudaGetLastError());
    }
    COEX_TRY(stage_end(c, 1));
    return 0;
}

// what needs the WHOLE packed table: float reciprocal half-norms, guarded half-norms, number of zero-variance rows
static int prepare_finish(coex_ctx *c) {
    COEX_CUDA(cudaMemsetAsync(c->d_ndeg.p, 0, sizeof(int), c->stream));
    icol_kernel<<<(unsigned)((c->N_pad + 255) / 256), 256, 0, c->stream>>>(c->sx.p, c->rawsum_pos.p, c->N, c->N_pad, c->icol.p, c->sxg.p, c->d_ndeg.p);
    COEX_CUDA(cudaGetLastError());
    c->launches += 1;
    // no synchronisation here: the count of zero-variance rows stays on the device (the count kernels read it there),
    // so the host can plan the next call while the rank transform is still running
    COEX_TRY(tc_invalidate(c));
    c->prepared = true;
    return 0;
}

static void local_outputs(coex_ctx *c, RankOut &ro, StatOut &so) {
    ro = RankOut{}; so = StatOut{};
    ro.rank2[0] = c->rank2.p; ro.hi[0] = c->u_hi.p; ro.lo[0] = c->u_lo.p; ro.suu[0] = c->suu.p; ro.sx[0] = c->sx.p; ro.n_dst = 1;
    so.sum[0] = c->raw_sum.p; so.mean[0] = c->raw_mean.p; so.xx[0] = c->raw_xx.p; so.pos[0] = c->rawsum_pos.p; so.n_dst = 1;
}

static int prepare_impl(coex_ctx *c) {
    COEX_TRY(prepare_reserve(c));
    RankOut ro; StatOut so;
    local_outputs(c, ro, so);
    COEX_TRY(prepare_rows(c, 0, c->N, ro, so));
    return prepare_finish(c);
}

// ------------------------------------------------------------------------------------------
// de-duplicated Spearman null: unique hit rows -> GEMM strips -> count_dedup_kernel
// ------------------------------------------------------------------------------------------
static int ensure_score_table(coex_ctx *c, const coex_params *prm) {
    const long long N = c->N;
    cudaStream_t st = c->stream;
    // Edge score of every possible bisect index, built ON THE HOST with the C library's log exactly as the reference evaluates
    // -math.log(p1, 10) = -(log(p1) / log(10)) (1-make-network.py:310; CPython divides two libm logs): the scores, the `>= threshold`
    // selections and the equal-density ties of the iterative removal (quirk Q7) are then decided on the reference's own doubles
    // (a device log10 differs in the last bit).  N + 1 values, rebuilt only when N or -a change.
    COEX_TRY(c->dd_tab.reserve((size_t)N + 1));
    if (c->tab_for_N != N || c->tab_for_anticor != prm->anticorrelation) {
        c->h_score_tab.resize((size_t)N + 1);
        const double len = (double)(N - 1), l10 = std::log(10.0);
        for (long long idx = 0; idx <= N; ++idx) {
            double p1 = 1.0 - (double)idx / len;
            if (prm->anticorrelation) p1 = std::min(p1, (double)idx / len);
            c->h_score_tab[(size_t)idx] = (p1 > 0.0) ? -(std::log(p1) / l10) : 0.0;     // math.log raises for p1 <= 0 -> score 0 (Q4)
        }
        COEX_CUDA(cudaMemcpyAsync(c->dd_tab.p, c->h_score_tab.data(), ((size_t)N + 1) * sizeof(double), cudaMemcpyHostToDevice, st));
        c->tab_for_N = N; c->tab_for_anticor = prm->anticorrelation;
    }
    return 0;
}

static int run_strip_gemm(coex_ctx *c, const int *d_rows, long long nq, long long nq_pad);
static int run_strip_gemm_pearson(coex_ctx *c, const int *d_rows, long long nq, long long nq_pad, long long a_rows);

static int spearman_dedup_path(coex_ctx *c, const coex_params *prm, int K, const long long *perm_off,
                               const int *d_hits, int Pmax, double *d_sc, long long *d_cn) {
    const long long H = perm_off[K];
    const long long N = c->N;
    cudaStream_t st = c->stream;
    const bool tdbg = getenv("COEX_HOST_DEBUG") != nullptr;
    auto now_us = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
    const double t_start = now_us();
    const bool pearson = (prm->measure == 1);
    const int D = c->pearson_digits;
    if (pearson) COEX_TRY(pearson_pack(c));
    bool rank_mode = (c->count_mode == 3) && !pearson;      // 3 = forced; count mode 0 decides below (cost model)
    const int Tcap = std::max(kDdBuckets / 2, (int)round_up(Pmax, 256));   // 2*Tcap >= buckets: the bucket cursors alias thr32|hist
    const int Qcap = kPlanQcap;
    {   // strips first (upper bound of the unique rows), so that the big buffers keep their place whatever the plan needs
        long long Us0 = (c->strip_mb << 20) / (c->N_pad * 4);
        Us0 = std::max<long long>(256, Us0 / 256 * 256);
        Us0 = std::min<long long>(Us0, round_up(std::min<long long>(H, N), 256));
        COEX_TRY(c->strip.reserve((size_t)Us0 * c->N_pad));
        COEX_TRY(c->a_hi.reserve((size_t)Us0 * c->n_pad)); COEX_TRY(c->a_lo.reserve((size_t)Us0 * c->n_pad));
    }
    // ---- the plan, on the device (plan.cuh): queries grouped by table row, work items ------------------------
    COEX_TRY(c->pl_cnt.reserve((size_t)N + 2)); COEX_TRY(c->pl_cursor.reserve((size_t)N));
    COEX_TRY(c->dd_uniq.reserve((size_t)std::min<long long>(H, N) + 1)); COEX_TRY(c->pl_ustart.reserve((size_t)std::min<long long>(H, N) + 2));
    COEX_TRY(c->dd_qlist.reserve((size_t)H)); COEX_TRY(c->dd_qperm.reserve((size_t)H));
    COEX_TRY(c->pl_meta.reserve(4)); COEX_TRY(c->pl_err.reserve(1));
    COEX_CUDA(cudaMemsetAsync(c->pl_cnt.p, 0, ((size_t)N + 2) * sizeof(int), st));
    COEX_CUDA(cudaMemsetAsync(c->pl_err.p, 0, sizeof(int), st));
    const unsigned hb = (unsigned)((H + 255) / 256);
    const bool sharded = (c->shard_phase == 1);
    const int sh_world = sharded ? c->shard_world : 1, sh_rank = sharded ? c->shard_rank : 0;
    double *const *perm_scores = sharded ? c->d_perm_scores.p : nullptr;
    plan_count_kernel<<<hb, 256, 0, st>>>(d_hits, c->d_perm_off.p, K, H, N, c->pl_cnt.p, c->dd_qperm.p, c->pl_err.p, sh_world, sh_rank);
    COEX_KCHECK(st);
    plan_rows_kernel<<<1, kPlanThreads, 0, st>>>(c->pl_cnt.p, N, c->dd_uniq.p, c->pl_ustart.p, c->pl_cursor.p, c->pl_meta.p);
    COEX_KCHECK(st);
    long long h_meta[2] = {0, 0};
    int h_err = 0;
    COEX_CUDA(cudaMemcpyAsync(h_meta, c->pl_meta.p, sizeof h_meta, cudaMemcpyDeviceToHost, st));
    COEX_CUDA(cudaMemcpyAsync(&h_err, c->pl_err.p, sizeof h_err, cudaMemcpyDeviceToHost, st));
    plan_scatter_kernel<<<hb, 256, 0, st>>>(d_hits, H, c->pl_cursor.p, c->dd_qlist.p, sh_world, sh_rank);
    COEX_KCHECK(st);
    COEX_CUDA(cudaStreamSynchronize(st));                  // the one read-back of the plan: U sizes the strips
    if (h_err) COEX_FAIL("hit row out of range");
    const long long U = h_meta[0];
    if (U == 0) return 0;                                  // (a rank of a shared job may own none of the hit rows)
    // ---- strips of unique rows ------------------------------------------------------------------------
    const long long ld = c->N_pad;
    long long Us = (c->strip_mb << 20) / (ld * (pearson ? 8 : 4));     // Pearson: the float r strip + the int32 Spearman strip of the statistics
    Us = std::max<long long>(256, Us / 256 * 256);
    Us = std::min<long long>(Us, round_up(U, 256));
    COEX_TRY(c->strip.reserve((size_t)Us * ld));
    COEX_TRY(c->a_hi.reserve((size_t)Us * c->n_pad)); COEX_TRY(c->a_lo.reserve((size_t)Us * c->n_pad));
    if (pearson) {
        COEX_TRY(c->strip_f32.reserve((size_t)Us * ld));
        COEX_TRY(c->za_planes.reserve((size_t)D * Us * c->n_pad)); COEX_TRY(c->za_inv_s.reserve((size_t)Us));
    }
    COEX_TRY(c->pl_nitems.reserve((size_t)U + 1)); COEX_TRY(c->pl_itemoff.reserve((size_t)U + 2));
    COEX_TRY(c->dd_items.reserve((size_t)(H + 1) * sizeof(DedupItem)));       // at most one item per query
    const unsigned ub = (unsigned)((U + 255) / 256);
    plan_items_kernel<false><<<ub, 256, 0, st>>>(c->pl_ustart.p, c->dd_qlist.p, c->dd_qperm.p, c->d_perm_off.p, (int)U, Tcap, Us,
                                                 c->pl_nitems.p, nullptr, nullptr, c->dd_qlist.p);
    COEX_KCHECK(st);
    plan_item_scan_kernel<<<1, kPlanThreads, 0, st>>>(c->pl_nitems.p, (int)U, c->pl_itemoff.p);
    COEX_KCHECK(st);
    plan_items_kernel<true><<<ub, 256, 0, st>>>(c->pl_ustart.p, c->dd_qlist.p, c->dd_qperm.p, c->d_perm_off.p, (int)U, Tcap, Us,
                                                nullptr, c->pl_itemoff.p, reinterpret_cast<DedupItem *>(c->dd_items.p), nullptr);
    COEX_KCHECK(st);
    const int n_strips = (int)((U + Us - 1) / Us);
    COEX_TRY(c->pl_bounds.reserve((size_t)n_strips + 1));
    plan_bounds_kernel<<<(unsigned)(n_strips / 256 + 1), 256, 0, st>>>(c->pl_itemoff.p, (int)U, Us, n_strips, c->pl_bounds.p);
    COEX_KCHECK(st);
    std::vector<long long> strip_item_begin((size_t)n_strips + 1);
    COEX_CUDA(cudaMemcpyAsync(strip_item_begin.data(), c->pl_bounds.p, ((size_t)n_strips + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    c->launches += 7;
    // hit-column bitmap (<= 32 KB: two CTAs per SM stay resident); Pearson null: a hit column's value is not its (Spearman) statistic
    const int bm_words = (!pearson && N <= 262144) ? (int)((N + 31) / 32) : 0;
    const int bm_words_rank = bm_words;
    if (tdbg) fprintf(stderr, "[coex host] dedupe plan (device): %.0f us on the host incl. one sync (H=%lld U=%lld)\n", now_us() - t_start, H, U);
    const int claim = (bm_words > 0 && bm_words <= 2048) ? 1 : 0;      // <= 8 KB for the second bitmap
    const size_t smem = dedup_smem_bytes(Tcap, Qcap, bm_words, claim);
    if (smem > 227 * 1024) COEX_FAIL("count kernel: too many statistics per query for shared memory");
    int per_sm = 1;
    if (pearson) {
        COEX_CUDA(cudaFuncSetAttribute(count_dedup_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        COEX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, count_dedup_kernel<1>, kDdThreads, smem));
    } else {
        COEX_CUDA(cudaFuncSetAttribute(count_dedup_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        COEX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, count_dedup_kernel<0>, kDdThreads, smem));
    }
    const int grid_max = std::max(1, per_sm) * c->sm_count;
    COEX_TRY(c->dd_gthr.reserve((size_t)grid_max * Tcap * 6));      // 3 x Tcap records of 16 bytes per CTA
    if (tdbg) fprintf(stderr, "[coex host] count kernel: %zu bytes of shared memory, %d CTA(s) per SM, Tcap %d, bitmap words %d\n", smem, per_sm, Tcap, bm_words);
    COEX_TRY(ensure_score_table(c, prm));
    // Spearman null, query items: count_fine_kernel (count_fine.cuh); what it declines goes to count_dedup_kernel<0> through a re-do list
    // (it wins where a row is long against the statistics of its work items -- FANTOM5: 200 000 columns, ~2 000 statistics, 55 ms
    //  against 72 ms for 201 hit sets; on a 20 000-row table its per-item set-up, alone on its SM, loses to the sorted-statistics
    //  kernel, 4.6 ms against 3.6 ms: profiles/r02_count_experiments.md)
    const bool fine_ok = !pearson && Tcap <= kCfTcap && Qcap <= kCfQcap;
    const bool fine = fine_ok && (c->count_mode == 4 || (c->count_mode == 0 && N >= c->fine_min_rows));
    if (c->count_mode == 4 && !fine_ok) COEX_FAIL("count mode 4: Spearman null and at most 4096 hits per set");
    int fine_nb = (N <= 32768) ? 16384 : 32768;
    if (const char *e = getenv("COEX_FINE_NB")) fine_nb = atoi(e);
    const size_t fsmem = fine_smem_bytes(fine_nb);
    long long fine_chunk = 0;
    double fine_sigmas = 6.0;                // centre of the bucket map in standard deviations of a null correlation (count_fine.cuh: CfMap;
                                             // measured at the FANTOM5 shape: 3 -> 170 ms, 4 -> 151, 6 -> 145, 8 -> 152, 12 -> 170)
    if (const char *e = getenv("COEX_FINE_SIGMAS")) fine_sigmas = atof(e);
    if (fine) {
        if (fine_nb < 1024 || fine_nb > 32768 || (fine_nb & (fine_nb - 1))) COEX_FAIL("COEX_FINE_NB: a power of two in 1024 .. 32768");
        const long long nch = (N + kCfChunkMax - 1) / kCfChunkMax;
        fine_chunk = round_up((N + nch - 1) / nch, 4 * kCfThreads);
        if (const char *e = getenv("COEX_FINE_CHUNK")) fine_chunk = std::min<long long>(kCfChunkMax, round_up(std::max(1LL, atoll(e)), 4 * kCfThreads));   // (tests: several passes on a small table)
        COEX_CUDA(cudaFuncSetAttribute(count_fine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
        COEX_TRY(c->dd_redo.reserve((size_t)H + 2));
    }
    DevBuf<unsigned long long> phase;
    const bool dbg = getenv("COEX_COUNT_DEBUG") != nullptr;
    if (dbg) { COEX_TRY(phase.reserve(16)); COEX_CUDA(cudaMemsetAsync(phase.p, 0, 16 * sizeof(unsigned long long), st)); }
    size_t s_idx = 0;
    int rank_grid_max = 1;
    long long c_sc_total = 0;
    for (int k = 0; k < K; ++k) c_sc_total += (perm_off[k + 1] - perm_off[k]) * (perm_off[k + 1] - perm_off[k]);
    for (long long u0 = 0; u0 < U; u0 += Us, ++s_idx) {
        const long long nu = std::min(Us, U - u0);
        COEX_TRY(run_strip_gemm(c, c->dd_uniq.p + u0, nu, round_up(nu, 256)));      // exact Spearman sums (statistics in both modes, Q2)
        if (pearson) COEX_TRY(run_strip_gemm_pearson(c, c->dd_uniq.p + u0, nu, round_up(nu, 256), Us));
        if (s_idx == 0) {
            COEX_CUDA(cudaStreamSynchronize(st));     // second (and last) read-back of the plan; the first GEMM is queued already
            // Regime switch: when the hit sets are so large that the queries of a row carry more statistics than the row has
            // columns, rank every strip row once (count_dedup_kernel<2>) and let the queries look their indices up.
            const long long SCtot = c_sc_total / sh_world;      // (a rank of a shared job looks up the rows it owns)
            const double cost_items = (double)strip_item_begin[(size_t)n_strips] * (double)N;
            // in streamed values: a rank item carries 8192 statistics (heavier set-up than a query item: x 1.5), a look-up of
            // rank_score_kernel costs about half a streamed value.  Measured on 132 hit sets of ~5 400 regions on a 20 000-row table
            // (config C5): query items 194 ms (1.4e10 values), rank mode 46 ms (1.2e9 values + 3.9e9 look-ups).
            const double cost_rank = 1.5 * (double)U * (double)((N + kRankTcap - 1) / kRankTcap) * (double)N + 0.6 * (double)SCtot;
            if (c->count_mode == 0 && !pearson && bm_words_rank > 0 && cost_rank < cost_items) rank_mode = true;
            if (rank_mode && bm_words_rank == 0) COEX_FAIL("rank mode needs the hit-column bitmap (table of at most 262144 rows)");
            if (tdbg) fprintf(stderr, "[coex host] count: %lld items, cost items %.3g vs rank %.3g -> %s\n", strip_item_begin[(size_t)n_strips],
                              cost_items, cost_rank, rank_mode ? "rank mode" : "query items");
        }
        if (rank_mode) {
            const int rclaim = (bm_words_rank <= 2048) ? 1 : 0;
            const size_t rsmem = dedup_smem_bytes(kRankTcap, Qcap, bm_words_rank, rclaim);
            if (s_idx == 0) {
                COEX_CUDA(cudaFuncSetAttribute(count_dedup_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
                int rper = 1;
                COEX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rper, count_dedup_kernel<2>, kDdThreads, rsmem));
                rank_grid_max = std::max(1, rper) * c->sm_count;
                COEX_TRY(c->dd_gthr.reserve((size_t)rank_grid_max * kRankTcap * 6));
                COEX_TRY(c->rank_strip.reserve((size_t)Us * ld));
            }
            const int nchunk = (int)((N + kRankTcap - 1) / kRankTcap);
            DedupArgs da{};
            da.strip = c->strip.p; da.ld = ld; da.items = nullptr; da.n_items = (int)(nu * nchunk);
            da.uniq_rows = c->dd_uniq.p + u0; da.qlist = c->dd_qlist.p; da.q_perm = c->dd_qperm.p; da.perm_off = c->d_perm_off.p; da.sc_off = c->d_sc_off.p;
            da.K = K; da.hit_rows = d_hits; da.sx = c->sx.p; da.sxg = c->sxg.p; da.icol = c->icol.p; da.rawsum_pos = c->rawsum_pos.p; da.N = N;
            da.n_degenerate = c->d_ndeg.p; da.anticor = prm->anticorrelation; da.scores = d_sc; da.counts = d_cn; da.perm_scores = perm_scores;
            da.Tcap = kRankTcap; da.Qcap = Qcap; da.bm_words = bm_words_rank; da.claim = rclaim;
            da.R = std::min(1.0, std::max(0.05, 8.0 / std::sqrt((double)std::max(c->n, 2)))); da.g_rec = reinterpret_cast<StatRec *>(c->dd_gthr.p);
            da.score_tab = c->dd_tab.p; da.phase_clk = nullptr; da.rank_out = c->rank_strip.p;
            COEX_TRY(stage_begin(c, ST_COUNT));
            count_dedup_kernel<2><<<std::min(rank_grid_max, std::max(da.n_items, 1)), kDdThreads, rsmem, st>>>(da);
            COEX_KCHECK(st);
            RankScoreArgs rs{};
            rs.strip = c->strip.p; rs.rank = c->rank_strip.p; rs.ld = ld; rs.uniq_rows = c->dd_uniq.p + u0; rs.ustart = c->pl_ustart.p + u0;
            rs.qlist = c->dd_qlist.p; rs.q_perm = c->dd_qperm.p; rs.u_count = nu; rs.perm_off = c->d_perm_off.p; rs.sc_off = c->d_sc_off.p;
            rs.hit_rows = d_hits; rs.sx = c->sx.p; rs.sxg = c->sxg.p; rs.rawsum_pos = c->rawsum_pos.p; rs.N = N; rs.anticor = prm->anticorrelation;
            rs.scores = d_sc; rs.counts = d_cn; rs.perm_scores = perm_scores; rs.score_tab = c->dd_tab.p;
            rank_score_kernel<<<(unsigned)std::min<long long>(H, (long long)c->sm_count * 16), 256, 0, st>>>(rs);
            COEX_KCHECK(st);
            COEX_TRY(stage_end(c, 2));
            continue;
        }
        const DedupItem *items_s = reinterpret_cast<const DedupItem *>(c->dd_items.p) + strip_item_begin[s_idx];
        const int n_items_s = (int)(strip_item_begin[s_idx + 1] - strip_item_begin[s_idx]);
        const int grid = std::min(grid_max, std::max(n_items_s, 1));
        const double R = std::min(1.0, std::max(0.05, 8.0 / std::sqrt((double)std::max(c->n, 2))));
        COEX_TRY(stage_begin(c, ST_COUNT));
        {
            DedupArgs da{};
            da.strip = c->strip.p; da.ld = ld; da.items = items_s; da.n_items = n_items_s;
            da.uniq_rows = c->dd_uniq.p + u0; da.qlist = c->dd_qlist.p; da.q_perm = c->dd_qperm.p;
            da.perm_off = c->d_perm_off.p; da.sc_off = c->d_sc_off.p; da.K = K; da.hit_rows = d_hits;
            da.sx = c->sx.p; da.sxg = c->sxg.p; da.icol = c->icol.p; da.rawsum_pos = c->rawsum_pos.p; da.N = N;
            da.n_degenerate = c->d_ndeg.p; da.anticor = prm->anticorrelation; da.scores = d_sc; da.counts = d_cn; da.perm_scores = perm_scores;
            da.Tcap = Tcap; da.Qcap = Qcap; da.bm_words = bm_words; da.claim = claim;
            da.R = R; da.g_rec = reinterpret_cast<StatRec *>(c->dd_gthr.p);
            da.score_tab = c->dd_tab.p; da.phase_clk = dbg ? phase.p : nullptr;
            if (!getenv("COEX_STATIC_ITEMS")) {
                COEX_TRY(c->dd_work.reserve(1));
                COEX_CUDA(cudaMemsetAsync(c->dd_work.p, 0, sizeof(int), st));
                da.work_counter = c->dd_work.p;
            }
            if (pearson) {
                da.eps_row = c->z_valid.p;
                da.strip_f32 = c->strip_f32.p; da.raw = c->raw; da.raw_sum = c->raw_sum.p; da.raw_mean = c->raw_mean.p;
                da.raw_xx = c->raw_xx.p; da.n = c->n; da.icol = c->z_valid.p; da.n_degenerate = c->z_ninv.p;
                count_dedup_kernel<1><<<grid, kDdThreads, smem, st>>>(da);
            } else if (fine) {
                // dd_redo: [0] number of declined items, [1] work counter of the re-do launch, [2 ..] the declined items
                COEX_CUDA(cudaMemsetAsync(c->dd_redo.p, 0, 2 * sizeof(int), st));
                DedupArgs fa = da;
                if (const char *e = getenv("COEX_FINE_DECLINE_EVERY")) fa.fine_decline_mod = atoi(e);      // (tests: the re-do list)
                fa.fine_nb = fine_nb; fa.fine_chunk = fine_chunk; fa.redo_count = c->dd_redo.p; fa.redo_list = c->dd_redo.p + 2;
                fa.R = std::min(0.5, std::max(0.02, fine_sigmas / std::sqrt((double)std::max(c->n, 2))));      // half-width of the map's centre
                fa.fine_map.r1 = (float)fa.R; fa.fine_map.half = (float)(fine_nb / 2); fa.fine_map.nbm1 = fine_nb - 1;
                fa.fine_map.st = (float)(0.125 * fine_nb / (1.0 - fa.R));
                fa.fine_map.k = (float)(0.375 * fine_nb / fa.R) - fa.fine_map.st;
                count_fine_kernel<<<std::min(c->sm_count, std::max(n_items_s, 1)), kCfThreads, fsmem, st>>>(fa);
                COEX_KCHECK(st);
                da.redo_count = c->dd_redo.p; da.redo_list = c->dd_redo.p + 2; da.work_counter = c->dd_redo.p + 1; da.phase_clk = nullptr;
                count_dedup_kernel<0><<<std::min(grid, 2 * c->sm_count), kDdThreads, smem, st>>>(da);
            } else {
                count_dedup_kernel<0><<<grid, kDdThreads, smem, st>>>(da);
            }
        }
        COEX_KCHECK(st);
        COEX_TRY(stage_end(c, fine ? 2 : 1));
    }
    if (dbg) {
        COEX_CUDA(cudaStreamSynchronize(st));
        unsigned long long h[16];
        COEX_CUDA(cudaMemcpy(h, phase.p, sizeof h, cudaMemcpyDeviceToHost));
        fprintf(stderr, "[coex count] items=%llu queued=%llu (beyond the queue %llu) cycles/item: meta=%llu stats=%llu rank=%llu stream=%llu refine=%llu score=%llu (Tcap=%d Qcap=%d grid<=%d)\n",
                h[6], h[7], h[8], h[0] / std::max(1ULL, h[6]), h[1] / std::max(1ULL, h[6]), h[2] / std::max(1ULL, h[6]), h[3] / std::max(1ULL, h[6]),
                h[4] / std::max(1ULL, h[6]), h[5] / std::max(1ULL, h[6]), Tcap, Qcap, grid_max);
        phase.release();
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// correlation strip: S[q, c] for queries q0..q0+nq of d_hits against the whole table
// ------------------------------------------------------------------------------------------
static int run_strip_gemm(coex_ctx *c, const int *d_rows, long long nq, long long nq_pad) {
    COEX_TRY(stage_begin(c, ST_GATHER));
    {
        const long long total = nq_pad * (c->n_pad >> 4);
        const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)c->sm_count * 16);
        gather_rows_kernel<<<blocks, 256, 0, c->stream>>>(c->u_hi.p, c->u_lo.p, c->n_pad, d_rows, nq, nq_pad,
                                                         c->a_hi.p, c->a_lo.p);
        COEX_CUDA(cudaGetLastError());
    }
    COEX_TRY(stage_end(c, 1));
    COEX_TRY(stage_begin(c, ST_GEMM));
    if (c->gemm_mode == 0) {
        COEX_TRY(tc_gemm_i8(c, nq_pad));
    } else {
        dim3 grid((unsigned)(c->N_pad / kSimtTile), (unsigned)(nq_pad / kSimtTile));
        gemm_i8_simt_kernel<<<grid, 256, 0, c->stream>>>(c->a_hi.p, c->a_lo.p, c->u_hi.p, c->u_lo.p, c->n_pad,
                                                         c->strip.p, c->N_pad);
        COEX_CUDA(cudaGetLastError());
    }
    COEX_TRY(stage_end(c, 1));
    return 0;
}

__global__ void gather_f64_kernel(const double *__restrict__ src, const int *__restrict__ rows, long long nq, long long nq_pad,
                                  double *__restrict__ dst) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nq_pad) dst[t] = (t < nq) ? src[rows[t]] : 0.0;
}

// Pearson r strip of arbitrary query rows on the tensor cores: gather their digit planes, D x D int8 products
static int run_strip_gemm_pearson(coex_ctx *c, const int *d_rows, long long nq, long long nq_pad, long long a_rows) {
    const int D = c->pearson_digits;
    const size_t plane = (size_t)c->N_pad * c->n_pad, a_plane = (size_t)a_rows * c->n_pad;
    COEX_TRY(stage_begin(c, ST_GATHER));
    const long long total = nq_pad * (c->n_pad >> 4);
    const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)c->sm_count * 16);
    for (int d = 0; d < D; d += 2) {
        const int d2 = std::min(d + 1, D - 1);
        gather_rows_kernel<<<blocks, 256, 0, c->stream>>>(c->z_planes.p + d * plane, c->z_planes.p + d2 * plane, c->n_pad, d_rows, nq, nq_pad,
                                                         c->za_planes.p + d * a_plane, c->za_planes.p + d2 * a_plane);
        COEX_CUDA(cudaGetLastError());
    }
    gather_f64_kernel<<<(unsigned)((nq_pad + 255) / 256), 256, 0, c->stream>>>(c->z_inv_s.p, d_rows, nq, nq_pad, c->za_inv_s.p);
    COEX_CUDA(cudaGetLastError());
    COEX_TRY(stage_end(c, (D + 1) / 2 + 1));
    COEX_TRY(stage_begin(c, ST_GEMM));
    if (D == 3)
        COEX_TRY(tc_gemm_digits_ab<3>(c, c->za_planes.p, (long long)a_plane, a_rows, c->za_inv_s.p, c->z_planes.p, (long long)plane, 0, nq_pad,
                                      nullptr, c->N_pad, c->z_inv_s.p, 0, c->strip_f32.p));
    else
        COEX_TRY(tc_gemm_digits_ab<4>(c, c->za_planes.p, (long long)a_plane, a_rows, c->za_inv_s.p, c->z_planes.p, (long long)plane, 0, nq_pad,
                                      nullptr, c->N_pad, c->z_inv_s.p, 0, c->strip_f32.p));
    COEX_TRY(stage_end(c, 1));
    return 0;
}

static int network_batch_impl(coex_ctx *c, const coex_params *prm, int K, const long long *perm_off,
                              const int *hit_rows, int *n_groups, int *group_of_hit, int *group_label,
                              int *group_winner, double *group_density, double *hit_score, double *scores,
                              long long *counts, int on_device) {
    if (!c->prepared) COEX_FAIL("coex_network_batch: call coex_prepare first");
    if (!c->have_features) COEX_FAIL("coex_network_batch: call coex_set_features first");
    if (c->glist_len <= 0) COEX_FAIL("coex_network_batch: call coex_set_global_list first");
    if (K <= 0) return 0;
    if (perm_off[0] != 0) COEX_FAIL("perm_off[0] must be 0");
    const long long H = perm_off[K];
    std::vector<long long> sc_off(K + 1, 0);
    int Pmax = 0;
    for (int k = 0; k < K; ++k) {
        const long long P = perm_off[k + 1] - perm_off[k];
        if (P < 0 || P > c->N) COEX_FAIL("permutation with an invalid number of hits");
        if (P > 8192) COEX_FAIL("more than 8192 hit regions in one permutation is not supported");
        Pmax = std::max<int>(Pmax, (int)P);
        sc_off[k + 1] = sc_off[k] + P * P;
    }
    const long long SC = sc_off[K];
    c->last_sc_off.clear();
    if (prm->measure == 0 && prm->use_specific_p && c->n > 1860)
        COEX_FAIL("Spearman tensor path supports up to 1860 samples (int32 dot products)");
    const int phase = c->shard_phase;              // 0 = plain call; 1 / 2 = the two halves of a job shared by several GPUs
    if (phase == 1) {
        if (!prm->use_specific_p || c->count_mode == 1 || (prm->measure == 1 && (c->pearson_mode != 0 || c->n > 1860)))
            COEX_FAIL("a job shared by several GPUs runs the de-duplicated node-specific count path only");
        if (counts) COEX_FAIL("bisect indices are not returned by a job shared by several GPUs");
    }
    if (phase == 2 && SC > (long long)c->shard_scores.cap) COEX_FAIL("coex_shard_finish: score buffer smaller than this rank's permutations");
    if (c->shard_phase != 2) stage_reset_from(c, ST_GATHER);      // (the second half of a shared job keeps the first half's stage times)
    cudaStream_t st = c->stream;

    // ---- buffers ----------------------------------------------------------------------------
    COEX_TRY(c->d_perm_off.reserve(K + 1)); COEX_TRY(c->d_sc_off.reserve(K + 1));
    COEX_CUDA(cudaMemcpyAsync(c->d_perm_off.p, perm_off, (K + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    COEX_CUDA(cudaMemcpyAsync(c->d_sc_off.p, sc_off.data(), (K + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    // sc_off is a stack-lifetime vector: the copy above must finish before we return
    const size_t Hs_alloc = (size_t)std::max<long long>(H, 1);
    const int *d_hits = hit_rows;
    int *d_ng = n_groups, *d_grp = group_of_hit, *d_lab = group_label, *d_win = group_winner;
    double *d_dens = group_density, *d_aps = hit_score, *d_sc = scores;
    long long *d_cn = counts;
    if (!on_device) {
        COEX_TRY(c->d_hits.reserve(Hs_alloc));
        COEX_CUDA(cudaMemcpyAsync(c->d_hits.p, hit_rows, H * sizeof(int), cudaMemcpyHostToDevice, st));
        d_hits = c->d_hits.p;
        COEX_TRY(c->d_ngroups.reserve(K)); COEX_TRY(c->d_grp.reserve(Hs_alloc)); COEX_TRY(c->d_label.reserve(Hs_alloc));
        COEX_TRY(c->d_win.reserve(Hs_alloc)); COEX_TRY(c->d_dens.reserve(Hs_alloc)); COEX_TRY(c->d_aps.reserve(Hs_alloc));
        d_ng = c->d_ngroups.p; d_grp = c->d_grp.p; d_lab = c->d_label.p; d_win = c->d_win.p;
        d_dens = c->d_dens.p; d_aps = c->d_aps.p;
        if (counts) { COEX_TRY(c->d_counts.reserve(std::max<long long>(SC, 1))); d_cn = c->d_counts.p; }
        d_sc = nullptr;
    }
    if (phase == 2) d_sc = c->shard_scores.p;      // written by every rank's count stage (coex_shard_count), complete after the caller's barrier
    else if (phase == 1) d_sc = nullptr;           // the score rows go to their owners through c->d_perm_scores
    else if (!d_sc) { COEX_TRY(c->d_scores.reserve(std::max<long long>(SC, 1))); d_sc = c->d_scores.p; }
    COEX_TRY(c->w_root.reserve(Hs_alloc)); COEX_TRY(c->w_pos.reserve(Hs_alloc)); COEX_TRY(c->w_s1.reserve(Hs_alloc));
    COEX_TRY(c->w_d0.reserve(Hs_alloc)); COEX_TRY(c->w_w.reserve(Hs_alloc)); COEX_TRY(c->w_best.reserve(Hs_alloc));
    COEX_TRY(c->w_cand.reserve(Hs_alloc)); COEX_TRY(c->w_state.reserve(K + 1));
    COEX_TRY(c->w_gsize.reserve(Hs_alloc)); COEX_TRY(c->w_cursor.reserve(Hs_alloc)); COEX_TRY(c->w_act.reserve(Hs_alloc));
    COEX_TRY(c->w_actn.reserve(K)); COEX_TRY(c->w_front.reserve(K));
    COEX_CUDA(cudaMemsetAsync(c->w_state.p, 0, (K + 1) * sizeof(int), st));
    // outputs of unused slots are defined
    COEX_CUDA(cudaMemsetAsync(d_grp, 0xff, H * sizeof(int), st));
    COEX_CUDA(cudaMemsetAsync(d_lab, 0xff, H * sizeof(int), st));
    COEX_CUDA(cudaMemsetAsync(d_win, 0xff, H * sizeof(int), st));
    COEX_CUDA(cudaMemsetAsync(d_dens, 0, H * sizeof(double), st));
    COEX_CUDA(cudaMemsetAsync(d_aps, 0, H * sizeof(double), st));

    const int T2max = next_pow2(std::max(Pmax, 2));
    if (phase == 1) {
        COEX_TRY(c->d_perm_scores.reserve((size_t)K));
        COEX_CUDA(cudaMemcpyAsync(c->d_perm_scores.p, c->h_perm_scores.data(), (size_t)K * sizeof(double *), cudaMemcpyHostToDevice, st));
    }
    if (H > 0 && Pmax >= 2 && phase != 2) {
        if (prm->use_specific_p) {
            // ---- node-specific null: strips of query rows x all table columns ---------------
            const long long ld = (prm->measure == 0) ? c->N_pad : c->N;
            const long long elt = (prm->measure == 0) ? 4 : 8;
            long long Hs = (c->strip_mb << 20) / (ld * elt);
            Hs = std::max<long long>(256, Hs / 256 * 256);
            Hs = std::min<long long>(Hs, round_up(H, 256));
            // -cm Pearson on the tensor cores: de-duplicated like Spearman, the tensor r is a filter and the reference's
            // in-order arithmetic decides every value it cannot (count mode 1 / pearson mode 1 keep the exact strip)
            // (rows of more than 1860 samples: the int32 Spearman strip the de-duplicated path reads its statistics from could
            //  overflow, so such tables take the exact strip + stat_rank_kernel, whose sums are 64-bit)
            const bool pearson_dedup = (prm->measure == 1 && c->count_mode != 1 && c->pearson_mode == 0 && c->n <= 1860);
            if (prm->measure == 0) {
                if (c->count_mode == 1) {
                    COEX_TRY(c->strip.reserve((size_t)Hs * ld));
                    COEX_TRY(c->a_hi.reserve((size_t)Hs * c->n_pad)); COEX_TRY(c->a_lo.reserve((size_t)Hs * c->n_pad));
                }
            } else {
                if (!pearson_dedup) {
                    COEX_TRY(c->strip_f64.reserve((size_t)Hs * ld));
                    StatArgs sa{c->d_perm_off.p, c->d_sc_off.p, d_hits, c->rank2.p, c->n, c->sx.p, c->rawsum_pos.p, d_sc};
                    COEX_TRY(stage_begin(c, ST_COUNT));
                    dim3 g((unsigned)std::min<long long>(((long long)Pmax * Pmax * 32 + 255) / 256, 4096), (unsigned)K);
                    stat_rank_kernel<<<g, 256, 0, st>>>(sa);
                    COEX_CUDA(cudaGetLastError());
                    COEX_TRY(stage_end(c, 1));
                }
            }
            if ((prm->measure == 0 && c->count_mode != 1) || pearson_dedup) {   // (count mode 3 with -cm Pearson: plain de-duplicated path)
                COEX_TRY(spearman_dedup_path(c, prm, K, perm_off, d_hits, Pmax, d_sc, d_cn));
            } else {
            const size_t smem = (size_t)T2max * 16;
            COEX_TRY(ensure_score_table(c, prm));
            COEX_CUDA(cudaFuncSetAttribute(count_score_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            COEX_CUDA(cudaFuncSetAttribute(count_score_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            for (long long q0 = 0; q0 < H; q0 += Hs) {
                const long long nq = std::min(Hs, H - q0);
                CountArgs ca{};
                ca.ld = ld; ca.q0 = q0; ca.nq = nq; ca.perm_off = c->d_perm_off.p; ca.sc_off = c->d_sc_off.p;
                ca.K = K; ca.hit_rows = d_hits; ca.sx = c->sx.p; ca.rawsum_pos = c->rawsum_pos.p; ca.N = c->N;
                ca.anticor = prm->anticorrelation; ca.scores = d_sc; ca.counts = d_cn; ca.T2max = T2max;
                ca.score_tab = c->dd_tab.p;
                if (prm->measure == 0) {
                    COEX_TRY(run_strip_gemm(c, d_hits + q0, nq, round_up(nq, 256)));
                    ca.strip = c->strip.p;
                    COEX_TRY(stage_begin(c, ST_COUNT));
                    count_score_kernel<false><<<(unsigned)nq, kCountThreads, smem, st>>>(ca);
                } else {
                    COEX_TRY(stage_begin(c, ST_GEMM));
                    dim3 g((unsigned)((c->N + 63) / 64), (unsigned)((nq + 15) / 16));
                    pearson_strip_kernel<<<g, 256, 0, st>>>(c->raw, c->n, c->N, d_hits + q0, nq, c->raw_sum.p,
                                                            c->raw_mean.p, c->raw_xx.p, c->strip_f64.p, ld);
                    COEX_CUDA(cudaGetLastError());
                    COEX_TRY(stage_end(c, 1));
                    ca.strip_f64 = c->strip_f64.p;
                    COEX_TRY(stage_begin(c, ST_COUNT));
                    count_score_kernel<true><<<(unsigned)nq, kCountThreads, smem, st>>>(ca);
                }
                COEX_CUDA(cudaGetLastError());
                COEX_TRY(stage_end(c, 1));
            }
            }
        } else {
            // ---- nsp == 0: whole-distribution p from the global list -------------------------
            StatArgs sa{c->d_perm_off.p, c->d_sc_off.p, d_hits, c->rank2.p, c->n, c->sx.p, c->rawsum_pos.p, d_sc};
            COEX_TRY(stage_begin(c, ST_COUNT));
            dim3 g((unsigned)std::min<long long>(((long long)Pmax * Pmax * 32 + 255) / 256, 4096), (unsigned)K);
            stat_rank_kernel<<<g, 256, 0, st>>>(sa);
            COEX_CUDA(cudaGetLastError());
            dim3 g2((unsigned)std::min<long long>(((long long)Pmax * Pmax + 255) / 256, 4096), (unsigned)K);
            global_from_stat_kernel<<<g2, 256, 0, st>>>(c->d_perm_off.p, c->d_sc_off.p, c->glist.p, c->glist_len,
                                                       prm->anticorrelation, d_sc, d_cn);
            COEX_CUDA(cudaGetLastError());
            COEX_TRY(stage_end(c, 2));
        }
    }
    // coex_set_score_readback: the score block of one permutation (the real data's) starts its way to the host NOW, on a second
    // stream, under the network stage
    bool readback = false;
    if (phase == 0 && c->rb_host && c->rb_k >= 0 && c->rb_k < K && H > 0) {
        const long long a = sc_off[(size_t)c->rb_k], b = sc_off[(size_t)c->rb_k + 1];
        if (b > a) {
            if (!c->rb_stream) { COEX_CUDA(cudaStreamCreateWithFlags(&c->rb_stream, cudaStreamNonBlocking)); COEX_CUDA(cudaEventCreateWithFlags(&c->rb_event, cudaEventDisableTiming)); }
            COEX_CUDA(cudaEventRecord(c->rb_event, st));
            COEX_CUDA(cudaStreamWaitEvent(c->rb_stream, c->rb_event, 0));
            COEX_CUDA(cudaMemcpyAsync(c->rb_host, d_sc + a, (size_t)(b - a) * sizeof(double), cudaMemcpyDeviceToHost, c->rb_stream));
            readback = true;
        }
    }
    if (phase == 1) {
        // every score row of my table rows is on its way to the rank that owns its permutation (stores into peer memory over
        // NVLink): they have landed when this stream is idle; the caller's barrier then tells the owners
        COEX_CUDA(cudaStreamSynchronize(st));
        if (c->gemm_mode == 0 || prm->measure == 1) COEX_TRY(tc_check_error(c));
        return 0;
    }
    // ---- K4: groups, winners, densities -----------------------------------------------------
    NetArgs na{};
    {
        na.K = K; na.perm_off = c->d_perm_off.p; na.sc_off = c->d_sc_off.p; na.hit_rows = d_hits; na.scores = d_sc;
        na.N = c->N; na.n = c->n; na.raw = c->raw; na.raw_sum = c->raw_sum.p; na.raw_mean = c->raw_mean.p;
        na.raw_xx = c->raw_xx.p; na.rank2 = c->rank2.p; na.sx = c->sx.p; na.rawsum_pos = c->rawsum_pos.p;
        na.chrom_id = c->chrom_id.p; na.name_rank = c->name_rank.p; na.feat_start = c->feat_start.p;
        na.feat_end = c->feat_end.p; na.glist = c->glist.p; na.glist_len = c->glist_len;
        na.measure = prm->measure; na.anticor = prm->anticorrelation; na.noiter = prm->noiter;
        na.useaverage = prm->useaverage; na.pvtj = prm->pval_to_join; na.thr = prm->selectionthreshold;
        na.softsep = prm->soft_separation;
        na.n_groups = d_ng; na.group_of_hit = d_grp; na.group_label = d_lab; na.group_winner = d_win;
        na.group_density = d_dens; na.hit_score = d_aps;
        na.w_root = c->w_root.p; na.w_pos = c->w_pos.p; na.w_s1 = c->w_s1.p; na.w_d0 = c->w_d0.p; na.w_w = c->w_w.p;
        na.w_best = c->w_best.p; na.w_cand = c->w_cand.p; na.w_state = c->w_state.p; na.w_flags = c->w_state.p + K;
        na.w_gsize = c->w_gsize.p; na.w_cursor = c->w_cursor.p; na.w_act = c->w_act.p; na.w_actn = c->w_actn.p; na.w_front = c->w_front.p;
        na.P2max = T2max;
    }
    std::vector<long long> warp_off(K + 1, 0);
    for (int k = 0; k < K; ++k) warp_off[k + 1] = warp_off[k] + std::max<long long>(1, (perm_off[k + 1] - perm_off[k] + kRsRows - 1) / kRsRows);
    COEX_TRY(c->d_warp_off.reserve(K + 1));
    COEX_CUDA(cudaMemcpyAsync(c->d_warp_off.p, warp_off.data(), (K + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    na.warp_off = c->d_warp_off.p;
    const unsigned rs_blocks = (unsigned)((warp_off[K] + kRsWarps - 1) / kRsWarps), rs_threads = kRsWarps * 32;
    auto pass2_round = [&](bool last) -> int {
        net_rowsum_kernel<1><<<rs_blocks, rs_threads, 0, st>>>(na);
        COEX_KCHECK(st);
        if (last) COEX_CUDA(cudaMemsetAsync(na.w_flags, 0, sizeof(int), st));
        net_commit_kernel<<<K, 256, 0, st>>>(na, 0);
        COEX_KCHECK(st);
        return 0;
    };
    auto density_tail = [&]() -> int {
        const size_t rsmem = (size_t)T2max * 8;
        net_rowsum_kernel<4><<<rs_blocks, rs_threads, 0, st>>>(na);      // allpromoterscores against the final winners
        COEX_KCHECK(st);
        net_rowsum_kernel<2><<<rs_blocks, rs_threads, 0, st>>>(na);
        COEX_KCHECK(st);
        net_removal_kernel<<<K, 256, rsmem, st>>>(na, 0);
        COEX_KCHECK(st);
        if (!prm->noiter) { net_rowsum_kernel<3><<<rs_blocks, rs_threads, 0, st>>>(na); COEX_KCHECK(st); }
        net_removal_kernel<<<K, 256, rsmem, st>>>(na, 1);
        COEX_KCHECK(st);
        return 0;
    };
    const int kRounds = 12;
    long long net_launches = 0;
    if (H > 0) {
        const size_t gsmem = (size_t)T2max * 24, lsmem = (size_t)T2max * 16;
        COEX_CUDA(cudaFuncSetAttribute(net_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
        COEX_CUDA(cudaFuncSetAttribute(net_label_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem));
        COEX_CUDA(cudaFuncSetAttribute(net_removal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(T2max * 8)));
        COEX_TRY(stage_begin(c, ST_NETWORK));
        net_sort_kernel<<<K, kNetThreads, gsmem, st>>>(na);
        COEX_KCHECK(st);
        net_pairs_kernel<<<(unsigned)((H * 32 + 255) / 256), 256, 0, st>>>(na, H);
        COEX_KCHECK(st);
        net_label_kernel<<<K, kNetThreads, lsmem, st>>>(na);
        COEX_KCHECK(st);
        net_rowsum_kernel<0><<<rs_blocks, rs_threads, 0, st>>>(na);
        COEX_KCHECK(st);
        net_commit_kernel<<<K, 256, 0, st>>>(na, 1);
        COEX_KCHECK(st);
        for (int r = 0; r < kRounds; ++r) COEX_TRY(pass2_round(r == kRounds - 1));
        COEX_TRY(density_tail());
        net_launches = 5 + 2 * kRounds + (prm->noiter ? 4 : 5);
        COEX_TRY(stage_end(c, net_launches));
    } else {
        COEX_CUDA(cudaMemsetAsync(d_ng, 0, K * sizeof(int), st));
    }
    if (!on_device) {
        COEX_CUDA(cudaMemcpyAsync(n_groups, d_ng, K * sizeof(int), cudaMemcpyDeviceToHost, st));
        COEX_CUDA(cudaMemcpyAsync(group_of_hit, d_grp, H * sizeof(int), cudaMemcpyDeviceToHost, st));
        COEX_CUDA(cudaMemcpyAsync(group_label, d_lab, H * sizeof(int), cudaMemcpyDeviceToHost, st));
        COEX_CUDA(cudaMemcpyAsync(group_winner, d_win, H * sizeof(int), cudaMemcpyDeviceToHost, st));
        COEX_CUDA(cudaMemcpyAsync(group_density, d_dens, H * sizeof(double), cudaMemcpyDeviceToHost, st));
        COEX_CUDA(cudaMemcpyAsync(hit_score, d_aps, H * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (scores) COEX_CUDA(cudaMemcpyAsync(scores, d_sc, SC * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (counts) COEX_CUDA(cudaMemcpyAsync(counts, d_cn, SC * sizeof(long long), cudaMemcpyDeviceToHost, st));
    }
    // perm_off / sc_off staging copies read host memory that dies with this frame
    COEX_CUDA(cudaStreamSynchronize(st));
    if (H > 0) {
        // pass 2 is a fixed-point iteration: kRounds rounds were queued without looking; in the rare
        // case that some permutation was still moving, iterate to convergence and redo the tail.
        int flag = 0;
        COEX_CUDA(cudaMemcpy(&flag, na.w_flags, sizeof(int), cudaMemcpyDeviceToHost));
        if (flag) {
            for (int guard = 0; flag && guard < 8192 + 2; ++guard) {
                COEX_TRY(pass2_round(true));
                c->launches += 2;
                COEX_CUDA(cudaStreamSynchronize(st));
                COEX_CUDA(cudaMemcpy(&flag, na.w_flags, sizeof(int), cudaMemcpyDeviceToHost));
            }
            COEX_TRY(density_tail());
            c->launches += 5;
            if (!on_device) {
                COEX_CUDA(cudaMemcpyAsync(group_winner, d_win, H * sizeof(int), cudaMemcpyDeviceToHost, st));
                COEX_CUDA(cudaMemcpyAsync(group_density, d_dens, H * sizeof(double), cudaMemcpyDeviceToHost, st));
                COEX_CUDA(cudaMemcpyAsync(hit_score, d_aps, H * sizeof(double), cudaMemcpyDeviceToHost, st));
            }
            COEX_CUDA(cudaStreamSynchronize(st));
        }
    }
    if (prm->use_specific_p && (c->gemm_mode == 0 || prm->measure == 1)) COEX_TRY(tc_check_error(c));   // (-cm Pearson runs the tcgen05 digit GEMM in every gemm mode)
    c->last_sc_off = sc_off; c->last_scores = d_sc;           // coex_read_scores
    if (readback) COEX_CUDA(cudaStreamSynchronize(c->rb_stream));
    return 0;
}

// ------------------------------------------------------------------------------------------
// on-device cross-check of the tcgen05 GEMM against the dp4a GEMM on the resident table
// ------------------------------------------------------------------------------------------
__global__ void compare_strips_kernel(const int *a, const int *b, long long rows, long long cols, long long ld,
                                      unsigned long long *nbad, long long *first) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < rows * cols;
         t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / cols, col = t - r * cols;
        if (a[r * ld + col] != b[r * ld + col]) {
            if (atomicAdd(nbad, 1ULL) == 0) { first[0] = r; first[1] = col; first[2] = a[r * ld + col]; first[3] = b[r * ld + col]; }
        }
    }
}

static int selftest_gemm_impl(coex_ctx *c, long long m_rows, long long *mismatches, long long *first4) {
    if (!c->prepared) COEX_FAIL("coex_selftest_gemm: call coex_prepare first");
    m_rows = std::min(round_up(std::max<long long>(m_rows, 1), 256), c->N_pad);
    const long long nq = std::min(m_rows, c->N);
    cudaStream_t st = c->stream;
    COEX_TRY(c->d_hits.reserve(m_rows));
    COEX_TRY(c->strip.reserve((size_t)m_rows * c->N_pad));
    COEX_TRY(c->a_hi.reserve((size_t)m_rows * c->n_pad)); COEX_TRY(c->a_lo.reserve((size_t)m_rows * c->n_pad));
    DevBuf<int> ref;
    DevBuf<unsigned long long> bad;
    DevBuf<long long> first;
    COEX_TRY(ref.reserve((size_t)m_rows * c->N_pad)); COEX_TRY(bad.reserve(1)); COEX_TRY(first.reserve(4));
    COEX_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(unsigned long long), st));
    COEX_CUDA(cudaMemsetAsync(first.p, 0xff, 4 * sizeof(long long), st));
    COEX_CUDA(cudaMemsetAsync(c->strip.p, 0x5a, (size_t)m_rows * c->N_pad * sizeof(int), st));
    iota_rows_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(c->d_hits.p, 0, nq);
    const long long total = m_rows * (c->n_pad >> 4);
    gather_rows_kernel<<<(int)std::min<long long>((total + 255) / 256, 4096), 256, 0, st>>>(
        c->u_hi.p, c->u_lo.p, c->n_pad, c->d_hits.p, nq, m_rows, c->a_hi.p, c->a_lo.p);
    COEX_CUDA(cudaGetLastError());
    COEX_TRY(tc_gemm_i8(c, m_rows));
    dim3 grid((unsigned)(c->N_pad / kSimtTile), (unsigned)(m_rows / kSimtTile));
    gemm_i8_simt_kernel<<<grid, 256, 0, st>>>(c->a_hi.p, c->a_lo.p, c->u_hi.p, c->u_lo.p, c->n_pad, ref.p, c->N_pad);
    COEX_CUDA(cudaGetLastError());
    compare_strips_kernel<<<1024, 256, 0, st>>>(c->strip.p, ref.p, m_rows, c->N_pad, c->N_pad, bad.p, first.p);
    COEX_CUDA(cudaGetLastError());
    unsigned long long hb = 0;
    COEX_CUDA(cudaMemcpyAsync(&hb, bad.p, sizeof hb, cudaMemcpyDeviceToHost, st));
    COEX_CUDA(cudaMemcpyAsync(first4, first.p, 4 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    COEX_CUDA(cudaStreamSynchronize(st));
    *mismatches = (long long)hb;
    ref.release(); bad.release(); first.release();
    c->launches += 5;
    return tc_check_error(c);
}

}  // namespace coex

using namespace coex;

// ============================================================================================
// C ABI
// ============================================================================================
#pragma GCC visibility push(default)
extern "C" {

const char *coex_last_error(void) { return g_last_error.c_str(); }
int coex_abi_version(void) { return 2; }

int coex_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int coex_create(coex_ctx **out, int device) {
    if (!out) COEX_FAIL("coex_create: null output pointer");
    *out = nullptr;
    int ndev = 0;
    COEX_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) COEX_FAIL("coex_create: no such CUDA device (there is no CPU fallback)");
    COEX_CUDA(cudaSetDevice(device));
    coex_ctx *c = new coex_ctx();
    c->device = device;
    cudaDeviceProp prop;
    COEX_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    COEX_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (const char *e = getenv("COEX_STRIP_MB")) { const long long mb = atoll(e); if (mb >= 16) c->strip_mb = mb; }   // profiling aid: same as coex_set_workspace_mb
    if (const char *e = getenv("COEX_FINE_MIN_ROWS")) c->fine_min_rows = atoll(e);
    if (const char *e = getenv("COEX_COUNT_MODE")) c->count_mode = atoi(e);      // profiling aid (scripts/): same as coex_set_count_mode
    *out = c;
    return 0;
}

int coex_destroy(coex_ctx *c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    stage_reset(c);
    tc_destroy(c);
    c->raw_own.release(); c->raw_sum.release(); c->raw_mean.release(); c->raw_xx.release(); c->rawsum_pos.release();
    c->rank2.release(); c->u_hi.release(); c->u_lo.release(); c->suu.release(); c->sx.release();
    c->z_valid.release(); c->z_ninv.release(); c->za_planes.release(); c->za_inv_s.release();
    c->st_thr.release(); c->st_w.release(); c->st_cnt.release(); c->st_itemoff.release(); c->st_itemT.release();
    c->st_cursor.release(); c->sxg.release(); c->dd_work.release(); c->dd_redo.release();
    c->z_planes.release(); c->z_ok.release(); c->z_l1.release(); c->z_inv_s.release();
    c->ap_hist.release(); c->feat_key.release(); c->feat_row.release(); c->feat_wrank.release(); c->feat_wrow.release(); c->pm_blob.release(); c->pm_counts.release(); c->pm_hits_tmp.release(); c->pm_off.release();
    c->chrom_id.release(); c->name_rank.release(); c->feat_start.release(); c->feat_end.release(); c->glist.release();
    c->a_hi.release(); c->a_lo.release(); c->strip.release(); c->strip_f64.release(); c->strip_f32.release(); c->rank_strip.release();
    c->d_perm_off.release(); c->d_sc_off.release(); c->d_warp_off.release(); c->d_hits.release(); c->d_ngroups.release(); c->d_grp.release();
    c->d_label.release(); c->d_win.release(); c->w_root.release(); c->w_pos.release(); c->d_dens.release();
    c->d_aps.release(); c->d_scores.release(); c->w_s1.release(); c->w_d0.release(); c->w_w.release();
    c->d_counts.release(); c->w_gsize.release(); c->w_cursor.release(); c->w_act.release(); c->w_actn.release(); c->w_front.release(); c->w_best.release(); c->w_cand.release(); c->w_state.release(); c->p_idxa.release(); c->p_idxb.release();
    c->p_out.release(); c->icol.release(); c->d_ndeg.release(); c->dd_uniq.release(); c->dd_qlist.release();
    c->dd_items.release(); c->dd_qperm.release(); c->pl_cnt.release(); c->pl_cursor.release(); c->pl_ustart.release(); c->pl_nitems.release();
    c->pl_itemoff.release(); c->pl_bounds.release(); c->pl_meta.release(); c->pl_err.release(); c->dd_gthr.release(); c->dd_tab.release();
    if (c->rb_stream) { cudaStreamDestroy(c->rb_stream); cudaEventDestroy(c->rb_event); }
    if (c->up_stream) {
        cudaStreamSynchronize(c->up_stream);
        cudaStreamDestroy(c->up_stream); cudaEventDestroy(c->up_start);
        for (int k = 0; k < coex_ctx::kUpChunks; ++k) cudaEventDestroy(c->up_event[k]);
    }
    for (void *p : c->peer_mapped) if (p) cudaIpcCloseMemHandle(p);
    for (void *p : c->tab_mapped) if (p) cudaIpcCloseMemHandle(p);
    c->shard_scores.release(); c->d_perm_scores.release();
    cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int coex_sync(coex_ctx *c) {
    COEX_TRY(check_ctx(c));
    COEX_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

void *coex_stream(coex_ctx *c) { return c ? (void *)c->stream : nullptr; }

int coex_set_expression(coex_ctx *c, const double *data, long long N, int n, int on_device) {
    COEX_TRY(check_ctx(c));
    if (!data || N <= 0 || n <= 0) COEX_FAIL("coex_set_expression: empty table");
    if (N > 2000000000LL) COEX_FAIL("coex_set_expression: too many rows");
    c->prepared = false;
    c->z_digits = 0;
    c->N = N; c->n = n;
    if (c->up_pending) {                        // a pipelined upload nobody prepared: the compute stream gets behind it first
        COEX_CUDA(cudaStreamWaitEvent(c->stream, c->up_event[c->up_pending - 1], 0));
        c->up_pending = 0;
    }
    if (on_device == 2) {
        // page-locked host memory, pipelined with the rank transform: kUpChunks row chunks on a stream of their own (behind
        // everything queued on the compute stream, which may still read the previous table), an event per chunk
        COEX_TRY(c->raw_own.reserve((size_t)N * n));
        if (!c->up_stream) {
            COEX_CUDA(cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking));
            COEX_CUDA(cudaEventCreateWithFlags(&c->up_start, cudaEventDisableTiming));
            for (int k = 0; k < coex_ctx::kUpChunks; ++k) COEX_CUDA(cudaEventCreateWithFlags(&c->up_event[k], cudaEventDisableTiming));
        }
        COEX_CUDA(cudaEventRecord(c->up_start, c->stream));
        COEX_CUDA(cudaStreamWaitEvent(c->up_stream, c->up_start, 0));
        c->up_rows = (N + coex_ctx::kUpChunks - 1) / coex_ctx::kUpChunks;
        int k = 0;
        for (long long r0 = 0; r0 < N; r0 += c->up_rows, ++k) {
            const long long r1 = std::min(N, r0 + c->up_rows);
            COEX_CUDA(cudaMemcpyAsync(c->raw_own.p + (size_t)r0 * n, data + (size_t)r0 * n, (size_t)(r1 - r0) * n * sizeof(double),
                                      cudaMemcpyHostToDevice, c->up_stream));
            COEX_CUDA(cudaEventRecord(c->up_event[k], c->up_stream));
        }
        c->up_pending = k;
        c->raw = c->raw_own.p;
    } else if (on_device) {
        c->raw = data;
    } else {
        COEX_TRY(c->raw_own.reserve((size_t)N * n));
        COEX_CUDA(cudaMemcpyAsync(c->raw_own.p, data, (size_t)N * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        COEX_CUDA(cudaStreamSynchronize(c->stream));
        c->raw = c->raw_own.p;
    }
    return 0;
}

int coex_prepare(coex_ctx *c) {
    COEX_TRY(check_ctx(c));
    return prepare_impl(c);
}

int coex_get_ranks(coex_ctx *c, double *out) {
    COEX_TRY(check_ctx(c));
    if (!c->prepared) COEX_FAIL("coex_get_ranks: call coex_prepare first");
    std::vector<uint16_t> h((size_t)c->N * c->n);
    COEX_CUDA(cudaMemcpyAsync(h.data(), c->rank2.p, h.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
    COEX_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < h.size(); ++i) out[i] = 0.5 * (double)h[i];
    return 0;
}

int coex_set_features(coex_ctx *c, const int *chrom_id, const long long *start, const long long *end,
                      const int *name_rank, long long N) {
    COEX_TRY(check_ctx(c));
    if (N != c->N) COEX_FAIL("coex_set_features: N differs from the expression table");
    for (long long i = 0; i < N; ++i)
        if (start[i] < 0 || start[i] >= (1LL << 39) || end[i] < 0 || end[i] >= (1LL << 39) || chrom_id[i] < 0 ||
            chrom_id[i] >= (1 << 20))
            COEX_FAIL("coex_set_features: coordinate out of range");
    COEX_TRY(c->chrom_id.reserve(N)); COEX_TRY(c->name_rank.reserve(N));
    COEX_TRY(c->feat_start.reserve(N)); COEX_TRY(c->feat_end.reserve(N));
    COEX_CUDA(cudaMemcpyAsync(c->chrom_id.p, chrom_id, N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    COEX_CUDA(cudaMemcpyAsync(c->name_rank.p, name_rank, N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    COEX_CUDA(cudaMemcpyAsync(c->feat_start.p, start, N * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    COEX_CUDA(cudaMemcpyAsync(c->feat_end.p, end, N * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    {   // index for the locus -> promoter join: rows ordered by (chromosome id, start, row)
        std::vector<std::pair<unsigned long long, int>> order((size_t)N);
        long long maxlen = 0;
        for (long long i = 0; i < N; ++i) {
            order[(size_t)i] = {((unsigned long long)chrom_id[i] << 40) | (unsigned long long)start[i], (int)i};
            maxlen = std::max(maxlen, end[i] - start[i]);
        }
        std::sort(order.begin(), order.end());
        std::vector<unsigned long long> keys((size_t)N);
        std::vector<int> rows((size_t)N);
        for (long long i = 0; i < N; ++i) { keys[(size_t)i] = order[(size_t)i].first; rows[(size_t)i] = order[(size_t)i].second; }
        COEX_TRY(c->feat_key.reserve(N)); COEX_TRY(c->feat_row.reserve(N));
        COEX_CUDA(cudaMemcpyAsync(c->feat_key.p, keys.data(), N * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
        COEX_CUDA(cudaMemcpyAsync(c->feat_row.p, rows.data(), N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        c->feat_maxlen = maxlen;
        COEX_CUDA(cudaStreamSynchronize(c->stream));
    }
    c->h_chrom_id.assign(chrom_id, chrom_id + N); c->h_feat_start.assign(start, start + N); c->h_feat_end.assign(end, end + N);
    c->h_bed_index.resize((size_t)N);
    for (long long i = 0; i < N; ++i) c->h_bed_index[(size_t)i] = (int)i;       // default: the feature BED lists the rows in table order
    COEX_TRY(upload_window_order(c));
    COEX_CUDA(cudaStreamSynchronize(c->stream));
    c->have_features = true;
    return 0;
}

int coex_set_feature_order(coex_ctx *c, const int *bed_index, long long N) {
    COEX_TRY(check_ctx(c));
    if (!c->have_features) COEX_FAIL("coex_set_feature_order: call coex_set_features first");
    if (N != c->N || !bed_index) COEX_FAIL("coex_set_feature_order: one line number per table row");
    c->h_bed_index.assign(bed_index, bed_index + N);
    return upload_window_order(c);
}

int coex_permute_map(coex_ctx *c, int mode, const int *snp_chrom, const long long *snp_addr, const long long *snp_back,
                     int M, const long long *shifts, int K, const long long *chrom_start, const long long *chrom_end,
                     const int *chrom_fid, int n_chrom, const int *back_chrom, const long long *back_addr,
                     long long n_back, long long window, long long *perm_off, int *hit_rows, long long capacity,
                     int on_device, int *hit_loci) {
    COEX_TRY(check_ctx(c));
    if (!c->have_features) COEX_FAIL("coex_permute_map: call coex_set_features first");
    if (mode != 0 && mode != 1 && mode != 2) COEX_FAIL("coex_permute_map: mode 0 = circular, 1 = backcirc, 2 = background");
    if (M <= 0 || K <= 0 || n_chrom <= 0 || !snp_chrom || !snp_addr || !shifts || !chrom_start || !chrom_end || !chrom_fid ||
        !perm_off || (!hit_rows && capacity > 0))
        COEX_FAIL("coex_permute_map: null or empty argument");
    if (mode >= 1 && (!snp_back || !back_chrom || !back_addr || n_back <= 0)) COEX_FAIL("coex_permute_map: backcirc / background need the background");
    const size_t n_sb = (mode == 2) ? (size_t)K * M : (size_t)M;     // background: one drawn line per (set, locus)
    if (window < 0) COEX_FAIL("coex_permute_map: negative window");
    for (int m = 0; m < M; ++m)
        if (snp_chrom[m] < 0 || snp_chrom[m] >= n_chrom) COEX_FAIL("coex_permute_map: chromosome index out of range");
    cudaStream_t st = c->stream;
    // ---- stage the (small) inputs in one blob ------------------------------------------------------
    auto al = [](size_t x) { return (x + 15) / 16 * 16; };
    const size_t o_sc = 0, o_sa = al(o_sc + (size_t)M * 4), o_sb = al(o_sa + (size_t)M * 8), o_sh = al(o_sb + n_sb * 8),
                 o_cs = al(o_sh + (size_t)K * 8), o_ce = al(o_cs + (size_t)n_chrom * 8), o_cf = al(o_ce + (size_t)n_chrom * 8),
                 o_bc = al(o_cf + (size_t)n_chrom * 4), o_ba = al(o_bc + (size_t)std::max<long long>(n_back, 0) * 4),
                 total = al(o_ba + (size_t)std::max<long long>(n_back, 0) * 8);
    std::vector<unsigned char> blob(total, 0);
    memcpy(&blob[o_sc], snp_chrom, (size_t)M * 4); memcpy(&blob[o_sa], snp_addr, (size_t)M * 8);
    if (snp_back) memcpy(&blob[o_sb], snp_back, n_sb * 8);
    memcpy(&blob[o_sh], shifts, (size_t)K * 8);
    memcpy(&blob[o_cs], chrom_start, (size_t)n_chrom * 8); memcpy(&blob[o_ce], chrom_end, (size_t)n_chrom * 8);
    memcpy(&blob[o_cf], chrom_fid, (size_t)n_chrom * 4);
    if (mode >= 1) {
        for (long long i = 0; i < n_back; ++i)
            if (back_chrom[i] < 0 || back_chrom[i] >= n_chrom) COEX_FAIL("coex_permute_map: background chromosome index out of range");
        memcpy(&blob[o_bc], back_chrom, (size_t)n_back * 4); memcpy(&blob[o_ba], back_addr, (size_t)n_back * 8);
    }
    COEX_TRY(c->pm_blob.reserve(total));
    COEX_CUDA(cudaMemcpyAsync(c->pm_blob.p, blob.data(), total, cudaMemcpyHostToDevice, st));
    PermuteArgs pa{};
    pa.mode = mode; pa.M = M; pa.K = K;
    pa.snp_chrom = reinterpret_cast<const int *>(c->pm_blob.p + o_sc);
    pa.snp_addr = reinterpret_cast<const long long *>(c->pm_blob.p + o_sa);
    pa.snp_back = reinterpret_cast<const long long *>(c->pm_blob.p + o_sb);
    pa.shifts = reinterpret_cast<const long long *>(c->pm_blob.p + o_sh);
    pa.chrom_start = reinterpret_cast<const long long *>(c->pm_blob.p + o_cs);
    pa.chrom_end = reinterpret_cast<const long long *>(c->pm_blob.p + o_ce);
    pa.chrom_fid = reinterpret_cast<const int *>(c->pm_blob.p + o_cf);
    pa.n_chrom = n_chrom;
    pa.back_chrom = reinterpret_cast<const int *>(c->pm_blob.p + o_bc);
    pa.back_addr = reinterpret_cast<const long long *>(c->pm_blob.p + o_ba);
    pa.n_back = n_back; pa.window = window;
    pa.fkey = c->feat_key.p; pa.frow = c->feat_row.p; pa.fstart = c->feat_start.p; pa.fend = c->feat_end.p;
    pa.wrank = c->feat_wrank.p; pa.wrow = c->feat_wrow.p;
    pa.N = c->N; pa.maxlen = c->feat_maxlen;
    COEX_TRY(c->pm_counts.reserve((size_t)2 * K + 1));
    COEX_CUDA(cudaMemsetAsync(c->pm_counts.p, 0, ((size_t)2 * K + 1) * sizeof(int), st));
    pa.cand_count = c->pm_counts.p; pa.hit_count = c->pm_counts.p + K; pa.overflow = c->pm_counts.p + 2 * K;
    // ---- pass 1: candidate pairs per set -> shared-memory capacity ------------------------------------
    const long long KM = (long long)K * M;
    permute_count_kernel<<<(unsigned)((KM + 255) / 256), 256, 0, st>>>(pa);
    COEX_KCHECK(st);
    std::vector<int> cand((size_t)K);
    COEX_CUDA(cudaMemcpyAsync(cand.data(), pa.cand_count, (size_t)K * sizeof(int), cudaMemcpyDeviceToHost, st));
    COEX_CUDA(cudaStreamSynchronize(st));
    int cmax = 1;
    for (int k = 0; k < K; ++k) cmax = std::max(cmax, cand[(size_t)k]);
    const int cap = std::max(256, next_pow2(cmax));
    if (cap > 8192) COEX_FAIL("coex_permute_map: more than 8192 (locus, feature) pairs in one set");
    pa.cap = cap;
    COEX_TRY(c->pm_hits_tmp.reserve((size_t)K * cap * (hit_loci ? 3 : 1)));
    pa.hits_tmp = c->pm_hits_tmp.p;
    pa.loci_tmp = hit_loci ? c->pm_hits_tmp.p + (size_t)K * cap : nullptr;
    // ---- pass 2: hits of every set, then offsets and the packed list ------------------------------------
    const size_t smem = (size_t)2 * cap * sizeof(unsigned long long);
    COEX_CUDA(cudaFuncSetAttribute(permute_map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    permute_map_kernel<<<K, 256, smem, st>>>(pa);
    COEX_KCHECK(st);
    long long *d_off = perm_off;
    int *d_rows = hit_rows;
    if (!on_device) { COEX_TRY(c->pm_off.reserve((size_t)K + 1)); d_off = c->pm_off.p; }
    permute_scan_kernel<<<1, 32, 0, st>>>(pa.hit_count, K, d_off);
    COEX_KCHECK(st);
    std::vector<long long> h_off((size_t)K + 1);
    COEX_CUDA(cudaMemcpyAsync(h_off.data(), d_off, ((size_t)K + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
    COEX_CUDA(cudaStreamSynchronize(st));
    const long long H = h_off[(size_t)K];
    if (H > capacity) COEX_FAIL("coex_permute_map: hit_rows capacity too small (needs " + std::to_string(H) + ")");
    int *d_loci = hit_loci;
    if (!on_device) {
        COEX_TRY(c->d_hits.reserve((size_t)std::max<long long>(H, 1) * (hit_loci ? 3 : 1)));
        d_rows = c->d_hits.p;
        if (hit_loci) d_loci = c->d_hits.p + std::max<long long>(H, 1);
    }
    if (H > 0) {
        permute_compact_kernel<<<K, 256, 0, st>>>(pa.hits_tmp, pa.loci_tmp, cap, d_off, K, d_rows, d_loci);
        COEX_KCHECK(st);
    }
    c->launches += 4;
    if (!on_device) {
        memcpy(perm_off, h_off.data(), ((size_t)K + 1) * sizeof(long long));
        if (H > 0) COEX_CUDA(cudaMemcpyAsync(hit_rows, d_rows, (size_t)H * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (H > 0 && hit_loci) COEX_CUDA(cudaMemcpyAsync(hit_loci, d_loci, (size_t)H * 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    COEX_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int coex_set_global_list(coex_ctx *c, const double *v, long long len) {
    COEX_TRY(check_ctx(c));
    if (len <= 0) COEX_FAIL("coex_set_global_list: empty list");
    for (long long i = 1; i < len; ++i)
        if (v[i] < v[i - 1]) COEX_FAIL("coex_set_global_list: list is not ascending");
    COEX_TRY(c->glist.reserve(len));
    COEX_CUDA(cudaMemcpyAsync(c->glist.p, v, len * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    COEX_CUDA(cudaStreamSynchronize(c->stream));
    c->glist_len = len;
    return 0;
}

int coex_pairs(coex_ctx *c, int which, const int *idxa, const int *idxb, long long npairs, double *out,
               int on_device) {
    COEX_TRY(check_ctx(c));
    if (!c->prepared) COEX_FAIL("coex_pairs: call coex_prepare first");
    if (npairs <= 0) return 0;
    const int *da = idxa, *db = idxb;
    double *dout = out;
    if (!on_device) {
        COEX_TRY(c->p_idxa.reserve(npairs)); COEX_TRY(c->p_idxb.reserve(npairs)); COEX_TRY(c->p_out.reserve(npairs));
        COEX_CUDA(cudaMemcpyAsync(c->p_idxa.p, idxa, npairs * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        COEX_CUDA(cudaMemcpyAsync(c->p_idxb.p, idxb, npairs * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        da = c->p_idxa.p; db = c->p_idxb.p; dout = c->p_out.p;
    }
    if (which == 1) {
        const long long blocks = (npairs * 32 + 255) / 256;
        pairs_rank_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(c->u_hi.p, c->u_lo.p, c->n_pad, da, db, npairs, c->sx.p, dout);
    } else {
        const long long blocks = (npairs + kPairsPerBlock - 1) / kPairsPerBlock;
        pairs_exact_kernel<<<(unsigned)blocks, kPairsPerBlock, 0, c->stream>>>(c->raw, c->n, da, db, npairs, c->raw_sum.p,
                                                                              c->raw_mean.p, c->raw_xx.p, dout);
    }
    COEX_CUDA(cudaGetLastError());
    c->launches += 1;
    if (!on_device) {
        COEX_CUDA(cudaMemcpyAsync(out, dout, npairs * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        COEX_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int coex_allpairs_histogram(coex_ctx *c, int which, long long row0, long long row1, int nbins,
                            unsigned long long *hist, unsigned long long *nan_count) {
    COEX_TRY(check_ctx(c));
    if (!c->prepared) COEX_FAIL("coex_allpairs_histogram: call coex_prepare first");
    return allpairs_histogram_impl(c, which, row0, row1, nbins, hist, nan_count, 1);
}

int coex_allpairs_accumulate(coex_ctx *c, int which, long long row0, long long row1, int nbins, int reset) {
    COEX_TRY(check_ctx(c));
    if (!c->prepared) COEX_FAIL("coex_allpairs_accumulate: call coex_prepare first");
    return allpairs_histogram_impl(c, which, row0, row1, nbins, nullptr, nullptr, reset);
}

int coex_allpairs_read(coex_ctx *c, unsigned long long *hist, unsigned long long *nan_count, void *device_copy) {
    COEX_TRY(check_ctx(c));
    return allpairs_read_impl(c, hist, nan_count, device_copy);
}

int coex_corr_block(coex_ctx *c, int which, long long row0, long long nrows, double *out) {
    COEX_TRY(check_ctx(c));
    if (!c->prepared) COEX_FAIL("coex_corr_block: call coex_prepare first");
    if (!out) COEX_FAIL("coex_corr_block: null output");
    return corr_block_impl(c, which, row0, nrows, out);
}

int coex_set_pearson(coex_ctx *c, int digits, int mode) {
    if (!c) COEX_FAIL("null context");
    if (digits != 3 && digits != 4) COEX_FAIL("coex_set_pearson: digits must be 3 or 4");
    if (mode != 0 && mode != 1) COEX_FAIL("coex_set_pearson: mode 0 = tensor-core digits, 1 = in-order double");
    c->pearson_digits = digits;
    c->pearson_mode = mode;
    return 0;
}

int coex_network_batch(coex_ctx *c, const coex_params *prm, int K, const long long *perm_off, const int *hit_rows,
                       int *n_groups, int *group_of_hit, int *group_label, int *group_winner, double *group_density,
                       double *hit_score, double *scores, long long *counts, int on_device) {
    COEX_TRY(check_ctx(c));
    if (!prm || !perm_off || (!hit_rows && perm_off[K] > 0)) COEX_FAIL("coex_network_batch: null argument");
    return network_batch_impl(c, prm, K, perm_off, hit_rows, n_groups, group_of_hit, group_label, group_winner,
                              group_density, hit_score, scores, counts, on_device);
}

// --------------------------------------------------------------------------------------------
// ONE job shared by several GPUs of a box (one process per GPU)
// --------------------------------------------------------------------------------------------
int coex_shard_scores_alloc(coex_ctx *c, int world, int rank, long long n_doubles, void *ipc_handle_out) {
    COEX_TRY(check_ctx(c));
    if (world < 1 || rank < 0 || rank >= world || n_doubles < 0 || !ipc_handle_out) COEX_FAIL("coex_shard_scores_alloc: bad argument");
    for (void *p : c->peer_mapped) if (p) cudaIpcCloseMemHandle(p);
    c->peer_mapped.assign((size_t)world, nullptr);
    c->peer_scores.assign((size_t)world, nullptr);
    c->shard_world = world; c->shard_rank = rank;
    COEX_TRY(c->shard_scores.reserve((size_t)std::max<long long>(n_doubles, 1)));
    c->peer_scores[(size_t)rank] = c->shard_scores.p;
    cudaIpcMemHandle_t h;
    COEX_CUDA(cudaIpcGetMemHandle(&h, c->shard_scores.p));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the header promises 64 bytes");
    memcpy(ipc_handle_out, &h, sizeof h);
    return 0;
}

int coex_shard_scores_open(coex_ctx *c, int peer_rank, const void *ipc_handle) {
    COEX_TRY(check_ctx(c));
    if (peer_rank < 0 || peer_rank >= c->shard_world || !ipc_handle) COEX_FAIL("coex_shard_scores_open: bad argument");
    if (peer_rank == c->shard_rank) return 0;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof h);
    void *p = nullptr;
    COEX_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    if (c->peer_mapped[(size_t)peer_rank]) cudaIpcCloseMemHandle(c->peer_mapped[(size_t)peer_rank]);
    c->peer_mapped[(size_t)peer_rank] = p;
    c->peer_scores[(size_t)peer_rank] = static_cast<double *>(p);
    return 0;
}

// The packed table of a shared job: every rank ranks its block of rows and stores the results into every rank's copy.
static const int kTabBufs = 9;
static void *tab_local(coex_ctx *c, int i) {
    switch (i) {
        case 0: return c->rank2.p; case 1: return c->u_hi.p; case 2: return c->u_lo.p; case 3: return c->suu.p; case 4: return c->sx.p;
        case 5: return c->raw_sum.p; case 6: return c->raw_mean.p; case 7: return c->raw_xx.p; default: return c->rawsum_pos.p;
    }
}

int coex_shard_table_export(coex_ctx *c, int world, int rank, void *ipc_handles_out) {
    COEX_TRY(check_ctx(c));
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || !ipc_handles_out) COEX_FAIL("coex_shard_table_export: bad argument (at most 8 ranks)");
    COEX_TRY(prepare_reserve(c));
    for (void *p : c->tab_mapped) if (p) cudaIpcCloseMemHandle(p);
    c->tab_mapped.assign((size_t)world * kTabBufs, nullptr);
    c->tab_peer.assign((size_t)world * kTabBufs, nullptr);
    c->shard_world = world; c->shard_rank = rank;
    c->tab_N = c->N; c->tab_n = c->n;
    for (int i = 0; i < kTabBufs; ++i) {
        c->tab_peer[(size_t)rank * kTabBufs + i] = tab_local(c, i);
        cudaIpcMemHandle_t h;
        COEX_CUDA(cudaIpcGetMemHandle(&h, tab_local(c, i)));
        memcpy(static_cast<unsigned char *>(ipc_handles_out) + 64 * i, &h, sizeof h);
    }
    return 0;
}

int coex_shard_table_open(coex_ctx *c, int peer_rank, const void *ipc_handles) {
    COEX_TRY(check_ctx(c));
    if (c->tab_peer.size() != (size_t)c->shard_world * kTabBufs) COEX_FAIL("coex_shard_table_open: call coex_shard_table_export first");
    if (peer_rank < 0 || peer_rank >= c->shard_world || !ipc_handles) COEX_FAIL("coex_shard_table_open: bad argument");
    if (peer_rank == c->shard_rank) return 0;
    for (int i = 0; i < kTabBufs; ++i) {
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const unsigned char *>(ipc_handles) + 64 * i, sizeof h);
        void *p = nullptr;
        COEX_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->tab_mapped[(size_t)peer_rank * kTabBufs + i] = p;
        c->tab_peer[(size_t)peer_rank * kTabBufs + i] = p;
    }
    return 0;
}

int coex_shard_prepare_rows(coex_ctx *c) {
    COEX_TRY(check_ctx(c));
    const int W = c->shard_world, r = c->shard_rank;
    if (c->tab_peer.size() != (size_t)W * kTabBufs || c->tab_N != c->N || c->tab_n != c->n)
        COEX_FAIL("coex_shard_prepare_rows: export / open the table buffers for this table shape first");
    for (int i = 0; i < kTabBufs; ++i)
        if (tab_local(c, i) != c->tab_peer[(size_t)r * kTabBufs + i]) COEX_FAIL("coex_shard_prepare_rows: table buffers were reallocated since the export");
    RankOut ro{}; StatOut so{};
    for (int d = 0; d < W; ++d) {
        void **t = &c->tab_peer[(size_t)d * kTabBufs];
        for (int i = 0; i < kTabBufs; ++i) if (!t[i]) COEX_FAIL("coex_shard_prepare_rows: the table of rank " + std::to_string(d) + " is not mapped");
        ro.rank2[d] = (uint16_t *)t[0]; ro.hi[d] = (int8_t *)t[1]; ro.lo[d] = (int8_t *)t[2]; ro.suu[d] = (long long *)t[3]; ro.sx[d] = (double *)t[4];
        so.sum[d] = (double *)t[5]; so.mean[d] = (double *)t[6]; so.xx[d] = (double *)t[7]; so.pos[d] = (unsigned char *)t[8];
    }
    ro.n_dst = so.n_dst = W;
    const long long chunk = round_up((c->N + W - 1) / W, 32);
    const long long row0 = std::min(c->N, (long long)r * chunk), row1 = std::min(c->N, row0 + chunk);
    COEX_TRY(prepare_rows(c, row0, row1, ro, so));
    COEX_CUDA(cudaStreamSynchronize(c->stream));            // my rows have landed everywhere; the caller's barrier tells the peers
    return 0;
}

int coex_shard_prepare_finish(coex_ctx *c) {
    COEX_TRY(check_ctx(c));
    return prepare_finish(c);
}

int coex_shard_count(coex_ctx *c, const coex_params *prm, int K, const long long *perm_off, const int *hit_rows,
                     const int *perm_owner, const long long *perm_score_off) {
    COEX_TRY(check_ctx(c));
    if (!prm || !perm_off || !perm_owner || !perm_score_off || (!hit_rows && perm_off[K] > 0)) COEX_FAIL("coex_shard_count: null argument");
    if (c->peer_scores.size() != (size_t)c->shard_world) COEX_FAIL("coex_shard_count: call coex_shard_scores_alloc first");
    c->h_perm_scores.assign((size_t)std::max(K, 1), nullptr);
    for (int k = 0; k < K; ++k) {
        const int o = perm_owner[k];
        if (o < 0 || o >= c->shard_world || !c->peer_scores[(size_t)o]) COEX_FAIL("coex_shard_count: the score buffer of rank " + std::to_string(o) + " is not mapped");
        c->h_perm_scores[(size_t)k] = c->peer_scores[(size_t)o] + perm_score_off[k];
    }
    c->shard_phase = 1;
    const int status = network_batch_impl(c, prm, K, perm_off, hit_rows, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0);
    c->shard_phase = 0;
    return status;
}

int coex_shard_finish(coex_ctx *c, const coex_params *prm, int K, const long long *perm_off, const int *hit_rows,
                      int *n_groups, int *group_of_hit, int *group_label, int *group_winner, double *group_density,
                      double *hit_score, double *scores, int on_device) {
    COEX_TRY(check_ctx(c));
    if (!prm || !perm_off || (!hit_rows && perm_off[K] > 0)) COEX_FAIL("coex_shard_finish: null argument");
    c->shard_phase = 2;
    const int status = network_batch_impl(c, prm, K, perm_off, hit_rows, n_groups, group_of_hit, group_label, group_winner,
                                          group_density, hit_score, scores, nullptr, on_device);
    c->shard_phase = 0;
    return status;
}

int coex_set_score_readback(coex_ctx *c, int k, double *host_out) {
    if (!c) COEX_FAIL("null context");
    c->rb_k = k; c->rb_host = host_out;
    return 0;
}

int coex_read_scores(coex_ctx *c, int k, double *out) {
    COEX_TRY(check_ctx(c));
    if (!out || c->last_sc_off.empty() || !c->last_scores) COEX_FAIL("coex_read_scores: no batch has been reduced by this context");
    if (k < 0 || k + 1 >= (int)c->last_sc_off.size()) COEX_FAIL("coex_read_scores: no such permutation in the last batch");
    const long long a = c->last_sc_off[(size_t)k], b = c->last_sc_off[(size_t)k + 1];
    if (b > a) COEX_CUDA(cudaMemcpyAsync(out, c->last_scores + a, (size_t)(b - a) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    COEX_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int coex_collate(coex_ctx *c, const double *pool, long long n_pool, int pool_on_device, const double *real, int G,
                 double *p_out, double *bh_out, double *by_out) {
    COEX_TRY(check_ctx(c));
    if (G < 0 || G > kCollateMax || n_pool < 0 || (G > 0 && (!real || !p_out || !bh_out || !by_out)) || (n_pool > 0 && !pool))
        COEX_FAIL("coex_collate: bad argument (at most 8192 real groups)");
    if (G == 0) return 0;
    cudaStream_t st = c->stream;
    const int G2 = std::max(2, next_pow2(G));
    DevBuf<double> d_pool, d_real, d_out;
    DevBuf<unsigned long long> d_hist;
    const double *dp = pool;
    if (!pool_on_device && n_pool > 0) {
        COEX_TRY(d_pool.reserve((size_t)n_pool));
        COEX_CUDA(cudaMemcpyAsync(d_pool.p, pool, (size_t)n_pool * sizeof(double), cudaMemcpyHostToDevice, st));
        dp = d_pool.p;
    }
    COEX_TRY(d_real.reserve((size_t)G)); COEX_TRY(d_out.reserve((size_t)3 * G)); COEX_TRY(d_hist.reserve((size_t)G + 1));
    COEX_CUDA(cudaMemcpyAsync(d_real.p, real, (size_t)G * sizeof(double), cudaMemcpyHostToDevice, st));
    COEX_CUDA(cudaMemsetAsync(d_hist.p, 0, ((size_t)G + 1) * sizeof(unsigned long long), st));
    const size_t sm1 = (size_t)G2 * 8 + ((size_t)G + 1) * 4, sm2 = (size_t)G2 * 16;
    COEX_CUDA(cudaFuncSetAttribute(collate_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
    COEX_CUDA(cudaFuncSetAttribute(collate_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
    if (n_pool > 0) {
        const int grid = (int)std::min<long long>((n_pool + 1023) / 1024, (long long)c->sm_count);
        collate_count_kernel<<<grid, 1024, sm1, st>>>(dp, n_pool, d_real.p, G, G2, d_hist.p);
        COEX_KCHECK(st);
    }
    collate_finish_kernel<<<1, 1024, sm2, st>>>(d_real.p, G, G2, d_hist.p, n_pool, d_out.p, d_out.p + G, d_out.p + 2 * G);
    COEX_KCHECK(st);
    COEX_CUDA(cudaMemcpyAsync(p_out, d_out.p, (size_t)G * sizeof(double), cudaMemcpyDeviceToHost, st));
    COEX_CUDA(cudaMemcpyAsync(bh_out, d_out.p + G, (size_t)G * sizeof(double), cudaMemcpyDeviceToHost, st));
    COEX_CUDA(cudaMemcpyAsync(by_out, d_out.p + 2 * G, (size_t)G * sizeof(double), cudaMemcpyDeviceToHost, st));
    COEX_CUDA(cudaStreamSynchronize(st));
    d_pool.release(); d_real.release(); d_out.release(); d_hist.release();
    return 0;
}

long long coex_launch_count(coex_ctx *c) { return c ? c->launches : 0; }

int coex_stage_ms(coex_ctx *c, int stage, double *ms, long long *launches) {
    COEX_TRY(check_ctx(c));
    if (stage < 0 || stage >= kNumStages) COEX_FAIL("coex_stage_ms: bad stage");
    COEX_TRY(stage_collect(c));
    if (ms) *ms = c->stage_ms[stage];
    if (launches) *launches = c->stage_launches[stage];
    return 0;
}

int coex_selftest_gemm(coex_ctx *c, long long m_rows, long long *mismatches, long long *first4) {
    COEX_TRY(check_ctx(c));
    if (!mismatches || !first4) COEX_FAIL("coex_selftest_gemm: null output");
    return selftest_gemm_impl(c, m_rows, mismatches, first4);
}

int coex_set_gemm_mode(coex_ctx *c, int mode) {
    if (!c) COEX_FAIL("null context");
    if (mode == 64 || mode == 128) { c->gemm_mode = 0; c->gemm_bn = mode; return 0; }   // tcgen05 with that column tile
    if (mode == 2) { c->gemm_mode = 0; c->gemm_bn = -64; return 0; }                    // persistent tcgen05 kernel, 128 x 64 tiles
    if (mode == 3) { c->gemm_mode = 0; c->gemm_bn = -128; return 0; }                   // persistent tcgen05 kernel, 128 x 128 tiles
    if (mode == 4) { c->gemm_mode = 0; c->gemm_bn = -256; return 0; }
    if (mode == 5) { c->gemm_mode = 0; c->gemm_bn = -129; return 0; }
    if (mode == 6) { c->gemm_mode = 0; c->gemm_bn = -257; return 0; }                   // digit-split CTA pairs (cta_group::2), double-buffered accumulators                   // 128 x 128 tiles, B multicast inside CTA pairs                   // CTA-pair tcgen05 kernel (cta_group::2), 256 x 64 tiles
    if (mode != 0 && mode != 1)
        COEX_FAIL("coex_set_gemm_mode: 0 = tcgen05 (persistent, tile chosen by row length), 1 = SIMT dp4a, 2 / 3 = persistent tcgen05 with "
                  "128 x 64 / 128 x 128 tiles, 5 = 128 x 128 with B multicast in CTA pairs, 6 = digit-split cta_group::2 pairs, "
                  "64 / 128 = one-tile-per-CTA tcgen05 with that column tile");
    c->gemm_mode = mode;
    if (mode == 0) c->gemm_bn = 0;
    return 0;
}

int coex_set_count_mode(coex_ctx *c, int mode) {
    if (!c) COEX_FAIL("null context");
    if (mode < 0 || mode > 4)
        COEX_FAIL("coex_set_count_mode: 0 = de-duplicated count kernels (bucket-table or sorted-statistics kernel by table length, row-rank mode by "
                  "cost), 1 = per-query kernel, 2 = sorted-statistics kernel, 3 = row-rank mode forced, 4 = bucket-table kernel forced");
    c->count_mode = mode;
    return 0;
}

int coex_set_workspace_mb(coex_ctx *c, long long strip_mb) {
    if (!c) COEX_FAIL("null context");
    if (strip_mb < 16) COEX_FAIL("coex_set_workspace_mb: at least 16 MB");
    c->strip_mb = strip_mb;
    return 0;
}

// --------------------------------------------------------------------------------------------
// Legacy drop-in (reference coexpression_v2.c:22-23).  Host pointers in, results written on
// return.  The symbol has no error channel, so any CUDA failure aborts loudly.
// --------------------------------------------------------------------------------------------
static void legacy_die(const char *what) {
    fprintf(stderr, "coexpression_v2.so (B200): getPearson failed: %s: %s\n", what, coex_last_error());
    abort();
}

void getPearson(double *dataAll, double *resultAll, int *idxa_all, int *idxb_all, int nPromPairs, int n) {
    if (nPromPairs <= 0) return;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    long long rows = 0;
    for (int x = 0; x < nPromPairs; ++x) {
        if (idxa_all[x] < 0 || idxb_all[x] < 0) { g_last_error = "negative row index"; legacy_die("arguments"); }
        rows = std::max<long long>(rows, std::max(idxa_all[x], idxb_all[x]) + 1LL);
    }
    coex_ctx *c = nullptr;
    if (coex_create(&c, 0)) legacy_die("no CUDA device (this library has no CPU path)");
    auto run = [&]() -> int {
        COEX_TRY(c->raw_own.reserve((size_t)rows * n));
        COEX_CUDA(cudaMemcpyAsync(c->raw_own.p, dataAll, (size_t)rows * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        c->raw = c->raw_own.p; c->N = rows; c->n = n;
        COEX_TRY(c->raw_sum.reserve(rows)); COEX_TRY(c->raw_mean.reserve(rows)); COEX_TRY(c->raw_xx.reserve(rows));
        StatOut so{};
        so.sum[0] = c->raw_sum.p; so.mean[0] = c->raw_mean.p; so.xx[0] = c->raw_xx.p; so.pos[0] = nullptr; so.n_dst = 1;
        rowstats_kernel<<<(unsigned)((rows + kStatRows - 1) / kStatRows), 256, 0, c->stream>>>(c->raw, rows, n, 0, so);
        COEX_CUDA(cudaGetLastError());
        c->prepared = true;               // only the raw statistics are needed by which == 0
        COEX_TRY(coex_pairs(c, 0, idxa_all, idxb_all, nPromPairs, resultAll, 0));
        return 0;
    };
    const int status = run();
    if (status) { coex_destroy(c); legacy_die("device computation"); }
    coex_destroy(c);
}

}  // extern "C"
#pragma GCC visibility pop
