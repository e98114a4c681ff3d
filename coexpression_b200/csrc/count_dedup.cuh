// count_dedup.cuh -- K3 epilogue stage, production version for the Spearman null.
//
// Same contract as count.cuh (reference 1-make-network.py:239-258, 273-327) with three changes
// that remove the work the ncu capture of round 1 showed to dominate (profiles/r01_summary.md):
//
//  1. De-duplication.  The null values of a table row do not depend on the permutation, only
//     the statistics (thresholds) do, and the same row is hit by several permutations.  Work
//     items are therefore (unique row, a set of its (permutation, hit position) queries): the
//     row is correlated and streamed ONCE and searched against the merged statistics of all
//     its queries.  Quirk Q1 (the null list of hit position i omits table column i) becomes a
//     per-query correction: idx = #{all columns below t} - [value at column i below t].
//  2. Bucket table instead of a 10..13-step binary search: statistics are counting-sorted into
//     8192 uniform buckets of [-R, R] (R ~ 8 standard deviations of a null correlation; values
//     outside land in the edge buckets) and ranked inside their bucket; a streamed value finds its
//     bucket with one multiply-add and searches only the statistics of that bucket: statistics and
//     values go through the SAME monotone float bucket function (dd_bucket32), so the order of two
//     numbers never contradicts the order of their buckets.
//  3. fp32 pre-filter with exact fp64 refinement: the value and the search are done in float;
//     whenever the float result could differ from the double one (a statistic within 5e-7 of the
//     value, < 1 % of values) the column is queued and redone in double with the reference's
//     exact arithmetic.  Counts are therefore still EXACT.
#pragma once
#include <math_constants.h>
#include "common.cuh"
#include "count.cuh"

namespace coex {

constexpr int kDdThreads = 512;
constexpr int kDdBuckets = 8192;
constexpr int kDdQueue = 2048;
constexpr int kRankTcap = 8192;          // columns ranked per work item in row-rank mode
constexpr float kDdDelta = 5e-7f;        // |v32 - v64| <= 3e-7 |v|, |t32 - t64| <= 6e-8 |t|  
constexpr int kDdSuper = 32768;          // columns streamed between two drains of the refinement queue (multiple of 4 * kDdThreads)

// The float errors are RELATIVE (int -> float, two rounded reciprocal half-norms, two products: 5 * 2^-24 |v|; the
// statistic is one rounding of its double: 2^-24 |t|).  If |t| < 1.13 |v| they add up to < 3.7e-7 |v|, otherwise
// |t - v| > 0.1 |v| dwarfs them: a float difference of at least 5e-7 |v| decides the order.  The floor keeps exact
// ties at zero out of the "certain" branch.
__device__ __forceinline__ float dd_delta(float v) { return fmaxf(kDdDelta * fabsf(v), 1e-30f); }

// One statistic in the L2-resident scratch of a work item: 16 bytes = ONE sector access where three parallel arrays
// (double statistic, origin, column) cost three -- the set-up phases are bound by the number of uncoalesced global accesses
// per statistic (~1 SM-cycle each, profiles/r01_count_split.md), not by their bytes.
struct __align__(16) StatRec {
    double r;            // the statistic (-99 = undefined, never sorted)
    unsigned int org;    // (query of the item << 13) | hit position j   (row-rank mode: column offset inside the run)
    int col;             // table column of the statistic
};

struct DedupItem {
    int urow;        // index into the strip (unique-row slot)
    int qb, qe;      // range in qlist
    int T;           // number of statistics of this item (upper bound: sum (P-1))
};

struct CfMap { float st, k, r1, half; int nbm1; };      // bucket map of count_fine_kernel (count_fine.cuh)

struct DedupArgs {
    const int *strip;               // [rows_in_strip, ld]
    long long ld;
    const DedupItem *items;         // items of this launch (plan.cuh)
    int n_items;
    const int *uniq_rows;           // table row of each unique-row slot of the strip
    const int *qlist;               // flat hit indices grouped by unique row
    const int *q_perm;              // [H] permutation of every flat hit index (plan_count_kernel)
    const long long *perm_off, *sc_off;
    int K;
    const int *hit_rows;
    const double *sx;               // [N_pad]
    const double *sxg;              // [N_pad] sx, NaN where the raw row sums to <= 0 (icol_kernel): the statistic's guard in one gather
    const float *icol;              // [N_pad] (float)(0.5 / sx), 0 for zero-variance / padding rows
    const unsigned char *rawsum_pos;
    long long N;
    const int *n_degenerate;        // [1] device: zero-variance table rows, their null value is -99 for every query
    int anticor;
    double *scores;
    long long *counts;
    double *const *perm_scores;     // nullptr, or [K]: where the P x P score block of every permutation lives -- on a multi-GPU run
                                    // the buffer of the rank that OWNS the permutation (peer memory over NVLink, mapped with CUDA IPC):
                                    // the score rows are stored straight into it, there is no exchange step afterwards
    int Tcap, Qcap;                 // shared-memory capacities (statistics, queries per item)
    double R;                       // bucket range [-R, R]
    int bm_words;                   // 32-bit words of the hit-column bitmap (0 = table too long, bitmap disabled)
    int claim;                      // 1: a second bitmap (claim mask) de-duplicates the hit columns (short tables only)
    StatRec *g_rec;                 // scratch [gridDim.x, 3, Tcap]: unsorted | grouped by bucket | sorted statistics
    const double *score_tab;        // [N+1] edge score of every possible bisect index (len = N-1)
    unsigned long long *phase_clk;  // optional [8]: summed clock64 per phase over all CTAs (profiling aid)
    int *work_counter;              // [1] zeroed before the launch: a CTA takes its first item by its index and every further one
                                    // from this counter (items differ 7-fold in their number of statistics: a static
                                    // round-robin leaves a tail); nullptr = static round-robin
    // ---- PEARSON == true (`-cm Pearson`: Pearson null, Spearman statistic -- quirk Q2) ---------------------------
    const float *strip_f32;         // [rows_in_strip, ld] tensor-core Pearson r of the unique rows (|error| <= eps); `strip` still holds the
                                    // exact Spearman sums the statistics are read from
    const double *raw;              // [N, n] raw table and its in-order row statistics: the reference's exact arithmetic
    const double *raw_sum, *raw_mean, *raw_xx;
    int n;
    unsigned int *rank_out;         // MODE 2: [rows_in_strip, ld] number of null values strictly below each column's value
    // ---- count_fine_kernel (count_fine.cuh) and its re-do list ----------------------------------------------------
    int fine_nb;                    // buckets of the shared-memory table (power of two, 1024 .. 32768)
    CfMap fine_map;
    int fine_decline_mod;           // test hook (COEX_FINE_DECLINE_EVERY): count_fine_kernel declines every item whose index is a multiple of it
    long long fine_chunk;           // columns per pass of its 16-bit counters (multiple of 4096, <= 57344)
    int *redo_list, *redo_count;    // count_fine_kernel: items it declines (appended); count_dedup_kernel<0>: if redo_count is set, the
                                    // items to work on are items[redo_list[0 .. *redo_count)]
    const float *eps_row;           // [N_pad] per-row share of the bound of |tensor r - exact r| (= icol in this mode; 0 = undefined row):
                                    // |error(u, c)| <= eps_row[u] + eps_row[c] (digit quantisation)
};

// Pearson r of two table rows exactly as coexpression_v2.c:89-113 computes it (in-order double sums); NaN -> -99.
__device__ __forceinline__ double pearson_exact_thread(const DedupArgs &a, long long ra, long long rb) {
    if (a.raw_sum[ra] == 0.0 || a.raw_sum[rb] == 0.0) return -99.0;
    const double *pa = a.raw + ra * a.n, *pb = a.raw + rb * a.n;
    const double ma = a.raw_mean[ra], mb = a.raw_mean[rb];
    double xy = 0.0;
    for (int c = 0; c < a.n; ++c) xy = __dadd_rn(xy, __dmul_rn(__dsub_rn(pa[c], ma), __dsub_rn(pb[c], mb)));
    const double r = __ddiv_rn(xy, __dmul_rn(sqrt(a.raw_xx[ra]), sqrt(a.raw_xx[rb])));
    return (r != r) ? -99.0 : r;
}

// Bucket of a value: ONE float function for statistics (rounded to float first), streamed float values and exactly
// refined doubles (rounded to float first).  Every step (double -> float, the fused multiply-add, the truncation, the
// clamps) is monotone non-decreasing, so  t <= v  =>  bucket(t) <= bucket(v)  and  t > v  =>  bucket(t) >= bucket(v)
// both for the float compare of the fast path and for the double compare of the refinement: the statistics <= v are
// all the statistics of the lower buckets plus some of v's OWN bucket -- a one-bucket search window with no
// "off by one bucket in float" case.
__device__ __forceinline__ int dd_bucket32(float x, float scale) {
    const int b = __float2int_rz(__fmaf_rn(x, scale, (float)(kDdBuckets / 2)));
    return min(max(b, 0), kDdBuckets - 1);
}
__device__ __forceinline__ int dd_bucket(double t, float scale) { return dd_bucket32(__double2float_rn(t), scale); }

// Sixteen of them per warp.  The reference adds the 1829 products strictly left to right: ONE refinement is a chain of dependent
// double additions, ~60 k cycles whatever the number of lanes working on it (a warp per column with staged products was measured:
// 64 k cycles; a thread per column walking its two rows with uncoalesced 8-byte loads: 300 k).  So a warp runs 16 chains at once, one
// per lane 0..15: per 8 samples, lane (g, s) = (lane >> 3, lane & 7) loads sample s of row u once and of the columns of chains
// 4 q + g, q = 0..3 (eight lanes share a 64-byte segment of a row; the next 8 samples are in flight meanwhile), the 16 x 8 products
// are staged in a warp-private tile of shared memory (rows padded to 9 doubles: conflict-free), and lane r adds row r in order.
// Samples behind the end of a row contribute +0.0 (x + 0.0 == x).  cols[r] < 0: no chain.  Returns lane r's r(u, cols[r]).
constexpr int kPxChains = 16, kPxTile = kPxChains * 9;
__device__ __forceinline__ double pearson_exact_16(const DedupArgs &a, long long ru, int my_col /* lane r < 16: column of chain r */,
                                                   int lane, double *tile) {
    const int g = lane >> 3, sidx = lane & 7;
    const double mu = a.raw_mean[ru];
    const double *pu = a.raw + ru * a.n;
    int cq[4];
    double mq[4];
    const double *pq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        cq[q] = __shfl_sync(0xffffffffu, my_col, 4 * q + g);
        mq[q] = (cq[q] >= 0) ? a.raw_mean[cq[q]] : 0.0;
        pq[q] = a.raw + (long long)max(cq[q], 0) * a.n;
    }
    double xy = 0.0;
    // two groups of 8 samples in flight (the rows come from HBM: the table does not fit L2)
    double vu0, vq0[4], vu1, vq1[4];
    auto fetch = [&](int c, double &vu, double (&vq)[4]) {
        vu = (c < a.n) ? pu[c] : mu;
#pragma unroll
        for (int q = 0; q < 4; ++q) vq[q] = (cq[q] >= 0 && c < a.n) ? pq[q][c] : mq[q];
    };
    auto chain8 = [&](const double vu, const double (&vq)[4]) {
        const double du = __dsub_rn(vu, mu);
        double prod[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) prod[q] = __dmul_rn(du, __dsub_rn(vq[q], mq[q]));
        __syncwarp();                                  // the previous tile has been read
#pragma unroll
        for (int q = 0; q < 4; ++q) tile[(4 * q + g) * 9 + sidx] = prod[q];
        __syncwarp();
        if (lane < kPxChains) {
            const double *row = tile + lane * 9;
#pragma unroll
            for (int l = 0; l < 8; ++l) xy = __dadd_rn(xy, row[l]);
        }
    };
    fetch(sidx, vu0, vq0);
    fetch(8 + sidx, vu1, vq1);
    for (int c0 = 0; c0 < a.n; c0 += 16) {
        const double tu0 = vu0; const double tq0[4] = {vq0[0], vq0[1], vq0[2], vq0[3]};
        fetch(c0 + 16 + sidx, vu0, vq0);
        chain8(tu0, tq0);
        const double tu1 = vu1; const double tq1[4] = {vq1[0], vq1[1], vq1[2], vq1[3]};
        fetch(c0 + 24 + sidx, vu1, vq1);
        chain8(tu1, tq1);                              // (behind the end of the row: eight times + 0.0)
    }
    double r = -99.0;
    if (lane < kPxChains && my_col >= 0 && a.raw_sum[ru] != 0.0 && a.raw_sum[my_col] != 0.0) {
        r = __ddiv_rn(xy, __dmul_rn(sqrt(a.raw_xx[ru]), sqrt(a.raw_xx[my_col])));
        if (r != r) r = -99.0;
    }
    return r;
}

// MODE 0: Spearman null, statistics of the item's queries.  MODE 1: Pearson null (tensor filter), same items.
// MODE 2: RANK of a whole strip row -- the "statistics" of an item are the null values of a run of Tcap table columns
//         themselves, and the result is, for each of them, the number of null values strictly below it (rank_out).  Used
//         when the hit sets are so large that a row's queries carry more statistics than the row has columns (config C5:
//         wide windows, thousands of hits per set): every (query, hit) then only LOOKS UP the rank of its column
//         (rank_score_kernel) instead of having the row streamed once per query.
template <int MODE>
__global__ void __launch_bounds__(kDdThreads, 2)
count_dedup_kernel(DedupArgs a) {
    constexpr bool PEARSON = (MODE == 1), RANKROW = (MODE == 2);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: q_vi[Qcap] | thr32[Tcap+4] | hist[Tcap+4] (thr32|hist = cursor[B] during step B) | lut[B+4] | queue[kDdQueue] | q_k q_i q_toff
    // (the double statistics and their origins live in an L2-resident per-CTA scratch: they are only
    //  touched by the rare exact refinements and by the final scoring)
    double *q_vi = reinterpret_cast<double *>(smem_raw);                      // [Qcap] exact null value at the skipped column
    float *thr32 = reinterpret_cast<float *>(q_vi + a.Qcap) + 1;             // [-1 .. Tcap+2]: thr32[-1] = -inf, thr32[T] = +inf (sentinels)
    unsigned int *hist = reinterpret_cast<unsigned int *>(thr32 + a.Tcap + 3);
    unsigned int *lut = hist + a.Tcap + 4;
    unsigned int *cursor = reinterpret_cast<unsigned int *>(thr32 - 1);   // bucket cursors alias thr32|hist (both dead until step C; 2*Tcap >= buckets)
    int *queue = reinterpret_cast<int *>(lut + kDdBuckets + 4);
    long long *q_sc = reinterpret_cast<long long *>(queue + kDdQueue);   // [Qcap] ADDRESS of the query's score row
    long long *q_cn = q_sc + a.Qcap;                                       // [Qcap] offset of the query's row in a.counts
    int *q_i = reinterpret_cast<int *>(q_cn + a.Qcap);                     // [Qcap] hit position
    int *q_off = q_i + a.Qcap;                   // [Qcap] offset of the query's permutation in hit_rows
    int *q_P = q_off + a.Qcap;                   // [Qcap] hits of that permutation
    int *q_toff = q_P + a.Qcap;                  // [Qcap+2] offset of the query's statistics in the item
    unsigned int *bitmap = reinterpret_cast<unsigned int *>(q_toff + a.Qcap + 2);   // [bm_words] hit columns of this item (skipped by the streaming pass) | [bm_words] claim mask if a.claim
    __shared__ int s_T, s_qn, s_next;
    __shared__ unsigned int s_warp[kDdThreads / 32];
    __shared__ double s_tile[PEARSON ? (kDdThreads / 32) * kPxTile : 1];        // staged products of a warp's 16 chains (pearson_exact_16)

    const int tid = threadIdx.x;
    const float bscale_f = (float)((double)kDdBuckets / (2.0 * a.R));
    StatRec *rec_tmp = a.g_rec + (long long)blockIdx.x * 3 * a.Tcap;   // statistic of flat index t
    StatRec *rec_grp = rec_tmp + a.Tcap;                             // grouped by bucket
    StatRec *srt_w = rec_grp + a.Tcap;                               // sorted
    const StatRec *srt = srt_w;

    const int n_deg = *a.n_degenerate;
    const int nchunk = RANKROW ? (int)((a.N + a.Tcap - 1) / a.Tcap) : 1;
    auto next_item = [&](int it) {
        if (!a.work_counter) return it + (int)gridDim.x;
        __syncthreads();
        if (threadIdx.x == 0) s_next = (int)gridDim.x + atomicAdd(a.work_counter, 1);
        __syncthreads();
        return s_next;
    };
    const int n_items = (MODE == 0 && a.redo_count) ? *a.redo_count : a.n_items;
    for (int it = blockIdx.x; it < n_items; it = next_item(it)) {
        DedupItem item;
        if (RANKROW) { item.urow = it / nchunk; item.qb = item.qe = 0; item.T = 0; }
        else item = a.items[(MODE == 0 && a.redo_count) ? a.redo_list[it] : it];
        const int rank_c0 = RANKROW ? (it % nchunk) * a.Tcap : 0;        // first column of the run
        const int nq = item.qe - item.qb;
        const int u = a.uniq_rows[item.urow];
        const int *srow = a.strip + (long long)item.urow * a.ld;
        const float *frow = PEARSON ? a.strip_f32 + (long long)item.urow * a.ld : nullptr;
        const double sxu = a.sx[u];
        const bool posu = a.rawsum_pos[u] != 0;
        __syncthreads();                                   // previous item fully retired
        long long tclk = clock64();
#define DD_PHASE(id) do { if (a.phase_clk && tid == 0) { const long long now = clock64(); atomicAdd(&a.phase_clk[id], (unsigned long long)(now - tclk)); tclk = now; } } while (0)
        int T = 0;
        {
        // ---- A. query metadata ----------------------------------------------------------------
        for (int ql = tid; ql < nq; ql += kDdThreads) {
            const long long q = a.qlist[item.qb + ql];
            const int lo = a.q_perm[q];                    // (was a binary search over perm_off: seven dependent global loads)
            const long long off = a.perm_off[lo];
            const int P = (int)(a.perm_off[lo + 1] - off);
            const int i = (int)(q - off);
            q_i[ql] = i;
            q_off[ql] = (int)off;
            q_P[ql] = P;
            double *sbase = a.perm_scores ? a.perm_scores[lo] : a.scores + a.sc_off[lo];
            q_sc[ql] = (long long)(size_t)(sbase + (long long)i * P);
            q_cn[ql] = a.sc_off[lo] + (long long)i * P;
            double vi = -99.0;                             // null value at table column i (quirk Q1)
            if (PEARSON) {
                vi = 0.0;                                  // filled in by the last refinement pass (phase E)
            } else {
                const double sxc = a.sx[i];
                if (sxu != 0.0 && sxc != 0.0) vi = __ddiv_rn(0.25 * (double)srow[i], __dmul_rn(sxu, sxc));
            }
            q_vi[ql] = vi;
        }
        if (tid == 0) { s_T = 0; s_qn = 0; }
        for (int b = tid; b <= kDdBuckets; b += kDdThreads) lut[b] = 0;
        for (int wd = tid; wd < (a.claim ? 2 : 1) * a.bm_words; wd += kDdThreads) bitmap[wd] = 0;
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int ql = 0; ql < nq; ++ql) { q_toff[ql] = acc; acc += q_P[ql]; }       // P (diagonal included)
            q_toff[nq] = acc;
        }
        __syncthreads();
        const int TP = RANKROW ? (int)min((long long)a.Tcap, a.N - rank_c0) : q_toff[nq];
        DD_PHASE(0);
        // ---- B. statistics: evaluated once (four per thread in flight), counted per bucket ------
        for (int t0 = tid; t0 < TP; t0 += 4 * kDdThreads) {
            int tt[4], qq[4], jj[4], cc4[4], sraw[4];
            double sxc4[4];
            bool live[4], guard[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                tt[e] = t0 + e * kDdThreads;
                live[e] = tt[e] < TP;
                qq[e] = 0; jj[e] = 0; cc4[e] = -1;
                if (RANKROW) {
                    if (live[e]) cc4[e] = rank_c0 + tt[e];
                } else if (live[e]) {
                    int lo = 0, hi = nq;
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (q_toff[mid] <= tt[e]) lo = mid; else hi = mid; }
                    qq[e] = lo; jj[e] = tt[e] - q_toff[lo];
                    if (jj[e] != q_i[lo]) cc4[e] = a.hit_rows[q_off[lo] + jj[e]];
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                guard[e] = RANKROW; sraw[e] = 0; sxc4[e] = 0.0;
                if (cc4[e] >= 0) {
                    sraw[e] = srow[cc4[e]];
                    sxc4[e] = RANKROW ? a.sx[cc4[e]] : a.sxg[cc4[e]];        // NaN = raw sum <= 0 (coexfunctions.py:389)
                    guard[e] = RANKROW || sxc4[e] == sxc4[e];
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (!live[e]) continue;
                double r = -99.0;
                if (RANKROW) {
                    // the null value of the column itself (no raw-sum guard: that belongs to the statistic, rank_score_kernel)
                    if (sxu != 0.0 && sxc4[e] != 0.0) r = __ddiv_rn(0.25 * (double)sraw[e], __dmul_rn(sxu, sxc4[e]));
                } else if (cc4[e] >= 0 && posu && guard[e]) {               // coexfunctions.py:389 (the statistic is Spearman in both modes, Q2)
                    r = __ddiv_rn(0.25 * (double)sraw[e], __dmul_rn(sxu, sxc4[e]));
                    if (r != r) r = -99.0;                                  // coexfunctions.py:406-408
                }
                rec_tmp[tt[e]] = StatRec{r, 0u, cc4[e]};
                if (r == -99.0) {                                           // diagonal, or 1-make-network.py:284-285
                    if (!RANKROW) {
                        reinterpret_cast<double *>((size_t)q_sc[qq[e]])[jj[e]] = (cc4[e] < 0) ? 0.0 : -log10(1.0);
                        if (a.counts) a.counts[q_cn[qq[e]] + jj[e]] = -1;
                    }
                } else {
                    atomicAdd(&lut[dd_bucket(r, bscale_f)], 1u);
                    // the null value of a hit column IS this statistic (same arithmetic): it is accounted
                    // exactly in step C' and skipped by the float streaming pass
                    if (a.bm_words) {
                        const unsigned int bit = 1u << (cc4[e] & 31);
                        atomicOr(&bitmap[cc4[e] >> 5], bit);
                        if (a.claim) atomicOr(&bitmap[a.bm_words + (cc4[e] >> 5)], bit);
                    }
                }
            }
        }
        __syncthreads();
        {   // exclusive scan of the bucket counts (each warp scans a contiguous slice, bank-conflict free)
            constexpr int kWarps = kDdThreads / 32, kPerWarp = kDdBuckets / kWarps;
            const int wid = tid >> 5, ln = tid & 31;
            unsigned int carry = 0;
            for (int c = 0; c < kPerWarp; c += 32) {
                const int b = wid * kPerWarp + c + ln;
                const unsigned int v = lut[b];
                unsigned int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned int x = __shfl_up_sync(0xffffffffu, incl, o); if (ln >= o) incl += x; }
                lut[b] = carry + incl - v;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (ln == 0) s_warp[wid] = carry;
            __syncthreads();
            unsigned int base = 0;
            for (int w = 0; w < wid; ++w) base += s_warp[w];
            for (int c = 0; c < kPerWarp; c += 32) {
                const int b = wid * kPerWarp + c + ln;
                lut[b] += base;
                cursor[b] = 0;
            }
            if (tid == kDdThreads - 1) { lut[kDdBuckets] = base + carry; s_T = (int)(base + carry); }
            __syncthreads();
        }
        // group by bucket
        for (int t = tid; t < TP; t += kDdThreads) {
            const StatRec x = rec_tmp[t];
            const double r = x.r;
            if (r != -99.0) {
                int lo = 0, hi = nq;
                if (!RANKROW) while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (q_toff[mid] <= t) lo = mid; else hi = mid; }
                const int b = dd_bucket(r, bscale_f);
                const unsigned int pos = lut[b] + atomicAdd(&cursor[b], 1u);
                rec_grp[pos] = StatRec{r, RANKROW ? (unsigned int)t : ((unsigned int)lo << 13) | (unsigned int)(t - q_toff[lo]), x.col};
            }
        }
        __syncthreads();
        T = s_T;
        // lut[b] = first statistic of bucket b | number of statistics in it << 16 (the cursors have counted them; T <= 8192):
        // one shared-memory load per streamed value gives its whole search window
        for (int b = tid; b < kDdBuckets; b += kDdThreads) lut[b] |= cursor[b] << 16;
        __syncthreads();                                    // the cursors (aliasing thr32 | hist) are dead from here on
        DD_PHASE(1);
        if (T == 0) continue;                               // every statistic undefined: scores already written
        // ---- C. rank inside the bucket -> fully sorted statistics -------------------------------------
        for (int e = tid; e < T; e += kDdThreads) {
            const StatRec x = rec_grp[e];
            const double t = x.r;
            const unsigned int pk = lut[dd_bucket(t, bscale_f)];
            const int lo = (int)(pk & 0xffffu), hi = lo + (int)(pk >> 16);
            int rank = 0;
            for (int m = lo; m < hi; ++m) {
                const double y = rec_grp[m].r;
                rank += (y < t) || (y == t && m < e);
            }
            srt_w[lo + rank] = x;
            thr32[lo + rank] = (float)t;
        }
        for (int p = tid; p <= T; p += kDdThreads) hist[p] = 0;
        if (tid == 0) { thr32[-1] = -CUDART_INF_F; thr32[T] = CUDART_INF_F; }
        __syncthreads();
        }
        // ---- C'. hit columns: value == statistic exactly.  A column hit by several queries of the item appears as
        // several EQUAL statistics (same row, same column, same arithmetic) and is added once: by the statistic that
        // claims its bit in a second bitmap (short tables), or by the first of the equal statistics in sorted order
        // (long tables, where a second bitmap would cost the second resident CTA).
        if (a.bm_words) {
            for (int s0 = 0; s0 < T; s0 += 4 * kDdThreads) {
              int c4[4];
              double t4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {                           // the thread's four global loads in flight together
                  const int sidx = s0 + tid + e * kDdThreads;
                  StatRec x = {0.0, 0u, 0};
                  if (sidx < T) x = srt[sidx];
                  c4[e] = x.col; t4[e] = x.r;
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int sidx = s0 + tid + e * kDdThreads;
                if (sidx >= T) continue;
                const int c = c4[e];
                const float tf = thr32[sidx];                         // shared-memory filter in front of the L2-resident doubles
                bool first = true;
                if (a.claim) {                                        // short table: the first statistic to claim the column's bit
                    const unsigned int bit = 1u << (c & 31);
                    first = (atomicAnd(&bitmap[a.bm_words + (c >> 5)], ~bit) & bit) != 0;
                }
                const double t = t4[e];
                if (!a.claim)
                    for (int e1 = sidx - 1; e1 >= 0 && thr32[e1] == tf && srt[e1].r == t; --e1)
                        if (srt[e1].col == c) { first = false; break; }
                if (first) {
                    int e2 = sidx + 1;
                    while (e2 < T && thr32[e2] == tf && srt[e2].r == t) ++e2;   // statistics <= value: through the run of equals
                    atomicAdd(&hist[e2], 1u);
                }
              }
            }
            __syncthreads();
        }
        DD_PHASE(2);
        // ---- D. stream the null values of table row u (fp32 fast path) -----------------------------
        const float iu = (float)(0.5 / sxu);
        const float eps_u = PEARSON ? a.eps_row[u] : 0.0f;
        const long long N = a.N;
        // Four columns per thread are handled in lockstep (independent shared-memory chains overlap)
        // and the next 16-byte strip / norm loads are issued before the current ones are consumed.
        // The row is streamed in super-chunks of kDdSuper columns; the queue of columns that need the reference's
        // double arithmetic is drained by the whole CTA after each of them (a long row of a zero-inflated table holds
        // thousands of values TIED with a statistic, far more than one queue; refining them in place -- one lane's
        // double division at a time -- is what it used to cost).
        for (long long cb = 0; cb < N; cb += kDdSuper) {
            const long long cend = (cb + kDdSuper < N) ? cb + kDdSuper : N;
            long long c0 = cb + (long long)tid * 4;
            bool have = c0 < cend;
            int4 s4n = make_int4(0, 0, 0, 0);
            float4 w4n = make_float4(0.f, 0.f, 0.f, 0.f);
            auto fetch = [&](long long cpos) {
                // Pearson: the four floats travel in the same registers as the four exact integer sums of the Spearman strip
                s4n = PEARSON ? *reinterpret_cast<const int4 *>(frow + cpos) : *reinterpret_cast<const int4 *>(srow + cpos);
                w4n = *reinterpret_cast<const float4 *>(a.icol + cpos);
            };
            if (have) fetch(c0);
            while (have) {
                const int sv[4] = {s4n.x, s4n.y, s4n.z, s4n.w};
                const float wv[4] = {w4n.x, w4n.y, w4n.z, w4n.w};
                const long long cc = c0;
                const unsigned int skip4 = a.bm_words ? (bitmap[cc >> 5] >> (cc & 31)) & 15u : 0u;   // cc % 4 == 0
                c0 += kDdThreads * 4;
                have = c0 < cend;
                if (have) fetch(c0);
                float v[4];
                int p[4], len[4];
                bool act[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    // zero-variance column (and the padding behind column N): mask 0 -> -99, counted in bulk; hit column:
                    // accounted exactly in step C'
                    act[e] = (wv[e] != 0.0f) && !((skip4 >> e) & 1u);
                    v[e] = PEARSON ? __int_as_float(sv[e]) : __int2float_rn(sv[e]) * (iu * wv[e]);
                    // the statistics <= v are those of the lower buckets and some of v's own bucket (dd_bucket32 is monotone)
                    const unsigned int pk = lut[dd_bucket32(v[e], bscale_f)];
                    p[e] = (int)(pk & 0xffffu);
                    len[e] = act[e] ? (int)(pk >> 16) : 0;
                }
                while ((len[0] | len[1] | len[2] | len[3]) > 0) {         // statistics <= v inside the bucket
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (len[e] > 0) {
                            const int half = len[e] >> 1;
                            if (thr32[p[e] + half] <= v[e]) { p[e] += half + 1; len[e] -= half + 1; } else len[e] = half;
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (!act[e]) continue;
                    // a statistic closer than delta cannot be ordered in float: the two neighbours of v in the sorted
                    // statistics decide (sentinels -inf / +inf at both ends)
                    const float below = v[e] - thr32[p[e] - 1];
                    const float above = thr32[p[e]] - v[e];
                    if (fminf(below, above) >= (PEARSON ? dd_delta(v[e]) + eps_u + wv[e] : dd_delta(v[e]))) {
                        atomicAdd(&hist[p[e]], 1u);
                    } else {
                        const int slot = atomicAdd(&s_qn, 1);
                        if (slot < kDdQueue) {
                            queue[slot] = (int)(cc + e);
                        } else {                                          // queue full: refine in place
                            const double vv = PEARSON ? pearson_exact_thread(a, u, cc + e)
                                                      : __ddiv_rn(0.25 * (double)srow[cc + e], __dmul_rn(sxu, a.sx[cc + e]));
                            const unsigned int pk = lut[dd_bucket(vv, bscale_f)];
                            int pp = (int)(pk & 0xffffu);
                            int ln = (int)(pk >> 16);
                            while (ln > 0) {
                                const int half = ln >> 1;
                                if (srt[pp + half].r <= vv) { pp += half + 1; ln -= half + 1; } else ln = half;
                            }
                            atomicAdd(&hist[pp], 1u);
                        }
                    }
                }
            }
            __syncthreads();
            // ---- E. exact refinement of the queued columns ------------------------------------------
            const int qn = min(s_qn, kDdQueue);
            if (PEARSON) {
                // A refinement pass costs one chain of n dependent additions whatever it holds (pearson_exact_16): the queue is
                // worked off when it is half full and behind the last column, together with the exact null values at the skipped
                // columns of the item's queries (quirk Q1), which phase F needs.
                const bool last = cend >= N;
                if (!last && qn < kDdQueue / 2) continue;              // (qn is the same in every thread: read behind the barrier)
                const int total = qn + (last ? nq : 0);
                for (int x0 = (tid >> 5) * kPxChains; x0 < total; x0 += (kDdThreads / 32) * kPxChains) {
                    const int lane = tid & 31, x = x0 + lane;
                    int c = -1;
                    if (lane < kPxChains && x < total) {
                        if (x < qn) c = queue[x];
                        else { const int i = q_i[x - qn]; c = (i < a.N) ? i : -1; }
                    }
                    const double vv = pearson_exact_16(a, u, c, lane, s_tile + (tid >> 5) * kPxTile);
                    if (lane < kPxChains && x < total) {
                        if (x >= qn) q_vi[x - qn] = vv;                // (c < 0: -99)
                        else {
                            const unsigned int pk = lut[dd_bucket(vv, bscale_f)];
                            int pp = (int)(pk & 0xffffu);
                            int len = (int)(pk >> 16);
                            while (len > 0) {
                                const int half = len >> 1;
                                if (srt[pp + half].r <= vv) { pp += half + 1; len -= half + 1; } else len = half;
                            }
                            atomicAdd(&hist[pp], 1u);
                        }
                    }
                }
            } else
            for (int x = tid; x < qn; x += kDdThreads) {
                const int c = queue[x];
                const double vv = __ddiv_rn(0.25 * (double)srow[c], __dmul_rn(sxu, a.sx[c]));
                const unsigned int pk = lut[dd_bucket(vv, bscale_f)];
                int pp = (int)(pk & 0xffffu);
                int len = (int)(pk >> 16);
                while (len > 0) {
                    const int half = len >> 1;
                    if (srt[pp + half].r <= vv) { pp += half + 1; len -= half + 1; } else len = half;
                }
                atomicAdd(&hist[pp], 1u);
            }
            if (a.phase_clk && tid == 0) atomicAdd(&a.phase_clk[7], (unsigned long long)s_qn);
            __syncthreads();
            if (tid == 0) s_qn = 0;
            __syncthreads();
        }
        DD_PHASE(3);
        DD_PHASE(4);
        // ---- F. inclusive prefix sum of hist[0..T), scores ----------------------------------------
        {
            const int per = (T + 1 + kDdThreads - 1) / kDdThreads;
            const int b0 = tid * per, b1 = min(b0 + per, T + 1);
            unsigned int local = 0;
            for (int p = b0; p < b1; ++p) local += hist[p];
            unsigned int incl = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned int x = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += x; }
            if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
            __syncthreads();
            unsigned int base = incl - local;
            for (int w = 0; w < (tid >> 5); ++w) base += s_warp[w];
            for (int p = b0; p < b1; ++p) { base += hist[p]; hist[p] = base; }
        }
        __syncthreads();
        if (RANKROW) {
            // all columns strictly below the column's own value (zero-variance columns: -99 < everything)
            unsigned int *rrow = a.rank_out + (long long)item.urow * a.ld + rank_c0;
            for (int s = tid; s < T; s += kDdThreads) rrow[srt[s].org] = hist[s] + (unsigned int)n_deg;
        } else
        for (int s0 = 0; s0 < T; s0 += 4 * kDdThreads) {
            // four statistics per thread: record loads, then score-table gathers, then stores -- each kind in flight together
            long long idx4[4], w4[4], wc4[4];                 // w4: address of the score (0 = none), wc4: offset in a.counts
            double sc4[4];
            bool tab4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int s = s0 + tid + e * kDdThreads;
                w4[e] = 0; wc4[e] = 0; idx4[e] = 0; tab4[e] = true;
                if (s < T) {
                    const StatRec x = srt[s];
                    const unsigned int o = x.org;
                    const int ql = (int)(o >> 13), j = (int)(o & 8191u);
                    const int i = q_i[ql];
                    // all columns strictly below t: streamed values + zero-variance columns (-99 < t always),
                    // minus the skipped column i of this query if it is below t (quirk Q1)
                    long long idx = (long long)hist[s] + n_deg;
                    if (i < N) { if (q_vi[ql] < x.r) idx -= 1; } else tab4[e] = false;     // no skipped column: N values
                    idx4[e] = idx;
                    w4[e] = q_sc[ql] + 8LL * j;
                    wc4[e] = q_cn[ql] + j;
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e)
                sc4[e] = (w4[e] == 0) ? 0.0 : (tab4[e] ? a.score_tab[idx4[e]] : edge_score_dev(idx4[e], N, a.anticor));
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (w4[e] == 0) continue;
                *reinterpret_cast<double *>((size_t)w4[e]) = sc4[e];
                if (a.counts) a.counts[wc4[e]] = idx4[e];
            }
        }
        DD_PHASE(5);
        if (a.phase_clk && tid == 0) atomicAdd(&a.phase_clk[6], 1ULL);
    }
}

// MODE 2 companion: every (query, hit) looks its bisect index up in the rank strip.  One CTA per query of the strip.
//   idx = rank[u][column of hit j]  (= #null values of row u strictly below the statistic: the statistic IS that column's
//   null value, same arithmetic) - [null value at the skipped column i below the statistic] (quirk Q1); guards, diagonal,
//   -99 and the score exactly as in phase B / F of the query modes.
struct RankScoreArgs {
    const int *strip;
    const unsigned int *rank;
    long long ld;
    const int *uniq_rows, *ustart, *qlist, *q_perm;   // unique rows of the strip (already offset), their query ranges
    long long u_count;
    const long long *perm_off, *sc_off;
    const int *hit_rows;
    const double *sx;
    const double *sxg;              // sx, NaN where the raw row sums to <= 0 (the statistic's guard in one gather)
    const unsigned char *rawsum_pos;
    long long N;
    int anticor;
    double *scores;
    long long *counts;
    double *const *perm_scores;     // as in DedupArgs
    const double *score_tab;
};

__global__ void __launch_bounds__(256)
rank_score_kernel(RankScoreArgs a) {
    __shared__ int s_slot;
    const long long q_begin = a.ustart[0], q_count = a.ustart[a.u_count] - q_begin;
    for (long long x = blockIdx.x; x < q_count; x += gridDim.x) {
        const long long q = a.qlist[q_begin + x];
        const int k = a.q_perm[q];
        const long long off = a.perm_off[k];
        const int P = (int)(a.perm_off[k + 1] - off);
        const int i = (int)(q - off);
        const int u = a.hit_rows[q];
        __syncthreads();
        if (threadIdx.x == 0) {                       // slot of row u among the strip's unique rows (ascending)
            long long lo = 0, hi = a.u_count;
            while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if (a.uniq_rows[mid] <= u) lo = mid; else hi = mid; }
            s_slot = (int)lo;
        }
        __syncthreads();
        const int *srow = a.strip + (long long)s_slot * a.ld;
        const unsigned int *rrow = a.rank + (long long)s_slot * a.ld;
        const double sxu = a.sx[u];
        const double sxu_g = a.sxg[u];
        double vi = -99.0;                            // null value at table column i (quirk Q1)
        if (i < a.N) { const double sxc = a.sx[i]; if (sxu != 0.0 && sxc != 0.0) vi = __ddiv_rn(0.25 * (double)srow[i], __dmul_rn(sxu, sxc)); }
        double *sc = (a.perm_scores ? a.perm_scores[k] : a.scores + a.sc_off[k]) + (long long)i * P;
        long long *cn = a.counts ? a.counts + a.sc_off[k] + (long long)i * P : nullptr;
        for (int j = threadIdx.x; j < P; j += blockDim.x) {
            if (j == i) { sc[j] = 0.0; if (cn) cn[j] = -1; continue; }
            const int c = a.hit_rows[off + j];
            // coexfunctions.py:389 (either raw sum <= 0 -> NaN through sxg) and :406-408 (NaN -> -99)
            double t = __ddiv_rn(0.25 * (double)srow[c], __dmul_rn(sxu_g, a.sxg[c]));
            if (t != t) t = -99.0;
            if (t == -99.0) { sc[j] = -log10(1.0); if (cn) cn[j] = -1; continue; }
            long long idx = rrow[c];
            long long len = a.N;
            if (i < a.N) { len = a.N - 1; if (vi < t) idx -= 1; }
            sc[j] = (len == a.N - 1) ? a.score_tab[idx] : edge_score_dev(idx, len, a.anticor);
            if (cn) cn[j] = idx;
        }
    }
}

// edge score of every possible bisect index, evaluated once per call by the same device function
__global__ void score_table_kernel(double *tab, long long N, int anticor) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx <= N) tab[idx] = edge_score_dev(idx, N - 1, anticor);
}

// float reciprocal half-norms for the fp32 fast path + number of zero-variance rows
// sxg: the half-norm with the raw-sum guard of the hit-vs-hit statistic folded in (coexfunctions.py:389): NaN for a row
// whose raw values sum to <= 0, so that the statistic's own division yields NaN -> -99 without a second gather.
__global__ void icol_kernel(const double *__restrict__ sx, const unsigned char *__restrict__ rawsum_pos, long long N,
                            long long N_pad, float *__restrict__ icol, double *__restrict__ sxg, int *__restrict__ n_degenerate) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int deg = 0;
    if (t < N_pad) {
        const double s = sx[t];
        const bool ok = (t < N) && (s != 0.0);
        icol[t] = ok ? (float)(0.5 / s) : 0.0f;
        sxg[t] = (t < N && rawsum_pos[t] != 0) ? s : CUDART_NAN;
        deg = (t < N && s == 0.0) ? 1 : 0;
    }
    deg = __syncthreads_count(deg);
    if (threadIdx.x == 0 && deg) atomicAdd(n_degenerate, deg);
}

inline size_t dedup_smem_bytes(int Tcap, int Qcap, int bm_words, int claim) {
    return (size_t)Qcap * 8 + (size_t)(Tcap + 4) * 4 + (size_t)(Tcap + 4) * 4 + (size_t)(kDdBuckets + 4) * 4 +
           (size_t)kDdQueue * 4 + (size_t)Qcap * 8 * 2 + (size_t)Qcap * 4 * 3 + (size_t)(Qcap + 2) * 4 + (size_t)bm_words * 4 * (claim ? 2 : 1) + 16;
}

}  // namespace coex
