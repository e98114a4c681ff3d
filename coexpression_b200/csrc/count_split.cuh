// count_split.cuh -- set-up half of the de-duplicated Spearman count stage (default path, count mode 0).
//
// Same contract as phases A - C of count_dedup_kernel<0> (reference 1-make-network.py:273-292, coexfunctions.py:386-409:
// the hit-vs-hit statistics of every query of a work item, undefined ones scored on the spot, the others sorted), done for
// ALL work items of a strip BEFORE the streaming kernel (count_dedup_kernel<3>) runs.
//
// Why a kernel of its own (profiles/r01_count_split.md): inside the fused kernel these phases were half of an item's cycles
// and purely latency bound -- the sort went through an L2-resident scratch (the shared memory belongs to the bucket table,
// the histogram and the bitmap of the streaming phase), every step a dependent global round trip, with only two CTAs per
// SM to hide it.  Here the statistics of an item live in shared memory from evaluation to the sorted write-out
// (8 B + 2 x 2 B per statistic + 16 KB of packed 16-bit bucket counters: three CTAs per SM at 4096 statistics), the only
// global round trips are the gathers of the statistic itself (hit row -> strip value, eight per thread in flight) and the
// coalesced write of the result.
//
// Output per item `it` of the launch: item_T[it] sorted statistics at item_soff[it] (a bump allocation from *st_cursor):
// st_rec (16-byte records: the statistic and its table column) for the streaming kernel, and st_w for score_scatter_kernel: the
// address of the statistic's edge score, with the two per-query facts of quirk Q1 already decided (bit 62: the null value
// at the skipped column lies below the statistic; bit 63: a column was skipped, the null list holds N - 1 values).
// The streaming kernel adds st_cnt (columns strictly below the statistic); score_scatter_kernel, one thread per
// statistic of the whole strip, turns that into bisect index and score -- the scattered 8-byte stores and the score-table
// gathers were another latency-bound tail of the fused kernel.
#pragma once
#include "count_dedup.cuh"

namespace coex {

constexpr int kSsThreads = 256;
constexpr int kPlanQcapSplit = 8;        // = kPlanQcap (plan.cuh): queries per work item
constexpr int kSsBatch = 8;              // statistics per thread in flight
constexpr int kSsBuckets = 4096;         // buckets of the sort (its own table: only monotonicity matters, not the streaming kernel's width)
static_assert(kPlanQcapSplit == 8, "locate() compares against seven query boundaries held in registers");

struct SortArgs {
    const int *strip;               // [rows_in_strip, ld] exact Spearman sums
    long long ld;
    const DedupItem *items;
    int n_items;
    const int *uniq_rows, *qlist, *q_perm;
    const long long *perm_off, *sc_off;
    const int *hit_rows;
    const double *sx;               // [N_pad] half-norms
    const double *sxg;              // [N_pad] half-norms, NaN where the raw row sums to <= 0 (icol_kernel)
    const unsigned char *rawsum_pos;
    long long N;
    double *scores;
    long long *counts;
    int Tcap;
    float bscale;                   // kDdBuckets / (2 R), the same float the streaming kernel uses
    StatRec *st_rec;                // {statistic, -, column}
    unsigned long long *st_w;
    unsigned long long st_cap;      // capacity of the arrays
    unsigned long long *st_cursor;  // [1] bump allocator, zeroed before the launch
    int *err;                       // [1] set when the capacity is exceeded (the host sizes it so that it cannot happen)
    long long *item_soff;
    int *item_T;
    unsigned long long *phase_clk;  // optional [8]: summed clock64 per phase over all CTAs (profiling aid), [7] = items
};

inline size_t sort_smem_bytes(int Tcap) {
    return (size_t)Tcap * 8 + (size_t)(kSsBuckets / 2) * 4 + (size_t)Tcap * 2 + 16;
}

// monotone bucket of a statistic (see dd_bucket32)
__device__ __forceinline__ int ss_bucket(double t, float scale) {
    const int b = __float2int_rz(__fmaf_rn(__double2float_rn(t), scale, (float)(kSsBuckets / 2)));
    return min(max(b, 0), kSsBuckets - 1);
}
// 16-bit counter b of the packed table
__device__ __forceinline__ unsigned int ss_half(const unsigned int *lutw, int b) { return (lutw[b >> 1] >> ((b & 1) << 4)) & 0xffffu; }

__global__ void __launch_bounds__(kSsThreads, 4)
stat_sort_kernel(SortArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *r64 = reinterpret_cast<double *>(smem_raw);                              // [Tcap] statistic of flat index t (-99 = undefined)
    unsigned int *lutw = reinterpret_cast<unsigned int *>(r64 + a.Tcap);             // [kSsBuckets / 2] two 16-bit counters per word (T <= 8192)
    unsigned short *g_t = reinterpret_cast<unsigned short *>(lutw + kSsBuckets / 2); // [Tcap] flat indices grouped by bucket
    __shared__ long long q_sc[kPlanQcapSplit];       // offset of the query's score row
    __shared__ double q_vi[kPlanQcapSplit];          // exact null value at the skipped column (quirk Q1)
    __shared__ int q_i[kPlanQcapSplit], q_off[kPlanQcapSplit], q_P[kPlanQcapSplit], q_toff[kPlanQcapSplit + 1];
    __shared__ unsigned int s_warp[kSsThreads / 32];
    __shared__ long long s_soff;

    const int tid = threadIdx.x;
    const float bscale_f = a.bscale * ((float)kSsBuckets / (float)kDdBuckets);
    for (int it = blockIdx.x; it < a.n_items; it += gridDim.x) {
        const DedupItem item = a.items[it];
        const int nq = item.qe - item.qb;
        const int u = a.uniq_rows[item.urow];
        const int *__restrict__ srow = a.strip + (long long)item.urow * a.ld;
        const double sxu = a.sx[u];
        const double sxu_g = a.sxg[u];                     // NaN if the hit row itself fails the raw-sum guard
        __syncthreads();                                   // previous item fully retired
        long long tclk = clock64();
#define SS_PHASE(id) do { if (a.phase_clk && tid == 0) { const long long now = clock64(); atomicAdd(&a.phase_clk[id], (unsigned long long)(now - tclk)); tclk = now; } } while (0)
        // ---- A. query metadata -------------------------------------------------------------------------
        if (tid < kPlanQcapSplit) {
            const int ql = tid;
            int P = 0;
            if (ql < nq) {
                const long long q = a.qlist[item.qb + ql];
                const int k = a.q_perm[q];
                const long long off = a.perm_off[k];
                P = (int)(a.perm_off[k + 1] - off);
                const int i = (int)(q - off);
                double vi = -99.0;                         // null value at table column i (quirk Q1)
                if (i < a.N) {
                    const double sxc = a.sx[i];
                    if (sxu != 0.0 && sxc != 0.0) vi = __ddiv_rn(0.25 * (double)srow[i], __dmul_rn(sxu, sxc));
                }
                q_i[ql] = i; q_off[ql] = (int)off; q_sc[ql] = a.sc_off[k] + (long long)i * P; q_vi[ql] = vi;
            }
            q_P[ql] = P;
        }
        for (int w = tid; w < kSsBuckets / 2; w += kSsThreads) lutw[w] = 0;
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int ql = 0; ql < kPlanQcapSplit; ++ql) { q_toff[ql] = acc; acc += q_P[ql]; }       // P (diagonal included)
            q_toff[kPlanQcapSplit] = acc;
        }
        __syncthreads();
        const int TP = q_toff[kPlanQcapSplit];
        // first flat index of queries 1..7 (queries beyond nq are empty: their boundary is TP)
        const int b1 = q_toff[1], b2 = q_toff[2], b3 = q_toff[3], b4 = q_toff[4], b5 = q_toff[5], b6 = q_toff[6], b7 = q_toff[7];
        auto locate = [&](int t) {                         // flat index -> query of the item (no memory access, no loop)
            return (int)(t >= b1) + (int)(t >= b2) + (int)(t >= b3) + (int)(t >= b4) + (int)(t >= b5) + (int)(t >= b6) + (int)(t >= b7);
        };
        SS_PHASE(0);
        // ---- B. statistics: evaluated once, counted per bucket -----------------------------------------
        for (int t0 = tid; t0 < TP; t0 += kSsBatch * kSsThreads) {
            int cc[kSsBatch], sraw[kSsBatch];
            double sxc[kSsBatch];
#pragma unroll
            for (int e = 0; e < kSsBatch; ++e) {
                const int t = t0 + e * kSsThreads;
                cc[e] = -2;                                // -2 = beyond the item, -1 = diagonal
                if (t < TP) {
                    const int ql = locate(t), j = t - q_toff[ql];
                    cc[e] = (j == q_i[ql]) ? -1 : a.hit_rows[q_off[ql] + j];
                }
            }
#pragma unroll
            for (int e = 0; e < kSsBatch; ++e) {
                sraw[e] = 0; sxc[e] = 0.0;
                if (cc[e] >= 0) { sraw[e] = srow[cc[e]]; sxc[e] = a.sxg[cc[e]]; }
            }
#pragma unroll
            for (int e = 0; e < kSsBatch; ++e) {
                if (cc[e] == -2) continue;
                const int t = t0 + e * kSsThreads;
                double r = -99.0;
                if (cc[e] >= 0) {
                    // coexfunctions.py:389 (either raw sum <= 0 -> NaN through sxg) and :406-408 (NaN -> -99)
                    r = __ddiv_rn(0.25 * (double)sraw[e], __dmul_rn(sxu_g, sxc[e]));
                    if (r != r) r = -99.0;
                }
                r64[t] = r;
                if (r == -99.0) {                                            // diagonal, or 1-make-network.py:284-285
                    const int ql = locate(t);
                    const long long w = q_sc[ql] + (t - q_toff[ql]);
                    a.scores[w] = (cc[e] < 0) ? 0.0 : -0.0;                  // -log10(1.0)
                    if (a.counts) a.counts[w] = -1;
                } else {
                    const int b = ss_bucket(r, bscale_f);
                    atomicAdd(&lutw[b >> 1], 1u << ((b & 1) << 4));
                }
            }
        }
        __syncthreads();
        SS_PHASE(1);
        // ---- exclusive scan of the 16-bit bucket counts (a warp scans a contiguous slice of words) -----
        unsigned int T;
        {
            constexpr int kWarps = kSsThreads / 32, kPerWarp = (kSsBuckets / 2) / kWarps;
            const int wid = tid >> 5, ln = tid & 31;
            unsigned int carry = 0;
            for (int c = 0; c < kPerWarp; c += 32) {
                const int w = wid * kPerWarp + c + ln;
                const unsigned int x = lutw[w];
                const unsigned int c0 = x & 0xffffu, c1 = x >> 16, v = c0 + c1;
                unsigned int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, incl, o); if (ln >= o) incl += y; }
                const unsigned int ex = carry + incl - v;
                lutw[w] = ex | ((ex + c0) << 16);
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (ln == 0) s_warp[wid] = carry;
            __syncthreads();
            unsigned int base = 0, tot = 0;
            for (int w = 0; w < kWarps; ++w) { const unsigned int x = s_warp[w]; if (w < wid) base += x; tot += x; }
            if (base)
                for (int c = 0; c < kPerWarp; c += 32) lutw[wid * kPerWarp + c + ln] += base * 0x10001u;
            T = tot;
            if (tid == 0) {
                long long soff = (long long)atomicAdd(a.st_cursor, (unsigned long long)T);
                if ((unsigned long long)soff + T > a.st_cap) { atomicOr(a.err, 2); soff = -1; }
                s_soff = soff;
                a.item_soff[it] = soff < 0 ? 0 : soff;
                a.item_T[it] = soff < 0 ? 0 : (int)T;
            }
            __syncthreads();
        }
        const long long soff = s_soff;
        SS_PHASE(2);
        if (a.phase_clk && tid == 0) atomicAdd(&a.phase_clk[7], 1ULL);
        if (T == 0 || soff < 0) continue;
        // ---- group by bucket: the counter of bucket b walks from its start to its end (= start of b + 1) ----
        for (int t = tid; t < TP; t += kSsThreads) {
            const double r = r64[t];
            if (r != -99.0) {
                const int b = ss_bucket(r, bscale_f);
                const int sh = (b & 1) << 4;
                const unsigned int old = atomicAdd(&lutw[b >> 1], 1u << sh);
                g_t[(old >> sh) & 0xffffu] = (unsigned short)t;
            }
        }
        __syncthreads();
        SS_PHASE(3);
        // ---- rank inside the bucket = sorted position (ties: grouped position); the statistic goes straight to its
        //      place in the item's slice of the output (scattered stores inside a few tens of KB: merged in L2) ----
        for (int e0 = tid; e0 < (int)T; e0 += 4 * kSsThreads) {
            int tt[4], ql4[4], col4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {                  // the column gathers of the thread's four statistics in flight together
                const int e = e0 + k * kSsThreads;
                tt[k] = -1; col4[k] = 0; ql4[k] = 0;
                if (e < (int)T) {
                    tt[k] = g_t[e];
                    ql4[k] = locate(tt[k]);
                    col4[k] = a.hit_rows[q_off[ql4[k]] + (tt[k] - q_toff[ql4[k]])];
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (tt[k] < 0) continue;
                const int e = e0 + k * kSsThreads;
                const double r = r64[tt[k]];
                const int b = ss_bucket(r, bscale_f);
                const int hi = (int)ss_half(lutw, b), lo = b ? (int)ss_half(lutw, b - 1) : 0;
                int rank = 0;
                if (hi - lo > 1)
                    for (int m = lo; m < hi; ++m) {
                        const double x = r64[g_t[m]];
                        rank += (x < r) || (x == r && m < e);
                    }
                const long long g = soff + lo + rank;
                const int ql = ql4[k];
                unsigned long long w = (unsigned long long)(q_sc[ql] + (tt[k] - q_toff[ql]));
                if (q_i[ql] < a.N) w |= (1ULL << 63) | ((q_vi[ql] < r) ? (1ULL << 62) : 0ULL);
                a.st_rec[g] = StatRec{r, 0u, col4[k]};
                a.st_w[g] = w;
            }
        }
        SS_PHASE(4);
    }
}

// One thread per sorted statistic of the strip: bisect index = columns strictly below the statistic (st_cnt: streamed + hit
// columns) + zero-variance columns (-99 < t always) - [the skipped column i lies below t] (quirk Q1); edge score from the
// table of every possible index (1-make-network.py:283-314 through edge_score_dev).
struct ScatterArgs {
    const unsigned long long *st_w;
    const unsigned int *st_cnt;
    const unsigned long long *st_total;   // [1] = the bump allocator after stat_sort_kernel
    const int *n_degenerate;
    long long N;
    int anticor;
    const double *score_tab;        // [N+1] edge score of every bisect index of a list of N - 1 values
    double *scores;
    long long *counts;
};

__global__ void __launch_bounds__(256)
score_scatter_kernel(ScatterArgs a) {
    const unsigned long long total = *a.st_total;
    const long long n_deg = *a.n_degenerate;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long ww = a.st_w[g];
        const long long w = (long long)(ww & ((1ULL << 62) - 1));
        const long long idx = (long long)a.st_cnt[g] + n_deg - (long long)((ww >> 62) & 1ULL);
        a.scores[w] = (ww >> 63) ? a.score_tab[idx] : edge_score_dev(idx, a.N, a.anticor);
        if (a.counts) a.counts[w] = idx;
    }
}

}  // namespace coex
