// count.cuh -- K3 epilogue stage: node-specific null counts and edge scores.
//
// Replaces, for one query (= hit position i of permutation k, table row u), the reference's
//   * NaN -> -99 and per-promoter sorted null list   1-make-network.py:239-258
//   * hit-vs-hit statistic r(i, j)                   1-make-network.py:281 (Spearman, quirk Q2)
//   * idx = bisect_left(null[i], r)                  1-make-network.py:287
//   * p1 = 1 - idx/len, -a: min(p1, idx/len)         1-make-network.py:287-292
//   * score = -log10(p1), p1 == 0 -> 0 (quirk Q4)    1-make-network.py:311-314
// without ever materialising or sorting the N-1 null values: bisect_left(list, r) is the count
// of list entries strictly below r, so the P-1 statistics of the query are sorted in shared
// memory and each streamed null value is binary-searched into them and histogrammed.
//
// Input is one row of the exact integer strip S[q, c] = sum u_q*u_c (int32); the null value is
// rebuilt in double exactly as the reference computes it on ranked rows:
//      r = (S/4) / (sqrt(Suu_q/4) * sqrt(Suu_c/4)),   NaN (zero-variance row) -> -99.
// Quirk Q1: table column c == i (the hit POSITION, not the hit's row) is left out of the list.
#pragma once
#include "common.cuh"

namespace coex {

constexpr int kCountThreads = 256;

struct CountArgs {
    const int *strip;               // [rows_in_strip, ld] exact integer S (Spearman null)
    const double *strip_f64;        // [rows_in_strip, ld] r values (Pearson null, NaN = undefined)
    long long ld;
    long long q0;                   // first global query of this strip
    long long nq;                   // queries in this strip
    const long long *perm_off;      // [K+1] device
    const long long *sc_off;        // [K+1] device, prefix of P_k^2
    int K;
    const int *hit_rows;            // [H]
    const double *sx;               // [N_pad]
    const unsigned char *rawsum_pos;
    long long N;
    int anticor;
    double *scores;                 // [sum P_k^2]
    long long *counts;              // optional
    int T2max;                      // shared-memory capacity (power of two)
    const double *score_tab;        // [N + 1] edge score of every bisect index of a list of N - 1 values (built on the host, coex.cu)
};

__device__ __forceinline__ double edge_score_dev(long long idx, long long len, int anticor) {
    double p1 = 1.0 - (double)idx / (double)len;
    if (anticor) {
        const double p2 = (double)idx / (double)len;
        p1 = fmin(p1, p2);
    }
    // -math.log(p1, 10) of 1-make-network.py:310 is -(log(p1) / log(10)); the device log is within 1 ulp of the C library's, so
    // this device function serves the side paths only (per-query cross-check kernel, global-list null): the production path
    // looks the score up in a table the HOST built with the C library (coex.cu)
    return (p1 > 0.0) ? -(log(p1) / log(10.0)) : 0.0;
}

// F64 == false: Spearman null from the integer strip, statistics taken from the same strip.
// F64 == true : Pearson null from a double strip; the (always Spearman, quirk Q2) statistics
//               were written into `scores` beforehand by stat_rank_kernel.
template <bool F64>
__global__ void __launch_bounds__(kCountThreads)
count_score_kernel(CountArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *thr = reinterpret_cast<double *>(smem_raw);                     // [T2]
    unsigned int *hist = reinterpret_cast<unsigned int *>(thr + a.T2max);   // [T2]
    int *origin = reinterpret_cast<int *>(hist + a.T2max);                  // [T2]
    __shared__ int s_cnt;
    __shared__ unsigned int s_warp[kCountThreads / 32];

    const int tid = threadIdx.x;
    const long long q = a.q0 + blockIdx.x;
    // locate permutation k with perm_off[k] <= q < perm_off[k+1]
    int lo = 0, hi = a.K;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a.perm_off[mid] <= q) lo = mid; else hi = mid;
    }
    const int k = lo;
    const long long off = a.perm_off[k];
    const int P = (int)(a.perm_off[k + 1] - off);
    const int i = (int)(q - off);
    const int *hits = a.hit_rows + off;
    const int u = hits[i];
    const double sxu = a.sx[u];
    const bool posu = a.rawsum_pos[u] != 0;
    const int *srow = F64 ? nullptr : a.strip + (long long)blockIdx.x * a.ld;
    const double *frow = F64 ? a.strip_f64 + (long long)blockIdx.x * a.ld : nullptr;
    double *sc = a.scores + a.sc_off[k] + (long long)i * P;
    long long *cn = a.counts ? a.counts + a.sc_off[k] + (long long)i * P : nullptr;

    if (tid == 0) s_cnt = 0;
    __syncthreads();
    // ---- the P-1 statistics of this query ---------------------------------------------------
    for (int j = tid; j < P; j += kCountThreads) {
        if (j == i) { sc[j] = 0.0; if (cn) cn[j] = -1; continue; }
        const int c = hits[j];
        double r = -99.0;
        if (F64) {
            r = sc[j];                                       // statistic from stat_rank_kernel
        } else if (posu && a.rawsum_pos[c]) {                // coexfunctions.py:389
            r = __ddiv_rn(0.25 * (double)srow[c], __dmul_rn(sxu, a.sx[c]));
            if (r != r) r = -99.0;                           // coexfunctions.py:406-408
        }
        if (r == -99.0) {                                    // 1-make-network.py:284-285
            sc[j] = -log10(1.0);
            if (cn) cn[j] = -1;
        } else {
            const int slot = atomicAdd(&s_cnt, 1);
            thr[slot] = r;
            origin[slot] = j;
        }
    }
    __syncthreads();
    const int T = s_cnt;
    if (T == 0) return;
    int T2 = 1;
    while (T2 < T + 1) T2 <<= 1;                             // at least one +inf sentinel
    for (int p = T + tid; p < T2; p += kCountThreads) {
        thr[p] = __longlong_as_double(0x7ff0000000000000LL);
        origin[p] = -1;
    }
    for (int p = tid; p < T2; p += kCountThreads) hist[p] = 0;
    __syncthreads();
    for (int kk = 2; kk <= T2; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (T2 >> 1); t += kCountThreads) {
                const int x = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int y = x | j;
                const double ka = thr[x], kb = thr[y];
                const int oa = origin[x], ob = origin[y];
                const bool gt = (ka > kb) || (ka == kb && (unsigned)oa > (unsigned)ob);
                if (gt == ((x & kk) == 0)) {
                    thr[x] = kb; thr[y] = ka;
                    origin[x] = ob; origin[y] = oa;
                }
            }
            __syncthreads();
        }
    }
    // ---- stream the null values of row u ----------------------------------------------------
    const long long N = a.N;
    for (long long c = tid; c < N; c += kCountThreads) {
        if (c == i) continue;                                // quirk Q1
        double v = -99.0;                                    // NaN -> -99 (1-make-network.py:242-243)
        if (F64) {
            const double f = frow[c];
            if (f == f) v = f;
        } else {
            const double sxc = a.sx[c];
            if (sxu != 0.0 && sxc != 0.0) v = __ddiv_rn(0.25 * (double)srow[c], __dmul_rn(sxu, sxc));
        }
        int pos = 0;                                         // number of statistics <= v
#pragma unroll 1
        for (int step = T2 >> 1; step > 0; step >>= 1)
            if (thr[pos + step - 1] <= v) pos += step;
        atomicAdd(&hist[pos], 1u);
    }
    __syncthreads();
    // ---- inclusive prefix sum of hist[0..T) -------------------------------------------------
    const int per = (T2 + kCountThreads - 1) / kCountThreads;
    const int b0 = tid * per, b1 = min(b0 + per, T2);
    unsigned int local = 0;
    for (int p = b0; p < b1; ++p) local += hist[p];
    unsigned int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    unsigned int base = incl - local;
    for (int w = 0; w < (tid >> 5); ++w) base += s_warp[w];
    for (int p = b0; p < b1; ++p) {
        base += hist[p];
        hist[p] = base;
    }
    __syncthreads();
    const long long len = (i < N) ? (N - 1) : N;
    for (int s = tid; s < T; s += kCountThreads) {
        const long long idx = hist[s];
        const int j = origin[s];
        sc[j] = (len == N - 1 && a.score_tab) ? a.score_tab[idx] : edge_score_dev(idx, len, a.anticor);
        if (cn) cn[j] = idx;
    }
}

// Hit-vs-hit statistic matrix (always Spearman on the raw-table guard, quirk Q2):
// stat[i][j] = get_correlation(prom_i, prom_j) of coexfunctions.py:386-409, -99 where undefined.
// One warp per ordered pair, permutation = blockIdx.y.  Written into the scores buffer.
struct StatArgs {
    const long long *perm_off, *sc_off;
    const int *hit_rows;
    const uint16_t *rank2;
    int n;
    const double *sx;
    const unsigned char *rawsum_pos;
    double *scores;
};

__device__ __forceinline__ long long lower_bound_dev(const double *v, long long len, double x) {
    long long lo = 0, hi = len;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (v[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
stat_rank_kernel(StatArgs a) {
    const int k = blockIdx.y;
    const long long off = a.perm_off[k];
    const int P = (int)(a.perm_off[k + 1] - off);
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < (long long)P * P; w += warps) {
        const int i = (int)(w / P), j = (int)(w % P);
        double *sc = a.scores + a.sc_off[k] + w;
        if (i == j) { if (lane == 0) *sc = 0.0; continue; }
        const int ra = a.hit_rows[off + i], rb = a.hit_rows[off + j];
        double r = -99.0;
        if (a.rawsum_pos[ra] && a.rawsum_pos[rb]) {
            const uint16_t *pa = a.rank2 + (long long)ra * a.n, *pb = a.rank2 + (long long)rb * a.n;
            long long acc = 0;
            for (int c = lane; c < a.n; c += 32)
                acc += (long long)((int)pa[c] - (a.n + 1)) * ((int)pb[c] - (a.n + 1));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            r = __ddiv_rn(0.25 * (double)acc, __dmul_rn(a.sx[ra], a.sx[rb]));
            if (r != r) r = -99.0;
        }
        if (lane == 0) *sc = r;
    }
}

// nsp == 0 (1-make-network.py:293-301): p from the global correlation list; no null pass.
// One thread per ordered hit pair; converts the statistic in place into the edge score.
__global__ void __launch_bounds__(256)
global_from_stat_kernel(const long long *perm_off, const long long *sc_off, const double *glist,
                        long long glist_len, int anticor, double *scores, long long *counts) {
    const int k = blockIdx.y;
    const int P = (int)(perm_off[k + 1] - perm_off[k]);
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < (long long)P * P;
         w += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(w / P), j = (int)(w % P);
        double *sc = scores + sc_off[k] + w;
        long long *cn = counts ? counts + sc_off[k] + w : nullptr;
        if (i == j) { *sc = 0.0; if (cn) *cn = -1; continue; }
        const double r = *sc;
        if (r == -99.0) { *sc = -log10(1.0); if (cn) *cn = -1; continue; }
        const long long idx = lower_bound_dev(glist, glist_len, r);
        *sc = edge_score_dev(idx, glist_len, anticor);
        if (cn) *cn = idx;
    }
}

// Pearson null strip: r(u_q, c) for every table column c, in-order double arithmetic
// bit-identical to coexpression_v2.c:105-113.  CTA tile = 16 queries x 64 columns, each thread
// owns 4 outputs and walks the samples left to right; operands are staged per 16-sample chunk.
__global__ void __launch_bounds__(256)
pearson_strip_kernel(const double *__restrict__ raw, int n, long long N, const int *__restrict__ qrows,
                     long long nq, const double *__restrict__ rsum, const double *__restrict__ rmean,
                     const double *__restrict__ rxx, double *__restrict__ out, long long ld) {
    __shared__ double sa[16][17], sb[64][17];
    __shared__ int srow[16];
    const int tid = threadIdx.x;
    const long long q0 = (long long)blockIdx.y * 16, c0 = (long long)blockIdx.x * 64;
    if (tid < 16) srow[tid] = (q0 + tid < nq) ? qrows[q0 + tid] : -1;
    __syncthreads();
    const int cl = tid & 63, qg = tid >> 6;            // column lane, query group (4 queries each)
    const long long c = c0 + cl;
    const double mb = (c < N) ? rmean[c] : 0.0;
    double ma[4], xy[4] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 4; ++e) { const int r = srow[qg * 4 + e]; ma[e] = (r >= 0) ? rmean[r] : 0.0; }
    const int lc = tid & 15, lr = tid >> 4;            // loader: 16 lanes along samples, 16 rows
    for (int k0 = 0; k0 < n; k0 += 16) {
        const int kk = k0 + lc;
        { const int r = srow[lr]; sa[lr][lc] = (r >= 0 && kk < n) ? raw[(long long)r * n + kk] : 0.0; }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long cc = c0 + lr + 16 * e;
            sb[lr + 16 * e][lc] = (cc < N && kk < n) ? raw[cc * n + kk] : 0.0;
        }
        __syncthreads();
        const int lim = min(16, n - k0);
        for (int j = 0; j < lim; ++j) {
            const double db = __dsub_rn(sb[cl][j], mb);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                xy[e] = __dadd_rn(xy[e], __dmul_rn(__dsub_rn(sa[qg * 4 + e][j], ma[e]), db));
        }
        __syncthreads();
    }
    if (c < N) {
        const double sumb = rsum[c], sqb = sqrt(rxx[c]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int r = srow[qg * 4 + e];
            if (r < 0) continue;
            double v;
            if (rsum[r] == 0.0 || sumb == 0.0) v = __longlong_as_double(0x7ff8000000000000LL);
            else v = __ddiv_rn(xy[e], __dmul_rn(sqrt(rxx[r]), sqb));
            out[(q0 + qg * 4 + e) * ld + c] = v;
        }
    }
}

}  // namespace coex
