// count_rows.cuh -- K3 epilogue stage, sort-free variant for the Spearman null (count mode 2): a third independent
// implementation of the counts, used by the cross-implementation parity tests.  On B200 it is slower than the default
// count_dedup.cuh (7.9 vs 4.6 ms per C2 step, profiles/r01_count_rows.md).
//
// Same contract as count.cuh (reference 1-make-network.py:239-258 null lists, :273-327 edge loop):
// for every query (permutation k, hit position i, table row u) and every other hit j of that
// permutation, idx = bisect_left(sorted null list of i, r(i, j)) = number of null values strictly
// below the statistic, then the edge score.  Work is de-duplicated over permutations exactly as in
// count_dedup.cuh: one work item = (unique table row u, a run of its queries), the row of the
// integer correlation strip is streamed ONCE for all of them.
//
// Difference to count_dedup.cuh: NOTHING IS SORTED and no value is searched.  Per chunk of 8192 null values
//   1. value -> float -> uniform bucket over [min statistic, max statistic] of the item (one subtract,
//      one multiply, one clamp); shared-memory histogram
//   2. exclusive scan of the 4096 bucket counts
//   3. counting-sort scatter of the chunk's column numbers by bucket
//   4. every statistic t, independently: (values in lower buckets) + (values of its own bucket window
//      that are below t).  The window is [bucket(t - delta), bucket(t + delta)]; the bucket function is
//      monotone, so a value outside the window differs from t by more than delta in float and its order
//      against t is certain; inside the window a value closer than delta is queued and re-evaluated in double
//      with the reference's exact arithmetic (its own hit column always is: value == statistic).  Windows made
//      long by tied values (sparse rows share identical correlations) are scanned by whole warps.  Counts are
//      therefore EXACT (tests/test_gpu_parity.py compares every index with the oracle).
// Float error budget: see cr_delta below (relative: 5e-7 |t|).
// Values below the item's lowest statistic are only counted (bucket 0), values above its highest one are
// ignored (last bucket): neither is histogrammed, so a narrow statistic range costs no atomic contention.
#pragma once
#include "common.cuh"
#include "count.cuh"
#include "count_dedup.cuh"

namespace coex {

constexpr int kCrThreads = 512;
constexpr int kCrBuckets = 4096;
constexpr int kCrChunk = 8192;          // null values staged in shared memory per pass
constexpr float kCrDelta = 5e-7f;
constexpr int kCrInline = 12;           // longer bucket windows are scanned by a whole warp
constexpr int kCrLong = 1024;
constexpr int kCrQueue = 3072;          // (statistic, column) pairs per chunk that need the reference's double arithmetic

struct RowCountArgs {
    const int *strip;               // [rows_in_strip, ld]
    long long ld;
    const DedupItem *items;         // items of this launch (plan.cuh)
    int n_items;
    const int *uniq_rows;           // table row of each unique-row slot of the strip
    const int *qlist;               // flat hit indices grouped by unique row
    const int *q_perm;              // [H] permutation of every flat hit index
    const long long *perm_off, *sc_off;
    const int *hit_rows;
    const double *sx;               // [N_pad]
    const float *icol;              // [N_pad] (float)(0.5 / sx), 0 for zero-variance / padding rows
    const unsigned char *rawsum_pos;
    long long N;
    const int *n_degenerate;        // [1] device: zero-variance table rows, their null value is -99 for every query
    int anticor;
    double *scores;
    long long *counts;
    int Tcap, Qcap;                 // shared-memory capacities: statistics and queries per item
    double *g_t64;                  // scratch [gridDim.x, Tcap]: statistics in double (exact refinements, scoring)
    const double *score_tab;        // [N+1] edge score of every possible bisect index (len = N-1)
    unsigned long long *phase_clk;  // optional [8]
};

__device__ __forceinline__ int cr_bucket(float v, float lo, float scale) {
    float f = __fmul_rn(__fsub_rn(v, lo), scale);        // monotone in v (no contraction: identical everywhere)
    f = fminf(fmaxf(f, 0.0f), (float)(kCrBuckets - 1));
    return (int)f;
}

// |v32 - v64| <= 5 * 2^-24 |v| (int -> float, two rounded reciprocal half-norms, two products: all relative) and
// |t32 - t64| <= 2^-24 |t|.  If |v| < 1.13 |t| the two errors add up to < 3.97e-7 |t|, otherwise |t - v| > 0.1 |v| dwarfs
// them; a float difference of at least 5e-7 |t| (4.4e-7 |t| after the rounding of t -+ delta in the window bounds)
// therefore decides the order.  The floor keeps exact ties at zero (d == 0) out of the "certain" branch.
__device__ __forceinline__ float cr_delta(float t) { return fmaxf(5e-7f * fabsf(t), 1e-30f); }

inline size_t rows_smem_bytes(int Tcap, int Qcap) {
    return (size_t)kCrChunk * 4 + (size_t)kCrBuckets * 4 + (size_t)kCrChunk * 2 + (size_t)Tcap * 8 + (size_t)kCrQueue * 4 +
           (size_t)Qcap * 16 + (size_t)Qcap * 8 + (size_t)(Qcap + 2) * 4 + 16;
}

__global__ void __launch_bounds__(kCrThreads, 2)
count_rows_kernel(RowCountArgs a) {
    const int Tcap = a.Tcap;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *chunkv = reinterpret_cast<float *>(smem_raw);                                  // [chunk] null values (NaN = not a list entry)
    unsigned int *cum = reinterpret_cast<unsigned int *>(chunkv + kCrChunk);               // [buckets]
    unsigned short *list = reinterpret_cast<unsigned short *>(cum + kCrBuckets);           // [chunk] columns grouped by bucket
    float *t32 = reinterpret_cast<float *>(list + kCrChunk);                               // [Tcap] statistics (NaN = undefined / diagonal)
    unsigned int *cnt = reinterpret_cast<unsigned int *>(t32 + Tcap);                      // [Tcap] in-range null values below the statistic
    unsigned int *queue = cnt + Tcap;                                                      // [kCrQueue] statistic << 13 | column in chunk
    double *q_vi = reinterpret_cast<double *>(queue + kCrQueue);                           // [Qcap] exact null value at the skipped column (quirk Q1)
    long long *q_sc = reinterpret_cast<long long *>(q_vi + a.Qcap);                        // [Qcap] offset of the query's score row
    int *q_i = reinterpret_cast<int *>(q_sc + a.Qcap);                                     // [Qcap] hit position
    int *q_off = q_i + a.Qcap;                                                             // [Qcap] offset of the query's permutation in hit_rows
    int *q_toff = q_off + a.Qcap;                                                          // [Qcap+2] offset of the query's statistics in the item
    __shared__ float s_min[kCrThreads / 32], s_max[kCrThreads / 32];
    __shared__ unsigned int s_warp[kCrThreads / 32];
    __shared__ float s_lo, s_scale;
    __shared__ int s_any, s_qn, s_ln;
    __shared__ unsigned short longq[kCrLong];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double *t64 = a.g_t64 + (long long)blockIdx.x * Tcap;
    const long long N = a.N;
    const float kNaN = __int_as_float(0x7fc00000);

    for (int it = blockIdx.x; it < a.n_items; it += gridDim.x) {
        const DedupItem item = a.items[it];
        const int nq = item.qe - item.qb;
        const int u = a.uniq_rows[item.urow];
        const int *srow = a.strip + (long long)item.urow * a.ld;
        const double sxu = a.sx[u];
        const bool posu = a.rawsum_pos[u] != 0;
        __syncthreads();                                   // previous item fully retired
        long long tclk = clock64();
#define CR_PHASE(id) do { if (a.phase_clk && tid == 0) { const long long now = clock64(); atomicAdd(&a.phase_clk[id], (unsigned long long)(now - tclk)); tclk = now; } } while (0)
        // ---- A. query metadata ----------------------------------------------------------------
        for (int ql = tid; ql < nq; ql += kCrThreads) {
            const long long q = a.qlist[item.qb + ql];
            const int k = a.q_perm[q];
            const long long off = a.perm_off[k];
            const int P = (int)(a.perm_off[k + 1] - off);
            const int i = (int)(q - off);
            q_i[ql] = i;
            q_off[ql] = (int)off;
            q_toff[ql + 1] = P;
            q_sc[ql] = a.sc_off[k] + (long long)i * P;
            double vi = -99.0;                             // null value at table column i (quirk Q1)
            const double sxc = a.sx[i];
            if (sxu != 0.0 && sxc != 0.0) vi = __ddiv_rn(0.25 * (double)srow[i], __dmul_rn(sxu, sxc));
            q_vi[ql] = vi;
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            q_toff[0] = 0;
            for (int ql = 0; ql < nq; ++ql) { acc += q_toff[ql + 1]; q_toff[ql + 1] = acc; }   // P each (diagonal included)
        }
        __syncthreads();
        const int TP = q_toff[nq];
        CR_PHASE(0);
        // ---- B. statistics r(i, j), evaluated once in double exactly as the reference does -------
        float tmin = __int_as_float(0x7f800000), tmax = __int_as_float(0xff800000);
        for (int t0 = tid; t0 < TP; t0 += 4 * kCrThreads) {
            int tt[4], qq[4], jj[4], cc4[4], sraw[4];
            double sxc4[4];
            bool live[4], guard[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                tt[e] = t0 + e * kCrThreads;
                live[e] = tt[e] < TP;
                qq[e] = 0; jj[e] = 0; cc4[e] = -1;
                if (live[e]) {
                    int lo = 0, hi = nq;
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (q_toff[mid] <= tt[e]) lo = mid; else hi = mid; }
                    qq[e] = lo; jj[e] = tt[e] - q_toff[lo];
                    if (jj[e] != q_i[lo]) cc4[e] = a.hit_rows[q_off[lo] + jj[e]];
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                guard[e] = false; sraw[e] = 0; sxc4[e] = 0.0;
                if (cc4[e] >= 0) { guard[e] = a.rawsum_pos[cc4[e]] != 0; sraw[e] = srow[cc4[e]]; sxc4[e] = a.sx[cc4[e]]; }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (!live[e]) continue;
                double r = -99.0;
                if (cc4[e] >= 0 && posu && guard[e]) {                      // coexfunctions.py:389
                    r = __ddiv_rn(0.25 * (double)sraw[e], __dmul_rn(sxu, sxc4[e]));
                    if (r != r) r = -99.0;                                  // coexfunctions.py:406-408
                }
                cnt[tt[e]] = 0;
                if (r == -99.0) {                                           // diagonal, or 1-make-network.py:284-285
                    const long long w = q_sc[qq[e]] + jj[e];
                    a.scores[w] = (cc4[e] < 0) ? 0.0 : -log10(1.0);
                    if (a.counts) a.counts[w] = -1;
                    t32[tt[e]] = kNaN;
                } else {
                    const float tf = (float)r;
                    t32[tt[e]] = tf;
                    t64[tt[e]] = r;
                    tmin = fminf(tmin, tf); tmax = fmaxf(tmax, tf);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
            tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        }
        if (lane == 0) { s_min[wid] = tmin; s_max[wid] = tmax; }
        __syncthreads();
        if (tid == 0) {
            float mn = s_min[0], mx = s_max[0];
            for (int w = 1; w < kCrThreads / 32; ++w) { mn = fminf(mn, s_min[w]); mx = fmaxf(mx, s_max[w]); }
            s_any = (mn <= mx) ? 1 : 0;
            s_qn = 0;
            // buckets 2 .. B-4 cover [min - pad, max + pad]; a range floor keeps the scale finite
            const float pad = 4.0f * kCrDelta;
            const float range = fmaxf(mx - mn, 1e-3f) + 2.0f * pad;
            const float scale = (float)(kCrBuckets - 6) / range;
            s_scale = scale;
            s_lo = mn - pad - 2.5f / scale;
        }
        __syncthreads();
        CR_PHASE(1);
        if (!s_any) continue;                               // every statistic undefined: scores already written
        const float blo = s_lo, bscale = s_scale;
        const float iu = (float)(0.5 / sxu);
        unsigned int n_below = 0;                           // values below every statistic of the item (bucket 0)
        // ---- C. chunks of null values ----------------------------------------------------------------
        for (long long c0 = 0; c0 < N; c0 += kCrChunk) {
            const int clen = (int)((N - c0 < kCrChunk) ? (N - c0) : kCrChunk);
            for (int b = tid; b < kCrBuckets; b += kCrThreads) cum[b] = 0;
            if (tid == 0) { s_qn = 0; s_ln = 0; }
            // strip row and column norms of the chunk: all loads issued before the first use
            constexpr int kIter = kCrChunk / (kCrThreads * 4);
            int4 s4[kIter];
            float4 w4[kIter];
#pragma unroll
            for (int r = 0; r < kIter; ++r) {
                const int base = (r * kCrThreads + tid) * 4;
                s4[r] = make_int4(0, 0, 0, 0); w4[r] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (base < clen) {                          // ld and icol are padded to a multiple of 256 columns
                    s4[r] = *reinterpret_cast<const int4 *>(srow + c0 + base);
                    w4[r] = *reinterpret_cast<const float4 *>(a.icol + c0 + base);
                }
            }
            __syncthreads();
            // 1. float value, bucket, histogram
#pragma unroll
            for (int r = 0; r < kIter; ++r) {
                const int base = (r * kCrThreads + tid) * 4;
                if (base >= clen) continue;
                const int sv[4] = {s4[r].x, s4[r].y, s4[r].z, s4[r].w};
                const float wv[4] = {w4[r].x, w4[r].y, w4[r].z, w4[r].w};
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    // zero-variance column: -99, counted in bulk (n_degenerate); padding column: nothing
                    const bool act = (wv[e] != 0.0f) && (base + e < clen);
                    v[e] = __int2float_rn(sv[e]) * (iu * wv[e]);
                    int b = act ? cr_bucket(v[e], blo, bscale) : (kCrBuckets - 1);
                    if (b == 0) ++n_below;
                    if (b == 0 || b == kCrBuckets - 1) v[e] = kNaN;
                    else atomicAdd(&cum[b], 1u);
                }
                *reinterpret_cast<float4 *>(chunkv + base) = make_float4(v[0], v[1], v[2], v[3]);
            }
            __syncthreads();
            {   // 2. exclusive scan of the bucket counts (each warp scans a contiguous slice, bank-conflict free)
                constexpr int kWarps = kCrThreads / 32, kPerWarp = kCrBuckets / kWarps;
                unsigned int carry = 0;
                for (int c = 0; c < kPerWarp; c += 32) {
                    const int b = wid * kPerWarp + c + lane;
                    const unsigned int v = cum[b];
                    unsigned int incl = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const unsigned int x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
                    cum[b] = carry + incl - v;
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
                if (lane == 0) s_warp[wid] = carry;
                __syncthreads();
                unsigned int base = 0;
                for (int w = 0; w < wid; ++w) base += s_warp[w];
                if (base)
                    for (int c = 0; c < kPerWarp; c += 32) cum[wid * kPerWarp + c + lane] += base;
                __syncthreads();
            }
            // 3. scatter the columns by bucket; afterwards cum[b] = end of bucket b = start of bucket b + 1
            for (int c = tid; c < clen; c += kCrThreads) {
                const float v = chunkv[c];
                if (v == v) {
                    const unsigned int pos = atomicAdd(&cum[cr_bucket(v, blo, bscale)], 1u);
                    list[pos] = (unsigned short)c;
                }
            }
            __syncthreads();
            // 4. every statistic against its bucket window.  A value that cannot be ordered in float is NOT resolved
            //    inline (one lane's double division and its two global loads would stall the other 31, and tied values
            //    -- sparse rows share a few hundred identical correlations -- pile up on single statistics): the
            //    (statistic, column) pair is queued and the whole CTA resolves the queue together.
            auto scan_entry = [&](int tt, float t, float dl, unsigned int m) -> unsigned int {
                const int c = list[m];
                const float d = t - chunkv[c];
                if (d >= dl) return 1u;
                if (d > -dl) {                                              // cannot be ordered in float
                    const int slot = atomicAdd(&s_qn, 1);
                    if (slot < kCrQueue) {
                        queue[slot] = ((unsigned int)tt << 13) | (unsigned int)c;
                    } else {                                                // queue full: reference arithmetic in place
                        const long long col = c0 + c;
                        const double vv = __ddiv_rn(0.25 * (double)srow[col], __dmul_rn(sxu, a.sx[col]));
                        return (vv < t64[tt]) ? 1u : 0u;
                    }
                }
                return 0u;
            };
            for (int tt = tid; tt < TP; tt += kCrThreads) {
                const float t = t32[tt];
                if (!(t == t)) continue;
                const float dl = cr_delta(t);
                const int bl = cr_bucket(t - dl, blo, bscale), bh = cr_bucket(t + dl, blo, bscale);
                const unsigned int s = cum[bl > 0 ? bl - 1 : 0], e = cum[bh];   // 2 <= bl <= bh <= B-4 by construction
                unsigned int less = s;
                if (e - s > (unsigned int)kCrInline) {
                    // tied values (sparse rows share identical correlations) make a few windows hundreds of entries long:
                    // those statistics are handed to whole warps so that one lane does not hold up the other 31
                    const int slot = atomicAdd(&s_ln, 1);
                    if (slot < kCrLong) { longq[slot] = (unsigned short)tt; cnt[tt] += less; continue; }
                }
                for (unsigned int m = s; m < e; ++m) less += scan_entry(tt, t, dl, m);
                cnt[tt] += less;
            }
            __syncthreads();
            {
                const int ln = min(s_ln, kCrLong);
                for (int x = wid; x < ln; x += kCrThreads / 32) {
                    const int tt = longq[x];
                    const float t = t32[tt];
                    const float dl = cr_delta(t);
                    const int bl = cr_bucket(t - dl, blo, bscale), bh = cr_bucket(t + dl, blo, bscale);
                    const unsigned int s = cum[bl > 0 ? bl - 1 : 0], e = cum[bh];
                    unsigned int less = 0;
                    for (unsigned int m = s + lane; m < e; m += 32) less += scan_entry(tt, t, dl, m);
                    less = __reduce_add_sync(0xffffffffu, less);
                    if (lane == 0) cnt[tt] += less;
                }
            }
            __syncthreads();
            const int qn = min(s_qn, kCrQueue);
            for (int x = tid; x < qn; x += kCrThreads) {
                const unsigned int ent = queue[x];
                const int tt = (int)(ent >> 13);
                const long long col = c0 + (ent & 8191u);
                const double vv = __ddiv_rn(0.25 * (double)srow[col], __dmul_rn(sxu, a.sx[col]));
                if (vv < t64[tt]) atomicAdd(&cnt[tt], 1u);
            }
            __syncthreads();
        }
        CR_PHASE(2);
        // ---- D. bisect index and edge score of every statistic ---------------------------------------
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_below += __shfl_xor_sync(0xffffffffu, n_below, o);
        if (lane == 0) s_warp[wid] = n_below;
        __syncthreads();
        long long below = *a.n_degenerate;                  // -99 < t always
        for (int w = 0; w < kCrThreads / 32; ++w) below += s_warp[w];
        for (int tt = tid; tt < TP; tt += kCrThreads) {
            const float tf = t32[tt];
            if (!(tf == tf)) continue;
            int lo = 0, hi = nq;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (q_toff[mid] <= tt) lo = mid; else hi = mid; }
            const int j = tt - q_toff[lo];
            const int i = q_i[lo];
            const double t = t64[tt];
            // all columns strictly below t, minus the skipped column i of this query if it is below t (quirk Q1)
            long long idx = (long long)cnt[tt] + below;
            long long len = N;
            if (i < N) { len = N - 1; if (q_vi[lo] < t) idx -= 1; }
            const long long w = q_sc[lo] + j;
            a.scores[w] = (len == N - 1) ? a.score_tab[idx] : edge_score_dev(idx, len, a.anticor);
            if (a.counts) a.counts[w] = idx;
        }
        CR_PHASE(3);
        if (a.phase_clk && tid == 0) atomicAdd(&a.phase_clk[6], 1ULL);
    }
}

// permutation of every flat hit index (one thread per hit; perm_off has K + 1 entries)
__global__ void query_perm_kernel(const long long *__restrict__ perm_off, int K, long long H, int *__restrict__ q_perm) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= H) return;
    int lo = 0, hi = K;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (perm_off[mid] <= q) lo = mid; else hi = mid; }
    q_perm[q] = lo;
}

}  // namespace coex
