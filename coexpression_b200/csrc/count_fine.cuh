// count_fine.cuh -- K3 epilogue stage, Spearman null, round-2 design: "a streamed value never looks at the statistics".
//
// Same contract and same work items as count_dedup_kernel<0> (reference 1-make-network.py:239-258, 273-327): for a unique hit
// row u and the statistics t_j of its queries, the number of null values of row u strictly below every t_j (bisect_left on the
// sorted null list), turned into the edge score.  What changed is the shape of the work (profiles/r02_count_experiments.md:
// the sorted-statistics kernel is bound by its ~70 instructions per streamed value and by a set-up that sorts the statistics):
//
//   * ONE table in shared memory, nb (16384 / 32768) buckets over the whole of [-1, 1] (piecewise linear, three quarters of them
//     for the centre: cf_bucket), one 32-bit word per bucket:
//         bits  0..17  number of streamed values of the current chunk of columns that fell into the bucket
//         bits 18..30  head of the chain of statistics that live in the bucket (node + 1, 0 = none)
//         bit  31      MULTI: more than one statistic, or a statistic of a NEIGHBOUR bucket that is closer than the float
//                      uncertainty to this bucket's edge
//     A streamed value costs one shared-memory atomic add on its bucket; the RETURNED word says whether a statistic lives
//     there.  No search, no sorted statistics, no neighbours: the statistics are never sorted at all.
//   * #values below t_j = (values in the buckets below the bucket of t_j: a prefix over the table, per chunk of <= 258048
//     columns so that 18 bits suffice) + (values of t_j's own bucket below t_j: counted by the values themselves -- 85 % of
//     the values see an empty bucket, most of the others one statistic and a float compare that is certain).
//   * Everything else -- MULTI buckets, a float compare that is not certain (|v - t| < 7e-7 |t|), exact ties -- goes through
//     per-warp queues (no barrier inside a chunk: a warp drains its own queue while the others stream) to a slow path with the
//     reference's double arithmetic; statistics that are the same double are counted once (alias); a statistic closer than the float
//     uncertainty to a bucket edge is ALSO chained into the neighbouring bucket, whose values then correct the prefix
//     (+-1) when the doubles disagree with the bucket order.  Counts are EXACT, as before.
//   * An item this kernel does not want (no room for the chains) is appended to a re-do list that count_dedup_kernel<0>
//     works off afterwards: there is no capacity limit that is not covered by the old kernel.
#pragma once
#include <math_constants.h>
#include "common.cuh"
#include "count_dedup.cuh"

namespace coex {

constexpr int kCfThreads = 1024;
constexpr int kCfTcap = 4096;            // statistics per work item
constexpr int kCfPool = 1024;            // extra chain nodes (statistics chained into a neighbour bucket)
constexpr int kCfWarpDrainAt = 64;       // a warp's queue of columns waiting for the slow path is drained when it holds this many ...
constexpr int kCfWarpQueue = kCfWarpDrainAt + 4 * 32;   // ... and one trip of the warp cannot add more than 4 x 32: it never overflows
constexpr int kCfQueue = kCfWarpQueue * (kCfThreads / 32);
constexpr int kCfQcap = 8;               // queries per work item (plan.cuh: kPlanQcap)
constexpr int kCfCntBits = 18;           // table word: value counter | head of the chain (13 bits: node + 1) | MULTI
constexpr unsigned int kCfCntMask = (1u << kCfCntBits) - 1u, kCfMulti = 0x80000000u, kCfHeadMask = 0x1fffu;
constexpr int kCfChunkMax = 258048;      // columns per pass of the counters (multiple of 4 * kCfThreads, < 2^18): one pass for FANTOM5
constexpr float kCfDelta = 7e-7f;        // see cf_delta

// The float value of a column and of a statistic are the SAME expression, fl(fl(int -> float) * fl(fl(0.5 / sx_u) * fl(0.5 / sx_c))):
// five roundings, |x32 - x64| <= 5 * 2^-24 |x| = 2.98e-7 |x| each.  With d = v32 - t32: the doubles are ordered like the floats
// whenever |d| > 3e-7 (|v| + |t|); |v| <= |t32| + |d| (up to 1 + 3e-7) makes |d| >= 6.1e-7 |t32| sufficient.  t32 == 0 is an exact
// zero (integer sum 0; the smallest non-zero |value| of a row of n <= 8192 samples is ~1e-12, far from the denormals) and any float
// order against it is the true one: no floor is needed, and v32 == t32 == 0 is "certain" and not below, which is right.
__device__ __forceinline__ float cf_delta(float t) { return kCfDelta * fabsf(t); }

// Bucket of a correlation: piecewise linear over the whole of [-1, 1] -- three quarters of the buckets for the centre |x| <= r1
// (6 standard deviations of a null correlation), one eighth for each tail (co-expressed rows are NOT null-distributed: a table
// with shared expression programmes has a fifth of its hit-vs-hit statistics out there, and a clamped uniform table would put
// them, and their values, into two edge buckets).  Monotone non-decreasing in x: the clamp, both fused multiply-adds (k >= 0,
// st > 0), the truncation and the final clamps all are, so  v32 <= t32  =>  bucket(v32) <= bucket(t32).
__device__ __forceinline__ int cf_bucket(float x, const CfMap &m) {
    const float xc = fminf(fmaxf(x, -m.r1), m.r1);
    const int b = __float2int_rz(__fmaf_rn(xc, m.k, __fmaf_rn(x, m.st, m.half)));
    return min(max(b, 0), m.nbm1);
}

// Shared-memory layout: the arrays of fixed size first (constant offsets), the table last.
//   t32 [kCfTcap] float statistic | cnt [kCfTcap] values below it | col [kCfTcap] its table column (-1: no statistic)
//   nxt [kCfTcap + kCfPool] next node + 1 | nprobe [kCfPool] statistic of a pool node | alias [kCfTcap] first of the statistics equal to it
//   queue [kCfQueue] | rowpre [kCfMaxBuckets / 32] prefix of the 32-bucket rows | tab [nb + 32] (tab[nb]: where inactive columns go)
constexpr int kCfMaxBuckets = 32768;
constexpr int kCfOffCnt = kCfTcap * 4, kCfOffCol = kCfTcap * 8, kCfOffNxt = kCfTcap * 12, kCfOffProbe = kCfOffNxt + (kCfTcap + kCfPool) * 2,
              kCfOffAlias = kCfOffProbe + kCfPool * 2, kCfOffQueue = kCfOffAlias + kCfTcap * 2, kCfOffRow = kCfOffQueue + kCfQueue * 4,
              kCfOffTab = kCfOffRow + (kCfMaxBuckets / 32) * 4;
inline size_t fine_smem_bytes(int nb) { return (size_t)kCfOffTab + (size_t)(nb + 32) * 4; }

struct CfCtx {                           // what the slow path needs besides the shared-memory arrays
    const int *srow;
    const double *sx;
    const float *icol;
    double sxu;
    float iu;
    const CfMap &m;
};
#define CF_SMEM_ARRAYS                                                                                              \
    extern __shared__ __align__(16) unsigned char smem_raw[];                                                       \
    float *const t32 = reinterpret_cast<float *>(smem_raw);                                                         \
    int *const cnt = reinterpret_cast<int *>(smem_raw + kCfOffCnt);                                                 \
    int *const col = reinterpret_cast<int *>(smem_raw + kCfOffCol);                                                 \
    unsigned short *const nxt = reinterpret_cast<unsigned short *>(smem_raw + kCfOffNxt);                           \
    unsigned short *const nprobe = reinterpret_cast<unsigned short *>(smem_raw + kCfOffProbe);                      \
    unsigned short *const alias = reinterpret_cast<unsigned short *>(smem_raw + kCfOffAlias);                       \
    int *const queue = reinterpret_cast<int *>(smem_raw + kCfOffQueue);                                             \
    unsigned int *const rowpre = reinterpret_cast<unsigned int *>(smem_raw + kCfOffRow);                            \
    unsigned int *const tab = reinterpret_cast<unsigned int *>(smem_raw + kCfOffTab)

// r(u, c1) < r(u, c2) with the reference's arithmetic (coexpression_v2.c:113 on rank rows: xy / (sqrt(xx) * sqrt(yy)), every
// partial sum exact in double -- DESIGN.md section 4); equal integer sums and equal norms are the same double
__device__ __forceinline__ bool cf_less_exact(const CfCtx &x, int c1, int c2) {
    const int s1 = x.srow[c1], s2 = x.srow[c2];
    const double n1 = x.sx[c1], n2 = x.sx[c2];
    if (s1 == s2 && n1 == n2) return false;
    const double r1 = __ddiv_rn(0.25 * (double)s1, __dmul_rn(x.sxu, n1));
    const double r2 = __ddiv_rn(0.25 * (double)s2, __dmul_rn(x.sxu, n2));
    return r1 < r2;
}

// one streamed column against every statistic chained into its bucket
__device__ __forceinline__ void cf_slow(const CfCtx &x, int c) {
    CF_SMEM_ARRAYS;
    (void)alias; (void)queue; (void)rowpre;
    const float v = __int2float_rn(x.srow[c]) * (x.iu * x.icol[c]);      // the streaming loop's expression: the same bits
    const int b = cf_bucket(v, x.m);
    int id = (int)((tab[b] >> kCfCntBits) & kCfHeadMask) - 1;
    while (id >= 0) {
        const int j = (id < kCfTcap) ? id : (int)nprobe[id - kCfTcap];
        const float t = t32[j];
        const float d = v - t;
        const int hb = cf_bucket(t, x.m);
        if (fabsf(d) >= cf_delta(t)) {
            // certain: a value of the statistic's own bucket is counted here, a value of another bucket by the prefix
            if (hb == b && d < 0.0f) atomicAdd(&cnt[j], 1);
        } else if (c != col[j]) {                       // (the statistic's own column: equal by construction)
            const bool less = cf_less_exact(x, c, col[j]);
            if (hb > b) { if (!less) atomicSub(&cnt[j], 1); }        // the prefix has counted it as below
            else if (less) atomicAdd(&cnt[j], 1);                    // the prefix has not
        }
        id = (int)nxt[id] - 1;
    }
}

__global__ void __launch_bounds__(kCfThreads, 1)
count_fine_kernel(DedupArgs a) {
    CF_SMEM_ARRAYS;
    const int nb = a.fine_nb;
    __shared__ double q_vi[kCfQcap];          // exact null value at the skipped column (quirk Q1), -99 = undefined
    __shared__ float q_vi32[kCfQcap];
    __shared__ long long q_sc[kCfQcap], q_cn[kCfQcap];
    __shared__ int q_i[kCfQcap], q_off[kCfQcap], q_toff[kCfQcap + 2];
    __shared__ int s_pool, s_decline, s_live, s_next;
    __shared__ unsigned int s_warp[kCfThreads / 32];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const CfMap &bm = a.fine_map;              // (kernel parameter: its members are constant-bank operands, not registers)
    const long long N = a.N;
    const int n_deg = *a.n_degenerate;
    const long long chunk = a.fine_chunk;

    auto next_item = [&](int it) {
        if (!a.work_counter) return it + (int)gridDim.x;
        __syncthreads();
        if (threadIdx.x == 0) s_next = (int)gridDim.x + atomicAdd(a.work_counter, 1);
        __syncthreads();
        return s_next;
    };
    for (int it = blockIdx.x; it < a.n_items; it = next_item(it)) {
        const DedupItem item = a.items[it];
        const int nq = item.qe - item.qb;
        const int u = a.uniq_rows[item.urow];
        const int *srow = a.strip + (long long)item.urow * a.ld;
        const double sxu = a.sx[u];
        const bool posu = a.rawsum_pos[u] != 0;
        const float iu = (sxu != 0.0) ? (float)(0.5 / sxu) : 0.0f;
        const CfCtx cx{srow, a.sx, a.icol, sxu, iu, bm};
        __syncthreads();                                   // previous item fully retired
        long long tclk = clock64();
#define CF_PHASE(id) do { if (a.phase_clk && tid == 0) { const long long now = clock64(); atomicAdd(&a.phase_clk[id], (unsigned long long)(now - tclk)); tclk = now; } } while (0)
        // ---- A. query metadata, empty table ---------------------------------------------------------
        if (tid < nq) {
            const long long q = a.qlist[item.qb + tid];
            const int k = a.q_perm[q];
            const long long off = a.perm_off[k];
            const int P = (int)(a.perm_off[k + 1] - off);
            const int i = (int)(q - off);
            q_i[tid] = i;
            q_off[tid] = (int)off;
            q_toff[tid + 1] = P;                            // turned into offsets below
            double *sbase = a.perm_scores ? a.perm_scores[k] : a.scores + a.sc_off[k];
            q_sc[tid] = (long long)(size_t)(sbase + (long long)i * P);
            q_cn[tid] = a.sc_off[k] + (long long)i * P;
            double vi = -99.0;                             // null value at table column i (quirk Q1)
            float vi32 = 0.0f;
            if (i < N) {
                const double sxc = a.sx[i];
                if (sxu != 0.0 && sxc != 0.0) {
                    const int s = srow[i];
                    vi = __ddiv_rn(0.25 * (double)s, __dmul_rn(sxu, sxc));
                    vi32 = __int2float_rn(s) * (iu * a.icol[i]);
                }
            }
            q_vi[tid] = vi; q_vi32[tid] = vi32;
        }
        if (tid == 0) { s_pool = 0; s_decline = 0; s_live = 0; }
        for (int w = tid; w < ((nb + 32) >> 2); w += kCfThreads) reinterpret_cast<uint4 *>(tab)[w] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int ql = 0; ql < nq; ++ql) { const int P = q_toff[ql + 1]; q_toff[ql] = acc; acc += P; }
            q_toff[nq] = acc;
        }
        __syncthreads();
        const int TP = q_toff[nq];
        if (TP > kCfTcap || (a.fine_decline_mod > 0 && it % a.fine_decline_mod == 0)) {      // (the plan cuts items at the host's Tcap <= kCfTcap)
            if (tid == 0) a.redo_list[atomicAdd(a.redo_count, 1)] = it;
            continue;
        }
        CF_PHASE(0);
        // ---- B. statistics: float value, chained into their bucket(s); undefined ones are scored at once ---------
        for (int t0 = tid; t0 < TP; t0 += 4 * kCfThreads) {
            int qq[4], jj[4], cc4[4], sraw[4];
            float w4[4];
            bool ok[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int tt = t0 + e * kCfThreads;
                qq[e] = 0; jj[e] = 0; cc4[e] = -1;
                if (tt < TP) {
                    int lo = 0, hi = nq;
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (q_toff[mid] <= tt) lo = mid; else hi = mid; }
                    qq[e] = lo; jj[e] = tt - q_toff[lo];
                    if (jj[e] != q_i[lo]) cc4[e] = a.hit_rows[q_off[lo] + jj[e]];
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                sraw[e] = 0; w4[e] = 0.0f; ok[e] = false;
                if (cc4[e] >= 0) {
                    sraw[e] = srow[cc4[e]];
                    w4[e] = a.icol[cc4[e]];                                   // 0: zero-variance row -> 0 / 0 -> NaN -> -99
                    ok[e] = a.rawsum_pos[cc4[e]] != 0;                        // coexfunctions.py:389
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int tt = t0 + e * kCfThreads;
                if (tt >= TP) continue;
                const bool defined = cc4[e] >= 0 && posu && ok[e] && w4[e] != 0.0f && sxu != 0.0;   // coexfunctions.py:389, 406-408
                if (!defined) {                                               // diagonal, or 1-make-network.py:284-285
                    col[tt] = -1;
                    reinterpret_cast<double *>((size_t)q_sc[qq[e]])[jj[e]] = (cc4[e] < 0) ? 0.0 : -log10(1.0);
                    if (a.counts) a.counts[q_cn[qq[e]] + jj[e]] = -1;
                    continue;
                }
                const float t = __int2float_rn(sraw[e]) * (iu * w4[e]);
                t32[tt] = t; cnt[tt] = sraw[e]; col[tt] = cc4[e];              // (cnt holds the integer sum until the aliases are known)
                const int hb = cf_bucket(t, bm);
                const unsigned int old = atomicExch(&tab[hb], (unsigned int)(tt + 1) << kCfCntBits);
                nxt[tt] = (unsigned short)(old >> kCfCntBits);
                s_live = 1;
            }
        }
        __syncthreads();
        if (!s_live) continue;                              // every statistic undefined: scores already written
        // Statistics that are the SAME double (same column, or equal integer sum and equal norm: rows of a zero-inflated table
        // with the same few non-zero samples share their rank vector) are counted once: alias = the first of them.  Without this
        // a run of L equal statistics meets a run of V equal values in the slow path, L x V exact comparisons.
        for (int tt = tid; tt < TP; tt += kCfThreads) {
            if (col[tt] < 0) continue;
            const float t = t32[tt];
            const int sraw = cnt[tt], c = col[tt];
            int first = tt;
            for (int id = (int)(tab[cf_bucket(t, bm)] >> kCfCntBits) - 1; id >= 0; id = (int)nxt[id] - 1) {
                if (id >= first || t32[id] != t || cnt[id] != sraw) continue;
                if (col[id] == c || a.sx[col[id]] == a.sx[c]) first = id;
            }
            alias[tt] = (unsigned short)first;
        }
        __syncthreads();
        for (int tt = tid; tt < TP; tt += kCfThreads) if (col[tt] >= 0) tab[cf_bucket(t32[tt], bm)] = 0u;
        __syncthreads();
        for (int tt = tid; tt < TP; tt += kCfThreads) {
            if (col[tt] < 0) continue;
            cnt[tt] = 0;
            if ((int)alias[tt] != tt) continue;
            const float t = t32[tt];
            const float dl = cf_delta(t);
            const int hb = cf_bucket(t, bm);
            const int b0 = cf_bucket(t - dl, bm), b1 = cf_bucket(t + dl, bm);
            for (int b = b0; b <= b1; ++b) {
                int node = tt;
                if (b != hb) {                             // closer to the neighbour bucket than the float uncertainty
                    const int pnode = atomicAdd(&s_pool, 1);
                    if (pnode >= kCfPool) { s_decline = 1; continue; }
                    nprobe[pnode] = (unsigned short)tt;
                    node = kCfTcap + pnode;
                }
                const unsigned int old = atomicExch(&tab[b], ((unsigned int)(node + 1) << kCfCntBits) | (b != hb ? kCfMulti : 0u));
                nxt[node] = (unsigned short)((old >> kCfCntBits) & kCfHeadMask);
                if (old >> kCfCntBits) atomicOr(&tab[b], kCfMulti);           // (keeps a MULTI bit the exchange has dropped)
            }
        }
        __syncthreads();
        CF_PHASE(1);
        if (s_decline) {
            if (tid == 0) a.redo_list[atomicAdd(a.redo_count, 1)] = it;
            continue;
        }
        // ---- C. stream the null values of table row u ---------------------------------------------------------------
        int wq = 0;                                        // columns in this warp's queue (the same in every lane)
        int *const myq = queue + wid * kCfWarpQueue;
        for (long long ch0 = 0; ch0 < N; ch0 += chunk) {
            const long long chend = (ch0 + chunk < N) ? ch0 + chunk : N;
            {
                const int sb = (int)ch0, send = (int)chend;                    // (N < 2^31: hit rows are ints)
                // every thread makes the same number of trips (a thread behind the end carries four inactive columns): the warp
                // stays converged for its queue push below
                int4 s4n = make_int4(0, 0, 0, 0);
                float4 w4n = make_float4(0.f, 0.f, 0.f, 0.f);
                if (sb + tid * 4 < send) {
                    s4n = *reinterpret_cast<const int4 *>(srow + sb + tid * 4); w4n = *reinterpret_cast<const float4 *>(a.icol + sb + tid * 4);
                }
                for (int cb = sb; cb < send; cb += 4 * kCfThreads) {
                    const int sv[4] = {s4n.x, s4n.y, s4n.z, s4n.w};
                    const float wv[4] = {w4n.x, w4n.y, w4n.z, w4n.w};
                    const int cc = cb + tid * 4;
                    const int cn = cc + 4 * kCfThreads;
                    w4n = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (cn < send) { s4n = *reinterpret_cast<const int4 *>(srow + cn); w4n = *reinterpret_cast<const float4 *>(a.icol + cn); }
                    float v[4];
                    unsigned int h[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        // zero-variance column (and the padding behind column N): -99, counted in bulk (n_deg)
                        // (they go to the spare word behind the table, whose upper half stays 0; the bucket is computed for every
                        //  lane -- the empty asm keeps the compiler from branching around seven instructions)
                        v[e] = __int2float_rn(sv[e]) * (iu * wv[e]);
                        int b = cf_bucket(v[e], bm);
                        asm volatile("" : "+r"(b));
                        b = (wv[e] != 0.0f) ? b : nb;
                        h[e] = atomicAdd(&tab[b], 1u) >> kCfCntBits;
                    }
                    // (a straight-line second stage for all four values -- statistic 0 read and discarded where the bucket has none --
                    //  executes fewer instructions but was measured 7 % slower: two more shared-memory loads per value)
                    unsigned int cm = 0u;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (h[e] == 0u) continue;
                        if (h[e] & (kCfMulti >> kCfCntBits)) { cm |= 1u << e; continue; }
                        const int j = (int)h[e] - 1;                          // the bucket's only statistic, at home
                        const float t = t32[j];
                        const float d = v[e] - t;
                        if (fabsf(d) >= cf_delta(t)) { if (d < 0.0f) atomicAdd(&cnt[j], 1); }
                        else if (cc + e != col[j]) cm |= 1u << e;
                    }
                    if (__any_sync(0xffffffffu, cm != 0u)) {
                        // the warp's own queue: exclusive prefix of the lanes' counts, no atomic, no barrier
                        const int mine = __popc(cm);
                        int incl = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
                        int slot = wq + incl - mine;
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if ((cm >> e) & 1u) myq[slot++] = cc + e;
                        wq += __shfl_sync(0xffffffffu, incl, 31);
                        if (wq >= kCfWarpDrainAt) {                           // (other warps keep streaming meanwhile)
                            __syncwarp();
                            for (int x = lane; x < wq; x += 32) cf_slow(cx, myq[x]);
                            __syncwarp();
                            wq = 0;
                        }
                    }
                }
            }
            if (wq > 0) {                                  // what is left in the warp's queue (the other warps are still streaming)
                __syncwarp();
                for (int x = lane; x < wq; x += 32) cf_slow(cx, myq[x]);
                wq = 0;
            }
            __syncthreads();                               // every value of the chunk is in the table
            CF_PHASE(3);
            // ---- D. prefix over the table: values of this chunk in the buckets below every statistic ----------------
            {
                // 32-bucket rows: a thread sums one row (its 32 words in an order rotated by the row number: the lanes of a warp hit
                // 32 different banks), a block-wide scan of the row totals, and every statistic adds the buckets of its own row in
                // front of it
                const int rows = nb >> 5;                                      // <= kCfThreads
                unsigned int mine = 0u;
                if (tid < rows) {
#pragma unroll 8
                    for (int i = 0; i < 32; ++i) mine += tab[(tid << 5) + ((i + tid) & 31)] & kCfCntMask;
                }
                unsigned int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
                if (lane == 31) s_warp[wid] = incl;
                __syncthreads();
                unsigned int base = 0u;
                for (int w = 0; w < wid; ++w) base += s_warp[w];
                if (tid < rows) rowpre[tid] = base + incl - mine;
                __syncthreads();
                for (int tt = tid; tt < TP; tt += kCfThreads) {
                    if (col[tt] < 0 || (int)alias[tt] != tt) continue;
                    const int hb = cf_bucket(t32[tt], bm);
                    unsigned int below = rowpre[hb >> 5];
                    for (int b = hb & ~31; b < hb; ++b) below += tab[b] & kCfCntMask;   // (eight 16-byte loads of the whole row: slower)
                    cnt[tt] += (int)below;
                }
                __syncthreads();
                if (chend < N)                                                 // another chunk follows: counters back to zero
                    for (int w = tid; w < nb; w += kCfThreads) tab[w] &= ~kCfCntMask;
                __syncthreads();
            }
            CF_PHASE(4);
        }
        // ---- F. scores ------------------------------------------------------------------------------------------
        for (int t0 = tid; t0 < TP; t0 += 4 * kCfThreads) {
            long long idx4[4], w4[4], wc4[4];
            bool tab4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int tt = t0 + e * kCfThreads;
                w4[e] = 0; wc4[e] = 0; idx4[e] = 0; tab4[e] = true;
                if (tt < TP && col[tt] >= 0) {
                    int lo = 0, hi = nq;
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (q_toff[mid] <= tt) lo = mid; else hi = mid; }
                    const int j = tt - q_toff[lo];
                    const int i = q_i[lo];
                    // all columns strictly below t: streamed values + zero-variance columns (-99 < t always),
                    // minus the skipped column i of this query if it is below t (quirk Q1)
                    long long idx = (long long)cnt[alias[tt]] + n_deg;
                    if (i < N) {
                        bool below;
                        if (q_vi[lo] == -99.0) below = true;
                        else {
                            const float t = t32[tt];
                            const float d = q_vi32[lo] - t;
                            if (fabsf(d) >= cf_delta(t)) below = d < 0.0f;
                            else below = (i != col[tt]) && cf_less_exact(cx, i, col[tt]);
                        }
                        if (below) idx -= 1;
                    } else tab4[e] = false;                                   // no skipped column: N values
                    idx4[e] = idx;
                    w4[e] = q_sc[lo] + 8LL * j;
                    wc4[e] = q_cn[lo] + j;
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (w4[e] == 0) continue;
                const double sc = tab4[e] ? a.score_tab[idx4[e]] : edge_score_dev(idx4[e], N, a.anticor);
                *reinterpret_cast<double *>((size_t)w4[e]) = sc;
                if (a.counts) a.counts[wc4[e]] = idx4[e];
            }
        }
        CF_PHASE(5);
        if (a.phase_clk && tid == 0) atomicAdd(&a.phase_clk[6], 1ULL);
    }
#undef CF_PHASE
}

}  // namespace coex
