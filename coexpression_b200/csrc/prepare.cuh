// prepare.cuh -- per-row work done once per table:
//   K1  rank_pack_kernel    average-tie ranks (scipy.stats.rankdata 'average' as called at
//                           reference 1-make-network.py:110-111), fused with the packing of
//                           u = 2*rank-(n+1) into two int8 digit planes for the exact integer
//                           correlation GEMM, sum(u^2) and sqrt(sum/4).
//   K1b rowstats_kernel     the per-row quantities reference coexpression_v2.c:74-110
//                           recomputes for every pair: left-to-right sum, mean = sum/n and
//                           sum((x-mean)^2), bit-identical to the reference's doubles.
#pragma once
#include "common.cuh"

namespace coex {

// -------------------------------------------------------------------------------------------
// K1: one CTA per row.  Shared-memory bitonic sort of (value, column) pairs, tie runs located
// by binary search in the sorted keys, 2*rank = first+last+2 scattered back to column order,
// then the packed outputs are written with 16-byte stores.
//
// HBM traffic per element: 8 B read (f64) + 2 B (u16 2*rank) + 2 B (two int8 digits) written.
// -------------------------------------------------------------------------------------------
constexpr int kRankThreads = 256;
constexpr int kMaxPeers = 8;

// Where a row's outputs go: one destination on a single GPU; on a job shared by several GPUs every rank ranks ITS block of rows
// and stores the results into its own copy AND into every peer's copy of the packed table (peer memory mapped with CUDA IPC,
// NVLink / NVSwitch stores issued by this kernel) -- the table is never ranked twice and there is no all-gather afterwards.
struct RankOut {
    uint16_t *rank2[kMaxPeers];
    int8_t *hi[kMaxPeers], *lo[kMaxPeers];
    long long *suu[kMaxPeers];
    double *sx[kMaxPeers];
    int n_dst;
};
struct StatOut {
    double *sum[kMaxPeers], *mean[kMaxPeers], *xx[kMaxPeers];
    unsigned char *pos[kMaxPeers];
    int n_dst;
};

__global__ void __launch_bounds__(kRankThreads)
rank_pack_kernel(const double *__restrict__ raw, long long N, int n, int n_pad, int P2, long long first_row, RankOut out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *keys = reinterpret_cast<double *>(smem_raw);                    // [P2]
    uint16_t *idx = reinterpret_cast<uint16_t *>(keys + P2);                // [P2]
    uint16_t *r2 = idx + P2;                                                // [n_pad]
    __shared__ long long red[kRankThreads / 32];

    const long long row = first_row + blockIdx.x;
    const int tid = threadIdx.x;
    const double *x = raw + row * (long long)n;

    // Expression rows are zero-inflated: every sample equal to the row minimum shares the average
    // rank (m0+1)/2, so only the m = n - m0 remaining samples need sorting (a 2.4x smaller bitonic
    // network for a half-zero FANTOM5 row).  Pass 1: row minimum.  Pass 2: compact the rest.
    __shared__ double s_min[kRankThreads / 32];
    __shared__ int s_cnt[kRankThreads / 32 + 1];
    double vmin = __longlong_as_double(0x7ff0000000000000LL);
    for (int p = tid; p < n; p += kRankThreads) vmin = fmin(vmin, x[p] + 0.0);   // +0.0: -0.0 ties 0.0 as a double compare does
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    if ((tid & 31) == 0) s_min[tid >> 5] = vmin;
    __syncthreads();
    vmin = s_min[0];
    for (int w = 1; w < kRankThreads / 32; ++w) vmin = fmin(vmin, s_min[w]);
    // order-preserving compaction (ballot + prefix over warps), chunk by chunk
    int m = 0;                                            // compacted so far (uniform)
    for (int p0 = 0; p0 < n; p0 += kRankThreads) {
        const int p = p0 + tid;
        const double v = (p < n) ? (x[p] + 0.0) : vmin;
        const bool keep = (p < n) && (v != vmin);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if ((tid & 31) == 0) s_cnt[tid >> 5] = __popc(bal);
        __syncthreads();
        int base = m, total = 0;
        for (int w = 0; w < kRankThreads / 32; ++w) { const int c = s_cnt[w]; if (w < (tid >> 5)) base += c; total += c; }
        if (keep) {
            const int slot = base + __popc(bal & ((1u << (tid & 31)) - 1u));
            keys[slot] = v;
            idx[slot] = (uint16_t)p;
        } else if (p < n) {
            r2[p] = 0;                                    // marked: rank of the minimum, filled in below
        }
        m += total;
        __syncthreads();
    }
    const int m0 = n - m;                                 // samples equal to the row minimum
    P2 = 32;                                              // at least one warp-wide block (see below)
    while (P2 < m) P2 <<= 1;
    for (int p = m + tid; p < P2; p += kRankThreads) {
        keys[p] = __longlong_as_double(0x7ff0000000000000LL);
        idx[p] = (uint16_t)0xFFFF;
    }
    __syncthreads();

    // Bitonic sort of (key, column).  The kernel is bound by shared-memory wavefronts (LSU 98 % of peak at the FANTOM5 shape,
    // profiles/r01_gemm_fullshape.md), so every compare-exchange whose partners are less than 32 apart is done on registers:
    // a warp owns 32 consecutive elements (one per lane, conflict-free loads), exchanges with shuffles for the strides 16..1
    // of a merge phase -- and for ALL strides of the phases k <= 32 -- and writes the block back once.  21 shared-memory
    // passes instead of 55 for 1024 elements.
    const int lane = tid & 31, wrp = tid >> 5;
    auto exchange = [&](double &key, int &col, int i, int k, int j) {      // compare-exchange with lane ^ j, element index i
        const double pk = __shfl_xor_sync(0xffffffffu, key, j);
        const int pc = __shfl_xor_sync(0xffffffffu, col, j);
        const bool gt = (key > pk) || (key == pk && col > pc);
        const bool lower = (i & j) == 0, up = (i & k) == 0;
        if ((lower == up) ? gt : !gt) { key = pk; col = pc; }
    };
    for (int b0 = wrp * 32; b0 < P2; b0 += kRankThreads) {                  // phases k = 2 .. 32 entirely inside the warp
        const int i = b0 + lane;
        double key = keys[i];
        int col = idx[i];
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) exchange(key, col, i, k, j);
        keys[i] = key;
        idx[i] = (uint16_t)col;
    }
    __syncthreads();
    for (int k = 64; k <= P2; k <<= 1) {
        for (int j = k >> 1; j >= 32; j >>= 1) {                            // partners in different warp blocks: shared memory
            for (int t = tid; t < (P2 >> 1); t += kRankThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const double ka = keys[i], kb = keys[l];
                const uint16_t ia = idx[i], ib = idx[l];
                const bool gt = (ka > kb) || (ka == kb && ia > ib);
                const bool up = ((i & k) == 0);
                if (gt == up) {
                    keys[i] = kb; keys[l] = ka;
                    idx[i] = ib;  idx[l] = ia;
                }
            }
            __syncthreads();
        }
        for (int b0 = wrp * 32; b0 < P2; b0 += kRankThreads) {              // strides 16 .. 1 of this phase on registers
            const int i = b0 + lane;
            double key = keys[i];
            int col = idx[i];
#pragma unroll
            for (int j = 16; j > 0; j >>= 1) exchange(key, col, i, k, j);
            keys[i] = key;
            idx[i] = (uint16_t)col;
        }
        __syncthreads();
    }

    // tie run of sorted position p = [lower_bound(key), upper_bound(key)) within the m sorted keys;
    // the m0 minimum samples occupy ranks 1..m0
    for (int p = tid; p < m; p += kRankThreads) {
        const double key = keys[p];
        int lo = 0, hi = p;                       // first position with keys[pos] >= key
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (keys[mid] < key) lo = mid + 1; else hi = mid;
        }
        const int first = lo;
        lo = p + 1; hi = m;                       // first position with keys[pos] > key
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (keys[mid] <= key) lo = mid + 1; else hi = mid;
        }
        const int last = lo - 1;
        r2[idx[p]] = (uint16_t)(2 * m0 + first + last + 2);   // 2 * average rank (1-based)
    }
    __syncthreads();
    for (int c = tid; c < n; c += kRankThreads)
        if (r2[c] == 0) r2[c] = (uint16_t)(m0 + 1);           // 2 * (1 + m0) / 2
    __syncthreads();

    // outputs: u16 2*rank, digits of u = 2*rank-(n+1) = 128*hi + lo, lo in [-64, 63]
    long long acc = 0;
    for (int c = tid; c < n; c += kRankThreads) {
        const uint16_t v = r2[c];
        for (int d = 0; d < out.n_dst; ++d) out.rank2[d][row * (long long)n + c] = v;
        const long long u = (long long)v - (n + 1);
        acc += u * u;
    }
    const long long prow = row * (long long)n_pad;
    for (int v = tid; v < (n_pad >> 4); v += kRankThreads) {
        alignas(16) int8_t h[16], l[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int c = (v << 4) + e;
            const int u = (c < n) ? ((int)r2[c] - (n + 1)) : 0;
            const int hh = (u + 64) >> 7;            // floor((u+64)/128): arithmetic shift
            h[e] = (int8_t)hh;
            l[e] = (int8_t)(u - (hh << 7));
        }
        for (int d = 0; d < out.n_dst; ++d) {
            *reinterpret_cast<int4 *>(out.hi[d] + prow + (v << 4)) = *reinterpret_cast<const int4 *>(h);
            *reinterpret_cast<int4 *>(out.lo[d] + prow + (v << 4)) = *reinterpret_cast<const int4 *>(l);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        long long s = 0;
        for (int w = 0; w < kRankThreads / 32; ++w) s += red[w];
        // reference: xx = sum((rank-mean)^2) = s/4 exactly; sqrt as coexpression_v2.c:113
        const double sxv = sqrt(0.25 * (double)s);
        for (int d = 0; d < out.n_dst; ++d) { out.suu[d][row] = s; out.sx[d][row] = sxv; }
    }
}

// -------------------------------------------------------------------------------------------
// K1b: in-order row statistics.  One thread owns one row and accumulates strictly left to
// right (the reference's order); a 32x32 tile is staged through shared memory so that global
// loads stay coalesced.  32 rows per CTA, 256 threads load, one warp accumulates.
// -------------------------------------------------------------------------------------------
constexpr int kStatRows = 32;

__global__ void __launch_bounds__(256)
rowstats_kernel(const double *__restrict__ data, long long N, int n, long long first_row, StatOut out) {
    __shared__ double tile[kStatRows][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long row0 = first_row + (long long)blockIdx.x * kStatRows;
    double s = 0.0, mean = 0.0, xx = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        for (int c0 = 0; c0 < n; c0 += 32) {
#pragma unroll
            for (int r = ty; r < kStatRows; r += 8) {
                const long long row = row0 + r;
                const int c = c0 + tx;
                tile[r][tx] = (row < N && c < n) ? data[row * (long long)n + c] : 0.0;
            }
            __syncthreads();
            if (ty == 0) {
                const int lim = min(32, n - c0);
                if (pass == 0) {
                    for (int j = 0; j < lim; ++j) s = __dadd_rn(s, tile[tx][j]);
                } else {
                    for (int j = 0; j < lim; ++j) {
                        const double d = __dsub_rn(tile[tx][j], mean);
                        xx = __dadd_rn(xx, __dmul_rn(d, d));
                    }
                }
            }
            __syncthreads();
        }
        if (pass == 0) mean = s / (double)n;
    }
    const long long row = row0 + tx;
    if (ty == 0 && row < N) {
        for (int d = 0; d < out.n_dst; ++d) {
            out.sum[d][row] = s;
            out.mean[d][row] = mean;
            out.xx[d][row] = xx;
            if (out.pos[d]) out.pos[d][row] = (s > 0.0) ? 1 : 0;
        }
    }
}

}  // namespace coex

namespace coex {

// -------------------------------------------------------------------------------------------
// K2 (Pearson): centre + normalise + pack.  z = (x - mean) / sqrt(sum (x-mean)^2) lies in
// [-1, 1]; q = round(z * kZScale) is split into D balanced base-128 int8 digits
// (q = sum_d digit_d * 128^d, low digits in [-64, 63]) written as D K-major planes for the
// exact-integer tensor-core GEMM.  r is then (sum_k q_a q_b) / (s_a s_b) with NO accumulation
// error; the only error is the rounding of z:  |r_q - r| <= (|z_a|_1 + |z_b|_1) / (2 s) + n / (4 s^2).
// One warp per row, 16-byte loads and stores.  Rows the reference maps to NaN (sum == 0 or zero
// variance, coexpression_v2.c:89-92,113) are packed as zeros and flagged in `ok`.
// -------------------------------------------------------------------------------------------
template <int D>
struct ZScale;
template <> struct ZScale<3> { static constexpr double s = 2040000.0; };          // 2.3 % below 127*128^2 + 63*128 + 63
template <> struct ZScale<4> { static constexpr double s = 261000000.0; };        // 2.3 % below 127*128^3 + ...

template <int D>
__global__ void __launch_bounds__(256)
pearson_pack_kernel(const double *__restrict__ raw, long long N, int n, int n_pad,
                    const double *__restrict__ rsum, const double *__restrict__ rmean,
                    const double *__restrict__ rxx, int8_t *__restrict__ planes, long long plane_stride,
                    unsigned char *__restrict__ ok_out, float *__restrict__ l1_out, double *__restrict__ inv_s_out,
                    float *__restrict__ valid_out, int *__restrict__ n_invalid) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    const double xx = rxx[row], mean = rmean[row];
    const bool ok = (rsum[row] != 0.0) && (xx > 0.0) && (xx == xx);
    // Per-row scale: expression rows are zero-inflated, so one value (the z-score of an exact 0)
    // is shared by a large fraction of the samples and its rounding error would add up coherently
    // over those samples.  The scale is nudged (by < 1.6 %) so that this value quantises EXACTLY;
    // the rounding errors of the remaining samples are independent.
    double srow = ZScale<D>::s;
    if (ok) {
        const double z0 = (0.0 - mean) / sqrt(xx);
        const double k0 = rint(z0 * srow);
        if (fabs(k0) >= 32.0) srow = k0 / z0;
    }
    const double inv = ok ? srow / sqrt(xx) : 0.0;
    const double *x = raw + row * (long long)n;
    double l1 = 0.0;
    for (int v = lane; v < (n_pad >> 3); v += 32) {
        alignas(8) int8_t dg[D][8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = (v << 3) + e;
            long long q = 0;
            if (ok && c < n) {
                const double zs = (x[c] - mean) * inv;
                q = llrint(zs);
                l1 += fabs(zs);
            }
#pragma unroll
            for (int d = 0; d < D - 1; ++d) {
                const long long dd = ((q + 64) & 127) - 64;
                dg[d][e] = (int8_t)dd;
                q = (q - dd) >> 7;
            }
            dg[D - 1][e] = (int8_t)q;
        }
#pragma unroll
        for (int d = 0; d < D; ++d)
            *reinterpret_cast<long long *>(planes + d * plane_stride + row * (long long)n_pad + (v << 3)) =
                *reinterpret_cast<const long long *>(dg[d]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    if (lane == 0) {
        ok_out[row] = ok ? 1 : 0;
        // the count kernel's column mask (like icol of the Spearman path) AND the row's share of the error bound of the
        // tensor-core r: |sum q_a q_b / (s_a s_b) - r| <= 0.5 |z_a|_1 / s_b + 0.5 |z_b|_1 / s_a + n / (4 s_a s_b), every s >= smin
        // = 0.98 ZScale.  0.51 |z|_1 / smin + n / (8 smin^2), rounded up, > 0 for every defined row.
        const double smin = 0.98 * ZScale<D>::s;
        valid_out[row] = ok ? (float)(1.02 * (0.51 * (l1 / srow) / smin + (double)n / (8.0 * smin * smin)) + 1e-12) : 0.0f;
        if (!ok) atomicAdd(n_invalid, 1);                   // rows whose Pearson null value is NaN -> -99 for every query
        l1_out[row] = (float)(l1 / srow);
        inv_s_out[row] = 1.0 / srow;
    }
}

}  // namespace coex
