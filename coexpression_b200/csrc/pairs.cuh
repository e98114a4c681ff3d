// pairs.cuh -- K3b: pair-list correlation (reference coexpression_v2.c:51-115).
//   pairs_exact_kernel   arbitrary double rows; xy is accumulated strictly left to right with
//                        separately rounded subtract / multiply / add, so the result is
//                        bit-identical to the reference's `xysum/(sqrt(xxsum)*sqrt(yysum))`.
//                        Per-row sum / mean / xx come from rowstats_kernel (they are functions
//                        of one row only; the reference recomputes them per pair).
//   pairs_rank_kernel    rows that are rank vectors: exact integer dot products of
//                        u = 2*rank-(n+1); r = (S/4) / (sqrt(Saa/4)*sqrt(Sbb/4)) is bit-identical
//                        to the reference on ranked input because every partial sum the
//                        reference forms on half-integers is exact in double.
#pragma once
#include "common.cuh"

namespace coex {

constexpr int kPairsPerBlock = 128;
constexpr int kPairChunk = 16;

__global__ void __launch_bounds__(kPairsPerBlock)
pairs_exact_kernel(const double *__restrict__ data, int n, const int *__restrict__ idxa,
                   const int *__restrict__ idxb, long long npairs,
                   const double *__restrict__ row_sum, const double *__restrict__ row_mean,
                   const double *__restrict__ row_xx, double *__restrict__ out) {
    __shared__ double ta[kPairsPerBlock][kPairChunk + 1];
    __shared__ double tb[kPairsPerBlock][kPairChunk + 1];
    __shared__ int sa[kPairsPerBlock], sb[kPairsPerBlock];
    const int tid = threadIdx.x;
    const long long p0 = (long long)blockIdx.x * kPairsPerBlock;
    const long long mine = p0 + tid;
    int ia = 0, ib = 0;
    if (mine < npairs) { ia = idxa[mine]; ib = idxb[mine]; }
    sa[tid] = ia; sb[tid] = ib;
    const double ma = row_mean[ia], mb = row_mean[ib];
    __syncthreads();
    double xy = 0.0;
    const int col = tid & (kPairChunk - 1), sub = tid / kPairChunk;   // 8 rows per sweep
    for (int c0 = 0; c0 < n; c0 += kPairChunk) {
        const int c = c0 + col;
#pragma unroll 4
        for (int r = sub; r < kPairsPerBlock; r += kPairsPerBlock / kPairChunk) {
            const bool ok = (c < n) && (p0 + r < npairs);
            ta[r][col] = ok ? data[(long long)sa[r] * n + c] : 0.0;
            tb[r][col] = ok ? data[(long long)sb[r] * n + c] : 0.0;
        }
        __syncthreads();
        const int lim = min(kPairChunk, n - c0);
        for (int j = 0; j < lim; ++j) {
            const double da = __dsub_rn(ta[tid][j], ma);
            const double db = __dsub_rn(tb[tid][j], mb);
            xy = __dadd_rn(xy, __dmul_rn(da, db));
        }
        __syncthreads();
    }
    if (mine < npairs) {
        double r;
        if (row_sum[ia] == 0.0 || row_sum[ib] == 0.0) {
            r = __longlong_as_double(0x7ff8000000000000LL);          // coexpression_v2.c:89-92
        } else {
            r = __ddiv_rn(xy, __dmul_rn(sqrt(row_xx[ia]), sqrt(row_xx[ib])));
        }
        out[mine] = r;
    }
}

// One warp per pair, on the PACKED rank rows (two int8 digit planes, u = 2*rank-(n+1) = 128*hi + lo, rows padded to a multiple
// of 128 bytes and zero-filled): 16-byte coalesced loads (a rank2 row of an odd number of samples is only 2-byte aligned) and
// four dp4a per four samples instead of scalar u16 loads; S = 16384 HH + 128 (HL + LH) + LL is the same exact integer.
__global__ void __launch_bounds__(256)
pairs_rank_kernel(const int8_t *__restrict__ hi, const int8_t *__restrict__ lo, int n_pad, const int *__restrict__ idxa,
                  const int *__restrict__ idxb, long long npairs, const double *__restrict__ sx,
                  double *__restrict__ out) {
    const long long pair = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pair >= npairs) return;
    const int ia = idxa[pair], ib = idxb[pair];
    const int4 *ah = reinterpret_cast<const int4 *>(hi + (long long)ia * n_pad), *al = reinterpret_cast<const int4 *>(lo + (long long)ia * n_pad);
    const int4 *bh = reinterpret_cast<const int4 *>(hi + (long long)ib * n_pad), *bl = reinterpret_cast<const int4 *>(lo + (long long)ib * n_pad);
    int hh = 0, hl = 0, ll = 0;                            // per lane: at most 60 samples x 64 x 64 (no overflow)
    for (int v = lane; v < (n_pad >> 4); v += 32) {
        const int4 x = ah[v], y = al[v], p = bh[v], q = bl[v];
        hh = __dp4a(x.x, p.x, hh); hh = __dp4a(x.y, p.y, hh); hh = __dp4a(x.z, p.z, hh); hh = __dp4a(x.w, p.w, hh);
        hl = __dp4a(x.x, q.x, hl); hl = __dp4a(x.y, q.y, hl); hl = __dp4a(x.z, q.z, hl); hl = __dp4a(x.w, q.w, hl);
        hl = __dp4a(y.x, p.x, hl); hl = __dp4a(y.y, p.y, hl); hl = __dp4a(y.z, p.z, hl); hl = __dp4a(y.w, p.w, hl);
        ll = __dp4a(y.x, q.x, ll); ll = __dp4a(y.y, q.y, ll); ll = __dp4a(y.z, q.z, ll); ll = __dp4a(y.w, q.w, ll);
    }
    long long acc = 16384LL * hh + 128LL * hl + ll;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        // 0/0 -> NaN for constant rows exactly as the reference's unguarded divide
        out[pair] = __ddiv_rn(0.25 * (double)acc, __dmul_rn(sx[ia], sx[ib]));
    }
}

}  // namespace coex
