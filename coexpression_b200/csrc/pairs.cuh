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

// one warp per pair; u16 2*rank rows
__global__ void __launch_bounds__(256)
pairs_rank_kernel(const uint16_t *__restrict__ rank2, int n, const int *__restrict__ idxa,
                  const int *__restrict__ idxb, long long npairs, const double *__restrict__ sx,
                  double *__restrict__ out) {
    const long long pair = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pair >= npairs) return;
    const int ia = idxa[pair], ib = idxb[pair];
    const uint16_t *a = rank2 + (long long)ia * n;
    const uint16_t *b = rank2 + (long long)ib * n;
    long long acc = 0;
    const int off = n + 1;
    for (int c = lane; c < n; c += 32) {
        const int ua = (int)a[c] - off, ub = (int)b[c] - off;
        acc += (long long)ua * ub;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        // 0/0 -> NaN for constant rows exactly as the reference's unguarded divide
        out[pair] = __ddiv_rn(0.25 * (double)acc, __dmul_rn(sx[ia], sx[ib]));
    }
}

}  // namespace coex
