// allpairs.cuh -- config C3: every unordered row pair of the resident table, reduced on the
// device to a histogram of r (the N x N matrix itself, 160 GB at the full FANTOM5 shape, is
// never stored).  Row-block strips reuse the correlation strip kernels of the network path:
//   which == 1 (ranks)  exact integer tcgen05 GEMM strip  -> r as in coexpression_v2.c:113
//   which == 0 (raw)    in-order double Pearson strip (bit-identical to the reference)
#pragma once
#include "common.cuh"
#include "count.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

namespace coex {

__device__ __forceinline__ int r_to_bin(double r, int nbins) {
    int b = (int)((r + 1.0) * 0.5 * (double)nbins);
    return b < 0 ? 0 : (b >= nbins ? nbins - 1 : b);
}

// rows [row0, row0+nq) of the strip against columns c > row; grid-stride over strip elements
template <bool F64>
__global__ void __launch_bounds__(256)
hist_strip_kernel(const int *__restrict__ strip, const double *__restrict__ strip_f64, long long ld, long long row0,
                  long long nq, long long N, const double *__restrict__ sx, int nbins,
                  unsigned long long *__restrict__ hist, unsigned long long *__restrict__ nan_count) {
    const long long q = blockIdx.y;
    if (q >= nq) return;
    const long long row = row0 + q;
    const double sxu = F64 ? 0.0 : sx[row];
    unsigned long long nans = 0;
    for (long long c = row + 1 + (long long)blockIdx.x * blockDim.x + threadIdx.x; c < N;
         c += (long long)gridDim.x * blockDim.x) {
        double r;
        if (F64) r = strip_f64[q * ld + c];
        else r = __ddiv_rn(0.25 * (double)strip[q * ld + c], __dmul_rn(sxu, sx[c]));
        if (r != r) { ++nans; continue; }
        atomicAdd(&hist[r_to_bin(r, nbins)], 1ULL);
    }
    if (nans) atomicAdd(nan_count, nans);
}

__global__ void iota_rows_kernel(int *rows, long long row0, long long nq) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nq) rows[t] = (int)(row0 + t);
}

static int allpairs_histogram_impl(coex_ctx *c, int which, long long row0, long long row1, int nbins,
                                   unsigned long long *hist, unsigned long long *nan_count) {
    if (row0 < 0 || row1 > c->N || row0 > row1 || nbins <= 0) COEX_FAIL("coex_allpairs_histogram: bad arguments");
    if (which == 1 && c->n > 1860) COEX_FAIL("Spearman tensor path supports up to 1860 samples");
    stage_reset(c);
    cudaStream_t st = c->stream;
    DevBuf<unsigned long long> dh;
    COEX_TRY(dh.reserve((size_t)nbins + 1));
    COEX_CUDA(cudaMemsetAsync(dh.p, 0, ((size_t)nbins + 1) * sizeof(unsigned long long), st));
    const long long ld = which == 1 ? c->N_pad : c->N;
    long long Hs = (c->strip_mb << 20) / (ld * (which == 1 ? 4 : 8));
    Hs = std::max<long long>(256, Hs / 256 * 256);
    Hs = std::min<long long>(Hs, round_up(std::max<long long>(row1 - row0, 1), 256));
    COEX_TRY(c->d_hits.reserve(Hs));
    if (which == 1) {
        COEX_TRY(c->strip.reserve((size_t)Hs * ld));
        COEX_TRY(c->a_hi.reserve((size_t)Hs * c->n_pad)); COEX_TRY(c->a_lo.reserve((size_t)Hs * c->n_pad));
    } else {
        COEX_TRY(c->strip_f64.reserve((size_t)Hs * ld));
    }
    for (long long r0 = row0; r0 < row1; r0 += Hs) {
        const long long nq = std::min(Hs, row1 - r0);
        iota_rows_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(c->d_hits.p, r0, nq);
        COEX_CUDA(cudaGetLastError());
        c->launches += 1;
        if (which == 1) {
            const long long nq_pad = round_up(nq, 256);
            COEX_TRY(stage_begin(c, ST_GATHER));
            const long long total = nq_pad * (c->n_pad >> 4);
            gather_rows_kernel<<<(int)std::min<long long>((total + 255) / 256, (long long)c->sm_count * 16), 256, 0, st>>>(
                c->u_hi.p, c->u_lo.p, c->n_pad, c->d_hits.p, nq, nq_pad, c->a_hi.p, c->a_lo.p);
            COEX_CUDA(cudaGetLastError());
            COEX_TRY(stage_end(c, 1));
            COEX_TRY(stage_begin(c, ST_GEMM));
            if (c->gemm_mode == 0) {
                COEX_TRY(tc_gemm_i8(c, nq_pad));
            } else {
                dim3 grid((unsigned)(c->N_pad / kSimtTile), (unsigned)(nq_pad / kSimtTile));
                gemm_i8_simt_kernel<<<grid, 256, 0, st>>>(c->a_hi.p, c->a_lo.p, c->u_hi.p, c->u_lo.p, c->n_pad, c->strip.p,
                                                          c->N_pad);
                COEX_CUDA(cudaGetLastError());
            }
            COEX_TRY(stage_end(c, 1));
        } else {
            COEX_TRY(stage_begin(c, ST_GEMM));
            dim3 g((unsigned)((c->N + 63) / 64), (unsigned)((nq + 15) / 16));
            pearson_strip_kernel<<<g, 256, 0, st>>>(c->raw, c->n, c->N, c->d_hits.p, nq, c->raw_sum.p, c->raw_mean.p,
                                                    c->raw_xx.p, c->strip_f64.p, ld);
            COEX_CUDA(cudaGetLastError());
            COEX_TRY(stage_end(c, 1));
        }
        COEX_TRY(stage_begin(c, ST_COUNT));
        dim3 hg((unsigned)std::min<long long>((c->N + 255) / 256, 64), (unsigned)nq);
        if (which == 1)
            hist_strip_kernel<false><<<hg, 256, 0, st>>>(c->strip.p, nullptr, ld, r0, nq, c->N, c->sx.p, nbins, dh.p, dh.p + nbins);
        else
            hist_strip_kernel<true><<<hg, 256, 0, st>>>(nullptr, c->strip_f64.p, ld, r0, nq, c->N, c->sx.p, nbins, dh.p, dh.p + nbins);
        COEX_CUDA(cudaGetLastError());
        COEX_TRY(stage_end(c, 1));
    }
    COEX_CUDA(cudaMemcpyAsync(hist, dh.p, (size_t)nbins * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    COEX_CUDA(cudaMemcpyAsync(nan_count, dh.p + nbins, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    COEX_CUDA(cudaStreamSynchronize(st));
    dh.release();
    if (which == 1 && c->gemm_mode == 0) COEX_TRY(tc_check_error(c));
    return 0;
}

}  // namespace coex
