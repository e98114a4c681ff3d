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
#include "prepare.cuh"

namespace coex {

__device__ __forceinline__ int r_to_bin(double r, int nbins) {
    int b = (int)((r + 1.0) * 0.5 * (double)nbins);
    return b < 0 ? 0 : (b >= nbins ? nbins - 1 : b);
}

// rows [row0, row0 + nq) of the strip against columns c > row.  The histogram is bound by same-address serialisation of the L2
// atomics (measured: 162 ms for 2e10 pairs into 2^20 bins, 307 ms into 2^16, 3.2 s into 2^12): correlations of unrelated rows
// pile up around r = 0.  So every CTA keeps the WINDOW of bins around r = 0 in shared memory (kHistWindow 32-bit counters, 192 KB,
// one persistent CTA per SM), sends only the values outside it to the global histogram and flushes its non-zero counters once.
constexpr int kHistWindow = 48 * 1024;
constexpr int kHistThreads = 1024;

template <bool F64>
__global__ void __launch_bounds__(kHistThreads, 1)
hist_strip_kernel(const int *__restrict__ strip, const double *__restrict__ strip_f64, long long ld, long long row0,
                  long long nq, long long N, const double *__restrict__ sx, const unsigned char *__restrict__ ok,
                  int nbins, unsigned long long *__restrict__ hist, unsigned long long *__restrict__ nan_count) {
    extern __shared__ unsigned int win[];                   // [W]
    const int W = nbins < kHistWindow ? nbins : kHistWindow;
    const int w0 = (nbins - W) / 2;                         // first bin of the window (centred on r = 0)
    for (int i = threadIdx.x; i < W; i += kHistThreads) win[i] = 0u;
    __syncthreads();
    unsigned long long nans = 0;
    for (long long q = blockIdx.x; q < nq; q += gridDim.x) {
        const long long row = row0 + q;
        const double sxu = F64 ? 0.0 : sx[row];
        for (long long c = row + 1 + threadIdx.x; c < N; c += kHistThreads) {
            double r;
            if (F64) {
                r = strip_f64[q * ld + c];
                if (ok && !(ok[row] && ok[c])) r = __longlong_as_double(0x7ff8000000000000LL);
            } else {
                r = __ddiv_rn(0.25 * (double)strip[q * ld + c], __dmul_rn(sxu, sx[c]));
            }
            if (r != r) { ++nans; continue; }
            const int b = r_to_bin(r, nbins);
            const unsigned int wb = (unsigned int)(b - w0);
            if (wb < (unsigned int)W) atomicAdd(&win[wb], 1u);      // (a CTA sees fewer than 2^32 pairs per launch: checked by the caller)
            else atomicAdd(&hist[b], 1ULL);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < W; i += kHistThreads) {
        const unsigned int v = win[i];
        if (v) atomicAdd(&hist[w0 + i], (unsigned long long)v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nans += __shfl_xor_sync(0xffffffffu, nans, o);
    if ((threadIdx.x & 31) == 0 && nans) atomicAdd(nan_count, nans);
}

__global__ void iota_rows_kernel(int *rows, long long row0, long long nq) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nq) rows[t] = (int)(row0 + t);
}

// lazily pack the Pearson digit planes (after coex_prepare)
static int pearson_pack(coex_ctx *c) {
    const int D = c->pearson_digits;
    if (c->z_digits == D) return 0;
    const size_t plane = (size_t)c->N_pad * c->n_pad;
    COEX_TRY(c->z_planes.reserve(plane * D)); COEX_TRY(c->z_ok.reserve(c->N_pad)); COEX_TRY(c->z_l1.reserve(c->N_pad));
    COEX_TRY(c->z_inv_s.reserve(c->N_pad)); COEX_TRY(c->z_valid.reserve(c->N_pad)); COEX_TRY(c->z_ninv.reserve(1));
    COEX_CUDA(cudaMemsetAsync(c->z_valid.p, 0, c->N_pad * sizeof(float), c->stream));
    COEX_CUDA(cudaMemsetAsync(c->z_ninv.p, 0, sizeof(int), c->stream));
    COEX_CUDA(cudaMemsetAsync(c->z_inv_s.p, 0, c->N_pad * sizeof(double), c->stream));
    COEX_CUDA(cudaMemsetAsync(c->z_planes.p, 0, plane * D, c->stream));
    COEX_CUDA(cudaMemsetAsync(c->z_ok.p, 0, c->N_pad, c->stream));
    COEX_CUDA(cudaMemsetAsync(c->z_l1.p, 0, c->N_pad * sizeof(float), c->stream));
    const unsigned blocks = (unsigned)((c->N * 32 + 255) / 256);
    if (D == 3)
        pearson_pack_kernel<3><<<blocks, 256, 0, c->stream>>>(c->raw, c->N, c->n, c->n_pad, c->raw_sum.p, c->raw_mean.p, c->raw_xx.p,
                                                            c->z_planes.p, (long long)plane, c->z_ok.p, c->z_l1.p, c->z_inv_s.p, c->z_valid.p, c->z_ninv.p);
    else
        pearson_pack_kernel<4><<<blocks, 256, 0, c->stream>>>(c->raw, c->N, c->n, c->n_pad, c->raw_sum.p, c->raw_mean.p, c->raw_xx.p,
                                                            c->z_planes.p, (long long)plane, c->z_ok.p, c->z_l1.p, c->z_inv_s.p, c->z_valid.p, c->z_ninv.p);
    COEX_KCHECK(c->stream);
    c->launches += 1;
    c->z_digits = D;
    return 0;
}

// Pearson r strip [m_rows, N_pad] (double) for table rows [row0, row0 + m_rows) on the tensor cores
static int pearson_tensor_strip(coex_ctx *c, long long row0, long long m_rows, int upper_only) {
    COEX_TRY(pearson_pack(c));
    const size_t plane = (size_t)c->N_pad * c->n_pad;
    if (c->z_digits == 3)
        return tc_gemm_digits<3>(c, c->z_planes.p, (long long)plane, row0, m_rows, c->strip_f64.p, c->N_pad, c->z_inv_s.p, upper_only);
    return tc_gemm_digits<4>(c, c->z_planes.p, (long long)plane, row0, m_rows, c->strip_f64.p, c->N_pad, c->z_inv_s.p, upper_only);
}

// reset != 0: the resident histogram is cleared first.  hist / nan_count == nullptr: nothing is read back (the caller
// accumulates several row blocks and fetches the result once, allpairs_read_impl).
static int allpairs_histogram_impl(coex_ctx *c, int which, long long row0, long long row1, int nbins,
                                   unsigned long long *hist, unsigned long long *nan_count, int reset) {
    if (row0 < 0 || row1 > c->N || row0 > row1 || nbins <= 0) COEX_FAIL("coex_allpairs_histogram: bad arguments");
    if (which == 1 && c->n > 1860) COEX_FAIL("Spearman tensor path supports up to 1860 samples");
    stage_reset(c);
    cudaStream_t st = c->stream;
    if (!reset && c->ap_nbins != nbins) COEX_FAIL("coex_allpairs_accumulate: nbins differs from the resident histogram (reset first)");
    DevBuf<unsigned long long> &dh = c->ap_hist;
    COEX_TRY(dh.reserve((size_t)nbins + 1));
    if (reset) COEX_CUDA(cudaMemsetAsync(dh.p, 0, ((size_t)nbins + 1) * sizeof(unsigned long long), st));
    c->ap_nbins = nbins;
    const bool ptensor = (which == 0 && c->pearson_mode == 0);
    const long long ld = (which == 1 || ptensor) ? c->N_pad : c->N;
    long long Hs = (c->strip_mb << 20) / (ld * (which == 1 ? 4 : 8));
    Hs = std::max<long long>(256, Hs / 256 * 256);
    Hs = std::min<long long>(Hs, round_up(std::max<long long>(row1 - row0, 1), 256));
    if (ptensor && row0 % 128) COEX_FAIL("coex_allpairs_histogram: row0 must be a multiple of 128 on the tensor path");
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
                c->gemm_col0 = r0;                          // only columns > row are histogrammed
                const int gs = tc_gemm_i8(c, nq_pad);
                c->gemm_col0 = 0;
                COEX_TRY(gs);
            } else {
                dim3 grid((unsigned)(c->N_pad / kSimtTile), (unsigned)(nq_pad / kSimtTile));
                gemm_i8_simt_kernel<<<grid, 256, 0, st>>>(c->a_hi.p, c->a_lo.p, c->u_hi.p, c->u_lo.p, c->n_pad, c->strip.p,
                                                          c->N_pad);
                COEX_CUDA(cudaGetLastError());
            }
            COEX_TRY(stage_end(c, 1));
        } else if (ptensor) {
            COEX_TRY(stage_begin(c, ST_GEMM));
            COEX_TRY(pearson_tensor_strip(c, r0, std::min(round_up(nq, 128), c->N_pad - r0), 1));
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
        if ((double)nq * (double)c->N >= 4.0e9) COEX_FAIL("coex_allpairs_histogram: strip of more than 2^32 pairs (lower the workspace)");
        const unsigned hg = (unsigned)std::min<long long>(nq, c->sm_count);
        const size_t hsm = (size_t)std::min(nbins, kHistWindow) * sizeof(unsigned int);
        if (which == 1) {
            COEX_CUDA(cudaFuncSetAttribute(hist_strip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm));
            hist_strip_kernel<false><<<hg, kHistThreads, hsm, st>>>(c->strip.p, nullptr, ld, r0, nq, c->N, c->sx.p, nullptr, nbins, dh.p, dh.p + nbins);
        } else {
            COEX_CUDA(cudaFuncSetAttribute(hist_strip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm));
            hist_strip_kernel<true><<<hg, kHistThreads, hsm, st>>>(nullptr, c->strip_f64.p, ld, r0, nq, c->N, c->sx.p, ptensor ? c->z_ok.p : nullptr, nbins, dh.p, dh.p + nbins);
        }
        COEX_CUDA(cudaGetLastError());
        COEX_TRY(stage_end(c, 1));
    }
    if (hist) COEX_CUDA(cudaMemcpyAsync(hist, dh.p, (size_t)nbins * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    if (nan_count) COEX_CUDA(cudaMemcpyAsync(nan_count, dh.p + nbins, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    if (hist || nan_count) {
        COEX_CUDA(cudaStreamSynchronize(st));
        if ((which == 1 && c->gemm_mode == 0) || ptensor) COEX_TRY(tc_check_error(c));
    }
    return 0;
}

static int allpairs_read_impl(coex_ctx *c, unsigned long long *hist, unsigned long long *nan_count, void *device_copy) {
    if (c->ap_nbins <= 0 || !c->ap_hist.p) COEX_FAIL("coex_allpairs_read: nothing accumulated");
    cudaStream_t st = c->stream;
    const size_t bytes = (size_t)c->ap_nbins * sizeof(unsigned long long);
    if (device_copy) COEX_CUDA(cudaMemcpyAsync(device_copy, c->ap_hist.p, bytes + sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    if (hist) COEX_CUDA(cudaMemcpyAsync(hist, c->ap_hist.p, bytes, cudaMemcpyDeviceToHost, st));
    if (nan_count) COEX_CUDA(cudaMemcpyAsync(nan_count, c->ap_hist.p + c->ap_nbins, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    COEX_CUDA(cudaStreamSynchronize(st));
    return tc_check_error(c);
}

// expand a strip into r values (NaN where the reference yields NaN)
__global__ void strip_to_r_kernel(const int *__restrict__ strip, const double *__restrict__ strip_f64, long long ld,
                                  long long row0, long long nq, long long N, const double *__restrict__ sx,
                                  const unsigned char *__restrict__ ok, double *__restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * N) return;
    const long long q = t / N, col = t - q * N;
    double r;
    if (strip) r = __ddiv_rn(0.25 * (double)strip[q * ld + col], __dmul_rn(sx[row0 + q], sx[col]));
    else {
        r = strip_f64[q * ld + col];
        if (ok && !(ok[row0 + q] && ok[col])) r = __longlong_as_double(0x7ff8000000000000LL);
    }
    out[t] = r;
}

// rows [row0, row0 + nrows) x all table rows: which 1 = Spearman (exact), 0 = Pearson on the tensor
// cores (quantised digits), 2 = Pearson in-order double (bit-exact).  out: host double [nrows, N].
static int corr_block_impl(coex_ctx *c, int which, long long row0, long long nrows, double *out) {
    if (row0 < 0 || nrows <= 0 || row0 + nrows > c->N) COEX_FAIL("coex_corr_block: bad row range");
    if (which != 2 && row0 % 128) COEX_FAIL("coex_corr_block: row0 must be a multiple of 128");
    if (which == 1 && c->n > 1860) COEX_FAIL("Spearman tensor path supports up to 1860 samples");
    cudaStream_t st = c->stream;
    const long long m_rows = std::min(round_up(nrows, 256), c->N_pad - row0);
    DevBuf<double> dout;
    COEX_TRY(dout.reserve((size_t)nrows * c->N));
    COEX_TRY(c->d_hits.reserve(m_rows));
    iota_rows_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(c->d_hits.p, row0, nrows);
    const unsigned eb = (unsigned)((nrows * c->N + 255) / 256);
    if (which == 1) {
        COEX_TRY(c->strip.reserve((size_t)m_rows * c->N_pad));
        COEX_TRY(c->a_hi.reserve((size_t)m_rows * c->n_pad)); COEX_TRY(c->a_lo.reserve((size_t)m_rows * c->n_pad));
        const long long total = m_rows * (c->n_pad >> 4);
        gather_rows_kernel<<<(int)std::min<long long>((total + 255) / 256, 4096), 256, 0, st>>>(
            c->u_hi.p, c->u_lo.p, c->n_pad, c->d_hits.p, nrows, m_rows, c->a_hi.p, c->a_lo.p);
        if (c->gemm_mode == 0) COEX_TRY(tc_gemm_i8(c, m_rows));
        else {
            dim3 grid((unsigned)(c->N_pad / kSimtTile), (unsigned)(m_rows / kSimtTile));
            gemm_i8_simt_kernel<<<grid, 256, 0, st>>>(c->a_hi.p, c->a_lo.p, c->u_hi.p, c->u_lo.p, c->n_pad, c->strip.p, c->N_pad);
        }
        strip_to_r_kernel<<<eb, 256, 0, st>>>(c->strip.p, nullptr, c->N_pad, row0, nrows, c->N, c->sx.p, nullptr, dout.p);
    } else if (which == 0) {
        COEX_TRY(c->strip_f64.reserve((size_t)m_rows * c->N_pad));
        COEX_TRY(pearson_tensor_strip(c, row0, m_rows, 0));
        strip_to_r_kernel<<<eb, 256, 0, st>>>(nullptr, c->strip_f64.p, c->N_pad, row0, nrows, c->N, c->sx.p, c->z_ok.p, dout.p);
    } else {
        COEX_TRY(c->strip_f64.reserve((size_t)m_rows * c->N));
        dim3 g((unsigned)((c->N + 63) / 64), (unsigned)((nrows + 15) / 16));
        pearson_strip_kernel<<<g, 256, 0, st>>>(c->raw, c->n, c->N, c->d_hits.p, nrows, c->raw_sum.p, c->raw_mean.p,
                                                c->raw_xx.p, c->strip_f64.p, c->N);
        strip_to_r_kernel<<<eb, 256, 0, st>>>(nullptr, c->strip_f64.p, c->N, row0, nrows, c->N, c->sx.p, nullptr, dout.p);
    }
    COEX_KCHECK(st);
    c->launches += 4;
    COEX_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)nrows * c->N * sizeof(double), cudaMemcpyDeviceToHost, st));
    COEX_CUDA(cudaStreamSynchronize(st));
    dout.release();
    if (which != 2) COEX_TRY(tc_check_error(c));
    return 0;
}

}  // namespace coex
