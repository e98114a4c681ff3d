// collate.cuh -- row A14: the consumer of the gathered density vectors (reference 2-collate-results.py:67-76,142,161-163).
//   perm_indm .......... every permuted density pooled (2-collate-results.py:68-76)
//   empiricalp ......... p = 1 - bisect_left(sorted(perm_indm), x) / len   (coexfunctions.py:379-383)
//   fdr_correction ..... Benjamini-Hochberg ('indep') and Benjamini-Yekutieli ('negcorr') adjusted p (coexfunctions.py:333-358)
// The pool is never sorted: bisect_left(sorted(pool), x) is the number of pool values strictly below x, so the pool is streamed
// once against the (few, sorted) real densities -- the count stage's algorithm in miniature -- and stays on the device it was
// gathered to.  Integer counts, then the reference's double operations in the reference's order: results are bit-identical.
#pragma once
#include "common.cuh"

namespace coex {

constexpr int kCollateMax = 8192;          // real groups (<= hit regions of a permutation)

// grid-stride over the pool; shared memory: sorted real densities [G2] (padded with +inf) + histogram [G + 1]
__global__ void __launch_bounds__(1024)
collate_count_kernel(const double *__restrict__ pool, long long n_pool, const double *__restrict__ real, int G, int G2,
                     unsigned long long *__restrict__ hist /* [G + 1], zeroed */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *xs = reinterpret_cast<double *>(smem_raw);
    unsigned int *h = reinterpret_cast<unsigned int *>(xs + G2);
    for (int i = threadIdx.x; i < G2; i += blockDim.x) xs[i] = (i < G) ? real[i] : __longlong_as_double(0x7ff0000000000000LL);
    for (int i = threadIdx.x; i <= G; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (int k = 2; k <= G2; k <<= 1)                       // bitonic sort, ascending
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (G2 >> 1); t += blockDim.x) {
                const int x = ((t & ~(j - 1)) << 1) | (t & (j - 1)), y = x | j;
                const double a = xs[x], b = xs[y];
                if ((a > b) == ((x & k) == 0)) { xs[x] = b; xs[y] = a; }
            }
            __syncthreads();
        }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pool; i += (long long)gridDim.x * blockDim.x) {
        const double v = pool[i];
        int lo = 0, hi = G;                                 // number of real densities <= v
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (xs[mid] <= v) lo = mid + 1; else hi = mid; }
        atomicAdd(&h[lo], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= G; i += blockDim.x) if (h[i]) atomicAdd(&hist[i], (unsigned long long)h[i]);
}

// one CTA: prefix of the histogram -> p of every real density; then the two FDR corrections on the ascending p
__global__ void __launch_bounds__(1024)
collate_finish_kernel(const double *__restrict__ real, int G, int G2, const unsigned long long *__restrict__ hist, long long n_pool,
                      double *__restrict__ p_out, double *__restrict__ bh_out, double *__restrict__ by_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *xs = reinterpret_cast<double *>(smem_raw);      // sorted real densities, then reused for ranked / ecdf
    double *pa = xs + G2;                                   // p ascending
    for (int i = threadIdx.x; i < G2; i += blockDim.x) xs[i] = (i < G) ? real[i] : __longlong_as_double(0x7ff0000000000000LL);
    __syncthreads();
    for (int k = 2; k <= G2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (G2 >> 1); t += blockDim.x) {
                const int x = ((t & ~(j - 1)) << 1) | (t & (j - 1)), y = x | j;
                const double a = xs[x], b = xs[y];
                if ((a > b) == ((x & k) == 0)) { xs[x] = b; xs[y] = a; }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0) {
        // pool values strictly below the i-th smallest real density = prefix of hist over bins 0 .. i (bin b holds the pool values
        // with exactly b real densities <= them; between equal densities the bin is empty, so they share their count).
        // p ascending = densities descending.
        unsigned long long below = 0;
        for (int i = 0; i < G; ++i) {
            below += hist[i];
            pa[G - 1 - i] = (n_pool > 0) ? 1.0 - (double)below / (double)n_pool : 1.0;          // coexfunctions.py:379-383
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < G; i += blockDim.x) p_out[i] = pa[i];
    // fdr_correction (coexfunctions.py:333-358): ecdf = k / m [/ sum 1/k], corrected = running minimum from the right of p / ecdf, <= 1
    if (threadIdx.x < 2) {
        double *out = threadIdx.x == 0 ? bh_out : by_out;
        double harmonic = 1.0;
        if (threadIdx.x == 1) { harmonic = 0.0; for (int k = 1; k <= G; ++k) harmonic += 1.0 / (double)k; }
        double run = __longlong_as_double(0x7ff0000000000000LL);
        for (int k = G; k >= 1; --k) {
            double ecdf = (double)k / (double)G;
            if (threadIdx.x == 1) ecdf = ecdf / harmonic;
            run = fmin(run, pa[k - 1] / ecdf);
            out[k - 1] = fmin(run, 1.0);
        }
    }
}

}  // namespace coex
