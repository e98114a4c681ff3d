// gemm_simt.cuh -- validation / bring-up path for the exact integer correlation GEMM.
//
//   S[q, c] = sum_k u_q[k] * u_c[k]     with u = 128*hi + lo stored as two int8 planes,
//           = 16384*(hi.hi) + 128*(hi.lo + lo.hi) + (lo.lo)
//
// computed with dp4a on CUDA cores.  The production path is the tcgen05/TMA kernel in
// gemm_tc.cuh (same operands, same int32 output, bit-identical result because the arithmetic
// is exact integer); this kernel exists so that the tensor-core kernel has an independent
// on-device check and so that a bring-up failure of the tensor path is visible, not silent.
#pragma once
#include "common.cuh"

namespace coex {

// gather the packed digit rows of the query list into a contiguous A operand
__global__ void gather_rows_kernel(const int8_t *__restrict__ hi, const int8_t *__restrict__ lo,
                                   int n_pad, const int *__restrict__ rows, long long nq,
                                   long long nq_pad, int8_t *__restrict__ a_hi,
                                   int8_t *__restrict__ a_lo) {
    const int vec = n_pad >> 4;
    const long long total = nq_pad * vec;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long q = t / vec;
        const int v = (int)(t - q * vec);
        int4 h = make_int4(0, 0, 0, 0), l = make_int4(0, 0, 0, 0);
        if (q < nq) {
            const long long src = (long long)rows[q] * n_pad + ((long long)v << 4);
            h = *reinterpret_cast<const int4 *>(hi + src);
            l = *reinterpret_cast<const int4 *>(lo + src);
        }
        *reinterpret_cast<int4 *>(a_hi + q * n_pad + ((long long)v << 4)) = h;
        *reinterpret_cast<int4 *>(a_lo + q * n_pad + ((long long)v << 4)) = l;
    }
}

constexpr int kSimtTile = 64;

// A planes [M_pad, n_pad], B planes [N_pad, n_pad] (both K-major), out int32 [M_pad, ldo]
__global__ void __launch_bounds__(256)
gemm_i8_simt_kernel(const int8_t *__restrict__ a_hi, const int8_t *__restrict__ a_lo,
                    const int8_t *__restrict__ b_hi, const int8_t *__restrict__ b_lo, int n_pad,
                    int *__restrict__ out, long long ldo) {
    __shared__ int sAh[kSimtTile][17], sAl[kSimtTile][17], sBh[kSimtTile][17], sBl[kSimtTile][17];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.y * kSimtTile;
    const long long c0 = (long long)blockIdx.x * kSimtTile;
    int hh[4][4] = {}, mid[4][4] = {}, ll[4][4] = {};
    const int lr = tid >> 2, lw = (tid & 3) << 2;          // row 0..63, word offset 0,4,8,12
    for (int k0 = 0; k0 < n_pad; k0 += 64) {
        const long long ao = (m0 + lr) * n_pad + k0 + (lw << 2);
        const long long bo = (c0 + lr) * n_pad + k0 + (lw << 2);
        const int4 vah = *reinterpret_cast<const int4 *>(a_hi + ao);
        const int4 val = *reinterpret_cast<const int4 *>(a_lo + ao);
        const int4 vbh = *reinterpret_cast<const int4 *>(b_hi + bo);
        const int4 vbl = *reinterpret_cast<const int4 *>(b_lo + bo);
        sAh[lr][lw] = vah.x; sAh[lr][lw + 1] = vah.y; sAh[lr][lw + 2] = vah.z; sAh[lr][lw + 3] = vah.w;
        sAl[lr][lw] = val.x; sAl[lr][lw + 1] = val.y; sAl[lr][lw + 2] = val.z; sAl[lr][lw + 3] = val.w;
        sBh[lr][lw] = vbh.x; sBh[lr][lw + 1] = vbh.y; sBh[lr][lw + 2] = vbh.z; sBh[lr][lw + 3] = vbh.w;
        sBl[lr][lw] = vbl.x; sBl[lr][lw + 1] = vbl.y; sBl[lr][lw + 2] = vbl.z; sBl[lr][lw + 3] = vbl.w;
        __syncthreads();
#pragma unroll 4
        for (int w = 0; w < 16; ++w) {
            int ah[4], al[4], bh[4], bl[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ah[i] = sAh[ty * 4 + i][w]; al[i] = sAl[ty * 4 + i][w];
                bh[i] = sBh[tx * 4 + i][w]; bl[i] = sBl[tx * 4 + i][w];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    hh[i][j] = __dp4a(ah[i], bh[j], hh[i][j]);
                    mid[i][j] = __dp4a(ah[i], bl[j], mid[i][j]);
                    mid[i][j] = __dp4a(al[i], bh[j], mid[i][j]);
                    ll[i][j] = __dp4a(al[i], bl[j], ll[i][j]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int4 v;
        int *pv = reinterpret_cast<int *>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            pv[j] = (int)(16384LL * hh[i][j] + 128LL * mid[i][j] + (long long)ll[i][j]);
        *reinterpret_cast<int4 *>(out + (m0 + ty * 4 + i) * ldo + c0 + tx * 4) = v;
    }
}

}  // namespace coex
