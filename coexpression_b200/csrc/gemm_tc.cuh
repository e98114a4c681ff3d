// gemm_tc.cuh -- K3: the correlation step as a dense contraction on the sm_100a tensor cores.
//
// Exact-integer Spearman:  u = 2*rank-(n+1) (|u| <= n-1) is split into two int8 digits,
// u = 128*hi + lo, and
//      S[q, c] = sum_k u_q[k]*u_c[k] = 16384*(hi.hi) + 128*(hi.lo + lo.hi) + (lo.lo)
// is formed from four int8 x int8 -> int32 products on tcgen05 (`kind::i8`, UTCIMMA), so the
// correlation is exact (no split-precision error at all): r = (S/4)/(sqrt(Saa/4)*sqrt(Sbb/4))
// reproduces the reference's double result bit for bit (coexpression_v2.c:105-113 on ranks).
//
// Three kernels share the pipeline below: gemm_i8_tc_persistent_kernel (default: one CTA per SM walks
// 128 x 64 tiles, accumulators double-buffered in TMEM so the epilogue overlaps the next tile's MMAs),
// gemm_i8_tc_kernel<128 / 64> (one tile per CTA) and gemm_digits_tc_kernel<D> (Pearson, D digits).
// Kernel shape of gemm_i8_tc_kernel<128> (one CTA per 128 x 128 output tile, 192 threads, warp-specialised):
//   warp 0      TMA producer: per 128-byte K block four SWIZZLE_128B boxes (A_hi, A_lo: 128 rows;
//               B_hi, B_lo: 128 rows each, stacked into ONE 256-row B operand) -> 64 KB / stage,
//               3 stages, mbarrier full/empty ring.
//   warp 1      TMEM allocator (512 columns) + single-thread MMA issuer: per 32-byte K step
//                 D1[128 x 256] += A_hi * [B_hi ; B_lo]^T      (columns 0..127 hi.hi, 128..255 hi.lo)
//                 D2[128 x 256] += A_lo * [B_hi ; B_lo]^T      (columns 0..127 lo.hi, 128..255 lo.lo)
//               i.e. two M128 N256 K32 UTCIMMA per step instead of four N128 ones: the stacked B
//               halves the shared-memory operand traffic per useful MAC.
//   warps 2..5  epilogue: tcgen05.ld the four partial sums, recombine in int64, store int32.
// Grid is rasterised with the query-row tile fastest so that one B column tile is reused from
// L2 by every row tile of the strip and the table streams from HBM once per strip.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace coex {

constexpr int kTcBM = 128, kTcBN = 128, kTcBK = 128;      // BK in bytes (= int8 elements)
constexpr int kTcThreads = 192;
constexpr unsigned kTcSpinLimit = 1u << 22;   // bounded mbarrier polls (~seconds) -> error, not a hang

struct TcState {
    PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    CUtensorMap map_bh, map_bl, map_ah, map_al;
    const void *bh = nullptr, *bl = nullptr, *ah = nullptr, *al = nullptr;
    long long b_rows = 0, a_rows = 0;
    int n_pad = 0;
    int bn = 0;
    int *d_err = nullptr;
    bool attr_set = false;
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must surface as an error code, never as a hung GPU
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int *err, int code) {
    for (unsigned spin = 0; spin < kTcSpinLimit; ++spin)
        if (mbar_try_wait(bar, parity)) return true;
    atomicExch(err, code);
    return false;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

// Epilogue store of 16 consecutive int32 columns of this lane's row.  A lane owns one strip row (the TMEM lane), so a plain 16-byte
// store per lane touches 32 different rows and writes HALF a 32-byte sector each: twice the L2 write requests of the bytes moved
// (the kernel is bound by the strip it writes when rows are short: L2 67 % busy at K = 256).  Lane pairs exchange halves through
// shuffles instead, so that every store instruction writes whole sectors: the even lane's row in two instructions, the odd
// lane's row in the other two.
__device__ __forceinline__ void store_chunk16_paired(int *orow_own, long long ldo, int cc, const int (&v)[16], int lane) {
    const bool odd = (lane & 1) != 0;
    int r[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int b = (q & 3) + 8 * (q >> 2);
        r[q] = __shfl_xor_sync(0xffffffffu, odd ? v[b] : v[b + 4], 1);   // even gets B[0..3], B[8..11]; odd gets A[4..7], A[12..15]
    }
    int *pown = orow_own + cc;
    int *ppar = pown + (odd ? -ldo : ldo);
    // instructions 0 / 1: the even lane's row (A); 2 / 3: the odd lane's row (B)
    // streaming stores: the strip is written once and read once, much later -- it should not displace the operand tiles in L2
    __stcs(reinterpret_cast<int4 *>(odd ? ppar + 4 : pown), odd ? make_int4(r[0], r[1], r[2], r[3]) : make_int4(v[0], v[1], v[2], v[3]));
    __stcs(reinterpret_cast<int4 *>(odd ? ppar + 12 : pown + 8), odd ? make_int4(r[4], r[5], r[6], r[7]) : make_int4(v[8], v[9], v[10], v[11]));
    __stcs(reinterpret_cast<int4 *>(odd ? pown + 4 : ppar), odd ? make_int4(v[4], v[5], v[6], v[7]) : make_int4(r[0], r[1], r[2], r[3]));
    __stcs(reinterpret_cast<int4 *>(odd ? pown + 12 : ppar + 8), odd ? make_int4(v[12], v[13], v[14], v[15]) : make_int4(r[4], r[5], r[6], r[7]));
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): start address >> 4 in
// bits [0,14), LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 1024 B (8 rows x 128 B)
// >> 4 in [32,46), descriptor version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ULL << 16) | (64ULL << 32) | (1ULL << 46) | (2ULL << 61);
}
// kind::i8 instruction descriptor: D = S32 (2 @ bit 4), A = B = signed int8 (1 @ bits 7, 10),
// both K-major (bits 15, 16 = 0), N >> 3 @ bit 17, M >> 4 @ bit 24.
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// BN = 128: one CTA per SM, 3 stages of 64 KB, all 512 TMEM columns.
// BN = 64 : two CTAs per SM (2 stages of 48 KB, 256 TMEM columns each): while one CTA drains its
//           accumulators the other one's MMAs keep the tensor pipe busy, which hides the epilogue
//           without a persistent scheduler.
template <int BN> struct TcCfg {
    static constexpr int kStage = 32768 + 2 * BN * 128;            // A_hi 16K | A_lo 16K | B_hi | B_lo
    static constexpr int kStages = (BN == 128) ? 3 : 2;
    static constexpr int kCols = 4 * BN;                            // TMEM columns: [hh|hl] [lh|ll]
    static constexpr int kSmem = kStages * kStage + 1024 + 256;
    static constexpr int kMinBlocks = (BN == 128) ? 1 : 2;
};

template <int BN>
__global__ void __launch_bounds__(kTcThreads, TcCfg<BN>::kMinBlocks)
gemm_i8_tc_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                  const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                  int *__restrict__ out, long long ldo, int nkb, int *err) {
    using C = TcCfg<BN>;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;          // SWIZZLE_128B: 1024-B aligned
    const uint32_t bars = base + C::kStages * C::kStage;                  // full[], empty[], tmem_full, slot
    const uint32_t bar_full = bars, bar_empty = bars + 8 * C::kStages, bar_tmem = bars + 16 * C::kStages;
    const uint32_t tmem_slot = bar_tmem + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kTcBM;           // query-row tile (fastest)
    const int c0 = blockIdx.y * BN;              // table-column tile

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_tmem, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(C::kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % C::kStages;
                if (kb >= C::kStages && !mbar_wait(bar_empty + 8 * s, ((kb / C::kStages) - 1) & 1, err, 1)) break;
                const uint32_t st = base + s * C::kStage;
                mbar_expect_tx(bar_full + 8 * s, C::kStage);
                tma_load_2d(st, &map_ah, bar_full + 8 * s, kb * kTcBK, m0);
                tma_load_2d(st + 16384, &map_al, bar_full + 8 * s, kb * kTcBK, m0);
                tma_load_2d(st + 32768, &map_bh, bar_full + 8 * s, kb * kTcBK, c0);
                tma_load_2d(st + 32768 + BN * 128, &map_bl, bar_full + 8 * s, kb * kTcBK, c0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(128, 2 * BN);
            bool ok = true;
            for (int kb = 0; kb < nkb && ok; ++kb) {
                const int s = kb % C::kStages;
                ok = mbar_wait(bar_full + 8 * s, (kb / C::kStages) & 1, err, 2);
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * C::kStage;
#pragma unroll
                for (int ks = 0; ks < kTcBK / 32; ++ks) {
                    const uint64_t a_hi = umma_desc_sw128(st + ks * 32);
                    const uint64_t a_lo = umma_desc_sw128(st + 16384 + ks * 32);
                    const uint64_t b_cat = umma_desc_sw128(st + 32768 + ks * 32);
                    const uint32_t acc = (kb | ks) ? 1u : 0u;
                    umma_i8(tmem_base, a_hi, b_cat, idesc, acc);               // [hi.hi | hi.lo]
                    umma_i8(tmem_base + 2 * BN, a_lo, b_cat, idesc, acc);      // [lo.hi | lo.lo]
                }
                umma_commit(bar_empty + 8 * s);        // frees the stage once these MMAs retire
            }
            umma_commit(bar_tmem);                     // accumulators complete
        }
    } else {
        // ===== epilogue: 4 warps, TMEM lane quarter = warp % 4 =====
        const int quarter = warp & 3;
        if (mbar_wait(bar_tmem, 0, err, 3)) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16);
            int *orow = out + (long long)(m0 + quarter * 32 + lane) * ldo + c0;
#pragma unroll 1
            for (int cc = 0; cc < BN; cc += 16) {
                uint32_t hh[16], hl[16], lh[16], ll[16];
                tmem_ld16(tl + cc, hh);
                tmem_ld16(tl + BN + cc, hl);
                tmem_ld16(tl + 2 * BN + cc, lh);
                tmem_ld16(tl + 3 * BN + cc, ll);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    int4 v;
                    v.x = (int)(16384LL * (int)hh[j] + 128LL * ((long long)(int)hl[j] + (int)lh[j]) + (int)ll[j]);
                    v.y = (int)(16384LL * (int)hh[j + 1] + 128LL * ((long long)(int)hl[j + 1] + (int)lh[j + 1]) + (int)ll[j + 1]);
                    v.z = (int)(16384LL * (int)hh[j + 2] + 128LL * ((long long)(int)hl[j + 2] + (int)lh[j + 2]) + (int)ll[j + 2]);
                    v.w = (int)(16384LL * (int)hh[j + 3] + 128LL * ((long long)(int)hl[j + 3] + (int)lh[j + 3]) + (int)ll[j + 3]);
                    *reinterpret_cast<int4 *>(orow + cc + j) = v;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::kCols));
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM walks a static list of 128 x 64 tiles.  The accumulators of a
// tile take 256 TMEM columns ([hh|hl] [lh|ll], 64 columns each), so TWO tiles fit in TMEM and the
// epilogue of tile t (warps 2..5) overlaps the MMAs of tile t+1 (warp 1) while warp 0 keeps the
// 4-stage TMA ring (48 KB / stage) running across tile boundaries.
//   barriers: full[4] / empty[4] (TMA <-> MMA), tmem_full[2] / tmem_empty[2] (MMA <-> epilogue)
// ---------------------------------------------------------------------------------------------
constexpr int kPsBN = 64, kPsStages = 4, kPsStage = 32768 + 2 * kPsBN * 128;
constexpr int kPsSmem = kPsStages * kPsStage + 1024 + 256;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(kTcThreads, 1)
gemm_i8_tc_persistent_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                             const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                             int *__restrict__ out, long long ldo, int nkb, int m_tiles, long long n_tiles, int col_tile0,
                             int *err) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t bars = base + kPsStages * kPsStage;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * kPsStages;
    const uint32_t bar_tfull = bars + 16 * kPsStages, bar_tempty = bar_tfull + 16;
    const uint32_t tmem_slot = bar_tempty + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kPsStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===== TMA producer: the stage ring runs continuously over all tiles of this CTA =====
        if (lane == 0) {
            long long g = 0;                                            // global K-block counter
            bool ok = true;
            for (long long tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
                const int m0 = (int)(tile % m_tiles) * kTcBM;           // query-row tile fastest
                const int c0 = ((int)(tile / m_tiles) + col_tile0) * kPsBN;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = (int)(g % kPsStages);
                    if (g >= kPsStages && !mbar_wait(bar_empty + 8 * s, (uint32_t)((g / kPsStages) - 1) & 1u, err, 1)) { ok = false; break; }
                    const uint32_t st = base + s * kPsStage;
                    mbar_expect_tx(bar_full + 8 * s, kPsStage);
                    tma_load_2d(st, &map_ah, bar_full + 8 * s, kb * kTcBK, m0);
                    tma_load_2d(st + 16384, &map_al, bar_full + 8 * s, kb * kTcBK, m0);
                    tma_load_2d(st + 32768, &map_bh, bar_full + 8 * s, kb * kTcBK, c0);
                    tma_load_2d(st + 32768 + kPsBN * 128, &map_bl, bar_full + 8 * s, kb * kTcBK, c0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: alternates between the two accumulator buffers =====
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(128, 2 * kPsBN);
            long long g = 0;
            int it = 0;
            bool ok = true;
            for (long long tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                if (it >= 2 && !mbar_wait(bar_tempty + 8 * buf, (uint32_t)((it >> 1) - 1) & 1u, err, 4)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc0 = tmem_base + buf * 256;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = (int)(g % kPsStages);
                    if (!mbar_wait(bar_full + 8 * s, (uint32_t)(g / kPsStages) & 1u, err, 2)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * kPsStage;
#pragma unroll
                    for (int ks = 0; ks < kTcBK / 32; ++ks) {
                        const uint64_t a_hi = umma_desc_sw128(st + ks * 32);
                        const uint64_t a_lo = umma_desc_sw128(st + 16384 + ks * 32);
                        const uint64_t b_cat = umma_desc_sw128(st + 32768 + ks * 32);
                        const uint32_t acc = (kb | ks) ? 1u : 0u;
                        umma_i8(acc0, a_hi, b_cat, idesc, acc);                    // [hi.hi | hi.lo]
                        umma_i8(acc0 + 2 * kPsBN, a_lo, b_cat, idesc, acc);        // [lo.hi | lo.lo]
                    }
                    umma_commit(bar_empty + 8 * s);
                }
                if (ok) umma_commit(bar_tfull + 8 * buf);
            }
        }
    } else {
        // ===== epilogue: drains buffer `buf` while the MMA warp fills the other one =====
        const int quarter = warp & 3;
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const int m0 = (int)(tile % m_tiles) * kTcBM;
            const int c0 = ((int)(tile / m_tiles) + col_tile0) * kPsBN;
            if (!mbar_wait(bar_tfull + 8 * buf, (uint32_t)(it >> 1) & 1u, err, 3)) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tl = tmem_base + buf * 256 + ((uint32_t)(quarter * 32) << 16);
            int *orow = out + (long long)(m0 + quarter * 32 + lane) * ldo + c0;
#pragma unroll 1
            for (int cc = 0; cc < kPsBN; cc += 16) {
                uint32_t hh[16], hl[16], lh[16], ll[16];
                tmem_ld16(tl + cc, hh);
                tmem_ld16(tl + kPsBN + cc, hl);
                tmem_ld16(tl + 2 * kPsBN + cc, lh);
                tmem_ld16(tl + 3 * kPsBN + cc, ll);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                int v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    v[j] = (int)(16384LL * (int)hh[j] + 128LL * ((long long)(int)hl[j] + (int)lh[j]) + (int)ll[j]);
                store_chunk16_paired(orow, ldo, cc, v, lane);
            }
            // this warp's TMEM reads have completed (wait::ld): hand the buffer back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent variant with 128 x 128 tiles for long rows (K >= 1024 bytes).  The 128 x 64 kernel above moves
// 28 KB through shared memory per 32-byte K step (12 KB written by TMA, 16 KB read by two M128 N128 MMAs = 219 B/clk at
// the tensor rate) and stalls on it (tensor pipe 61 % at the FANTOM5 shape); with N = 256 MMAs it is 40 KB per TWO such
// steps of time (156 B/clk), which the one-tile-per-CTA kernel sustains at the full tensor rate -- but that kernel pays
// TMEM allocation, barrier set-up, the first TMA latency and a 4-warp epilogue per tile with nothing overlapped (44 % of
// a tile).  A tile's four partial sums fill all 512 TMEM columns, so the accumulators cannot be double-buffered; here
//   * the TMA producer keeps the 3-stage ring (64 KB / stage) running across tile boundaries, so the operands of tile
//     t+1 are in shared memory when the epilogue of tile t hands the accumulators back,
//   * EIGHT epilogue warps drain them (two per TMEM lane quarter, 64 columns each), halving the only serial part,
//   * allocation and barrier set-up happen once per CTA.
//   barriers: full[3] / empty[3] (TMA <-> MMA), tmem_full (MMA -> epilogue), tmem_empty (8 epilogue warps -> MMA)
// ---------------------------------------------------------------------------------------------
constexpr int kP8EpiWarps = 16;                           // epilogue warps: four per TMEM lane quarter, 32 columns each
constexpr int kP8Threads = 64 + 32 * kP8EpiWarps, kP8BN = 128;

__global__ void __launch_bounds__(kP8Threads, 1)
gemm_i8_tc_persistent128_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                                const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                                int *__restrict__ out, long long ldo, int nkb, int m_tiles, long long n_tiles, int col_tile0,
                                int *err) {
    using C = TcCfg<128>;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t bars = base + C::kStages * C::kStage;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * C::kStages;
    const uint32_t bar_tfull = bars + 16 * C::kStages, bar_tempty = bar_tfull + 8;
    const uint32_t tmem_slot = bar_tempty + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, kP8EpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===== TMA producer: the stage ring runs continuously over all tiles of this CTA =====
        if (lane == 0) {
            long long g = 0;
            bool ok = true;
            for (long long tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
                const int m0 = (int)(tile % m_tiles) * kTcBM;           // query-row tile fastest
                const int c0 = ((int)(tile / m_tiles) + col_tile0) * kP8BN;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = (int)(g % C::kStages);
                    if (g >= C::kStages && !mbar_wait(bar_empty + 8 * s, (uint32_t)((g / C::kStages) - 1) & 1u, err, 1)) { ok = false; break; }
                    const uint32_t st = base + s * C::kStage;
                    mbar_expect_tx(bar_full + 8 * s, C::kStage);
                    tma_load_2d(st, &map_ah, bar_full + 8 * s, kb * kTcBK, m0);
                    tma_load_2d(st + 16384, &map_al, bar_full + 8 * s, kb * kTcBK, m0);
                    tma_load_2d(st + 32768, &map_bh, bar_full + 8 * s, kb * kTcBK, c0);
                    tma_load_2d(st + 32768 + kP8BN * 128, &map_bl, bar_full + 8 * s, kb * kTcBK, c0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one accumulator set, handed back by the epilogue before the next tile starts =====
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(128, 2 * kP8BN);
            long long g = 0;
            int it = 0;
            bool ok = true;
            for (long long tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++it) {
                if (it >= 1 && !mbar_wait(bar_tempty, (uint32_t)(it - 1) & 1u, err, 4)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = (int)(g % C::kStages);
                    if (!mbar_wait(bar_full + 8 * s, (uint32_t)(g / C::kStages) & 1u, err, 2)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * C::kStage;
#pragma unroll
                    for (int ks = 0; ks < kTcBK / 32; ++ks) {
                        const uint64_t a_hi = umma_desc_sw128(st + ks * 32);
                        const uint64_t a_lo = umma_desc_sw128(st + 16384 + ks * 32);
                        const uint64_t b_cat = umma_desc_sw128(st + 32768 + ks * 32);
                        const uint32_t acc = (kb | ks) ? 1u : 0u;
                        umma_i8(tmem_base, a_hi, b_cat, idesc, acc);                    // [hi.hi | hi.lo]
                        umma_i8(tmem_base + 2 * kP8BN, a_lo, b_cat, idesc, acc);        // [lo.hi | lo.lo]
                    }
                    umma_commit(bar_empty + 8 * s);
                }
                if (ok) umma_commit(bar_tfull);
            }
        }
    } else {
        // ===== epilogue: kP8EpiWarps warps; TMEM lane quarter = warp % 4, column part = (warp - 2) / 4 =====
        // (the accumulators fill all 512 TMEM columns, so this phase is serial with the MMAs of the next tile: the more warps drain
        //  them in parallel, the shorter it is -- 16 instead of 8 warps, measured)
        constexpr int kCols = kP8BN / (kP8EpiWarps / 4);
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int m0 = (int)(tile % m_tiles) * kTcBM;
            const int c0 = ((int)(tile / m_tiles) + col_tile0) * kP8BN;
            if (!mbar_wait(bar_tfull, (uint32_t)it & 1u, err, 3)) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16);
            int *orow = out + (long long)(m0 + quarter * 32 + lane) * ldo + c0;
#pragma unroll 1
            for (int cc = half * kCols; cc < half * kCols + kCols; cc += 16) {
                uint32_t hh[16], hl[16], lh[16], ll[16];
                tmem_ld16(tl + cc, hh);
                tmem_ld16(tl + kP8BN + cc, hl);
                tmem_ld16(tl + 2 * kP8BN + cc, lh);
                tmem_ld16(tl + 3 * kP8BN + cc, ll);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                int v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    v[j] = (int)(16384LL * (int)hh[j] + 128LL * ((long long)(int)hl[j] + (int)lh[j]) + (int)ll[j]);
                store_chunk16_paired(orow, ldo, cc, v, lane);
            }
            // this warp's TMEM reads have completed (wait::ld): hand the accumulators back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// =============================================================================================
// 128 x 128 tiles, TWO CTAs of a cluster on neighbouring query-row tiles of the SAME table-column tile: the B operand of a K block
// is fetched from L2 once per cluster -- rank 0 loads the B_hi plane, rank 1 the B_lo plane, each with a multicast TMA into both
// CTAs -- so a CTA reads 48 KB instead of 64 KB per K block.  Why: at the FANTOM5 shape the 128 x 128 kernel pulls 2.25 TB through
// L2 -> shared memory in 189 ms (11.9 TB/s): it is bound by that traffic, not by the tensor pipe (77 %) or its epilogue.
// Same MMAs (cta_group::1), same epilogue.  A stage may be refilled when BOTH CTAs have consumed it (the partner's multicast
// writes into it too): every MMA thread commits to the empty barrier of both CTAs (count 2).
// =============================================================================================
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, unsigned short mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kP8Threads, 1)
gemm_i8_tc_cluster128_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                             const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                             int *__restrict__ out, long long ldo, int nkb, int m_tiles, long long n_tiles, int col_tile0,
                             int *err) {
    using C = TcCfg<128>;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t bars = base + C::kStages * C::kStage;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * C::kStages;
    const uint32_t bar_tfull = bars + 16 * C::kStages, bar_tempty = bar_tfull + 8;
    const uint32_t tmem_slot = bar_tempty + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 2); }
        mbar_init(bar_tfull, 1);
        mbar_init(bar_tempty, kP8EpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // the partner's barriers exist
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // tile pair p of the cluster: tiles 2 p and 2 p + 1 in query-row-fastest order (m_tiles is even: the same column tile)
    const long long first = 2 * (long long)(blockIdx.x >> 1) + rank, step = 2 * (long long)(gridDim.x >> 1);

    if (warp == 0) {
        if (lane == 0) {
            long long g = 0;
            bool ok = true;
            for (long long tile = first; tile < n_tiles && ok; tile += step) {
                const int m0 = (int)(tile % m_tiles) * kTcBM;
                const int c0 = ((int)(tile / m_tiles) + col_tile0) * kP8BN;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = (int)(g % C::kStages);
                    if (g >= C::kStages && !mbar_wait(bar_empty + 8 * s, (uint32_t)((g / C::kStages) - 1) & 1u, err, 1)) { ok = false; break; }
                    const uint32_t st = base + s * C::kStage;
                    mbar_expect_tx(bar_full + 8 * s, C::kStage);
                    tma_load_2d(st, &map_ah, bar_full + 8 * s, kb * kTcBK, m0);
                    tma_load_2d(st + 16384, &map_al, bar_full + 8 * s, kb * kTcBK, m0);
                    if (rank == 0) tma_load_2d_mc(st + 32768, &map_bh, bar_full + 8 * s, kb * kTcBK, c0, 3);
                    else tma_load_2d_mc(st + 32768 + kP8BN * 128, &map_bl, bar_full + 8 * s, kb * kTcBK, c0, 3);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(128, 2 * kP8BN);
            long long g = 0;
            int it = 0;
            bool ok = true;
            for (long long tile = first; tile < n_tiles && ok; tile += step, ++it) {
                if (it >= 1 && !mbar_wait(bar_tempty, (uint32_t)(it - 1) & 1u, err, 4)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = (int)(g % C::kStages);
                    if (!mbar_wait(bar_full + 8 * s, (uint32_t)(g / C::kStages) & 1u, err, 2)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * C::kStage;
#pragma unroll
                    for (int ks = 0; ks < kTcBK / 32; ++ks) {
                        const uint64_t a_hi = umma_desc_sw128(st + ks * 32);
                        const uint64_t a_lo = umma_desc_sw128(st + 16384 + ks * 32);
                        const uint64_t b_cat = umma_desc_sw128(st + 32768 + ks * 32);
                        const uint32_t acc = (kb | ks) ? 1u : 0u;
                        umma_i8(tmem_base, a_hi, b_cat, idesc, acc);                    // [hi.hi | hi.lo]
                        umma_i8(tmem_base + 2 * kP8BN, a_lo, b_cat, idesc, acc);        // [lo.hi | lo.lo]
                    }
                    umma_commit_mc(bar_empty + 8 * s, 3);                              // both CTAs' producers
                }
                if (ok) umma_commit(bar_tfull);
            }
        }
    } else {
        constexpr int kCols = kP8BN / (kP8EpiWarps / 4);
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        int it = 0;
        for (long long tile = first; tile < n_tiles; tile += step, ++it) {
            const int m0 = (int)(tile % m_tiles) * kTcBM;
            const int c0 = ((int)(tile / m_tiles) + col_tile0) * kP8BN;
            if (!mbar_wait(bar_tfull, (uint32_t)it & 1u, err, 3)) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16);
            int *orow = out + (long long)(m0 + quarter * 32 + lane) * ldo + c0;
#pragma unroll 1
            for (int cc = half * kCols; cc < half * kCols + kCols; cc += 16) {
                uint32_t hh[16], hl[16], lh[16], ll[16];
                tmem_ld16(tl + cc, hh);
                tmem_ld16(tl + kP8BN + cc, hl);
                tmem_ld16(tl + 2 * kP8BN + cc, lh);
                tmem_ld16(tl + 3 * kP8BN + cc, ll);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                int v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    v[j] = (int)(16384LL * (int)hh[j] + 128LL * ((long long)(int)hl[j] + (int)lh[j]) + (int)ll[j]);
                store_chunk16_paired(orow, ldo, cc, v, lane);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // (the partner may still be multicasting into this CTA's shared memory / arriving on its barriers)
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// =============================================================================================
// cta_group::2 helpers (two CTAs of a cluster = one TPC driving ONE tcgen05.mma of M = 256)
// =============================================================================================

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((unsigned short)3) : "memory");
}

// =============================================================================================
// DIGIT-SPLIT CTA pair (cta_group::2): both CTAs of a cluster work on the SAME 128 x 128 output tile; the leader stages the A_hi
// plane of the 128 query rows and the B_hi plane of the 128 table rows, its peer A_lo and B_lo, and ONE tcgen05.mma.cta_group::2
// of M = 256, N = 256 per K step computes [A_hi ; A_lo] x [B_hi ; B_lo]^T: the leader's TMEM receives [hi.hi | hi.lo], the peer's
// [lo.hi | lo.lo] -- 256 columns per tile and CTA, so the accumulators are DOUBLE-BUFFERED and the epilogue of tile t runs under the
// MMAs of tile t + 1, with the N = 256 MMAs the other double-buffered shapes cannot have.  The price is a
// combine step: S = (16384 hi.hi + 128 hi.lo) + (128 lo.hi + lo.lo), the two brackets live in different CTAs.  Each is formed
// modulo 2^32 (S itself fits 32 bits); the leader finishes columns 0..63 of the tile, its peer columns 64..127, and each sends the
// other half of its bracket to the partner's shared memory (st.shared::cluster, [column][row] layout: conflict-free).
//   full[s]     leader's barrier, armed with the bytes of both CTAs       empty[s]   per CTA, multicast commit
//   tfull[a]    per CTA, multicast commit after a tile's last K block     tempty[a]  leader's, 2 x 16 arrivals (the peer's remotely)
//   recv[a]     per CTA: the partner's 8 sender warps have written their half bracket into this CTA's staging buffer a
//   sendok[a]   per CTA: the partner's 8 keeper warps have read staging buffer a of THEIR CTA (it may be written again)
// =============================================================================================
constexpr int kSpStages = 4, kSpStage = 16384 + 16384, kSpStaging = 64 * 128 * 4;
constexpr int kSpSmem = kSpStages * kSpStage + 2 * kSpStaging + 1024 + 256;

// wait with acquire semantics at CLUSTER scope: what the partner CTA wrote into this CTA's shared memory before its (release.cluster)
// arrival is visible afterwards
__device__ __forceinline__ bool mbar_wait_cluster(uint32_t bar, uint32_t parity, int *err, int code) {
    for (unsigned spin = 0; spin < kTcSpinLimit; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    atomicExch(err, code);
    return false;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kP8Threads, 1)
gemm_i8_tc_split_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
                        const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl,
                        int *__restrict__ out, long long ldo, int nkb, int m_tiles, long long n_tiles, int col_tile0, int *err) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t staging = base + kSpStages * kSpStage;                     // [2][64 columns][128 rows] u32
    const uint32_t bars = staging + 2 * kSpStaging;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * kSpStages;
    const uint32_t bar_tfull = bars + 16 * kSpStages, bar_tempty = bar_tfull + 16;
    const uint32_t bar_recv = bar_tempty + 16, bar_sendok = bar_recv + 16;
    const uint32_t tmem_slot = bar_sendok + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kSpStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 2 * kP8EpiWarps);
            mbar_init(bar_recv + 8 * a, kP8EpiWarps / 2); mbar_init(bar_sendok + 8 * a, kP8EpiWarps / 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const long long cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

    if (warp == 0) {
        // ===== TMA producer: this CTA's digit plane of the query rows and of the table rows =====
        if (lane == 0) {
            const CUtensorMap *ma = leader ? &map_ah : &map_al, *mb = leader ? &map_bh : &map_bl;
            long long g = 0;
            bool ok = true;
            for (long long tile = cid; tile < n_tiles && ok; tile += ncl) {
                const int m0 = (int)(tile % m_tiles) * kTcBM;                         // query-row tile fastest
                const int c0 = ((int)(tile / m_tiles) + col_tile0) * kP8BN;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = (int)(g % kSpStages);
                    if (g >= kSpStages && !mbar_wait(bar_empty + 8 * s, (uint32_t)((g / kSpStages) - 1) & 1u, err, 1)) { ok = false; break; }
                    const uint32_t st = base + s * kSpStage;
                    const uint32_t lf = mapa_rank(bar_full + 8 * s, 0);
                    if (leader) mbar_expect_tx(bar_full + 8 * s, 2 * kSpStage);
                    tma_load_2d_pair(st, ma, lf, kb * kTcBK, m0);
                    tma_load_2d_pair(st + 16384, mb, lf, kb * kTcBK, c0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader only): one M = 256, N = 256 MMA per K step, two accumulator sets =====
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = umma_idesc_i8(256, 2 * kP8BN);
            long long g = 0;
            int it = 0;
            bool ok = true;
            for (long long tile = cid; tile < n_tiles && ok; tile += ncl, ++it) {
                const int as = it & 1;
                if (it >= 2 && !mbar_wait(bar_tempty + 8 * as, (uint32_t)((it >> 1) - 1) & 1u, err, 4)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d0 = tmem_base + (uint32_t)as * 256u;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = (int)(g % kSpStages);
                    if (!mbar_wait(bar_full + 8 * s, (uint32_t)(g / kSpStages) & 1u, err, 2)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * kSpStage;
#pragma unroll
                    for (int ks = 0; ks < kTcBK / 32; ++ks)
                        umma_i8_pair(d0, umma_desc_sw128(st + ks * 32), umma_desc_sw128(st + 16384 + ks * 32), idesc, (kb | ks) ? 1u : 0u);
                    umma_commit_pair(bar_empty + 8 * s);
                }
                if (ok) umma_commit_pair(bar_tfull + 8 * as);
            }
        }
    } else {
        // ===== epilogue: TMEM lane quarter = warp % 4 (32 tile rows), 32 of the 128 tile columns per warp =====
        const int quarter = warp & 3, cp = (warp - 2) >> 2;                    // cp 0, 1: columns 0..63, cp 2, 3: columns 64..127
        const bool keeper = leader ? (cp < 2) : (cp >= 2);                     // this CTA finishes and stores these columns
        const uint32_t peer = rank ^ 1u;
        const uint32_t leader_tempty0 = mapa_rank(bar_tempty, 0);
        const uint32_t peer_staging = mapa_rank(staging, peer), peer_recv = mapa_rank(bar_recv, peer), peer_sendok = mapa_rank(bar_sendok, peer);
        const int row = quarter * 32 + lane;
        const int scol0 = (cp & 1) * 32;                                       // first column of the warp inside a 64-column half
        const unsigned int w_hi = leader ? 16384u : 128u, w_lo = leader ? 128u : 1u;   // this CTA's bracket: w_hi * [x.hi] + w_lo * [x.lo]
        int it = 0;
        for (long long tile = cid; tile < n_tiles; tile += ncl, ++it) {
            const int as = it & 1;
            const int m0 = (int)(tile % m_tiles) * kTcBM;
            const int c0 = ((int)(tile / m_tiles) + col_tile0) * kP8BN;
            if (!mbar_wait(bar_tfull + 8 * as, (uint32_t)(it >> 1) & 1u, err, 3)) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tl = tmem_base + (uint32_t)as * 256u + ((uint32_t)(quarter * 32) << 16);
            bool ok = true;
            if (!keeper) {
                // the partner's staging buffer `as` has been read (tile it - 2)
                if (it >= 2 && !mbar_wait_cluster(bar_sendok + 8 * as, (uint32_t)((it >> 1) - 1) & 1u, err, 5)) ok = false;
            } else {
                if (!mbar_wait_cluster(bar_recv + 8 * as, (uint32_t)(it >> 1) & 1u, err, 6)) ok = false;
            }
            if (!ok) break;
            int *orow = out + (long long)(m0 + row) * ldo + c0 + (leader ? 0 : 64);
#pragma unroll 1
            for (int cc = 0; cc < 32; cc += 16) {
                uint32_t xh[16], xl[16];
                tmem_ld16(tl + cp * 32 + cc, xh);
                tmem_ld16(tl + kP8BN + cp * 32 + cc, xl);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (!keeper) {
                    const uint32_t dst = peer_staging + (uint32_t)as * kSpStaging + ((uint32_t)(scol0 + cc) * 128u + (uint32_t)row) * 4u;
#pragma unroll
                    for (int j = 0; j < 16; ++j) st_cluster_u32(dst + (uint32_t)j * 512u, w_hi * xh[j] + w_lo * xl[j]);
                } else {
                    const uint32_t src = staging + (uint32_t)as * kSpStaging + ((uint32_t)(scol0 + cc) * 128u + (uint32_t)row) * 4u;
                    int v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        uint32_t other;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(other) : "r"(src + (uint32_t)j * 512u));
                        v[j] = (int)(w_hi * xh[j] + w_lo * xl[j] + other);
                    }
                    store_chunk16_paired(orow, ldo, scol0 + cc, v, lane);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (!keeper) mbar_arrive_cluster(peer_recv + 8 * as);           // release.cluster: the stores above are visible to the partner
                else mbar_arrive_cluster(peer_sendok + 8 * as);
                mbar_arrive_cluster(leader_tempty0 + 8 * as);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// =============================================================================================
// Generic D-digit variant (Pearson): z-scores quantised to D base-128 int8 digits
// (prepare.cuh: pearson_pack_kernel).  One CTA per 128 x 32 output tile; per 32-byte K step
//      D_i[128 x (D*32)] += A_i * [B_{D-1} ; ... ; B_0]^T            i = 0 .. D-1
// so TMEM holds the D*D partial products (D*D*32 <= 512 columns) and the epilogue recombines
//      S = sum_{i,l} 128^(i+l) * P_il          (int64, exact),     r = S / s^2      (double).
// Column tiles entirely at or below the diagonal are skipped when `upper_only` is set
// (all-pairs: r(a,b) = r(b,a)).
// =============================================================================================
template <int D> struct TcDigits {
    static constexpr int BN = 32;
    static constexpr int N = D * BN;                           // UMMA N (96 or 128)
    static constexpr int kStage = D * 16384 + D * BN * 128;    // A digits + stacked B digits
    static constexpr int kStages = (D == 3) ? 3 : 2;
    static constexpr int kSmem = kStages * kStage + 1024 + 256;
};

struct TcMaps8 { CUtensorMap a[4]; CUtensorMap b[4]; };

template <int D>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_digits_tc_kernel(const __grid_constant__ TcMaps8 maps, double *__restrict__ out, long long ldo, int nkb,
                      long long a_row0, const double *__restrict__ inv_s_a, const double *__restrict__ inv_s, int upper_only,
                      float *__restrict__ out32, int *err) {
    using T = TcDigits<D>;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t bars = base + T::kStages * T::kStage;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * T::kStages, bar_tmem = bars + 16 * T::kStages;
    const uint32_t tmem_slot = bar_tmem + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kTcBM;                 // row tile inside the strip (fastest)
    const int c0 = blockIdx.y * T::BN;                 // table-column tile
    if (upper_only && (long long)c0 + T::BN <= a_row0 + m0) return;      // whole tile is below the diagonal

    if (threadIdx.x == 0) {
        for (int s = 0; s < T::kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_tmem, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % T::kStages;
                if (kb >= T::kStages && !mbar_wait(bar_empty + 8 * s, ((kb / T::kStages) - 1) & 1, err, 1)) break;
                const uint32_t st = base + s * T::kStage;
                mbar_expect_tx(bar_full + 8 * s, T::kStage);
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    tma_load_2d(st + d * 16384, &maps.a[d], bar_full + 8 * s, kb * kTcBK, (int)(a_row0 + m0));
                    // stacked B operand: highest digit first is irrelevant, only the order must match the epilogue
                    tma_load_2d(st + D * 16384 + d * T::BN * 128, &maps.b[d], bar_full + 8 * s, kb * kTcBK, c0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(128, T::N);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % T::kStages;
                if (!mbar_wait(bar_full + 8 * s, (kb / T::kStages) & 1, err, 2)) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * T::kStage;
#pragma unroll
                for (int ks = 0; ks < kTcBK / 32; ++ks) {
                    const uint64_t b_cat = umma_desc_sw128(st + D * 16384 + ks * 32);
                    const uint32_t acc = (kb | ks) ? 1u : 0u;
#pragma unroll
                    for (int d = 0; d < D; ++d)
                        umma_i8(tmem_base + d * T::N, umma_desc_sw128(st + d * 16384 + ks * 32), b_cat, idesc, acc);
                }
                umma_commit(bar_empty + 8 * s);
            }
            umma_commit(bar_tmem);
        }
    } else {
        const int quarter = warp & 3;
        if (mbar_wait(bar_tmem, 0, err, 3)) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16);
            double *orow = out + (long long)(m0 + quarter * 32 + lane) * ldo + c0;
            float *orow32 = out32 + (long long)(m0 + quarter * 32 + lane) * ldo + c0;      // out32 != nullptr: float strip instead
            const double inv_a = inv_s_a[a_row0 + m0 + quarter * 32 + lane];
#pragma unroll 1
            for (int cc = 0; cc < T::BN; cc += 16) {
                long long S[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) S[j] = 0;
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int l = 0; l < D; ++l) {
                        uint32_t v[16];
                        tmem_ld16(tl + i * T::N + l * T::BN + cc, v);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int j = 0; j < 16; ++j) S[j] += (long long)(int)v[j] * (1LL << (7 * (i + l)));
                    }
                if (out32) {
                    // float strip: whole sectors per store instruction (see store_chunk16_paired), not 8 bytes per lane and row
                    int vf[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) vf[j] = __float_as_int((float)((double)S[j] * (inv_a * inv_s[c0 + cc + j])));
                    store_chunk16_paired(reinterpret_cast<int *>(orow32), ldo, cc, vf, lane);
                } else {
                    // double strip, the same way: the even lane's row in four store instructions, the odd lane's row in the other
                    // four, each lane pair writing one whole 32-byte sector per instruction
                    double w[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) w[j] = (double)S[j] * (inv_a * inv_s[c0 + cc + j]);
                    const bool odd = (lane & 1) != 0;
                    double r[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int b = 4 * (q >> 1) + (q & 1);
                        r[q] = __shfl_xor_sync(0xffffffffu, odd ? w[b] : w[b + 2], 1);   // even gets B[4k], B[4k+1]; odd gets A[4k+2], A[4k+3]
                    }
                    double *pown = orow + cc;
                    double *ppar = pown + (odd ? -ldo : ldo);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {          // row A (the even lane's)
                        *reinterpret_cast<double2 *>(odd ? ppar + 4 * k + 2 : pown + 4 * k) =
                            odd ? make_double2(r[2 * k], r[2 * k + 1]) : make_double2(w[4 * k], w[4 * k + 1]);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {          // row B (the odd lane's)
                        *reinterpret_cast<double2 *>(odd ? pown + 4 * k + 2 : ppar + 4 * k) =
                            odd ? make_double2(w[4 * k + 2], w[4 * k + 3]) : make_double2(r[2 * k], r[2 * k + 1]);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- host side ----------------------------------------------------------------------------------
static int tc_state(coex_ctx *c, TcState **out) {
    if (!c->tc) {
        TcState *t = new TcState();
        cudaDriverEntryPointQueryResult qres;
        void *fn = nullptr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            delete t;
            COEX_FAIL("cuTensorMapEncodeTiled is not available from the CUDA driver");
        }
        t->encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
        if (cudaMalloc(&t->d_err, sizeof(int)) != cudaSuccess) { delete t; COEX_FAIL("cudaMalloc(error flag)"); }
        cudaMemset(t->d_err, 0, sizeof(int));
        c->tc = t;
    }
    *out = static_cast<TcState *>(c->tc);
    return 0;
}

static int tc_encode(TcState *t, CUtensorMap *map, const void *ptr, long long rows, int n_pad, int box_rows = 128) {
    cuuint64_t dims[2] = {(cuuint64_t)n_pad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)n_pad};
    cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = t->encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) COEX_FAIL("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 0;
}

inline int tc_invalidate(coex_ctx *c) {
    if (c->tc) {
        TcState *t = static_cast<TcState *>(c->tc);
        t->bh = t->bl = t->ah = t->al = nullptr;
    }
    return 0;
}

inline void tc_destroy(coex_ctx *c) {
    if (c->tc) {
        TcState *t = static_cast<TcState *>(c->tc);
        if (t->d_err) cudaFree(t->d_err);
        delete t;
        c->tc = nullptr;
    }
}

// S strip [m_rows, N_pad] = A (gathered query digits, c->a_hi / c->a_lo) x table digits^T
inline int tc_gemm_i8_ptrs(coex_ctx *c, const int8_t *a_hi, const int8_t *a_lo, long long a_rows_alloc,
                           long long m_rows, int *out, long long ldo) {
    TcState *t = nullptr;
    COEX_TRY(tc_state(c, &t));
    if (m_rows % kTcBM) COEX_FAIL("tc_gemm: query rows must be padded to 128");
    // gemm_bn: 0 = persistent (128 x 64 tiles; 128 x 128 tiles for long rows), -64 / -128 = that persistent kernel forced,
    //          64 / 128 = one-tile-per-CTA kernel with that table-column tile
    const bool persistent = (c->gemm_bn <= 0);
    const bool split = (c->gemm_bn == -257) && (c->sm_count >= 2);                            // digit-split CTA pairs, 128 x 128 tiles
    const bool wide = persistent && (split || c->gemm_bn == -128 || c->gemm_bn == -129 || (c->gemm_bn == 0 && c->n_pad >= 1024));
    const bool clustered = wide && c->gemm_bn == -129 && (m_rows % 256 == 0) && c->sm_count >= 2;      // B multicast inside CTA pairs
    const int BN = persistent ? (wide ? kP8BN : kPsBN) : c->gemm_bn;
    if (t->bh != c->u_hi.p || t->bl != c->u_lo.p || t->b_rows != c->N_pad || t->n_pad != c->n_pad || t->bn != BN) {
        COEX_TRY(tc_encode(t, &t->map_bh, c->u_hi.p, c->N_pad, c->n_pad, BN));
        COEX_TRY(tc_encode(t, &t->map_bl, c->u_lo.p, c->N_pad, c->n_pad, BN));
        t->bh = c->u_hi.p; t->bl = c->u_lo.p; t->b_rows = c->N_pad; t->n_pad = c->n_pad; t->bn = BN;
        t->ah = t->al = nullptr;
    }
    if (t->ah != a_hi || t->al != a_lo || t->a_rows != a_rows_alloc) {
        COEX_TRY(tc_encode(t, &t->map_ah, a_hi, a_rows_alloc, c->n_pad));
        COEX_TRY(tc_encode(t, &t->map_al, a_lo, a_rows_alloc, c->n_pad));
        t->ah = a_hi; t->al = a_lo; t->a_rows = a_rows_alloc;
    }
    if (!t->attr_set) {
        COEX_CUDA(cudaFuncSetAttribute(gemm_i8_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::kSmem));
        COEX_CUDA(cudaFuncSetAttribute(gemm_i8_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<64>::kSmem));
        COEX_CUDA(cudaFuncSetAttribute(gemm_i8_tc_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPsSmem));
        COEX_CUDA(cudaFuncSetAttribute(gemm_i8_tc_persistent128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::kSmem));
        COEX_CUDA(cudaFuncSetAttribute(gemm_i8_tc_cluster128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::kSmem));
        COEX_CUDA(cudaFuncSetAttribute(gemm_i8_tc_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpSmem));
        t->attr_set = true;
    }
    if (split) {
        const int m_tiles = (int)(m_rows / kTcBM);
        const int col_tile0 = (int)std::min<long long>(c->gemm_col0 / kP8BN, c->N_pad / kP8BN - 1);
        const long long n_tiles = (long long)m_tiles * (c->N_pad / kP8BN - col_tile0);
        const unsigned ctas = 2u * (unsigned)std::min<long long>(n_tiles, c->sm_count / 2);
        gemm_i8_tc_split_kernel<<<ctas, kP8Threads, kSpSmem, c->stream>>>(t->map_ah, t->map_al, t->map_bh, t->map_bl, out, ldo,
                                                                         c->n_pad / kTcBK, m_tiles, n_tiles, col_tile0, t->d_err);
        COEX_CUDA(cudaGetLastError());
        return 0;
    }
    if (wide) {
        const int m_tiles = (int)(m_rows / kTcBM);
        const int col_tile0 = (int)std::min<long long>(c->gemm_col0 / kP8BN, c->N_pad / kP8BN - 1);
        const long long n_tiles = (long long)m_tiles * (c->N_pad / kP8BN - col_tile0);
        if (clustered) {
            const unsigned cl = 2u * (unsigned)std::min<long long>(n_tiles / 2, c->sm_count / 2);
            gemm_i8_tc_cluster128_kernel<<<cl, kP8Threads, TcCfg<128>::kSmem, c->stream>>>(t->map_ah, t->map_al, t->map_bh, t->map_bl,
                                                                                           out, ldo, c->n_pad / kTcBK, m_tiles, n_tiles,
                                                                                           col_tile0, t->d_err);
            COEX_CUDA(cudaGetLastError());
            return 0;
        }
        const unsigned ctas = (unsigned)std::min<long long>(n_tiles, c->sm_count);
        gemm_i8_tc_persistent128_kernel<<<ctas, kP8Threads, TcCfg<128>::kSmem, c->stream>>>(t->map_ah, t->map_al, t->map_bh, t->map_bl,
                                                                                          out, ldo, c->n_pad / kTcBK, m_tiles, n_tiles,
                                                                                          col_tile0, t->d_err);
        COEX_CUDA(cudaGetLastError());
        return 0;
    }
    if (persistent) {
        const int m_tiles = (int)(m_rows / kTcBM);
        const int col_tile0 = (int)std::min<long long>(c->gemm_col0 / kPsBN, c->N_pad / kPsBN - 1);   // upper-triangle callers skip the columns before their rows
        const long long n_tiles = (long long)m_tiles * (c->N_pad / kPsBN - col_tile0);
        const unsigned ctas = (unsigned)std::min<long long>(n_tiles, c->sm_count);
        gemm_i8_tc_persistent_kernel<<<ctas, kTcThreads, kPsSmem, c->stream>>>(t->map_ah, t->map_al, t->map_bh, t->map_bl, out, ldo,
                                                                               c->n_pad / kTcBK, m_tiles, n_tiles, col_tile0, t->d_err);
        COEX_CUDA(cudaGetLastError());
        return 0;
    }
    dim3 grid((unsigned)(m_rows / kTcBM), (unsigned)(c->N_pad / BN));
    if (BN == 128)
        gemm_i8_tc_kernel<128><<<grid, kTcThreads, TcCfg<128>::kSmem, c->stream>>>(t->map_ah, t->map_al, t->map_bh, t->map_bl, out,
                                                                                  ldo, c->n_pad / kTcBK, t->d_err);
    else
        gemm_i8_tc_kernel<64><<<grid, kTcThreads, TcCfg<64>::kSmem, c->stream>>>(t->map_ah, t->map_al, t->map_bh, t->map_bl, out,
                                                                                ldo, c->n_pad / kTcBK, t->d_err);
    COEX_CUDA(cudaGetLastError());
    return 0;
}

inline int tc_gemm_i8(coex_ctx *c, long long m_rows) {
    return tc_gemm_i8_ptrs(c, c->a_hi.p, c->a_lo.p, (long long)(c->a_hi.cap / c->n_pad), m_rows, c->strip.p, c->N_pad);
}

// r strip [m_rows, ldo] (double) = rows [a_row0, a_row0 + m_rows) of the A digit planes x all rows of the table's planes.
// a_planes == planes: rows of the table itself (all-pairs strips); otherwise a gathered copy of arbitrary query rows
// ([D][a_rows, n_pad], a_plane_stride apart) with its own per-row 1 / scale.
template <int D>
inline int tc_gemm_digits_ab(coex_ctx *c, const int8_t *a_planes, long long a_plane_stride, long long a_rows,
                             const double *inv_s_a, const int8_t *planes, long long plane_stride, long long a_row0,
                             long long m_rows, double *out, long long ldo, const double *inv_s, int upper_only,
                             float *out32 = nullptr) {
    using T = TcDigits<D>;
    TcState *t = nullptr;
    COEX_TRY(tc_state(c, &t));
    if (m_rows % kTcBM) COEX_FAIL("tc_gemm_digits: rows must be a multiple of 128");
    TcMaps8 maps;
    for (int d = 0; d < D; ++d) {
        COEX_TRY(tc_encode(t, &maps.a[d], a_planes + d * a_plane_stride, a_rows, c->n_pad, 128));
        COEX_TRY(tc_encode(t, &maps.b[d], planes + d * plane_stride, c->N_pad, c->n_pad, T::BN));
    }
    COEX_CUDA(cudaFuncSetAttribute(gemm_digits_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::kSmem));
    dim3 grid((unsigned)(m_rows / kTcBM), (unsigned)(c->N_pad / T::BN));
    gemm_digits_tc_kernel<D><<<grid, kTcThreads, T::kSmem, c->stream>>>(maps, out, ldo, c->n_pad / kTcBK, a_row0, inv_s_a, inv_s,
                                                                       upper_only, out32, t->d_err);
    COEX_CUDA(cudaGetLastError());
    return 0;
}

template <int D>
inline int tc_gemm_digits(coex_ctx *c, const int8_t *planes, long long plane_stride, long long a_row0, long long m_rows,
                          double *out, long long ldo, const double *inv_s, int upper_only) {
    return tc_gemm_digits_ab<D>(c, planes, plane_stride, c->N_pad, inv_s, planes, plane_stride, a_row0, m_rows, out, ldo, inv_s,
                                upper_only);
}

// non-zero if any tensor-core kernel of this context hit its bounded-wait limit
inline int tc_check_error(coex_ctx *c) {
    if (!c->tc) return 0;
    TcState *t = static_cast<TcState *>(c->tc);
    int h = 0;
    COEX_CUDA(cudaMemcpyAsync(&h, t->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    COEX_CUDA(cudaStreamSynchronize(c->stream));
    if (h) {
        cudaMemsetAsync(t->d_err, 0, sizeof(int), c->stream);
        COEX_FAIL("tcgen05 GEMM pipeline timed out waiting on mbarrier (role code " + std::to_string(h) + ")");
    }
    return 0;
}

}  // namespace coex
