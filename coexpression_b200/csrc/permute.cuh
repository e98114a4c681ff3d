// permute.cuh -- row N1 of SURVEY.md section 8(f): permutation generation + locus -> promoter mapping on the device.
//
// Replaces, for K locus sets at once, what reference 0-prepare-coex.py does once per permutation with Python dict
// loops and one `windowBed` subprocess each:
//   get_gwad ............ reference coexfunctions.py:102-112
//   shift + get_address . reference 0-prepare-coex.py:548-563 (circular, and every mode's real data: shift == 0),
//                         coexfunctions.py:115-128 (`while gwad > genome_max`, `start < gwad <= end`, else chr1:0)
//   backcirc ............ reference 0-prepare-coex.py:579-593 with listitem, coexfunctions.py:319-325 (C round())
//   background .......... reference 0-prepare-coex.py:564-578: every locus of a permuted set is a line of the background
//                         file drawn by `random.sample` (the draw itself stays on the host: it is the reference's RNG
//                         stream); chromosome = column 0, address = column 2 of that line
//   windowBed -w W ...... bedtools window: A = [address, address + 1) widened by W (start clamped at 0), every
//                         feature with a.start' < b.end and b.start < a.end' on the same chromosome
//   promoters ........... first-appearance order over (locus in input order, feature in the order `bedtools window` reports the
//                         B records of one A record: UCSC-bin level, bin, line of the B file -- coex.cu upload_window_order),
//                         the order a Python-3 dict gives to the reference's snp_map (1-make-network.py:131-177)
// The features are the rows of the resident table (coex_set_features): a BED feature that is not in the expression
// table is dropped by the reference anyway (1-make-network.py:142-146).
// Output = perm_off / hit_rows exactly as coex_network_batch takes them, optionally left on the device.
#pragma once
#include "common.cuh"

#include <algorithm>

namespace coex {

struct PermuteArgs {
    int mode;                        // 0 circular, 1 backcirc, 2 background
    int M, K;
    const int *snp_chrom;            // [M] index into the chromosome list
    const long long *snp_addr;       // [M] address on that chromosome
    const long long *snp_back;       // mode 1: [M] index of the SNP in the sorted background; mode 2: [K, M] drawn background lines
    const long long *shifts;         // [K]
    const long long *chrom_start, *chrom_end;   // [n_chrom] genome-wide range of every chromosome of the list
    const int *chrom_fid;            // [n_chrom] id of the chromosome among the feature chromosomes (-1 = none)
    int n_chrom;
    const int *back_chrom;           // [n_back] sorted background: chromosome list index
    const long long *back_addr;      // [n_back]
    long long n_back;
    long long window;
    // feature index (sorted by chromosome id, start)
    const unsigned long long *fkey;  // [N] chrom_id << 40 | start
    const int *frow;                 // [N] table row
    const long long *fstart, *fend;  // [N] by table row
    const int *wrank, *wrow;         // [N] rank of a table row in bedtools' reporting order, and the row of a rank
    long long N;
    long long maxlen;
    // outputs / scratch
    int *cand_count;                 // [K] candidate (locus, feature) pairs of each set
    int *hit_count;                  // [K]
    int *hits_tmp;                   // [K, cap]
    int *loci_tmp;                   // [K, cap, 2] or nullptr: first and last locus (input order) mapping to each hit
    int cap;                         // power of two
    int *overflow;
};

// position of locus m in set k: feature chromosome id (or -1) and address
__device__ __forceinline__ void permuted_position(const PermuteArgs &a, int k, int m, int &fid, long long &addr) {
    const long long shift = a.shifts[k];
    int ci;
    long long ad;
    if (a.mode == 2 && shift != 0) {
        const long long d = a.snp_back[(long long)k * a.M + m];     // the line random.sample drew for this locus
        if (d < 0 || d >= a.n_back) { fid = -1; addr = 0; return; }
        ci = a.back_chrom[d];
        ad = a.back_addr[d];
    } else if (a.mode == 1 && shift != 0) {
        // listitem(sortedbackground, index + shift), coexfunctions.py:319-325
        const double n = (double)a.n_back;
        const double x = (double)(a.snp_back[m] + shift) / n;
        const double b = x - (double)(long long)round(x);
        long long d = (long long)round(b * n);
        if (d < 0) d += a.n_back;                                   // Python indexes from the end
        if (d < 0 || d >= a.n_back) { fid = -1; addr = 0; return; } // IndexError in the reference
        ci = a.back_chrom[d];
        ad = a.back_addr[d];
    } else {
        ci = a.snp_chrom[m];
        const long long ad0 = a.snp_addr[m];
        // get_gwad: None (the reference then fails) unless the address is on the chromosome or negative
        if (!((a.chrom_start[ci] + ad0) < a.chrom_end[ci] || ad0 < 0)) { fid = -1; addr = 0; return; }
        long long g = a.chrom_start[ci] + ad0 + shift;
        const long long gmax = a.chrom_end[a.n_chrom - 1];
        if (g > gmax) g -= ((g - gmax - 1) / gmax + 1) * gmax;      // while gwad > genome_max: gwad -= genome_max
        ci = 0; ad = 0;                                              // "chr1", 0 when no range matches
        for (int c = 0; c < a.n_chrom; ++c)
            if (g > a.chrom_start[c] && g <= a.chrom_end[c]) { ci = c; ad = g - a.chrom_start[c]; break; }
    }
    fid = a.chrom_fid[ci];
    addr = ad;
}

__device__ __forceinline__ long long lower_bound_u64(const unsigned long long *v, long long n, unsigned long long key) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (v[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// visits every feature row overlapping the widened locus, in sorted-feature order
template <class F>
__device__ __forceinline__ void for_each_overlap(const PermuteArgs &a, int fid, long long addr, F f) {
    if (fid < 0) return;
    const long long a0 = (addr - a.window > 0) ? addr - a.window : 0;
    const long long a1 = addr + 1 + a.window;
    const long long s0 = (a0 - a.maxlen > 0) ? a0 - a.maxlen : 0;
    const unsigned long long base = (unsigned long long)fid << 40;
    long long i = lower_bound_u64(a.fkey, a.N, base | (unsigned long long)s0);
    const unsigned long long stop = base | (unsigned long long)a1;            // start < a1
    for (; i < a.N && a.fkey[i] < stop; ++i) {
        const int row = a.frow[i];
        if (a.fend[row] > a0) f(row);
    }
}

__global__ void permute_count_kernel(PermuteArgs a) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)a.K * a.M) return;
    const int k = (int)(t / a.M), m = (int)(t - (long long)k * a.M);
    int fid; long long addr;
    permuted_position(a, k, m, fid, addr);
    int n = 0;
    for_each_overlap(a, fid, addr, [&](int) { ++n; });
    if (n) atomicAdd(&a.cand_count[k], n);
}

__device__ void bitonic_keys(unsigned long long *keys, int n2) {
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
                const int x = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int y = x | j;
                const unsigned long long ka = keys[x], kb = keys[y];
                if ((ka > kb) == ((x & k) == 0)) { keys[x] = kb; keys[y] = ka; }
            }
            __syncthreads();
        }
}

// one CTA per locus set: candidates -> first appearances -> (locus, bedtools reporting order) order
__global__ void __launch_bounds__(256)
permute_map_kernel(PermuteArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);     // [2 * cap]
    __shared__ int s_n;
    const int k = blockIdx.x;
    if (threadIdx.x == 0) s_n = 0;
    for (int i = threadIdx.x; i < a.cap; i += blockDim.x) keys[i] = ~0ULL;
    __syncthreads();
    for (int m = threadIdx.x; m < a.M; m += blockDim.x) {
        int fid; long long addr;
        permuted_position(a, k, m, fid, addr);
        for_each_overlap(a, fid, addr, [&](int row) {
            const int slot = atomicAdd(&s_n, 1);
            if (slot < a.cap) keys[slot] = ((unsigned long long)(unsigned)row << 32) | (unsigned)m;
        });
    }
    __syncthreads();
    if (s_n > a.cap) { if (threadIdx.x == 0) atomicOr(a.overflow, 1); return; }
    bitonic_keys(keys, a.cap);                                       // by (row, locus)
    // first appearance of every row; everything else becomes a sentinel
    unsigned long long *keys2 = keys + a.cap;
    for (int i = threadIdx.x; i < a.cap; i += blockDim.x) {
        const unsigned long long key = keys[i];
        unsigned long long out = ~0ULL;
        if (key != ~0ULL && (i == 0 || (keys[i - 1] >> 32) != (key >> 32)))
            out = ((key & 0xffffffffULL) << 32) | (unsigned)a.wrank[(int)(key >> 32)];       // (locus, reporting rank of the row)
        keys2[i] = out;
    }
    __syncthreads();
    // (`keys` stays sorted by (row, locus): the last locus of a row -- the reference's real-data map file keeps the LAST
    //  windowBed line of a promoter, 0-prepare-coex.py:396-399 -- is looked up there after the second sort)
    bitonic_keys(keys2, a.cap);                                      // by (locus, reporting rank); sentinels last
    int n = 0;
    const int ncand = min(s_n, a.cap);
    for (int i = threadIdx.x; i < a.cap; i += blockDim.x)
        if (keys2[i] != ~0ULL) {
            const int row = a.wrow[(int)(keys2[i] & 0xffffffffULL)];
            a.hits_tmp[(long long)k * a.cap + i] = row;
            if (a.loci_tmp) {
                // `keys` is still sorted by (row, locus): the last candidate of this row
                int lo = 0, hi = ncand;                       // first index with key.row > row
                while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)(keys[mid] >> 32) <= row) lo = mid + 1; else hi = mid; }
                a.loci_tmp[((long long)k * a.cap + i) * 2] = (int)(keys2[i] >> 32);
                a.loci_tmp[((long long)k * a.cap + i) * 2 + 1] = (int)(keys[lo - 1] & 0xffffffffULL);
            }
            ++n;
        }
    if (n) atomicAdd(&a.hit_count[k], n);
}

__global__ void permute_scan_kernel(const int *hit_count, int K, long long *perm_off) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        long long acc = 0;
        perm_off[0] = 0;
        for (int k = 0; k < K; ++k) { acc += hit_count[k]; perm_off[k + 1] = acc; }
    }
}

__global__ void permute_compact_kernel(const int *hits_tmp, const int *loci_tmp, int cap, const long long *perm_off, int K, int *hit_rows,
                                       int *hit_loci) {
    const int k = blockIdx.x;
    const long long off = perm_off[k];
    const int P = (int)(perm_off[k + 1] - off);
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        hit_rows[off + i] = hits_tmp[(long long)k * cap + i];
        if (hit_loci) {
            hit_loci[2 * (off + i)] = loci_tmp[((long long)k * cap + i) * 2];
            hit_loci[2 * (off + i) + 1] = loci_tmp[((long long)k * cap + i) * 2 + 1];
        }
    }
}

}  // namespace coex
