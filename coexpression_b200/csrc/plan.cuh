// plan.cuh -- de-duplication plan of the node-specific null, built ON THE DEVICE.
//
// The null values of a table row do not depend on the permutation, so the count kernels work per UNIQUE hit row
// (count_dedup.cuh).  The plan groups the H queries (= flat hit indices) by table row and cuts every row's queries
// into work items whose statistics fit the kernels' shared memory.  It used to be a host-side counting sort
// (0.5 - 0.7 ms per C2 step with the GPU idle, profiles/r01_final_summary.md); here it is seven small kernels and TWO
// tiny read-backs (the number of unique rows sizes the strips; the first item of every strip sizes the launches):
//   plan_count   cnt[row + 1] += 1 for every query (and the permutation of every query)
//   plan_rows    single-CTA scan over the table rows: unique rows, start of each row's queries, scatter cursors
//   plan_scatter queries grouped by row (the order inside a row is irrelevant: every statistic is counted on its own)
//   plan_items   per unique row: number of items / single-CTA scan / the items themselves
#pragma once
#include "common.cuh"
#include "count_dedup.cuh"

namespace coex {

constexpr int kPlanThreads = 1024;
constexpr int kPlanQcap = 8;             // queries per work item (shared-memory capacity of the count kernels)

// world > 1: ONE job shared by several GPUs -- this rank plans only the queries whose table row it owns (row % world == rank),
// so that a row is correlated and streamed once whatever the number of GPUs (sharding by permutation would repeat it on every
// rank whose permutations hit it).
__global__ void plan_count_kernel(const int *__restrict__ hit_rows, const long long *__restrict__ perm_off, int K,
                                  long long H, long long N, int *__restrict__ cnt, int *__restrict__ q_perm,
                                  int *__restrict__ err, int world, int rank) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= H) return;
    const int r = hit_rows[q];
    if (r < 0 || r >= N) { atomicOr(err, 1); return; }
    if (world > 1 && r % world != rank) return;
    atomicAdd(&cnt[r + 1], 1);
    int lo = 0, hi = K;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (perm_off[mid] <= q) lo = mid; else hi = mid; }
    q_perm[q] = lo;
}

// block-wide exclusive scan of one int per thread (kPlanThreads threads); returns the exclusive prefix, total in *tot
__device__ __forceinline__ int plan_block_scan(int v, int *s_warp, int *tot) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = s_warp[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += x; }
        s_warp[lane] = wi - w;                       // exclusive prefix of the warp totals
        if (lane == 31) *tot = wi;
    }
    __syncthreads();
    const int out = s_warp[wid] + incl - v;
    __syncthreads();
    return out;
}

// rows with queries -> uniq[], ustart[] (start of the row's queries in qlist), cursor[row] (scatter position)
__global__ void __launch_bounds__(kPlanThreads)
plan_rows_kernel(const int *__restrict__ cnt, long long N, int *__restrict__ uniq, int *__restrict__ ustart,
                 int *__restrict__ cursor, long long *__restrict__ out_U /* [0] = U, [1] = H' */) {
    __shared__ int s_warp[32];
    __shared__ int s_tot;
    int base_u = 0, base_q = 0;
    for (long long r0 = 0; r0 < N; r0 += kPlanThreads) {
        const long long r = r0 + threadIdx.x;
        const int c = (r < N) ? cnt[r + 1] : 0;
        const int pu = plan_block_scan(c > 0 ? 1 : 0, s_warp, &s_tot);
        const int tot_u = s_tot;
        const int pq = plan_block_scan(c, s_warp, &s_tot);
        const int tot_q = s_tot;
        if (c > 0) {
            uniq[base_u + pu] = (int)r;
            ustart[base_u + pu] = base_q + pq;
            cursor[r] = base_q + pq;
        }
        base_u += tot_u;
        base_q += tot_q;
    }
    if (threadIdx.x == 0) { ustart[base_u] = base_q; out_U[0] = base_u; out_U[1] = base_q; }
}

__global__ void plan_scatter_kernel(const int *__restrict__ hit_rows, long long H, int *__restrict__ cursor,
                                    int *__restrict__ qlist, int world, int rank) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= H) return;
    if (world > 1 && hit_rows[q] % world != rank) return;
    qlist[atomicAdd(&cursor[hit_rows[q]], 1)] = (int)q;
}

// items of unique row u: runs of its queries with at most Tcap statistics and kPlanQcap queries.
// WRITE == false: number of items of every row; WRITE == true: the items at item_off[u].
template <bool WRITE>
__global__ void plan_items_kernel(const int *__restrict__ ustart, const int *__restrict__ qlist,
                                  const int *__restrict__ q_perm, const long long *__restrict__ perm_off, int U, int Tcap,
                                  long long Us, int *__restrict__ n_items, const long long *__restrict__ item_off,
                                  DedupItem *__restrict__ items, int *__restrict__ qlist_rw) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= U) return;
    int qb = ustart[u];
    const int qend = ustart[u + 1];
    if (!WRITE && qend - qb <= 64) {
        // the scatter left the row's queries in arrival order: ascending order makes the plan (and the kernels'
        // memory access pattern) reproducible.  Rows have a handful of queries; longer ones stay as they are.
        for (int i = qb + 1; i < qend; ++i) {
            const int v = qlist_rw[i];
            int j = i - 1;
            while (j >= qb && qlist_rw[j] > v) { qlist_rw[j + 1] = qlist_rw[j]; --j; }
            qlist_rw[j + 1] = v;
        }
    }
    int n = 0;
    long long w = WRITE ? item_off[u] : 0;
    while (qb < qend) {
        int qe = qb, T = 0;
        while (qe < qend && qe - qb < kPlanQcap) {
            const int k = q_perm[qlist[qe]];
            const int P = (int)(perm_off[k + 1] - perm_off[k]);
            if (qe > qb && T + P > Tcap) break;
            T += P;
            ++qe;
        }
        if (WRITE) items[w++] = DedupItem{(int)(u % Us), qb, qe, T};
        ++n;
        qb = qe;
    }
    if (!WRITE) n_items[u] = n;
}

// exclusive scan of n_items[0..U) into item_off[0..U] (single CTA)
__global__ void __launch_bounds__(kPlanThreads)
plan_item_scan_kernel(const int *__restrict__ n_items, int U, long long *__restrict__ item_off) {
    __shared__ int s_warp[32];
    __shared__ int s_tot;
    long long base = 0;
    for (int u0 = 0; u0 < U; u0 += kPlanThreads) {
        const int u = u0 + threadIdx.x;
        const int v = (u < U) ? n_items[u] : 0;
        const int p = plan_block_scan(v, s_warp, &s_tot);
        if (u < U) item_off[u] = base + p;
        base += s_tot;
    }
    if (threadIdx.x == 0) item_off[U] = base;
}

// first item of every strip (strips of Us unique rows) + the total, for the host: bounds[s] = item_off[min(s * Us, U)]
__global__ void plan_bounds_kernel(const long long *__restrict__ item_off, int U, long long Us, int n_strips,
                                   long long *__restrict__ bounds) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_strips) return;
    const long long u = (long long)s * Us;
    bounds[s] = item_off[u < U ? u : U];
}

}  // namespace coex
