// network.cuh -- K4: one CTA per permutation turns the P x P edge-score matrix into network
// densities.  Restates, with explicit orders (see oracle/network_oracle.py header):
//   join_nearby + merge_ranges ........ reference coexfunctions.py:277-316, 412-425
//   winners pass 1 / pass 2 ........... reference 1-make-network.py:365-375 / 378-434
//   indjoined ......................... reference 1-make-network.py:437-452
//   indmeasure (density) .............. reference 1-make-network.py:456-462
//   iterative removal ................. reference 1-make-network.py:464-482
// Every floating-point sum is a plain left-to-right double accumulation in the same order the
// oracle uses (hit order for pass 1, group order elsewhere), so ties and the winner choice are
// decided on bit-identical numbers.
#pragma once
#include "common.cuh"
#include "count.cuh"

namespace coex {

constexpr int kNetThreads = 1024;   // net_sort_kernel / net_label_kernel: one CTA per permutation

struct NetArgs {
    int K;
    const long long *perm_off, *sc_off;
    const int *hit_rows;
    const double *scores;
    // table
    long long N;
    int n;
    const double *raw;
    const double *raw_sum, *raw_mean, *raw_xx;
    const uint16_t *rank2;
    const double *sx;
    const unsigned char *rawsum_pos;
    const int *chrom_id, *name_rank;
    const long long *feat_start, *feat_end;
    const double *glist;
    long long glist_len;
    // options
    int measure, anticor, noiter, useaverage;
    double pvtj, thr;
    long long softsep;
    // outputs
    int *n_groups, *group_of_hit, *group_label, *group_winner;
    double *group_density, *hit_score;
    // workspace, all [H]
    int *w_root, *w_pos;
    double *w_s1, *w_d0, *w_w;
    unsigned long long *w_best, *w_cand;
    int *w_gsize;                // [H] members of group g (slot off + g)
    int *w_cursor;               // [H] scratch of the active-row list builder
    int *w_act;                  // [H] hit positions of the members of multi-member groups, grouped by ascending group
    int *w_actn;                 // [K] length of that list
    int *w_front;                // [K] pass-2 frontier: groups <= front are final (see net_commit_kernel)
    int *w_state;                // [K] 1 while the permutation's pass-2 iteration has not converged
    int *w_flags;                // [1] bit 0: some permutation changed in the last committed round
    const long long *warp_off;   // [K+1] prefix of ceil(max(P_k, 1) / 32): warp -> permutation map of the row-sum kernels
    int P2max;
};

__device__ void bitonic_u64(unsigned long long *keys, int *idx, int P2) {
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P2 >> 1); t += blockDim.x) {
                const int x = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int y = x | j;
                const unsigned long long ka = keys[x], kb = keys[y];
                const int ia = idx[x], ib = idx[y];
                const bool gt = (ka > kb) || (ka == kb && ia > ib);
                if (gt == ((x & k) == 0)) {
                    keys[x] = kb; keys[y] = ka;
                    idx[x] = ib; idx[y] = ia;
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int perm_of(const long long *perm_off, int K, long long q) {
    int lo = 0, hi = K;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (perm_off[mid] <= q) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int uf_find(volatile int *parent, int x) {
    int p = parent[x];
    while (p != x) { x = p; p = parent[x]; }
    return x;
}

// get_correlation(pa, pb, ed, cm) of coexfunctions.py:386-409 for one pair, whole warp.
__device__ double warp_correlation(const NetArgs &a, int ra, int rb, int lane) {
    if (!(a.rawsum_pos[ra] && a.rawsum_pos[rb])) return -99.0;
    double r;
    if (a.measure == 0) {
        const uint16_t *pa = a.rank2 + (long long)ra * a.n, *pb = a.rank2 + (long long)rb * a.n;
        long long acc = 0;
        for (int c = lane; c < a.n; c += 32)
            acc += (long long)((int)pa[c] - (a.n + 1)) * ((int)pb[c] - (a.n + 1));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        r = __ddiv_rn(0.25 * (double)acc, __dmul_rn(a.sx[ra], a.sx[rb]));
    } else {
        // in-order Pearson on raw rows, bit-identical to coexpression_v2.c:105-113
        const double *pa = a.raw + (long long)ra * a.n, *pb = a.raw + (long long)rb * a.n;
        const double ma = a.raw_mean[ra], mb = a.raw_mean[rb];
        double xy = 0.0;
        for (int c0 = 0; c0 < a.n; c0 += 32) {
            const int c = c0 + lane;
            double prod = 0.0;
            if (c < a.n) prod = __dmul_rn(__dsub_rn(pa[c], ma), __dsub_rn(pb[c], mb));
            const int lim = min(32, a.n - c0);
            for (int l = 0; l < lim; ++l) xy = __dadd_rn(xy, __shfl_sync(0xffffffffu, prod, l));
        }
        if (a.raw_sum[ra] == 0.0 || a.raw_sum[rb] == 0.0) r = __longlong_as_double(0x7ff8000000000000LL);
        else r = __ddiv_rn(xy, __dmul_rn(sqrt(a.raw_xx[ra]), sqrt(a.raw_xx[rb])));
    }
    if (r != r) r = -99.0;
    return r;
}

// ===========================================================================================
// K4a  join_nearby (coexfunctions.py:277-316) in three kernels, so that the in-cluster correlations -- the expensive part,
//      and very uneven: the real-data hit set is 2-3x larger than a permuted one -- are spread over the whole GPU:
//   net_sort_kernel   one CTA per permutation: hits sorted by (chromosome, start); merge_ranges as two scans
//   net_pairs_kernel  one WARP per sorted hit of ANY permutation: its later cluster partners -> correlation ->
//                     bisect into the global list -> lock-free union-find (global memory), smaller name wins the root
//   net_label_kernel  one CTA per permutation: roots, groups numbered by ascending label
// ===========================================================================================
// merge_ranges (coexfunctions.py:412-425) walks the sorted hits keeping `stop` = the largest end + soft_separation of the
// current cluster and starts a new cluster when start > stop.  Because the hits are sorted by start and every earlier
// cluster ended below the start that closed it, `stop` can be replaced by the prefix maximum over ALL earlier hits:
// boundary after s-1  <=>  key[s] > max(ekey[0..s-1]).  That turns the sequential walk into a prefix-max scan and the
// cluster ends into a suffix-min scan.
__global__ void __launch_bounds__(kNetThreads)
net_sort_kernel(NetArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = blockIdx.x;
    const int tid = threadIdx.x;
    const long long off = a.perm_off[k];
    const int P = (int)(a.perm_off[k + 1] - off);
    if (P < 2) {                     // individualscores empty -> no groups (1-make-network.py:333-338)
        if (tid == 0) { a.n_groups[k] = 0; a.w_state[k] = 0; }
        return;
    }
    int P2 = 1;
    while (P2 < P) P2 <<= 1;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);       // [P2]
    unsigned long long *pm = keys + a.P2max;                                            // [P2] prefix max of ekey
    int *order = reinterpret_cast<int *>(pm + a.P2max);                                 // [P2]
    int *nb = order + a.P2max;                                                          // [P2] next boundary
    constexpr int kPer = 8192 / kNetThreads;                                            // scan values per thread (P2 <= 8192)
    const int *hits = a.hit_rows + off;

    for (int p = tid; p < P2; p += kNetThreads) {
        if (p < P) {
            const int row = hits[p];
            keys[p] = ((unsigned long long)a.chrom_id[row] << 40) | (unsigned long long)a.feat_start[row];
            order[p] = p;
        } else {
            keys[p] = ~0ULL;
            order[p] = 0x7fffffff;
        }
    }
    __syncthreads();
    bitonic_u64(keys, order, P2);
    for (int p = tid; p < P2; p += kNetThreads) {
        unsigned long long e = 0;
        if (p < P) {
            const int row = hits[order[p]];
            e = ((unsigned long long)a.chrom_id[row] << 40) | (unsigned long long)(a.feat_end[row] + a.softsep);
        }
        pm[p] = e;
    }
    __syncthreads();
    // inclusive prefix maximum (Hillis-Steele in place: read, barrier, write, barrier)
    for (int o = 1; o < P2; o <<= 1) {
        unsigned long long v[kPer];
#pragma unroll
        for (int e = 0; e < kPer; ++e) {
            const int p = tid + e * kNetThreads;
            v[e] = (p < P2) ? ((p >= o) ? max(pm[p], pm[p - o]) : pm[p]) : 0ULL;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < kPer; ++e) { const int p = tid + e * kNetThreads; if (p < P2) pm[p] = v[e]; }
        __syncthreads();
    }
    // boundary after s  <=>  s == P-1 or key[s+1] > prefix max up to s;  nb[s] = s if boundary else "infinity"
    for (int p = tid; p < P2; p += kNetThreads)
        nb[p] = (p < P && (p == P - 1 || keys[p + 1] > pm[p])) ? p : 0x7fffffff;
    __syncthreads();
    // suffix minimum: first boundary at or after s
    for (int o = 1; o < P2; o <<= 1) {
        int v[kPer];
#pragma unroll
        for (int e = 0; e < kPer; ++e) {
            const int p = tid + e * kNetThreads;
            v[e] = (p < P2) ? ((p + o < P2) ? min(nb[p], nb[p + o]) : nb[p]) : 0;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < kPer; ++e) { const int p = tid + e * kNetThreads; if (p < P2) nb[p] = v[e]; }
        __syncthreads();
    }
    int *ns = nb;
    for (int p = tid; p < P; p += kNetThreads) {
        a.w_act[off + p] = order[p];                  // sorted position -> hit position
        a.w_cursor[off + p] = ns[p] + 1;              // end (exclusive) of the cluster of sorted position p
        a.w_gsize[off + p] = p;                       // union-find parent, indexed by hit position
    }
}

__global__ void __launch_bounds__(256)
net_pairs_kernel(NetArgs a, long long H) {
    const int lane = threadIdx.x & 31;
    const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // one warp per sorted hit
    if (q >= H) return;
    const int k = perm_of(a.perm_off, a.K, q);
    const long long off = a.perm_off[k];
    const int P = (int)(a.perm_off[k + 1] - off);
    if (P < 2) return;
    const int s = (int)(q - off);
    const int cend = a.w_cursor[off + s];
    if (cend <= s + 1) return;                        // last hit of its cluster
    const int *hits = a.hit_rows + off;
    const int *order = a.w_act + off;
    volatile int *parent = a.w_gsize + off;
    const int ha = order[s];
    const int ra = hits[ha];
    // the warp walks its partners t in batches of 32: the correlations of the batch are formed cooperatively, then
    // every lane bisects its own pair into the global list (32 independent chains) and links the two trees
    for (int t0 = s + 1; t0 < cend; t0 += 32) {
        const int nbat = min(32, cend - t0);
        double rmine = 0.0;
        for (int b = 0; b < nbat; ++b) {
            const double r = warp_correlation(a, ra, hits[order[t0 + b]], lane);
            if (lane == b) rmine = r;
        }
        if (lane < nbat) {
            const int hb = order[t0 + lane];
            const long long idx = lower_bound_dev(a.glist, a.glist_len, rmine);
            const double p1 = 1.0 - (double)idx / (double)a.glist_len;
            double p = p1;
            if (a.anticor) p = fmin(p1, (double)idx / (double)a.glist_len);
            if (p < a.pvtj) {
                int x = ha, y = hb;
                while (true) {
                    x = uf_find(parent, x);
                    y = uf_find(parent, y);
                    if (x == y) break;
                    // attach the root with the larger name under the smaller one
                    if (a.name_rank[hits[x]] < a.name_rank[hits[y]]) { const int tmp = x; x = y; y = tmp; }
                    if (atomicCAS(const_cast<int *>(&parent[x]), x, y) == x) break;
                }
            }
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(kNetThreads)
net_label_kernel(NetArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = blockIdx.x;
    const int tid = threadIdx.x;
    const long long off = a.perm_off[k];
    const int P = (int)(a.perm_off[k + 1] - off);
    if (P < 2) return;
    int P2 = 1;
    while (P2 < P) P2 <<= 1;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);       // [P2]
    int *order = reinterpret_cast<int *>(keys + a.P2max);                               // [P2]
    int *gnum = order + a.P2max;                                                        // [P2] group number of a root
    __shared__ int s_G;
    const int *hits = a.hit_rows + off;
    int *grp = a.group_of_hit + off;
    int *label = a.group_label + off;
    int *root = a.w_root + off;
    volatile int *parent = a.w_gsize + off;
    for (int p = tid; p < P; p += kNetThreads) root[p] = uf_find(parent, p);
    __syncthreads();
    // groups numbered in ascending label order (label = smallest member name = the root)
    for (int p = tid; p < P2; p += kNetThreads) {
        const bool is_root = (p < P) && (root[p] == p);
        keys[p] = is_root ? (unsigned long long)(unsigned)a.name_rank[hits[p]] : ~0ULL;
        order[p] = (p < P) ? p : 0x7fffffff;
    }
    if (tid == 0) s_G = 0;
    __syncthreads();
    bitonic_u64(keys, order, P2);
    for (int p = tid; p < P; p += kNetThreads) {
        if (keys[p] != ~0ULL) {
            label[p] = order[p];
            gnum[order[p]] = p;                // group number of a root, indexed by hit position
            atomicAdd(&s_G, 1);
        }
    }
    __syncthreads();
    const int G = s_G;
    for (int p = tid; p < P; p += kNetThreads) grp[p] = gnum[root[p]];
    for (int g = tid; g < P; g += kNetThreads) { a.w_best[off + g] = 0ULL; a.w_cand[off + g] = ~0ULL; }
    if (tid == 0) { a.n_groups[k] = G; a.w_state[k] = 1; }
}

// ===========================================================================================
// K4b..  grid-wide kernels: one warp per hit (or per group slot) of ANY permutation, so that a
// large hit set is spread over the whole GPU instead of one SM.  All sums are left-to-right:
// the 32 lanes fetch 32 consecutive addends, lane order is then folded in with shuffles.
// Skipped addends are replaced by +0.0, which leaves a non-negative running sum bit-identical.
// ===========================================================================================
// Ordered row sums, shared by pass 1, pass 2 and the density / removal sums.  A warp owns kRsRows
// consecutive rows of ONE permutation.  For every chunk of 32 columns the warp fetches, row by
// row, the 32 addends (lane = column: coalesced for pass 1, a near-contiguous gather of winner
// columns otherwise) into a 32 x 33 shared-memory tile; then lane l adds row l left to right.
// Skipped addends are stored as +0.0, which leaves a non-negative running sum bit-identical.
//   MODE 0  pass 1  (1-make-network.py:365-375)  columns = all hits, hit order
//   MODE 1  pass 2  (1-make-network.py:378-434)  columns = winners, group order (fixed-point round); rows = the
//           members of multi-member groups behind the permutation's frontier only (a one-member group cannot change
//           its winner, a group at or before the frontier is final)
//   MODE 2  density (1-make-network.py:437-462)  rows = groups (their winner's row)
//   MODE 3  whittled density (1-make-network.py:471-478)
//   MODE 4  pass 2 sums of EVERY hit against the final winners = allpromoterscores (1-make-network.py:424)
constexpr int kRsWarps = 8;
constexpr int kRsRows = 4;          // rows per warp: more warps in flight than 32-row blocks
constexpr int kRsStages = 4;        // register pipeline depth: 32-column chunks of addends in flight per warp

template <int MODE>
__global__ void __launch_bounds__(kRsWarps * 32)
net_rowsum_kernel(NetArgs a) {
    __shared__ double tiles[kRsWarps][kRsRows * 33];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long w = (long long)blockIdx.x * kRsWarps + wib;
    if (w >= a.warp_off[a.K]) return;
    constexpr bool kHitRows = (MODE == 0 || MODE == 1 || MODE == 4);        // rows are hits (else groups)
    constexpr bool kPass2 = (MODE == 1 || MODE == 4);
    const int k = perm_of(a.warp_off, a.K, w);
    if (MODE == 1 && !a.w_state[k]) return;
    const long long off = a.perm_off[k];
    const int P = (int)(a.perm_off[k + 1] - off);
    if (P < 2) return;
    const int G = a.n_groups[k];
    const int nrows = (MODE == 1) ? a.w_actn[k] : (kHitRows ? P : G);
    const int ncols = (MODE == 0) ? P : G;
    const int i0 = (int)(w - a.warp_off[k]) * kRsRows;
    if (i0 >= nrows) return;
    double *tile = tiles[wib];
    const int *grp = a.group_of_hit + off;
    const int *win = a.group_winner + off, *w1 = a.w_pos + off, *vis = a.w_root + off;
    const double *Sc = a.scores + a.sc_off[k];
    const int slot = i0 + lane;                                   // this lane's row slot
    bool mine_ok = (lane < kRsRows) && (slot < nrows);
    const int mine = (MODE == 1) ? (mine_ok ? a.w_act[off + slot] : 0) : slot;   // hit position (or group) of the row
    // row attributes held by lane r for row slot i0 + r
    int my_g = mine_ok ? (kHitRows ? grp[mine] : mine) : -1;                 // group of the row
    if (MODE == 1) {
        if (my_g <= a.w_front[k]) { my_g = -1; mine_ok = false; }            // final already
        if (!__any_sync(0xffffffffu, mine_ok)) return;
    }
    const long long my_base = mine_ok ? (long long)(kHitRows ? mine : win[mine]) * P : 0;
    const int my_vis = (MODE == 3 && mine_ok) ? vis[mine] : 0;
    double s = 0.0;
    // Register pipeline over the 32-column chunks: the kRsRows x 32 addends of kRsStages - 1 chunks are in flight and the
    // column attributes (winner indices) one chunk further ahead, so the critical path per chunk is the 32 dependent
    // additions and not two memory latencies plus them.  (An 8-byte cp.async gather into shared-memory stages was
    // measured 1.9x SLOWER than this, profiles/r01_network.md.)
    struct ColIdx { int col_a, col_b, gcol, vcol; bool ok; };
    auto load_idx = [&](int chunk) -> ColIdx {
        ColIdx ix{0, 0, -2, 0, false};
        const int c = chunk * 32 + lane;                           // this lane's column
        ix.ok = c < ncols;
        if (ix.ok) {
            if (MODE == 0) { ix.col_a = c; ix.gcol = grp[c]; }
            else { ix.col_a = win[c]; ix.col_b = kPass2 ? w1[c] : ix.col_a; ix.gcol = c; if (MODE == 3) ix.vcol = vis[c]; }
        }
        return ix;
    };
    auto load_x = [&](const ColIdx &ix, int chunk, double (&x)[kRsRows]) {
        const int c = chunk * 32 + lane;
#pragma unroll
        for (int r = 0; r < kRsRows; ++r) {
            const int rg = __shfl_sync(0xffffffffu, my_g, r);
            const long long rb = __shfl_sync(0xffffffffu, my_base, r);
            const int rvis = (MODE == 3) ? __shfl_sync(0xffffffffu, my_vis, r) : 0;
            bool use = ix.ok && rg >= 0 && ix.gcol != rg;
            if (MODE == 0) use = use && (c != i0 + r);
            if (MODE == 3) use = use && !(ix.vcol < rvis);
            const int col = (kPass2 && !(ix.gcol < rg)) ? ix.col_b : ix.col_a;
            x[r] = use ? Sc[rb + col] : 0.0;
        }
    };
    auto consume = [&](const double (&x)[kRsRows]) {
        __syncwarp();                                              // the previous chunk's additions have read the tile
#pragma unroll
        for (int r = 0; r < kRsRows; ++r) {
            double v = x[r];
            if (!kPass2 && !(v >= a.thr)) v = 0.0;             // selection threshold (not applied in pass 2)
            tile[r * 33 + lane] = v;
        }
        __syncwarp();
        if (mine_ok) {
            // masked and out-of-range addends are +0.0: a non-negative running sum stays bit-identical
            double v[32];
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) v[cc] = tile[lane * 33 + cc];
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) s = __dadd_rn(s, v[cc]);
        }
    };
    {
        const int nch = (ncols + 31) >> 5;
        double x[kRsStages][kRsRows];
        ColIdx ix_next;
        {
            ColIdx pre[kRsStages];
#pragma unroll
            for (int j = 0; j < kRsStages; ++j) pre[j] = load_idx(j);
#pragma unroll
            for (int j = 0; j < kRsStages - 1; ++j) load_x(pre[j], j, x[j]);
            ix_next = pre[kRsStages - 1];
        }
        for (int base = 0; base < nch; base += kRsStages) {
#pragma unroll
            for (int j = 0; j < kRsStages; ++j) {
                const int i = base + j;
                if (i >= nch) break;                               // warp-uniform
                load_x(ix_next, i + kRsStages - 1, x[(j + kRsStages - 1) % kRsStages]);
                ix_next = load_idx(i + kRsStages);
                consume(x[j]);
            }
        }
    }
    if (mine_ok) {
        if (MODE <= 1) {
            a.w_s1[off + mine] = s;
            atomicMax(&a.w_best[off + my_g], (unsigned long long)__double_as_longlong(s));
        } else if (MODE == 4) {
            a.hit_score[off + mine] = s;
        } else {
            const double d = a.useaverage ? s / (double)G : s;
            if (MODE == 3) a.w_w[off + mine] = d; else a.w_d0[off + mine] = d;
        }
    }
}

// one CTA per permutation: pick and commit the winners.  Pick = argmax with the tie-break of the oracle: among the
// members reaching the group's best sum (w_best, atomicMax of the row-sum kernel), the smallest name.
// first != 0: pass-1 winners (w1 + working copy); it also counts the members of every group and lists the hits of
// the multi-member groups by ascending group (the only rows pass 2 has to iterate on).
// Otherwise one fixed-point round of pass 2.  Pass 2 is sequential in the reference (Gauss-Seidel over groups, quirk
// Q6): W*[g] = f_g(W*[< g], W1[> g]) with W1 the pass-1 winners, a lower-triangular rule.  Iterating
// W^r[g] = f_g(W^(r-1)[< g], W1[> g]) from W^0 = W1 reaches W* as its unique fixed point; if c is the FIRST group whose
// winner moved in round r, every group <= c has unchanged inputs in round r + 1 and is final: front[k] = c, the next
// round only recomputes the groups behind it, and state[k] drops to 0 when nothing moved.
__global__ void __launch_bounds__(256)
net_commit_kernel(NetArgs a, int first) {
    const int k = blockIdx.x;
    if (!a.w_state[k]) return;
    const long long off = a.perm_off[k];
    const int P = (int)(a.perm_off[k + 1] - off);
    if (P < 2) { if (threadIdx.x == 0) a.w_state[k] = 0; return; }
    const int G = a.n_groups[k];
    int *win = a.group_winner + off, *w1 = a.w_pos + off;
    int *gsize = a.w_gsize + off, *cursor = a.w_cursor + off;
    const int *grp = a.group_of_hit + off;
    __shared__ int s_part[256];
    __shared__ int s_first;
    {   // pick: pass 1 looks at every hit, a pass-2 round at the rows it recomputed (multi-member groups behind the frontier)
        const int nrows = first ? P : a.w_actn[k];
        const int front0 = first ? -1 : a.w_front[k];
        for (int i = threadIdx.x; i < nrows; i += blockDim.x) {
            const int p = first ? i : a.w_act[off + i];
            const int gi = grp[p];
            if (gi <= front0) continue;
            if ((unsigned long long)__double_as_longlong(a.w_s1[off + p]) == a.w_best[off + gi])
                atomicMin(&a.w_cand[off + gi],
                          ((unsigned long long)(unsigned)a.name_rank[a.hit_rows[off + p]] << 32) | (unsigned)p);
        }
        __threadfence_block();
        __syncthreads();
    }
    if (first) {
        for (int g = threadIdx.x; g < G; g += blockDim.x) {
            const int w = (int)(a.w_cand[off + g] & 0xffffffffULL);
            w1[g] = w; win[g] = w;
            a.w_best[off + g] = 0ULL;
            a.w_cand[off + g] = ~0ULL;
            gsize[g] = 0;
        }
        __syncthreads();
        for (int p = threadIdx.x; p < P; p += blockDim.x) atomicAdd(&gsize[grp[p]], 1);
        __syncthreads();
        // exclusive scan over the groups of (members of multi-member groups): contiguous slice per thread
        const int per = (G + blockDim.x - 1) / blockDim.x;
        const int g0 = min(G, (int)threadIdx.x * per), g1 = min(G, g0 + per);
        int local = 0;
        for (int g = g0; g < g1; ++g) local += (gsize[g] >= 2) ? gsize[g] : 0;
        s_part[threadIdx.x] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int t = 0; t < (int)blockDim.x; ++t) { const int v = s_part[t]; s_part[t] = acc; acc += v; }
            a.w_actn[k] = acc;
            a.w_front[k] = -1;
        }
        __syncthreads();
        int base = s_part[threadIdx.x];
        for (int g = g0; g < g1; ++g) { cursor[g] = base; base += (gsize[g] >= 2) ? gsize[g] : 0; }
        __syncthreads();
        for (int p = threadIdx.x; p < P; p += blockDim.x) {
            const int g = grp[p];
            if (gsize[g] >= 2) a.w_act[off + atomicAdd(&cursor[g], 1)] = p;
        }
        return;
    }
    const int front = a.w_front[k];
    if (threadIdx.x == 0) s_first = 0x7fffffff;
    __syncthreads();
    for (int g = front + 1 + (int)threadIdx.x; g < G; g += blockDim.x) {
        if (gsize[g] < 2) continue;
        const int w = (int)(a.w_cand[off + g] & 0xffffffffULL);
        if (w != win[g]) { win[g] = w; atomicMin(&s_first, g); }
        a.w_best[off + g] = 0ULL;
        a.w_cand[off + g] = ~0ULL;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int c = s_first;
        if (c == 0x7fffffff) {
            a.w_state[k] = 0;
        } else {
            a.w_front[k] = c;
            atomicOr(a.w_flags, 1);
        }
    }
}

// iterative removal bookkeeping, one CTA per permutation.
// stage 0: visiting order by (density, label) descending (1-make-network.py:467)
// stage 1: tie reuse, quirk Q7 (1-make-network.py:468-469) -> final densities
__global__ void __launch_bounds__(256)
net_removal_kernel(NetArgs a, int stage) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *d = reinterpret_cast<double *>(smem_raw);
    const int k = blockIdx.x;
    const long long off = a.perm_off[k];
    if (a.perm_off[k + 1] - off < 2) return;
    const int G = a.n_groups[k];
    for (int g = threadIdx.x; g < G; g += blockDim.x) d[g] = a.w_d0[off + g];
    __syncthreads();
    if (a.noiter) {
        if (stage == 1) for (int g = threadIdx.x; g < G; g += blockDim.x) a.group_density[off + g] = d[g];
        return;
    }
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const double dg = d[g];
        if (stage == 0) {
            int before = 0;                            // groups visited before g
            for (int h = 0; h < G; ++h) {
                const double dh = d[h];
                before += (dh > dg) || (dh == dg && h > g);     // larger label == larger group number
            }
            a.w_root[off + g] = before;
        } else {
            int head = g;                              // first visited group of the run of equal densities
            for (int h = g + 1; h < G; ++h) if (d[h] == dg) head = h;
            a.group_density[off + g] = a.w_w[off + head];
        }
    }
}

}  // namespace coex
