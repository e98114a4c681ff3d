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

constexpr int kNetThreads = 256;

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

__global__ void __launch_bounds__(kNetThreads)
network_kernel(NetArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long off = a.perm_off[k];
    const int P = (int)(a.perm_off[k + 1] - off);
    if (P < 2) {                     // individualscores empty -> no groups (1-make-network.py:333-338)
        if (tid == 0) a.n_groups[k] = 0;
        return;
    }
    int P2 = 1;
    while (P2 < P) P2 <<= 1;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);       // [P2]
    unsigned long long *ekey = keys + a.P2max;                                          // [P2]
    int *order = reinterpret_cast<int *>(ekey + a.P2max);                               // [P2]
    int *clend = order + a.P2max;                                                       // [P2]
    int *parent = reinterpret_cast<int *>(ekey);                                        // reuse after clustering
    __shared__ int s_G, s_changed;

    const int *hits = a.hit_rows + off;
    const double *Sc = a.scores + a.sc_off[k];
    int *grp = a.group_of_hit + off;
    int *label = a.group_label + off, *win = a.group_winner + off;
    double *dens = a.group_density + off, *aps = a.hit_score + off;
    int *root = a.w_root + off, *pos = a.w_pos + off;
    double *s1 = a.w_s1 + off, *d0 = a.w_d0 + off, *wown = a.w_w + off;
    unsigned long long *best = a.w_best + off, *cand = a.w_cand + off;

    // ---- (1) sort hits by (chromosome, start) ----------------------------------------------
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
    for (int p = tid; p < P; p += kNetThreads) {
        const int row = hits[order[p]];
        ekey[p] = ((unsigned long long)a.chrom_id[row] << 40) |
                  (unsigned long long)(a.feat_end[row] + a.softsep);
    }
    __syncthreads();
    // ---- (2) merge_ranges: clusters of consecutive sorted hits -----------------------------
    if (tid == 0) {
        unsigned long long stop = ekey[0];
        for (int s = 1; s < P; ++s) {
            if (keys[s] > stop) { clend[s - 1] = -1; stop = ekey[s]; }   // boundary after s-1
            else { clend[s - 1] = 0; stop = max(stop, ekey[s]); }
        }
        int end = P;
        for (int s = P - 1; s >= 0; --s) {
            const bool boundary_after = (s == P - 1) || (clend[s] == -1);
            if (boundary_after) end = s + 1;
            clend[s] = end;
        }
    }
    __syncthreads();
    for (int p = tid; p < P; p += kNetThreads) parent[p] = p;     // indexed by hit position
    __syncthreads();
    // ---- (3) in-cluster pairs -> edges -> union-find (smaller name wins the root) ----------
    for (int s = warp; s < P; s += kNetThreads / 32) {
        const int ha = order[s];
        const int ra = hits[ha];
        for (int t = s + 1; t < clend[s]; ++t) {
            const int hb = order[t];
            const int rb = hits[hb];
            const double r = warp_correlation(a, ra, rb, lane);
            if (lane == 0) {
                const long long idx = lower_bound_dev(a.glist, a.glist_len, r);
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
                        if (atomicCAS(&parent[x], x, y) == x) break;
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int p = tid; p < P; p += kNetThreads) root[p] = uf_find(parent, p);
    __syncthreads();
    // ---- (4) groups numbered in ascending label order --------------------------------------
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
            clend[order[p]] = p;               // group number of a root, indexed by hit position
            atomicAdd(&s_G, 1);
        }
    }
    __syncthreads();
    const int G = s_G;
    for (int p = tid; p < P; p += kNetThreads) grp[p] = clend[root[p]];
    for (int g = tid; g < G; g += kNetThreads) { best[g] = 0ULL; cand[g] = ~0ULL; }
    __syncthreads();
    // ---- (5) pass 1: member score = sum over other-group hits, hit order -------------------
    // One warp owns 32 rows; a 32 x 16 tile of the score matrix is staged through shared memory
    // (coalesced 128-byte row segments in, one lane per row walking left to right out).
    {
        double *tile = reinterpret_cast<double *>(smem_raw) + warp * (32 * 17);   // [32][17] per warp
        int *gcol = reinterpret_cast<int *>(reinterpret_cast<double *>(smem_raw) + (kNetThreads / 32) * (32 * 17)) + warp * 16;
        const int half = lane >> 4, col = lane & 15;
        for (int i0 = warp * 32; i0 < P; i0 += kNetThreads) {
            const int i = i0 + lane;
            const int gi = (i < P) ? grp[i] : -1;
            double s = 0.0;
            for (int j0 = 0; j0 < P; j0 += 16) {
                __syncwarp();
#pragma unroll 4
                for (int r = half; r < 32; r += 2) {
                    const int row = i0 + r, j = j0 + col;
                    tile[r * 17 + col] = (row < P && j < P) ? Sc[(long long)row * P + j] : -1.0;
                }
                if (lane < 16) gcol[lane] = (j0 + lane < P) ? grp[j0 + lane] : gi;
                __syncwarp();
                if (i < P) {
                    const int lim = min(16, P - j0);
                    for (int c = 0; c < lim; ++c) {
                        const double v = tile[lane * 17 + c];
                        if ((j0 + c) != i && v >= a.thr && gcol[c] != gi) s = __dadd_rn(s, v);
                    }
                }
            }
            if (i < P) {
                s1[i] = s;
                atomicMax(&best[gi], (unsigned long long)__double_as_longlong(s));
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < P; i += kNetThreads) {
        const int gi = grp[i];
        if ((unsigned long long)__double_as_longlong(s1[i]) == best[gi])
            atomicMin(&cand[gi], ((unsigned long long)(unsigned)a.name_rank[hits[i]] << 32) | (unsigned)i);
    }
    __syncthreads();
    // pass-1 winners w1 (kept in pos[]) and the working copy win[]
    for (int g = tid; g < G; g += kNetThreads) { const int w = (int)(cand[g] & 0xffffffffULL); pos[g] = w; win[g] = w; }
    __syncthreads();
    // ---- (6) pass 2: Gauss-Seidel over groups in group order -------------------------------
    // The sequential rule "group g is scored against the FINAL winners of groups < g and the
    // pass-1 winners of groups > g" is lower-triangular, so iterating
    //     win_new[g] = argmax_m sum_h Sc[m][ h < g ? win_old[h] : w1[h] ]
    // from win = w1 reaches the sequential result as its unique fixed point; after r rounds
    // the first r groups are final, in practice a handful of rounds suffice.  The sums of the
    // last round are exactly the sequential ones (allpromoterscores).
    const int *w1 = pos;
    for (int round = 0; round <= G; ++round) {
        for (int g = tid; g < G; g += kNetThreads) { best[g] = 0ULL; cand[g] = ~0ULL; }
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int i = tid; i < P; i += kNetThreads) {
            const int gi = grp[i];
            const double *row = Sc + (long long)i * P;
            double s = 0.0;
            int h = 0;
            for (; h + 4 <= gi; h += 4) {                  // groups before gi: current winners
                const double v0 = row[win[h]], v1 = row[win[h + 1]], v2 = row[win[h + 2]], v3 = row[win[h + 3]];
                s = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(s, v0), v1), v2), v3);
            }
            for (; h < gi; ++h) s = __dadd_rn(s, row[win[h]]);
            h = gi + 1;
            for (; h + 4 <= G; h += 4) {                   // groups after gi: pass-1 winners
                const double v0 = row[w1[h]], v1 = row[w1[h + 1]], v2 = row[w1[h + 2]], v3 = row[w1[h + 3]];
                s = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(s, v0), v1), v2), v3);
            }
            for (; h < G; ++h) s = __dadd_rn(s, row[w1[h]]);
            s1[i] = s;
            atomicMax(&best[gi], (unsigned long long)__double_as_longlong(s));
        }
        __syncthreads();
        for (int i = tid; i < P; i += kNetThreads) {
            const int gi = grp[i];
            if ((unsigned long long)__double_as_longlong(s1[i]) == best[gi])
                atomicMin(&cand[gi], ((unsigned long long)(unsigned)a.name_rank[hits[i]] << 32) | (unsigned)i);
        }
        __syncthreads();
        int changed = 0;
        for (int g = tid; g < G; g += kNetThreads) {
            const int w = (int)(cand[g] & 0xffffffffULL);
            if (w != win[g]) { win[g] = w; changed = 1; }
        }
        if (changed) s_changed = 1;
        __syncthreads();
        if (!s_changed) break;
        __syncthreads();
    }
    for (int i = tid; i < P; i += kNetThreads) aps[i] = s1[i];
    __syncthreads();
    // ---- (7) indjoined row sums = density ---------------------------------------------------
    const double gdiv = (double)G;
    for (int g = tid; g < G; g += kNetThreads) {
        const double *row = Sc + (long long)win[g] * P;
        double s = 0.0;
        for (int h = 0; h < G; ++h) {
            if (h == g) continue;
            const double v = row[win[h]];
            if (v >= a.thr) s = __dadd_rn(s, v);
        }
        d0[g] = a.useaverage ? s / gdiv : s;
    }
    __syncthreads();
    if (a.noiter) {
        for (int g = tid; g < G; g += kNetThreads) dens[g] = d0[g];
    } else {
        // ---- (8) iterative removal: order by (density, label) descending -------------------
        for (int g = tid; g < G; g += kNetThreads) {
            const double dg = d0[g];
            int before = 0;                         // groups visited before g
            for (int h = 0; h < G; ++h) {
                const double dh = d0[h];
                before += (dh > dg) || (dh == dg && h > g);   // larger label == larger group number
            }
            root[g] = before;
        }
        __syncthreads();
        for (int g = tid; g < G; g += kNetThreads) {
            const double *row = Sc + (long long)win[g] * P;
            const int pg = root[g];
            double s = 0.0;
            for (int h = 0; h < G; ++h) {
                if (h == g || root[h] < pg) continue;          // h already cut
                const double v = row[win[h]];
                if (v >= a.thr) s = __dadd_rn(s, v);
            }
            wown[g] = a.useaverage ? s / gdiv : s;
        }
        __syncthreads();
        // tie reuse (quirk Q7): a group whose un-whittled density equals its predecessor's
        // inherits the predecessor's whittled value -> value of the head of its run.
        for (int g = tid; g < G; g += kNetThreads) {
            const double dg = d0[g];
            int head = g;
            for (int h = g + 1; h < G; ++h) if (d0[h] == dg) head = h;   // largest label in the run
            // the first visited group compares against lastscore = -1, never equal
            dens[g] = wown[head];
        }
    }
    if (tid == 0) a.n_groups[k] = G;
}

}  // namespace coex
