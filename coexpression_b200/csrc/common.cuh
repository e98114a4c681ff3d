// common.cuh -- context, error handling and small device helpers shared by every kernel file.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/coexpression_b200.h"

namespace coex {

// ------------------------------------------------------------------------------------------
// errors: thread-local message, int status codes on the batched API
// ------------------------------------------------------------------------------------------
extern thread_local std::string g_last_error;

inline int fail(const char *file, int line, const std::string &msg) {
    char buf[64];
    snprintf(buf, sizeof buf, " (%s:%d)", file, line);
    g_last_error = msg + buf;
    return 1;
}

#define COEX_FAIL(msg) return ::coex::fail(__FILE__, __LINE__, (msg))
#define COEX_CUDA(expr)                                                                         \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return ::coex::fail(__FILE__, __LINE__,                                             \
                                std::string(#expr " -> ") + cudaGetErrorString(_e));            \
    } while (0)
// after a kernel launch: always check the launch; with COEX_DEBUG_SYNC=1 also wait for it, so an
// asynchronous fault is reported at the launch that caused it
#define COEX_KCHECK(stream)                                                                     \
    do {                                                                                        \
        COEX_CUDA(cudaGetLastError());                                                          \
        if (::coex::debug_sync()) COEX_CUDA(cudaStreamSynchronize(stream));                     \
    } while (0)
#define COEX_TRY(expr)                                                                          \
    do {                                                                                        \
        int _s = (expr);                                                                        \
        if (_s) return _s;                                                                      \
    } while (0)

inline bool debug_sync() {
    static const bool on = getenv("COEX_DEBUG_SYNC") != nullptr;
    return on;
}

constexpr int kNumStages = 7;
enum Stage { ST_RANK = 0, ST_ROWSTATS = 1, ST_GATHER = 2, ST_GEMM = 3, ST_COUNT = 4, ST_NETWORK = 5, ST_STATSORT = 6 };

struct StageSpan {
    int stage;
    cudaEvent_t a, b;
};

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;   // elements
    int reserve(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e != cudaSuccess)
            return fail(__FILE__, __LINE__, std::string("cudaMalloc of ") + std::to_string(n * sizeof(T)) +
                                                " bytes -> " + cudaGetErrorString(e));
        cap = n;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace coex

// The opaque handle of include/coexpression_b200.h
struct coex_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;

    // ---- resident table -----------------------------------------------------------------
    long long N = 0;       // rows
    int n = 0;             // samples
    int n_pad = 0;         // samples padded to a multiple of 128 (one SW128 TMA chunk of int8)
    long long N_pad = 0;   // rows padded to a multiple of 256 (GEMM tile)
    const double *raw = nullptr;  // [N, n] device
    coex::DevBuf<double> raw_own;
    bool prepared = false;

    // in-order row statistics of the raw table (reference coexpression_v2.c:74-110 per row)
    coex::DevBuf<double> raw_sum, raw_mean, raw_xx;
    coex::DevBuf<unsigned char> rawsum_pos;   // sum(row) > 0 (coexfunctions.py:389)

    // rank transform: 2*rank (u16), u = 2*rank-(n+1) as two int8 digits, sum u^2, sqrt(sum/4)
    coex::DevBuf<uint16_t> rank2;             // [N, n]
    coex::DevBuf<int8_t> u_hi, u_lo;          // [N_pad, n_pad] each, K-major
    coex::DevBuf<long long> suu;              // [N_pad]
    coex::DevBuf<double> sx;                  // [N_pad]  sqrt(0.25*suu); 0 for padding

    // Pearson packing: z-scores quantised to D base-128 int8 digits, D K-major planes
    coex::DevBuf<int8_t> z_planes;            // [D, N_pad, n_pad]
    coex::DevBuf<unsigned char> z_ok;         // [N_pad] 0 where the reference yields NaN
    coex::DevBuf<float> z_l1;                 // [N_pad] |z|_1 (error bound of the quantisation)
    coex::DevBuf<double> z_inv_s;             // [N_pad] 1 / per-row quantisation scale
    coex::DevBuf<float> z_valid;              // [N_pad] 1 where the Pearson null value of the row is defined, else 0
    coex::DevBuf<int> z_ninv;                 // [1] rows with an undefined Pearson null value
    coex::DevBuf<int8_t> za_planes;           // gathered query rows of the digit planes [D][Us, n_pad]
    coex::DevBuf<double> za_inv_s;            // [Us]
    int z_digits = 0;                         // digits currently packed (0 = none)
    int pearson_digits = 4;                   // requested: 4 (default, |r - ref| ~ 1e-8) or 3 (faster, ~ 1e-6)
    int pearson_mode = 0;                     // 0 = tensor-core digits, 1 = in-order double strip (bit-exact)

    // ---- join_nearby inputs -------------------------------------------------------------
    coex::DevBuf<int> chrom_id, name_rank;
    coex::DevBuf<long long> feat_start, feat_end;
    bool have_features = false;
    // feature index for the locus -> promoter join (permute.cuh): rows sorted by (chromosome id, start)
    coex::DevBuf<unsigned long long> feat_key;
    coex::DevBuf<int> feat_row;
    long long feat_maxlen = 0;
    // ---- ONE job shared by several GPUs (coex_shard_*): this rank's score buffer (exported with CUDA IPC), the peers' buffers
    // mapped into this process, and the per-permutation base pointers of the running call
    int shard_world = 1, shard_rank = 0, shard_phase = 0;      // phase: 0 = plain call, 1 = count stage of my ROWS, 2 = network stage of my PERMUTATIONS
    coex::DevBuf<double> shard_scores;
    std::vector<double *> peer_scores;                         // [world] (own buffer at shard_rank)
    std::vector<void *> peer_mapped;                           // what cudaIpcOpenMemHandle returned (closed in coex_destroy)
    coex::DevBuf<double *> d_perm_scores;
    int rb_k = -1; double *rb_host = nullptr;                  // coex_set_score_readback: block to copy out during the network stage
    cudaStream_t rb_stream = nullptr; cudaEvent_t rb_event = nullptr;
    // coex_set_expression(on_device = 2): the table arrives in row chunks on up_stream; coex_prepare ranks a chunk when its event fires
    static constexpr int kUpChunks = 8;
    cudaStream_t up_stream = nullptr; cudaEvent_t up_event[kUpChunks] = {}, up_start = nullptr;
    int up_pending = 0;                       // chunks in flight (0: the table is resident)
    long long up_rows = 0;                    // rows per chunk
    std::vector<long long> last_sc_off;                        // score-block offsets of the last reduced batch and where its scores
    const double *last_scores = nullptr;                       // live (the context's buffer, or the caller's device buffer)
    std::vector<double> h_score_tab;                           // host-built edge-score table (see spearman_dedup_path)
    long long tab_for_N = -1; int tab_for_anticor = -1;
    std::vector<void *> tab_peer, tab_mapped;                  // [world][9] packed-table buffers of every rank (own at shard_rank) / IPC mappings
    long long tab_N = -1; int tab_n = -1;                      // table shape the buffers were exported for
    std::vector<double *> h_perm_scores;
    // order in which `bedtools window` reports the B records of one A record (permute.cuh): rank of every table row and its inverse
    coex::DevBuf<int> feat_wrank, feat_wrow;
    std::vector<int> h_chrom_id, h_bed_index;
    std::vector<long long> h_feat_start, h_feat_end;
    coex::DevBuf<unsigned char> pm_blob;      // staged inputs of coex_permute_map
    coex::DevBuf<int> pm_counts, pm_hits_tmp;
    coex::DevBuf<long long> pm_off;
    coex::DevBuf<double> glist;
    long long glist_len = 0;

    // ---- workspaces -----------------------------------------------------------------------
    coex::DevBuf<int8_t> a_hi, a_lo;          // gathered query rows [Hs, n_pad]
    coex::DevBuf<int> strip;                  // S strip [Hs, N_pad] int32
    coex::DevBuf<double> strip_f64;           // Pearson null strip [Hs, N]
    coex::DevBuf<unsigned int> rank_strip;    // row-rank mode: number of null values strictly below each value of the strip [Us, N_pad]
    coex::DevBuf<float> strip_f32;            // Pearson null strip of the de-duplicated path [Us, N_pad] (tensor r as float)
    coex::DevBuf<long long> d_perm_off, d_sc_off, d_warp_off;
    coex::DevBuf<int> d_hits, d_ngroups, d_grp, d_label, d_win, w_root, w_pos, w_state;
    coex::DevBuf<int> w_gsize, w_cursor, w_act, w_actn, w_front;   // pass-2 active rows (network.cuh)
    coex::DevBuf<double> d_dens, d_aps, d_scores, w_s1, w_d0, w_w;
    coex::DevBuf<long long> d_counts;
    coex::DevBuf<unsigned long long> w_best, w_cand;
    coex::DevBuf<float> icol;                 // [N_pad] (float)(0.5/sx), 0 for zero-variance rows
    coex::DevBuf<int> d_ndeg;                 // [1]
    coex::DevBuf<int> dd_uniq, dd_qlist, dd_qperm;   // de-duplicated query rows; permutation of every flat hit index
    coex::DevBuf<unsigned char> dd_items;     // DedupItem[]
    coex::DevBuf<int> pl_cnt, pl_cursor, pl_ustart, pl_nitems, pl_err;   // device-side plan (plan.cuh)
    coex::DevBuf<long long> pl_itemoff, pl_bounds, pl_meta;
    // sorted statistics of a strip's work items (count_split.cuh: stat_sort_kernel -> count_dedup_kernel<3>)
    coex::DevBuf<double> st_thr, sxg;
    coex::DevBuf<unsigned int> st_cnt;
    coex::DevBuf<int> st_itemT, dd_work, dd_redo;
    coex::DevBuf<long long> st_itemoff;
    coex::DevBuf<unsigned long long> st_cursor, st_w;
    coex::DevBuf<double> dd_gthr, dd_tab;
    int count_mode = 0;                       // coex_set_count_mode: 0 = de-duplicated (count_fine for long tables / count_dedup, row-rank by cost), 1 = per-query, 2 = sorted statistics (count_dedup<0>), 3 = row-rank forced, 4 = bucket table (count_fine) forced
    long long fine_min_rows = 65536;          // mode 0: tables at least this long use count_fine_kernel (COEX_FINE_MIN_ROWS)
    coex::DevBuf<int> p_idxa, p_idxb;         // coex_pairs staging
    coex::DevBuf<double> p_out;
    long long strip_mb = 4096;                // correlation-strip workspace in MB: fewer, larger GEMM + count launches (C2: 6.15 -> 6.06 ms,
                                              // FANTOM5 shape with 201 hit sets: 222 -> 210 ms; 8192 is slower again)
    int gemm_mode = 0;
    long long gemm_col0 = 0;                  // persistent tcgen05 kernel: first table column to compute (0 except for upper-triangle strips)
    coex::DevBuf<unsigned long long> ap_hist; // all-pairs histogram [nbins + 1] (+ NaN pairs), resident across calls
    int ap_nbins = 0;
    int gemm_bn = 0;                          // tcgen05 kernel: 0 = persistent, 128 x 64 tiles or 128 x 128 for rows of >= 1024 samples (default); -64 / -128 = that persistent kernel forced; 128 / 64 = one tile per CTA

    // ---- bookkeeping ------------------------------------------------------------------------
    long long launches = 0;
    std::vector<coex::StageSpan> spans;
    double stage_ms[coex::kNumStages] = {0};
    long long stage_launches[coex::kNumStages] = {0};
    void *tc = nullptr;                       // tensor-core GEMM state (tensor maps), gemm_tc.cuh
};

namespace coex {

// RAII-free helpers to bracket a stage with events on the context's stream
inline int stage_begin(coex_ctx *c, int stage) {
    StageSpan s;
    s.stage = stage;
    COEX_CUDA(cudaEventCreate(&s.a));
    COEX_CUDA(cudaEventCreate(&s.b));
    COEX_CUDA(cudaEventRecord(s.a, c->stream));
    c->spans.push_back(s);
    return 0;
}
inline int stage_end(coex_ctx *c, long long launches) {
    StageSpan &s = c->spans.back();
    COEX_CUDA(cudaEventRecord(s.b, c->stream));
    c->launches += launches;
    c->stage_launches[s.stage] += launches;
    return 0;
}
inline void stage_reset(coex_ctx *c) {
    for (auto &s : c->spans) {
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    c->spans.clear();
    for (int i = 0; i < kNumStages; ++i) {
        c->stage_ms[i] = 0;
        c->stage_launches[i] = 0;
    }
}
// forget the spans and totals of stages >= first (a batch call keeps the spans of a prepare that nobody collected yet)
inline void stage_reset_from(coex_ctx *c, int first) {
    std::vector<StageSpan> keep;
    for (auto &s : c->spans) {
        if (s.stage >= first) {
            cudaEventDestroy(s.a);
            cudaEventDestroy(s.b);
        } else {
            keep.push_back(s);
        }
    }
    c->spans.swap(keep);
    for (int i = first; i < kNumStages; ++i) {
        c->stage_ms[i] = 0;
        c->stage_launches[i] = 0;
    }
}
inline int stage_collect(coex_ctx *c) {
    COEX_CUDA(cudaStreamSynchronize(c->stream));
    for (auto &s : c->spans) {
        float ms = 0.f;
        COEX_CUDA(cudaEventElapsedTime(&ms, s.a, s.b));
        c->stage_ms[s.stage] += ms;
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    c->spans.clear();
    return 0;
}

__host__ __device__ inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

__device__ __forceinline__ int next_pow2_dev(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}
inline int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

}  // namespace coex
