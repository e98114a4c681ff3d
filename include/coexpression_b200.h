/*
 * coexpression_b200.h -- C ABI of the B200-native drop-in for the network-density hot path
 * of baillielab/coexpression.  The shared library is built as `coexpression_v2.so` (the file
 * name the reference loads with ctypes.cdll.LoadLibrary, reference 1-make-network.py:66-67,
 * 0-prepare-coex.py:162,223).  Plain C types only: pointers, sizes, ints, doubles.
 *
 * Two groups of entry points:
 *   1. the LEGACY symbol `getPearson`, bit-for-bit the reference's signature, for unmodified
 *      reference scripts (call sites 1-make-network.py:229-234 and 0-prepare-coex.py:692-697);
 *   2. a handle-based BATCHED API (new; the reference has no counterpart because it re-uploads
 *      and re-ranks the whole table in every one of its per-permutation processes,
 *      1-make-network.py:100-122) in which the table is made resident once and K permutations
 *      are reduced to network densities per call.
 *
 * Every function of group 2 returns 0 on success and a non-zero code on failure; the message
 * is available from coex_last_error().  There is NO CPU fallback: without a CUDA device every
 * compute entry point fails (group 2) or aborts loudly (getPearson, which has no error channel,
 * reference coexpression_v2.c:22 returns void).
 */
#ifndef COEXPRESSION_B200_H
#define COEXPRESSION_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * 1. Legacy drop-in.  Replaces reference coexpression_v2.c:22-129.
 *    dataAll  : host, double [rows * n] row-major (rows is implied: max index + 1)
 *    resultAll: host, double [nPromPairs], written on return (synchronous)
 *    idxa_all, idxb_all: host, int [nPromPairs], 0-based row numbers
 *    Result per pair is bit-identical to the reference's: left-to-right double sums,
 *    NaN when either row sums to exactly 0.0 (coexpression_v2.c:89-92) or has zero variance.
 * ---------------------------------------------------------------------------------------- */
void getPearson(double *dataAll, double *resultAll, int *idxa_all, int *idxb_all,
                int nPromPairs, int n);

/* ------------------------------------------------------------------------------------------
 * 2. Batched API.
 * ---------------------------------------------------------------------------------------- */
typedef struct coex_ctx coex_ctx;

const char *coex_last_error(void);
int coex_abi_version(void);
int coex_device_count(void);            /* 0 when no CUDA device / driver is present */

int coex_create(coex_ctx **out, int device);
int coex_destroy(coex_ctx *ctx);
int coex_sync(coex_ctx *ctx);
void *coex_stream(coex_ctx *ctx);       /* the cudaStream_t every kernel of ctx is launched on */

/* A3 "matrix marshal" (1-make-network.py:100-122): one bulk copy instead of N*n Python stores.
 * data: double [N * n] row-major in table order.  `on_device` = 1: data is a device pointer (used in place).
 * `on_device` = 2: data is PAGE-LOCKED host memory and the call returns while the copy is in flight -- row chunks on a stream of
 * their own, and coex_prepare ranks every chunk as it lands (the rank transform hides under the PCIe copy); the caller must leave
 * `data` alone until a synchronising call (coex_sync, coex_network_batch, ...) has returned. */
int coex_set_expression(coex_ctx *ctx, const double *data, long long N, int n, int on_device);

/* A2 rank transform (scipy.stats.rankdata 'average', 1-make-network.py:108-111) + per-row
 * statistics + packing for the correlation kernels.  Must follow coex_set_expression. */
int coex_prepare(coex_ctx *ctx);
/* ranks as doubles [N * n] (host), for parity checks against scipy.stats.rankdata */
int coex_get_ranks(coex_ctx *ctx, double *ranks_out);

/* Genomic coordinates of every table row, for join_nearby (coexfunctions.py:277-316):
 * chrom_id orders like the chromosome string, name_rank orders like the row label. */
int coex_set_features(coex_ctx *ctx, const int *chrom_id, const long long *start,
                      const long long *end, const int *name_rank, long long N);
/* Optional: bed_index[row] = line number of the row's feature in the feature BED (`-q`, the B file of `windowBed`).  `bedtools
 * window` reports the B records of one A record by (UCSC-bin level, bin, line of the B file), and that order decides the order of a
 * permutation's promoters (1-make-network.py:131-177, quirk Q1).  Default after coex_set_features: the BED lists the rows in table order. */
int coex_set_feature_order(coex_ctx *ctx, const int *bed_index, long long N);
/* ascending global correlation list (<expr>.spearmanlist, 1-make-network.py:76-79) */
int coex_set_global_list(coex_ctx *ctx, const double *sorted_values, long long len);

/* Row N1: K permuted locus sets and their promoters in one call, replacing the per-permutation
 * Python loops of 0-prepare-coex.py:548-593 (get_gwad / shift / get_address, coexfunctions.py:102-128; backcirc with
 * listitem, :319-325) and the `windowBed -w window` subprocess that follows each of them (0-prepare-coex.py:357,609).
 *   mode 0 = circular, 1 = backcirc, 2 = background (0-prepare-coex.py:564-578); shifts[k] == 0 is the real data
 *   (always the circular arithmetic, as in the reference).  snp_chrom[m] indexes the chromosome list (chromrange order, 0-prepare-coex.py:319-330) whose
 *   genome-wide ranges are chrom_start / chrom_end; chrom_fid[c] is the chrom_id that chromosome has in
 *   coex_set_features (-1 if no feature lies on it).  backcirc: snp_back[m] = index of SNP m in the background
 *   sorted by (genome-wide address, key) (0-prepare-coex.py:582), back_chrom / back_addr its entries.
 *   background: back_chrom / back_addr = column 0 / column 2 of every line of the background file in FILE order and
 *   snp_back[k * M + m] = the line `random.sample(range(n_back), M)` drew for locus m of set k (the draw stays with the
 *   caller: it is the reference's RNG stream); any non-zero shifts[k] marks a permuted set.
 * Output: perm_off[K+1] and hit_rows (table rows, first-appearance order over the loci in input order and, within a locus, the
 * order in which `bedtools window` reports the features: coex_set_feature_order) exactly as coex_network_batch takes them;
 * host arrays, or device arrays when on_device != 0.  hit_loci (optional, int [2 * capacity]): for every hit the first and the
 * last locus (input order) that map to it -- what a map-file writer needs for column 3 (0-prepare-coex.py:396-399 keeps the
 * last windowBed line of a promoter for the real data).  Fails if capacity is too small. */
int coex_permute_map(coex_ctx *ctx, int mode, const int *snp_chrom, const long long *snp_addr,
                     const long long *snp_back, int M, const long long *shifts, int K,
                     const long long *chrom_start, const long long *chrom_end, const int *chrom_fid, int n_chrom,
                     const int *back_chrom, const long long *back_addr, long long n_back, long long window,
                     long long *perm_off, int *hit_rows, long long capacity, int on_device, int *hit_loci);

/* A1 on the resident table.  which: 0 = raw values (Pearson), 1 = ranks (Spearman).
 * idxa/idxb/out are host pointers (on_device == 0) or device pointers. */
int coex_pairs(coex_ctx *ctx, int which, const int *idxa, const int *idxb, long long npairs,
               double *out, int on_device);

/* All-pairs correlation of the resident table reduced on the device (config C3): histogram of
 * r over `nbins` equal bins on [-1, 1] for row pairs (a, b), a in [row0, row1), b > a; NaN pairs
 * counted in nan_count.  which as in coex_pairs.  hist: host, unsigned long long [nbins]. */
int coex_allpairs_histogram(coex_ctx *ctx, int which, long long row0, long long row1, int nbins,
                            unsigned long long *hist, unsigned long long *nan_count);
/* The same reduction with the histogram RESIDENT on the device: row blocks are accumulated call after call (reset != 0
 * clears it first) and fetched once -- to host arrays and / or to a device buffer of nbins + 1 counters (the last one
 * is the NaN-pair count) that a multi-GPU caller hands to its all-reduce. */
int coex_allpairs_accumulate(coex_ctx *ctx, int which, long long row0, long long row1, int nbins, int reset);
int coex_allpairs_read(coex_ctx *ctx, unsigned long long *hist, unsigned long long *nan_count, void *device_copy);

/* A block of the correlation matrix: table rows [row0, row0 + nrows) against every table row.
 * which: 1 = Spearman (exact integer tensor path, bit-identical to the reference on ranks),
 *        0 = Pearson on the tensor cores (z-scores quantised to 3 or 4 int8 digits, exact integer
 *            accumulation; |r - reference| is bounded by the quantisation only),
 *        2 = Pearson in-order double arithmetic (bit-identical to the reference).
 * row0 must be a multiple of 128 for which 0 / 1.  out: host, double [nrows * N], NaN where the
 * reference yields NaN. */
int coex_corr_block(coex_ctx *ctx, int which, long long row0, long long nrows, double *out);
/* Pearson tensor path: digits = 4 (default; measured |r - reference| <= 2e-8) or 3 (9 instead of 16
 * MMAs per K step; measured <= 3e-6); mode 0 = tensor cores, 1 = in-order double
 * (used by coex_allpairs_histogram for which == 0). */
int coex_set_pearson(coex_ctx *ctx, int digits, int mode);

typedef struct coex_params {
    int measure;                /* -cm: 0 Spearman, 1 Pearson (null list + join_nearby; the   */
                                /*      hit-vs-hit statistic is always Spearman, quirk Q2)     */
    int anticorrelation;        /* -a  (1-make-network.py:289-292, coexfunctions.py:305-307)   */
    int noiter;                 /* -i  (1-make-network.py:465)                                 */
    int useaverage;             /* -u  (1-make-network.py:457)                                 */
    int use_specific_p;         /* nsp > 1/N (1-make-network.py:189-193); 0 = global list      */
    int reserved;
    double pval_to_join;        /* -j                                                          */
    double selectionthreshold;  /* -st, or -log10(pval_to_join) under -m                       */
    long long soft_separation;  /* -ss                                                         */
} coex_params;

/* A4-A12 for K permutations in one call (1-make-network.py:198-482 once per permutation).
 *   perm_off : long long [K+1], hits of permutation k are hit_rows[perm_off[k] .. perm_off[k+1])
 *   hit_rows : int [H], table rows in map-file first-appearance order (Q1: the null list of
 *              hit position i omits table column i)
 * Outputs (H = perm_off[K]; per-group arrays use the first n_groups[k] slots of permutation k's
 * segment; groups are numbered in ascending label order):
 *   n_groups        int    [K]
 *   group_of_hit    int    [H]   group number of each hit
 *   group_label     int    [H]   per group: hit position (within the permutation) of its label
 *   group_winner    int    [H]   per group: hit position of its winner (1-make-network.py:431-434)
 *   group_density   double [H]   per group: indmeasure (after iterative removal unless noiter)
 *   hit_score       double [H]   allpromoterscores (1-make-network.py:424)
 *   scores          double [sum P_k^2] or NULL: individualscores, row-major per permutation
 *   counts          long long [sum P_k^2] or NULL: bisect_left indices (-1 where not computed)
 * on_device != 0: every pointer (inputs and outputs) is a device pointer and the call is
 * asynchronous on coex_stream(ctx). */
int coex_network_batch(coex_ctx *ctx, const coex_params *params, int K,
                       const long long *perm_off, const int *hit_rows,
                       int *n_groups, int *group_of_hit, int *group_label, int *group_winner,
                       double *group_density, double *hit_score,
                       double *scores, long long *counts, int on_device);

/* ------------------------------------------------------------------------------------------
 * 3. ONE job shared by the GPUs of a box (one process and one coex_ctx per GPU; the reference's counterpart is runlocal.py's
 *    pool of per-permutation processes, runlocal.py:19,50-55,78-79).
 *    Splitting the PERMUTATIONS over the ranks would repeat the work of the node-specific null: a table row hit by permutations
 *    of several ranks would be correlated against the whole table and streamed on each of them.  So the count stage
 *    (A4 - A7) is sharded by TABLE ROW -- rank r handles the queries of EVERY permutation whose hit row u has u % world == r,
 *    each row still exactly once -- and the network stage (A8 - A12) by permutation.  The edge scores move from the rank that
 *    computes a score row to the rank that owns its permutation inside the count kernel: plain stores into the owner's score
 *    buffer, mapped into this process with CUDA IPC (peer memory over NVLink / NVSwitch).  There is no exchange step and no
 *    collective on the data path; the caller provides one barrier between the two calls (all stores have landed when every
 *    rank's coex_shard_count has returned) and gathers the density vectors afterwards (2-collate-results.py:68-76).
 *
 *    coex_shard_scores_alloc  this rank's score buffer of n_doubles (= sum of P_k^2 over ITS permutations) and its 64-byte
 *                             CUDA IPC handle, to be exchanged by the caller (e.g. torch.distributed.all_gather)
 *    coex_shard_scores_open   maps the buffer of a peer rank
 *    coex_shard_count         K = ALL permutations of the job (host arrays); perm_owner[k] = rank that owns permutation k,
 *                             perm_score_off[k] = offset (in doubles) of its P_k x P_k block inside the owner's buffer
 *    coex_shard_finish        K = THIS rank's permutations, in the order their blocks have in its buffer; outputs as
 *                             coex_network_batch (scores: optional host copy of the blocks; on_device as there)
 * ---------------------------------------------------------------------------------------- */
/* The rank transform of a shared job (optional; every rank may also call coex_prepare on the whole table): rank r ranks and
 * packs ITS block of rows only, and rank_pack_kernel / rowstats_kernel store every result into its own AND into every peer's
 * packed table (nine buffers per rank, mapped with CUDA IPC) -- no rank ranks a row twice, nothing is gathered afterwards.
 *   coex_shard_table_export   after coex_set_expression: sizes the packed table, -> 9 x 64 bytes of IPC handles (at most 8 ranks)
 *   coex_shard_table_open     maps the nine buffers of a peer
 *   coex_shard_prepare_rows   my rows -> everybody (returns when the stores have landed); then ONE barrier by the caller
 *   coex_shard_prepare_finish what needs the whole table (reciprocal norms); the context is then prepared as after coex_prepare */
int coex_shard_table_export(coex_ctx *ctx, int world, int rank, void *ipc_handles_out);
int coex_shard_table_open(coex_ctx *ctx, int peer_rank, const void *ipc_handles);
int coex_shard_prepare_rows(coex_ctx *ctx);
int coex_shard_prepare_finish(coex_ctx *ctx);
int coex_shard_scores_alloc(coex_ctx *ctx, int world, int rank, long long n_doubles, void *ipc_handle_out);
int coex_shard_scores_open(coex_ctx *ctx, int peer_rank, const void *ipc_handle);
int coex_shard_count(coex_ctx *ctx, const coex_params *params, int K, const long long *perm_off, const int *hit_rows,
                     const int *perm_owner, const long long *perm_score_off);
int coex_shard_finish(coex_ctx *ctx, const coex_params *params, int K, const long long *perm_off, const int *hit_rows,
                      int *n_groups, int *group_of_hit, int *group_label, int *group_winner, double *group_density,
                      double *hit_score, double *scores, int on_device);

/* The P_k x P_k edge-score block (individualscores, 1-make-network.py:316-320) of permutation k of the LAST batch this context
 * reduced, copied to the host: a batch of permuted sets only needs its densities back (2-collate-results.py:68-76 reads nothing
 * else of them), the real data (k = 0) also its scores (:79-109).  Valid until the next coex_network_batch / coex_shard_* call. */
int coex_read_scores(coex_ctx *ctx, int k, double *out);
/* The same block copied out BY coex_network_batch itself, on a second stream under the network stage (host_out: pinned memory of
 * P_k^2 doubles for the copy to overlap; filled when coex_network_batch returns).  host_out == NULL switches it off. */
int coex_set_score_readback(coex_ctx *ctx, int k, double *host_out);

/* A14, the consumer of the gathered density vectors (2-collate-results.py:67-76,142,161-163; coexfunctions.py:379-383 empiricalp,
 * :333-358 fdr_correction): pool = every permuted density (device pointer if pool_on_device, e.g. the gathered block), real = the
 * G densities of the real data (host).  p_out = sorted(empiricalp(x, pool) for x in real) (ascending), bh_out / by_out = the
 * Benjamini-Hochberg ('indep') / Benjamini-Yekutieli ('negcorr') adjusted values aligned with p_out; host arrays of G doubles,
 * bit-identical to the reference's arithmetic. */
int coex_collate(coex_ctx *ctx, const double *pool, long long n_pool, int pool_on_device, const double *real, int G,
                 double *p_out, double *bh_out, double *by_out);

/* Introspection for bench.py: kernels launched by ctx so far, and per-stage device time (ms,
 * CUDA events on coex_stream) of the last coex_network_batch / coex_prepare call.
 * stage ids: 0 rank+pack, 1 row stats, 2 gather, 3 correlation GEMM, 4 count+score, 5 network, 6 unused (kept for ABI stability) */
long long coex_launch_count(coex_ctx *ctx);
int coex_stage_ms(coex_ctx *ctx, int stage, double *ms, long long *launches);
/* 0 = tcgen05/TMA tensor-core GEMM (default: persistent kernel; 128 x 64 tiles with double-buffered TMEM accumulators, or
 * 128 x 128 tiles with eight epilogue warps when a row has >= 1024 samples), 1 = SIMT dp4a GEMM (validation path),
 * 2 / 3 = the persistent tcgen05 kernel with 128 x 64 / 128 x 128 tiles explicitly,
 * 5 = 128 x 128 tiles with the B planes multicast inside clusters of two CTAs, 6 = digit-split CTA pairs (tcgen05.mma.cta_group::2:
 * the leader holds the hi digits, its peer the lo digits of the same tile; double-buffered accumulators) -- both exact, both as fast as
 * the default and no faster: the contraction runs at the board's power cap,
 * 64 / 128 = one-tile-per-CTA tcgen05 kernel with that table-column tile */
int coex_set_gemm_mode(coex_ctx *ctx, int mode);
/* 0 = de-duplicated count kernels (default).  Spearman null: the bucket-table kernel of count_fine.cuh for tables of >= 65536 rows
 * (a streamed null value costs one shared-memory atomic and never looks at the statistics; what it declines is re-done by the
 * sorted-statistics kernel), the sorted-statistics kernel of count_dedup.cuh below that; `-cm Pearson`: the sorted-statistics kernel.
 * Mode 0 switches by itself to ROW-RANK mode (every strip row ranked once, the queries look their bisect indices up) when the hit
 * sets are so large that this is cheaper (config C5: thousands of hits per set); 3 forces that mode.
 * 1 = per-query binary-search kernel (count.cuh), 2 = sorted-statistics kernel forced, 4 = bucket-table kernel forced: independent
 * implementations of the same counts, cross-checked against each other and against the oracle by the tests. */
int coex_set_count_mode(coex_ctx *ctx, int mode);
int coex_set_workspace_mb(coex_ctx *ctx, long long strip_mb);
/* On-device cross-check: the first m_rows table rows against the whole table through the tcgen05
 * kernel and through the dp4a kernel; *mismatches = differing int32 dot products,
 * first4 = {row, column, tensor value, dp4a value} of one of them (or -1s). */
int coex_selftest_gemm(coex_ctx *ctx, long long m_rows, long long *mismatches, long long *first4);

#ifdef __cplusplus
}
#endif
#endif /* COEXPRESSION_B200_H */
