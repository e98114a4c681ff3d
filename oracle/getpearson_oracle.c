/*
 * oracle/getpearson_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement ("port") of the reference's pair-list correlation kernel,
 * getPearson (reference coexpression_v2.c:22-129).  It exists only so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg have a
 * checker on the GPU box, where /root/reference is absent.  The product
 * library (coexpression_b200/csrc) never links, loads or calls this file.
 *
 * Parity status: PINNED.  tests/test_oracle.py compares this restatement
 * bit-for-bit against the unmodified reference C file compiled by
 * oracle/Makefile (`make ref`) on seeded inputs, and against the committed
 * vectors in tests/golden/getpearson_*.npz that were generated from that
 * reference build (tests/golden/make_golden.py).
 *
 * Arithmetic contract restated from the reference (each step cites it):
 *   - rows a = data[ia*n .. ia*n+n), b = data[ib*n ..)   coexpression_v2.c:69-76
 *   - sums are accumulated left to right in double          :83-84
 *   - either sum exactly 0.0  ->  result NaN                :89-92
 *   - mean = (left-to-right sum) / n                        :95-101
 *   - xy += (a-ma)*(b-mb); xx += (a-ma)^2; yy += (b-mb)^2,
 *     left to right, each product rounded before the add    :105-110
 *   - r = xy / (sqrt(xx) * sqrt(yy))                        :113
 *   - thread count: OMP_NUM_THREADS, else 4                 :29-33
 *   - indices are C ints, row offset computed in int        :69-75
 * A constant non-zero row gives xx == 0 and r = 0/0 = NaN without a branch.
 */
#include <math.h>
#include <stdlib.h>
#include <omp.h>

static double ordered_sum(const double *row, int n)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += row[i];
    return s;
}

static double one_pair(const double *a, const double *b, int n)
{
    const double sa = ordered_sum(a, n);
    const double sb = ordered_sum(b, n);
    if (sa == 0.0 || sb == 0.0) return NAN;
    const double ma = sa / n;
    const double mb = sb / n;
    double xy = 0.0, xx = 0.0, yy = 0.0;
    for (int i = 0; i < n; ++i) {
        const double da = a[i] - ma;
        const double db = b[i] - mb;
        xy += da * db;
        xx += da * da;
        yy += db * db;
    }
    return xy / (sqrt(xx) * sqrt(yy));
}

/* Same C ABI as the reference symbol (coexpression_v2.c:22-23), under a
 * different name so that both can be loaded into one process. */
void oracle_getPearson(const double *data, double *result, const int *idxa,
                       const int *idxb, int npairs, int n)
{
    if (!getenv("OMP_NUM_THREADS")) omp_set_num_threads(4);
#pragma omp parallel for schedule(static)
    for (int x = 0; x < npairs; ++x) {
        const int ia = idxa[x], ib = idxb[x];
        result[x] = one_pair(data + ia * n, data + ib * n, n);
    }
}

/* 64-bit-indexed variant used by the bench's cpu_baseline on large tables. */
void oracle_getPearson64(const double *data, double *result, const long long *idxa,
                         const long long *idxb, long long npairs, int n)
{
    if (!getenv("OMP_NUM_THREADS")) omp_set_num_threads(4);
#pragma omp parallel for schedule(static)
    for (long long x = 0; x < npairs; ++x)
        result[x] = one_pair(data + idxa[x] * (long long)n, data + idxb[x] * (long long)n, n);
}

int oracle_num_threads(void)
{
    int t = 1;
    if (!getenv("OMP_NUM_THREADS")) omp_set_num_threads(4);
#pragma omp parallel
    {
#pragma omp single
        t = omp_get_num_threads();
    }
    return t;
}
