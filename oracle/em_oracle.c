/*
 * oracle/em_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's per-gene-group EM (bambu 3.3.5):
 *     src/em.cpp:9-71    em_theta
 *     src/em.cpp:77-110  emWithL1
 *     R/bambu-quantify_utilityFunctions.R:401-431  run_parallel / getAMat
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or as
 * the CPU baseline.  Nothing under bambu_b200/ links or calls it.
 *
 * The reference itself cannot be compiled here (src/em.cpp:1 includes
 * RcppArmadillo.h; neither R, Rcpp nor Armadillo is installed), so every
 * Armadillo expression is restated by hand.  The one implementation-defined
 * detail is the summation order inside arma::sum, restated in
 * arma_accumulate() below.  PARITY IS PINNED: tests/test_oracle_golden.py
 * checks this file against the ten estOutput_* tables of R/sysdata.rda
 * (tests/testthat/test_quant.R:12-27) bit for bit and against the bundled
 * chr9 end-to-end assays.
 *
 * Two evaluations of the same arithmetic are provided:
 *   dense  - literal: dense column-major M x N matrices, the same sweeps,
 *            temporaries and transposes the Armadillo expressions create
 *            ("B1", the reference-faithful CPU baseline);
 *   sparse - the same floating-point operations in the same order on the
 *            non-zeros only ("B2", the strong CPU baseline).  Dense zeros only
 *            ever add +0.0, so both give identical doubles; the tests check
 *            that on every case the dense code can hold in memory.
 *
 * Build: make -C oracle   (gcc -O2, no -ffast-math, -ffp-contract=off)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------ */
/* arma::sum over a contiguous run (arrayops::accumulate): two interleaved
 * accumulators, even 0-based positions into the first, odd into the second,
 * result = acc1 + acc2.  Used by every arma::sum at src/em.cpp:33,36,47,50,
 * 98,100-102. */
static double arma_accumulate(const double *p, int64_t n)
{
    double acc1 = 0.0, acc2 = 0.0;
    int64_t j;
    for (j = 1; j < n; j += 2) {
        acc1 += p[j - 1];
        acc2 += p[j];
    }
    if ((j - 1) < n) acc1 += p[j - 1];
    return acc1 + acc2;
}

static double *dalloc(int64_t n)
{
    double *p = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    return p;
}

/* out (n_cols x n_rows, column-major) = in^T, in is n_rows x n_cols column-major */
static void transpose(const double *in, int64_t n_rows, int64_t n_cols, double *out)
{
    for (int64_t c = 0; c < n_cols; ++c)
        for (int64_t r = 0; r < n_rows; ++r)
            out[c + r * n_cols] = in[r + c * n_rows];
}

/* (X.t() * diagmat(theta)).t()  -- src/em.cpp:46,98,100-102.
 * Armadillo materialises X.t() (N x M), scales its column i by theta_i
 * (glue_times_diag), then transposes back.  Element = X_ij * theta_i. */
static void scale_rows_via_transposes(const double *X, const double *theta, int64_t M, int64_t N,
                                      double *tmp_NM, double *out_MN)
{
    transpose(X, M, N, tmp_NM);
    for (int64_t i = 0; i < M; ++i) {
        double t = theta[i];
        double *col = tmp_NM + i * N;
        for (int64_t j = 0; j < N; ++j) col[j] = col[j] * t;
    }
    transpose(tmp_NM, N, M, out_MN);
}

/* ------------------------------------------------------------------------ */
/* em_theta, dense and literal: src/em.cpp:9-71.
 * X is M x N column-major.  Returns iter; theta_out has M entries.
 * keep_trace != 0 additionally allocates the M x maxiter theta_trace matrix
 * (src/em.cpp:23), writes one column per update (:52) and copies the whole
 * matrix once at return, as Rcpp::wrap does for ret["theta_trace"] (:68). */
int oracle_em_theta_dense(const double *X, const double *Y, int M, int N, int maxiter,
                          double minvalue, double conv, int keep_trace, double *theta_out,
                          double *final_delta)
{
    int64_t MN = (int64_t)M * N;
    double *theta = dalloc(M), *before = dalloc(M);
    double *rs = dalloc(M), *c = dalloc(N), *ratio = dalloc(N);
    double *tmp = dalloc(MN), *adjP = dalloc(MN), *lmat = dalloc(MN), *lt = dalloc(MN);
    double *trace = NULL;
    if (keep_trace) trace = dalloc((int64_t)M * maxiter);

    for (int i = 0; i < M; ++i) theta[i] = 1.0;               /* :20-21 */
    if (trace) memcpy(trace, theta, sizeof(double) * M);      /* :25 */

    transpose(X, M, N, tmp);                                  /* :33  sum(Xb.t(),0) */
    for (int i = 0; i < M; ++i) rs[i] = arma_accumulate(tmp + (int64_t)i * N, N);

    int iter = 0;
    double delta = 1.0;                                       /* :38 */
    while (delta > conv && iter < (maxiter - 1)) {            /* :42 */
        iter++;                                               /* :45 */
        memcpy(before, theta, sizeof(double) * M);
        scale_rows_via_transposes(X, theta, M, N, tmp, adjP); /* :46 */
        for (int j = 0; j < N; ++j)                           /* :47 */
            c[j] = arma_accumulate(adjP + (int64_t)j * M, M);
        for (int j = 0; j < N; ++j) ratio[j] = Y[j] / c[j];   /* :48  Y / summed */
        for (int j = 0; j < N; ++j) {                         /* :48  adjP * diagmat(.) */
            double r = ratio[j];
            const double *a = adjP + (int64_t)j * M;
            double *l = lmat + (int64_t)j * M;
            for (int i = 0; i < M; ++i) l[i] = a[i] * r;
        }
        for (int64_t k = 0; k < MN; ++k)                      /* :49  replace(nan,0) */
            if (isnan(lmat[k])) lmat[k] = 0.0;
        transpose(lmat, M, N, lt);                            /* :50  sum(lmat.t(),0)/rs */
        for (int i = 0; i < M; ++i)
            theta[i] = arma_accumulate(lt + (int64_t)i * N, N) / rs[i];
        if (trace) memcpy(trace + (int64_t)iter * M, theta, sizeof(double) * M); /* :52 */
        /* :54-63  max over {i : after_i > minvalue} of |after-before|/after, 0 if empty */
        int any = 0;
        double mx = 0.0;
        for (int i = 0; i < M; ++i) {
            if (theta[i] > minvalue) {
                double d = fabs(theta[i] - before[i]) / theta[i];
                if (!any || d > mx) mx = d;
                any = 1;
            }
        }
        delta = any ? mx : 0.0;
    }
    memcpy(theta_out, theta, sizeof(double) * M);
    if (final_delta) *final_delta = delta;
    if (trace) {                                              /* :68  wrap() copies */
        double *copy = dalloc((int64_t)M * maxiter);
        memcpy(copy, trace, sizeof(double) * (size_t)M * maxiter);
        free(copy);
        free(trace);
    }
    free(theta); free(before); free(rs); free(c); free(ratio);
    free(tmp); free(adjP); free(lmat); free(lt);
    return iter;
}

/* emWithL1, dense and literal: src/em.cpp:77-110.
 * est_out is 3 x M, row k contiguous (est_out[k*M + i]): counts,
 * fullLengthCounts, uniqueCounts.  Returns the iteration count that the
 * reference computes and then discards (src/em.cpp:93-94). */
int oracle_emWithL1_dense(const double *A, const double *A_full, const double *A_unique,
                          const double *Y, const double *K, int M, int N, int maxiter,
                          double minvalue, double conv, int keep_trace, double *est_out,
                          double *final_delta)
{
    int64_t MN = (int64_t)M * N;
    double *theta = dalloc(M);
    int iter = oracle_em_theta_dense(A, Y, M, N, maxiter, minvalue, conv, keep_trace, theta,
                                     final_delta);

    double *tmp = dalloc(MN), *sc = dalloc(MN), *prod = dalloc(MN), *pt = dalloc(MN);
    double *base = dalloc(N);
    scale_rows_via_transposes(A, theta, M, N, tmp, sc);        /* :98 */
    for (int j = 0; j < N; ++j) {
        double cj = arma_accumulate(sc + (int64_t)j * M, M);
        double b = (K[j] * Y[j]) / cj;                         /* K % Y / sum */
        base[j] = isnan(b) ? 0.0 : b;                          /* :99 */
    }
    const double *mats[3] = {A, A_full, A_unique};
    for (int k = 0; k < 3; ++k) {                              /* :100-102 */
        scale_rows_via_transposes(mats[k], theta, M, N, tmp, sc);
        for (int j = 0; j < N; ++j) {
            double b = base[j];
            const double *s = sc + (int64_t)j * M;
            double *p = prod + (int64_t)j * M;
            for (int i = 0; i < M; ++i) p[i] = s[i] * b;
        }
        transpose(prod, M, N, pt);
        for (int i = 0; i < M; ++i)
            est_out[(int64_t)k * M + i] = arma_accumulate(pt + (int64_t)i * N, N);
    }
    free(theta); free(tmp); free(sc); free(prod); free(pt); free(base);
    return iter;
}

/* ------------------------------------------------------------------------ */
/* The same arithmetic on the non-zeros only.
 *
 * Group slice in CSR: rows 0..M-1, rp[M+1] offsets into ci/av (relative to the
 * slice start), ci = column index local to the group, strictly increasing
 * within a row.  A CSC copy is built once.  Every reduction adds its terms in
 * increasing dense index order, term at even dense position -> acc1, odd ->
 * acc2 (the dense zeros in between would add +0.0 and are skipped). */
typedef struct {
    int M, N;
    int64_t nnz;
    const int64_t *rp; /* M+1, absolute offsets into ci/av/fl */
    const int32_t *ci;
    const double *av;
    const uint8_t *fl; /* bit0: equal */
} slice_t;

/* Column sums c_j = sum_i a_ij*theta_i as the dense code sees them.  If P rows
 * carry a non-finite theta, a dense column is NaN as soon as one of those rows
 * has a structural zero in it (0*NaN, 0*Inf); entries that are present compute
 * naturally.   */
static void sparse_colsums(int N, const int64_t *cp, const int32_t *ri, const double *cv,
                           const double *theta, int M, double *c)
{
    int P = 0;
    for (int i = 0; i < M; ++i) if (!isfinite(theta[i])) P++;
    for (int j = 0; j < N; ++j) {
        double a1 = 0.0, a2 = 0.0;
        int cnt = 0;
        for (int64_t k = cp[j]; k < cp[j + 1]; ++k) {
            double t = cv[k] * theta[ri[k]];
            if (!isfinite(theta[ri[k]])) cnt++;
            if (ri[k] & 1) a2 += t; else a1 += t;
        }
        c[j] = (cnt < P) ? NAN : (a1 + a2);
    }
}

static int em_group_sparse(const slice_t *s, const double *Y, const double *K,
                           const uint8_t *col_flags, int maxiter, double minvalue, double conv,
                           double *est0, double *est1, double *est2, int *status_out)
{
    int M = s->M, N = s->N;
    int64_t base0 = s->rp[0], nnz = s->rp[M] - base0;
    /* CSC copy, rows increasing inside each column */
    int64_t *cp = (int64_t *)calloc((size_t)N + 1, sizeof(int64_t));
    int32_t *ri = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    double *cv = dalloc(nnz);
    for (int64_t k = 0; k < nnz; ++k) cp[s->ci[base0 + k] + 1]++;
    for (int j = 0; j < N; ++j) cp[j + 1] += cp[j];
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1));
    memcpy(fill, cp, sizeof(int64_t) * (size_t)N);
    for (int i = 0; i < M; ++i)
        for (int64_t k = s->rp[i]; k < s->rp[i + 1]; ++k) {
            int j = s->ci[k];
            ri[fill[j]] = i;
            cv[fill[j]] = s->av[k];
            fill[j]++;
        }
    free(fill);

    double *theta = dalloc(M), *rs = dalloc(M), *ratio = dalloc(N);
    int status = 0;
    for (int i = 0; i < M; ++i) {                     /* src/em.cpp:33 */
        double a1 = 0.0, a2 = 0.0;
        for (int64_t k = s->rp[i]; k < s->rp[i + 1]; ++k) {
            if (s->ci[k] & 1) a2 += s->av[k]; else a1 += s->av[k];
        }
        rs[i] = a1 + a2;
        theta[i] = 1.0;
    }
    int iter = 0;
    double delta = 1.0;
    while (delta > conv && iter < (maxiter - 1)) {    /* :42 */
        iter++;
        sparse_colsums(N, cp, ri, cv, theta, M, ratio);          /* :46-47 */
        for (int j = 0; j < N; ++j) ratio[j] = Y[j] / ratio[j];  /* :48 */
        int any = 0;
        double mx = 0.0;
        for (int i = 0; i < M; ++i) {
            double a1 = 0.0, a2 = 0.0;
            double th = theta[i];
            for (int64_t k = s->rp[i]; k < s->rp[i + 1]; ++k) {
                double t = (s->av[k] * th) * ratio[s->ci[k]];    /* :46,48 */
                if (isnan(t)) t = 0.0;                           /* :49 */
                if (s->ci[k] & 1) a2 += t; else a1 += t;
            }
            double after = (a1 + a2) / rs[i];                    /* :50 */
            if (after > minvalue) {                              /* :56 */
                double d = fabs(after - th) / after;
                if (!any || d > mx) mx = d;
                any = 1;
            }
            theta[i] = after; /* row i is not read again in this sweep */
        }
        delta = any ? mx : 0.0;                                  /* :58-63 */
    }
    if (iter >= maxiter - 1 && delta > conv) status |= 1;

    /* epilogue, src/em.cpp:98-102 */
    sparse_colsums(N, cp, ri, cv, theta, M, ratio);
    int Q = 0; /* columns whose baseSum is +-Inf: a dense row with a structural
                  zero there gets 0*Inf = NaN in all three estimates */
    for (int j = 0; j < N; ++j) {
        double b = (K[j] * Y[j]) / ratio[j];
        ratio[j] = isnan(b) ? 0.0 : b;
        if (isinf(ratio[j])) Q++;
    }
    for (int i = 0; i < M; ++i) {
        double t1 = 0, t2 = 0, f1 = 0, f2 = 0, u1 = 0, u2 = 0, th = theta[i];
        int cnt = 0;
        int64_t len = s->rp[i + 1] - s->rp[i];
        for (int64_t k = s->rp[i]; k < s->rp[i + 1]; ++k) {
            int j = s->ci[k];
            double b = ratio[j];
            double a = s->av[k];
            double af = (s->fl[k] & 1) ? a : a * 0.0;            /* fullAval = aval*equal */
            double au = (col_flags[j] & 1) ? a * 0.0 : a;        /* uniqueAval = aval*!multi */
            double t = (a * th) * b, f = (af * th) * b, u = (au * th) * b;
            if (isinf(b)) cnt++;
            if (j & 1) { t2 += t; f2 += f; u2 += u; } else { t1 += t; f1 += f; u1 += u; }
        }
        est0[i] = t1 + t2; est1[i] = f1 + f2; est2[i] = u1 + u2;
        if ((!isfinite(th) && len < N) || cnt < Q) { est0[i] = NAN; est1[i] = NAN; est2[i] = NAN; }
        if (!isfinite(est0[i]) || !isfinite(est1[i]) || !isfinite(est2[i])) status |= 2;
    }
    free(cp); free(ri); free(cv); free(theta); free(rs); free(ratio);
    *status_out = status;
    return iter;
}

/* ------------------------------------------------------------------------ */
/* abundance_quantification over a whole arena
 * (R/bambu-quantify_utilityFunctions.R:379-431).  Same arena layout as the
 * product's C-ABI (include/bambu_em.h).
 *   mode 0 = dense literal (B1), mode 1 = dense literal + theta_trace,
 *   mode 2 = sparse (B2).
 * Single-transcript groups take the closed form of :410-414, with R's
 * long-double sum().  n_threads <= 0 -> all cores.  Returns 0. */
int oracle_em_batched(int mode, int n_threads, int64_t n_groups,
                      const int64_t *grp_row_ptr, const int64_t *grp_col_ptr,
                      const int64_t *row_nnz_ptr, const int32_t *col_idx, const double *aval,
                      const uint8_t *nz_flags, const double *Y, const double *K,
                      const uint8_t *col_flags, int maxiter, double minvalue, double conv,
                      double *est /* 3 x total_rows */, int32_t *iters, int32_t *status)
{
    int64_t total_rows = grp_row_ptr[n_groups];
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (int64_t g = 0; g < n_groups; ++g) {
        int64_t r0 = grp_row_ptr[g], c0 = grp_col_ptr[g];
        int M = (int)(grp_row_ptr[g + 1] - r0), N = (int)(grp_col_ptr[g + 1] - c0);
        const double *Yg = Y + c0, *Kg = K + c0;
        const uint8_t *cf = col_flags + c0;
        double *e0 = est + r0, *e1 = est + total_rows + r0, *e2 = est + 2 * total_rows + r0;
        iters[g] = 0;
        status[g] = 0;
        if (M == 0) continue;
        if (M == 1) { /* :410-414  sum(K * n.obs * mat), R sums in long double */
            long double s0 = 0, s1 = 0, s2 = 0;
            int64_t k = row_nnz_ptr[r0], kend = row_nnz_ptr[r0 + 1];
            for (int j = 0; j < N; ++j) {
                double a = 0.0, af = 0.0, au = 0.0;
                if (k < kend && col_idx[k] == j) {
                    a = aval[k];
                    af = (nz_flags[k] & 1) ? a : a * 0.0;
                    au = (cf[j] & 1) ? a * 0.0 : a;
                    ++k;
                }
                double ky = Kg[j] * Yg[j];
                s0 += (long double)(ky * a);
                s1 += (long double)(ky * af);
                s2 += (long double)(ky * au);
            }
            e0[0] = (double)s0; e1[0] = (double)s1; e2[0] = (double)s2;
            continue;
        }
        if (mode == 2) {
            slice_t s = {M, N, 0, row_nnz_ptr + r0, col_idx, aval, nz_flags};
            int st = 0;
            iters[g] = em_group_sparse(&s, Yg, Kg, cf, maxiter, minvalue, conv, e0, e1, e2, &st);
            status[g] = st;
        } else {
            int64_t MN = (int64_t)M * N;
            double *A = (double *)calloc((size_t)MN, sizeof(double));
            double *Af = (double *)calloc((size_t)MN, sizeof(double));
            double *Au = (double *)calloc((size_t)MN, sizeof(double));
            double *eo = dalloc(3 * (int64_t)M);
            for (int i = 0; i < M; ++i)
                for (int64_t k = row_nnz_ptr[r0 + i]; k < row_nnz_ptr[r0 + i + 1]; ++k) {
                    int j = col_idx[k];
                    double a = aval[k];
                    A[i + (int64_t)j * M] = a;
                    Af[i + (int64_t)j * M] = (nz_flags[k] & 1) ? a : a * 0.0;
                    Au[i + (int64_t)j * M] = (cf[j] & 1) ? a * 0.0 : a;
                }
            double fd = 0.0;
            iters[g] = oracle_emWithL1_dense(A, Af, Au, Yg, Kg, M, N, maxiter, minvalue, conv,
                                             mode == 1, eo, &fd);
            for (int i = 0; i < M; ++i) { e0[i] = eo[i]; e1[i] = eo[M + i]; e2[i] = eo[2 * (int64_t)M + i]; }
            int st = 0;
            if (iters[g] >= maxiter - 1 && fd > conv) st |= 1;
            for (int64_t i = 0; i < 3 * (int64_t)M; ++i) if (!isfinite(eo[i])) st |= 2;
            status[g] = st;
            free(A); free(Af); free(Au); free(eo);
        }
    }
    return 0;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
