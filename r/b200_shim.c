/* b200_shim.c -- .Call entry points that bind libbambu_em.so (include/bambu_em.h) into bambu's native library.
 *
 * Sits next to src/RcppExports.cpp: add this file to src/, `PKG_LIBS += -L<dir> -lbambu_em` to src/Makevars
 * (src/Makevars:1) and register the entries below next to _bambu_em_theta / _bambu_emWithL1
 * (src/RcppExports.cpp:48-52).  R and its headers are not installed in the image this repository is developed in, so
 * this file is shipped uncompiled; the C-ABI it calls is what the Python tests exercise (tests/test_abi.py checks
 * that every symbol used here is exported with this signature).
 *
 *   bambu_b200_em_batched   replaces the lapply over gene groups of abundance_quantification()
 *                           (R/bambu-quantify_utilityFunctions.R:379-391): ONE call for all groups of a sample
 *   bambu_b200_em_batched_wire  the same through the wire form v2 (a third of the bytes over PCIe)
 *   bambu_b200_emWithL1     drop-in for .Call("_bambu_emWithL1", ...) (R/RcppExports.R:12-14), one dense group
 *   bambu_b200_quantDT      replaces the body of bambu.quantDT() (R/bambu-quantify.R:53-74)
 *   bambu_b200_last_error   text of the last failing call
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>

#include "bambu_em.h"

static int64_t *as_i64(SEXP x)        /* R has no int64: offsets arrive as doubles (exact below 2^53) */
{
    R_xlen_t n = XLENGTH(x);
    int64_t *p = (int64_t *) R_alloc(n ? n : 1, sizeof(int64_t));
    if (TYPEOF(x) == INTSXP) for (R_xlen_t i = 0; i < n; ++i) p[i] = (int64_t) INTEGER(x)[i];
    else for (R_xlen_t i = 0; i < n; ++i) p[i] = (int64_t) REAL(x)[i];
    return p;
}

static void check(int rc)             /* what END_RCPP does for C++ exceptions (src/RcppExports.cpp:45) */
{
    if (rc == BAMBU_EM_OK) return;
    const char *m = bambu_em_last_error();
    if (!m || !m[0]) m = bambu_quant_last_error();
    Rf_error("libbambu_em (%d): %s", rc, m);
}

SEXP bambu_b200_last_error(void)
{
    const char *m = bambu_em_last_error();
    if (!m || !m[0]) m = bambu_quant_last_error();
    return mkString(m ? m : "");
}

/* arena of one sample (or of many: groups back to back), see include/bambu_em.h */
SEXP bambu_b200_em_batched(SEXP grp_row, SEXP grp_col, SEXP row_nnz, SEXP col_idx, SEXP aval, SEXP nz_flags,
                           SEXP Y, SEXP K, SEXP col_flags, SEXP maxiter, SEXP minvalue, SEXP conv)
{
    const int64_t ng = (int64_t) XLENGTH(grp_row) - 1;
    if (ng < 0) Rf_error("grp_row_ptr is empty");
    int64_t *g = as_i64(grp_row), *c = as_i64(grp_col), *r = as_i64(row_nnz);
    const int64_t rows = g[ng];
    SEXP est = PROTECT(allocMatrix(REALSXP, 3, (int) rows));
    SEXP iters = PROTECT(allocVector(INTSXP, ng)), status = PROTECT(allocVector(INTSXP, ng));
    double *tmp = (double *) R_alloc(3 * rows + 1, sizeof(double));
    int rc = bambu_em_batched(ng, g, c, r, INTEGER(col_idx), REAL(aval), RAW(nz_flags), REAL(Y), REAL(K), RAW(col_flags),
                              asInteger(maxiter), asReal(minvalue), asReal(conv), tmp, INTEGER(iters), INTEGER(status));
    if (rc != BAMBU_EM_OK) { UNPROTECT(3); check(rc); }
    for (int64_t i = 0; i < rows; ++i)            /* library: est[k * rows + i]; the R matrix 3 x rows is column-major */
        for (int k = 0; k < 3; ++k) REAL(est)[3 * i + k] = tmp[k * rows + i];
    const char *names[] = {"est", "iters", "status", ""};
    SEXP out = PROTECT(mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, est); SET_VECTOR_ELT(out, 1, iters); SET_VECTOR_ELT(out, 2, status);
    UNPROTECT(4);
    return out;
}

/* The same call with the wire form v2 on the bus: the arena R built is converted by the library's native packer
 * (bambu_wire_create: dead columns and zero entries dropped, rs precomputed, 16-bit indices, bit flags -- 36 % of the
 * bytes) and quantified by bambu_em_batched_v2.  Same arguments, same result, identical doubles. */
SEXP bambu_b200_em_batched_wire(SEXP grp_row, SEXP grp_col, SEXP row_nnz, SEXP col_idx, SEXP aval, SEXP nz_flags,
                                SEXP Y, SEXP K, SEXP col_flags, SEXP maxiter, SEXP minvalue, SEXP conv)
{
    const int64_t ng = (int64_t) XLENGTH(grp_row) - 1;
    if (ng < 0) Rf_error("grp_row_ptr is empty");
    int64_t *g = as_i64(grp_row), *c = as_i64(grp_col), *r = as_i64(row_nnz);
    const int64_t rows = g[ng];
    bambu_wire *w = NULL;
    int rc = bambu_wire_create(&w, ng, g, c, r, INTEGER(col_idx), REAL(aval), RAW(nz_flags), REAL(Y), REAL(K), RAW(col_flags));
    if (rc != BAMBU_EM_OK) Rf_error("libbambu_em (%d): %s", rc, bambu_wire_last_error());
    const int64_t *wg, *wc, *we;
    const uint32_t *row_end, *eq_bits, *multi_bits;
    const double *rs, *wav;
    const void *cidx, *colY, *colK;
    int32_t col_mode = 0;
    bambu_wire_sizes(w, NULL, NULL, NULL, NULL, NULL, &col_mode, NULL);
    bambu_wire_arrays(w, &wg, &wc, &we, &row_end, &rs, &wav, &cidx, &eq_bits, &colY, &colK, &multi_bits, NULL);
    SEXP est = PROTECT(allocMatrix(REALSXP, 3, (int) rows));
    SEXP iters = PROTECT(allocVector(INTSXP, ng)), status = PROTECT(allocVector(INTSXP, ng));
    double *tmp = (double *) R_alloc(3 * rows + 1, sizeof(double));
    rc = bambu_em_batched_v2(ng, wg, wc, we, row_end, rs, wav, cidx, eq_bits, col_mode, colY, colK, multi_bits,
                             asInteger(maxiter), asReal(minvalue), asReal(conv), tmp, INTEGER(iters), INTEGER(status));
    bambu_wire_free(w);
    if (rc != BAMBU_EM_OK) { UNPROTECT(3); check(rc); }
    for (int64_t i = 0; i < rows; ++i)
        for (int k = 0; k < 3; ++k) REAL(est)[3 * i + k] = tmp[k * rows + i];
    const char *names[] = {"est", "iters", "status", ""};
    SEXP out = PROTECT(mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, est); SET_VECTOR_ELT(out, 1, iters); SET_VECTOR_ELT(out, 2, status);
    UNPROTECT(4);
    return out;
}

/* same arguments and result as _bambu_emWithL1 (src/RcppExports.cpp:31-46): List(theta = 3 x M) */
SEXP bambu_b200_emWithL1(SEXP A, SEXP A_full, SEXP A_unique, SEXP Y, SEXP K, SEXP maxiter, SEXP minvalue, SEXP conv)
{
    const int M = nrows(A), N = ncols(A);
    SEXP theta = PROTECT(allocMatrix(REALSXP, 3, M));
    int iters = 0;
    int rc = bambu_emWithL1(REAL(A), REAL(A_full), REAL(A_unique), REAL(Y), REAL(K), M, N, asInteger(maxiter),
                            asReal(minvalue), asReal(conv), REAL(theta), &iters);
    if (rc != BAMBU_EM_OK) { UNPROTECT(1); check(rc); }
    const char *names[] = {"theta", "iter", ""};
    SEXP out = PROTECT(mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, theta);
    SET_VECTOR_ELT(out, 1, ScalarInteger(iters));
    UNPROTECT(2);
    return out;
}

/* the long table of bambu.quantDT (integer / numeric / logical columns as R holds them) -> data for
 * data.table(txid, counts, fullLengthCounts, uniqueCounts) */
SEXP bambu_b200_quantDT(SEXP gene_code, SEXP txid, SEXP eqClassId, SEXP nobs, SEXP equal, SEXP rcWidth, SEXP txlen,
                        SEXP degradation_bias, SEXP maxiter, SEXP minvalue, SEXP conv)
{
    const int n = (int) XLENGTH(txid);
    SEXP t_out = PROTECT(allocVector(INTSXP, n)), c0 = PROTECT(allocVector(REALSXP, n));
    SEXP c1 = PROTECT(allocVector(REALSXP, n)), c2 = PROTECT(allocVector(REALSXP, n));
    int n_result = 0, rc = 0, bias = asLogical(degradation_bias), mi = asInteger(maxiter);
    double mv = asReal(minvalue), cv = asReal(conv), d_rate = NA_REAL;
    bambu_quantDT_dotC(&n, INTEGER(gene_code), INTEGER(txid), INTEGER(eqClassId), REAL(nobs), LOGICAL(equal), REAL(rcWidth),
                       REAL(txlen), &bias, &mi, &mv, &cv, INTEGER(t_out), REAL(c0), REAL(c1), REAL(c2), &n_result, &d_rate, &rc);
    if (rc != BAMBU_EM_OK) { UNPROTECT(4); check(rc); }
    const char *names[] = {"txid", "counts", "fullLengthCounts", "uniqueCounts", "d_rate", ""};
    SEXP out = PROTECT(mkNamed(VECSXP, names));
    SET_VECTOR_ELT(out, 0, lengthgets(t_out, n_result)); SET_VECTOR_ELT(out, 1, lengthgets(c0, n_result));
    SET_VECTOR_ELT(out, 2, lengthgets(c1, n_result)); SET_VECTOR_ELT(out, 3, lengthgets(c2, n_result));
    SET_VECTOR_ELT(out, 4, ScalarReal(d_rate));
    UNPROTECT(5);
    return out;
}

SEXP bambu_b200_shutdown(void)        /* .onUnload (R/bambu-quantify_utilityFunctions.R:505-507) */
{
    bambu_em_shutdown();
    return R_NilValue;
}

/* to be merged into CallEntries (src/RcppExports.cpp:48-52) */
static const R_CallMethodDef b200_CallEntries[] = {
    {"bambu_b200_em_batched", (DL_FUNC) &bambu_b200_em_batched, 12},
    {"bambu_b200_em_batched_wire", (DL_FUNC) &bambu_b200_em_batched_wire, 12},
    {"bambu_b200_emWithL1", (DL_FUNC) &bambu_b200_emWithL1, 8},
    {"bambu_b200_quantDT", (DL_FUNC) &bambu_b200_quantDT, 11},
    {"bambu_b200_last_error", (DL_FUNC) &bambu_b200_last_error, 0},
    {"bambu_b200_shutdown", (DL_FUNC) &bambu_b200_shutdown, 0},
    {NULL, NULL, 0}
};

void bambu_b200_register(DllInfo *dll)     /* call from R_init_bambu (src/RcppExports.cpp:54-57) */
{
    R_registerRoutines(dll, NULL, b200_CallEntries, NULL, NULL);
}
