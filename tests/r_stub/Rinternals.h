/* Minimal stand-in for R's Rinternals.h: just enough declarations to type-check r/b200_shim.c with gcc -fsyntax-only
 * in an image without R (tests/test_abi.py::test_r_shim_typechecks).  Not R, not usable for linking. */
#ifndef R_STUB_RINTERNALS_H
#define R_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
extern SEXP R_NilValue;
extern double R_NaReal;
#define NA_REAL R_NaReal
int TYPEOF(SEXP);
R_xlen_t XLENGTH(SEXP);
int *INTEGER(SEXP);
int *LOGICAL(SEXP);
double *REAL(SEXP);
Rbyte *RAW(SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP Rf_allocMatrix(int, int, int);
SEXP Rf_allocVector(int, R_xlen_t);
SEXP Rf_mkNamed(int, const char **);
SEXP Rf_mkString(const char *);
SEXP Rf_ScalarInteger(int);
SEXP Rf_ScalarReal(double);
SEXP Rf_lengthgets(SEXP, R_xlen_t);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
void Rf_error(const char *, ...) __attribute__((noreturn));
char *R_alloc(size_t, int);
#define allocMatrix Rf_allocMatrix
#define allocVector Rf_allocVector
#define mkNamed Rf_mkNamed
#define mkString Rf_mkString
#define ScalarInteger Rf_ScalarInteger
#define ScalarReal Rf_ScalarReal
#define lengthgets Rf_lengthgets
#define asInteger Rf_asInteger
#define asLogical Rf_asLogical
#define asReal Rf_asReal
#define nrows Rf_nrows
#define ncols Rf_ncols
#endif
