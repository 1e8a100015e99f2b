/* TEST INFRASTRUCTURE: stub of Rinternals.h (see R.h in this directory). */
#ifndef CB200_RINTERNALS_STUB_H
#define CB200_RINTERNALS_STUB_H
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef void (*R_CFinalizer_t)(SEXP);
extern SEXP R_NilValue, R_NamesSymbol;
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
int *INTEGER(SEXP);
double *REAL(SEXP);
int LENGTH(SEXP);
R_xlen_t XLENGTH(SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_install(const char *);
SEXP Rf_mkChar(const char *);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP R_do_slot(SEXP, SEXP);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
Rboolean Rf_isNull(SEXP);
int Rf_nrows(SEXP);
int Rf_length(SEXP);
int R_IsNA(double);
#define ISNAN(x) ((x) != (x))   /* R_ext/Arith.h: true for NA_real_ and NaN */
int Rf_ncols(SEXP);
void Rf_error(const char *, ...) __attribute__((noreturn));
SEXP R_MakeExternalPtr(void *, SEXP, SEXP);
void *R_ExternalPtrAddr(SEXP);
SEXP R_ExternalPtrProtected(SEXP);
void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#endif
