/* Minimal stand-in for R's headers: ONLY for syntax-checking shim.c where R is not installed. */
#ifndef CPIMH_R_STUB_H
#define CPIMH_R_STUB_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
typedef int R_xlen_t;
#define INTSXP 13
#define REALSXP 14
#define RAWSXP 24
#define VECSXP 19
extern SEXP R_NilValue;
double* REAL(SEXP); int* INTEGER(SEXP); unsigned char* RAW(SEXP);
int LENGTH(SEXP); int Rf_nrows(SEXP); int Rf_ncols(SEXP);
int Rf_asInteger(SEXP); double Rf_asReal(SEXP); int Rf_asLogical(SEXP); int Rf_isNull(SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t); SEXP Rf_allocMatrix(unsigned int, int, int); SEXP Rf_alloc3DArray(unsigned int, int, int, int);
SEXP Rf_protect(SEXP); void Rf_unprotect(int);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP); void* R_ExternalPtrAddr(SEXP); void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP); void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, int);
void Rf_error(const char*, ...);
char* R_alloc(size_t, int);
void GetRNGstate(void); void PutRNGstate(void); double unif_rand(void); double norm_rand(void);
int R_registerRoutines(DllInfo*, const void*, const R_CallMethodDef*, const void*, const void*);
int R_useDynamicSymbols(DllInfo*, int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#define TRUE 1
#define FALSE 0
#endif
