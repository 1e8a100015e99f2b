/* TEST INFRASTRUCTURE: a stub of the slice of R's C API that r-pkg/src/rshim.c uses, so that the shim can be
 * compile-checked (types, arity, missing declarations) in an image that has no R.  Declarations follow
 * Rinternals.h / R_ext/Rdynload.h of R 4.4; nothing here is ever linked or run. */
#ifndef CB200_R_STUB_H
#define CB200_R_STUB_H
#include <stddef.h>
void *R_chk_calloc(size_t, size_t);
void R_chk_free(void *);
char *R_alloc(size_t, int);
#define R_Calloc(n, t) ((t *)R_chk_calloc((size_t)(n), sizeof(t)))
#define R_Free(p) (R_chk_free((void *)(p)), (p) = NULL)
#endif
