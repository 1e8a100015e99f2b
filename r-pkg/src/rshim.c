/*
 * rshim.c -- .Call shim of the cellulaB200 R package.
 *
 * SEXP unpacking / packing only; every entry forwards to the plain C ABI declared in
 * include/cellula_b200.h.  No arithmetic, no Rcpp.  Rf_error() is only reached after the
 * library call has returned (the C ABI never throws or longjmps), so no C++ frame that owns
 * resources is ever skipped.  Not compiled in the build image (no R there); see
 * INTEGRATION.md.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>

#include "cellula_b200.h"

static void ctx_finalizer(SEXP ptr)
{
    cb200_ctx *ctx = (cb200_ctx *)R_ExternalPtrAddr(ptr);
    if (ctx != NULL) {
        cb200_free(ctx);
        R_ClearExternalPtr(ptr);
    }
}

static cb200_ctx *get_ctx(SEXP ptr)
{
    cb200_ctx *ctx = (cb200_ctx *)R_ExternalPtrAddr(ptr);
    if (ctx == NULL)
        Rf_error("cellulaB200: the device context has been released");
    return ctx;
}

#define CB_OK(ctx, call)                                               \
    do {                                                               \
        if ((call) != 0)                                               \
            Rf_error("%s", cb200_last_error(ctx));                     \
    } while (0)

SEXP cb200r_init(SEXP device)
{
    cb200_ctx *ctx = NULL;
    if (cb200_init(Rf_asInteger(device), NULL, &ctx) != 0)
        Rf_error("%s", cb200_last_error(NULL));
    SEXP ptr = PROTECT(R_MakeExternalPtr(ctx, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* counts: a dgCMatrix.  Its slots are read in place (never copied, never written). */
SEXP cb200r_load_counts(SEXP sctx, SEXP counts)
{
    cb200_ctx *ctx = get_ctx(sctx);
    SEXP p = R_do_slot(counts, Rf_install("p")), i = R_do_slot(counts, Rf_install("i"));
    SEXP x = R_do_slot(counts, Rf_install("x")), dim = R_do_slot(counts, Rf_install("Dim"));
    int64_t n_genes = INTEGER(dim)[0], n_cells = INTEGER(dim)[1];
    CB_OK(ctx, cb200_load_csc_i32(ctx, n_genes, n_cells, INTEGER(p), INTEGER(i), REAL(x), 0, n_cells));
    return R_NilValue;
}

/* sf_in: NULL (library-size factors) or a numeric vector that is only re-centred. */
SEXP cb200r_size_factors(SEXP sctx, SEXP sf_in, SEXP n_cells)
{
    cb200_ctx *ctx = get_ctx(sctx);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(n_cells)));
    int rc = cb200_size_factors(ctx, Rf_isNull(sf_in) ? NULL : REAL(sf_in), REAL(out));
    UNPROTECT(1);
    if (rc != 0)
        Rf_error("%s", cb200_last_error(ctx));
    return out;
}

/* returns list(x = logcounts@x, mean = , var = ) */
SEXP cb200r_lognorm_genestats(SEXP sctx, SEXP nnz, SEXP n_genes)
{
    cb200_ctx *ctx = get_ctx(sctx);
    SEXP lx = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)Rf_asReal(nnz)));
    SEXP mean = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(n_genes)));
    SEXP var = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(n_genes)));
    int rc = cb200_lognorm_genestats(ctx, REAL(lx), REAL(mean), REAL(var));
    if (rc != 0) {
        UNPROTECT(3);
        Rf_error("%s", cb200_last_error(ctx));
    }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3)), nm = PROTECT(Rf_allocVector(STRSXP, 3));
    SET_VECTOR_ELT(out, 0, lx);
    SET_VECTOR_ELT(out, 1, mean);
    SET_VECTOR_ELT(out, 2, var);
    SET_STRING_ELT(nm, 0, Rf_mkChar("x"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("mean"));
    SET_STRING_ELT(nm, 2, Rf_mkChar("var"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(5);
    return out;
}

/* hvg_idx: 0-based integer rows in getTopHVGs order.  returns list(scores, rotation, varExplained, totalVar) */
SEXP cb200r_pca(SEXP sctx, SEXP hvg_idx, SEXP ndims, SEXP n_cells)
{
    cb200_ctx *ctx = get_ctx(sctx);
    int d = Rf_asInteger(ndims), h = LENGTH(hvg_idx), n = Rf_asInteger(n_cells);
    SEXP scores = PROTECT(Rf_allocMatrix(REALSXP, n, d));
    SEXP rot = PROTECT(Rf_allocMatrix(REALSXP, h, d));
    SEXP ve = PROTECT(Rf_allocVector(REALSXP, d));
    SEXP tv = PROTECT(Rf_allocVector(REALSXP, 1));
    int rc = cb200_pca(ctx, INTEGER(hvg_idx), h, d, REAL(scores), REAL(rot), REAL(ve), REAL(tv));
    if (rc != 0) {
        UNPROTECT(4);
        Rf_error("%s", cb200_last_error(ctx));
    }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 4)), nm = PROTECT(Rf_allocVector(STRSXP, 4));
    SET_VECTOR_ELT(out, 0, scores);
    SET_VECTOR_ELT(out, 1, rot);
    SET_VECTOR_ELT(out, 2, ve);
    SET_VECTOR_ELT(out, 3, tv);
    SET_STRING_ELT(nm, 0, Rf_mkChar("scores"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("rotation"));
    SET_STRING_ELT(nm, 2, Rf_mkChar("varExplained"));
    SET_STRING_ELT(nm, 3, Rf_mkChar("totalVar"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(6);
    return out;
}

/* space: numeric matrix N x d (column-major, as R stores it).  Returns list(from, to, weight), 1-based,
 * in libscran's emission order -- exactly what neighborsToSNNGraph hands to igraph::make_graph. */
SEXP cb200r_snn_graph(SEXP sctx, SEXP space, SEXP k, SEXP type)
{
    cb200_ctx *ctx = get_ctx(sctx);
    int64_t n = Rf_nrows(space);
    int d = Rf_ncols(space), kk = Rf_asInteger(k);
    CB_OK(ctx, cb200_knn(ctx, REAL(space), n, d, kk, 0, n, NULL));
    int64_t ne = 0;
    /* count, size the R vectors, emit straight into them (copies overlap the emission) */
    CB_OK(ctx, cb200_snn_count_device(ctx, NULL, n, kk, Rf_asInteger(type), 0, n, &ne));
    SEXP sf = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)ne)), st = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)ne));
    SEXP sw = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)ne));
    if (cb200_snn_emit(ctx, INTEGER(sf), INTEGER(st), REAL(sw)) != 0) {
        UNPROTECT(3);
        Rf_error("%s", cb200_last_error(ctx));
    }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3)), nm = PROTECT(Rf_allocVector(STRSXP, 3));
    SET_VECTOR_ELT(out, 0, sf);
    SET_VECTOR_ELT(out, 1, st);
    SET_VECTOR_ELT(out, 2, sw);
    SET_STRING_ELT(nm, 0, Rf_mkChar("from"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("to"));
    SET_STRING_ELT(nm, 2, Rf_mkChar("weight"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(5);
    return out;
}

/* findKNN()$index equivalent: integer matrix N x k, 1-based */
SEXP cb200r_knn(SEXP sctx, SEXP space, SEXP k)
{
    cb200_ctx *ctx = get_ctx(sctx);
    int64_t n = Rf_nrows(space);
    int d = Rf_ncols(space), kk = Rf_asInteger(k);
    SEXP idx = PROTECT(Rf_allocMatrix(INTSXP, (int)n, kk));
    int rc = cb200_knn(ctx, REAL(space), n, d, kk, 0, n, INTEGER(idx));
    UNPROTECT(1);
    if (rc != 0)
        Rf_error("%s", cb200_last_error(ctx));
    return idx;
}

/* bluster::approxSilhouette(x, clusters): codes = 0-based position of every cell's label in
 * sort(unique(clusters)).  Returns list(width, other) (other: 0-based code of the closest other cluster). */
SEXP cb200r_approx_silhouette(SEXP sctx, SEXP space, SEXP codes, SEXP nclust)
{
    cb200_ctx *ctx = get_ctx(sctx);
    int64_t n = Rf_nrows(space);
    int d = Rf_ncols(space);
    SEXP width = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)n)), other = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)n));
    int rc = cb200_approx_silhouette(ctx, REAL(space), n, d, INTEGER(codes), Rf_asInteger(nclust), REAL(width),
                                     INTEGER(other));
    if (rc != 0) {
        UNPROTECT(2);
        Rf_error("%s", cb200_last_error(ctx));
    }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2)), nm = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_VECTOR_ELT(out, 0, width);
    SET_VECTOR_ELT(out, 1, other);
    SET_STRING_ELT(nm, 0, Rf_mkChar("width"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("other"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(4);
    return out;
}

/* bluster::pairwiseModularity(g, clusters, get.weights = TRUE): edges = as_edgelist(g) + E(g)$weight.
 * Returns list(observed, expected, total); the K x K matrices are filled row-major by the library, so they
 * are transposed into R's column-major layout here (lower/upper triangle swap handled by t() in R). */
SEXP cb200r_pairwise_modularity(SEXP sctx, SEXP from, SEXP to, SEXP weight, SEXP codes, SEXP nclust)
{
    cb200_ctx *ctx = get_ctx(sctx);
    int K = Rf_asInteger(nclust);
    SEXP obs = PROTECT(Rf_allocMatrix(REALSXP, K, K)), expd = PROTECT(Rf_allocMatrix(REALSXP, K, K));
    SEXP tot = PROTECT(Rf_allocVector(REALSXP, 1));
    int rc = cb200_pairwise_modularity(ctx, INTEGER(from), INTEGER(to), REAL(weight), (int64_t)XLENGTH(from), INTEGER(codes),
                                       (int64_t)XLENGTH(codes), K, REAL(obs), REAL(expd), REAL(tot));
    if (rc != 0) {
        UNPROTECT(3);
        Rf_error("%s", cb200_last_error(ctx));
    }
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3)), nm = PROTECT(Rf_allocVector(STRSXP, 3));
    SET_VECTOR_ELT(out, 0, obs);
    SET_VECTOR_ELT(out, 1, expd);
    SET_VECTOR_ELT(out, 2, tot);
    SET_STRING_ELT(nm, 0, Rf_mkChar("observed_t"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("expected_t"));
    SET_STRING_ELT(nm, 2, Rf_mkChar("total"));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(5);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"cb200r_init", (DL_FUNC)&cb200r_init, 1},
    {"cb200r_load_counts", (DL_FUNC)&cb200r_load_counts, 2},
    {"cb200r_size_factors", (DL_FUNC)&cb200r_size_factors, 3},
    {"cb200r_lognorm_genestats", (DL_FUNC)&cb200r_lognorm_genestats, 3},
    {"cb200r_pca", (DL_FUNC)&cb200r_pca, 4},
    {"cb200r_snn_graph", (DL_FUNC)&cb200r_snn_graph, 4},
    {"cb200r_knn", (DL_FUNC)&cb200r_knn, 3},
    {"cb200r_approx_silhouette", (DL_FUNC)&cb200r_approx_silhouette, 4},
    {"cb200r_pairwise_modularity", (DL_FUNC)&cb200r_pairwise_modularity, 6},
    {NULL, NULL, 0}};

void R_init_cellulaB200(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
