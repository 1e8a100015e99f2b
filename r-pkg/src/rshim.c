/*
 * rshim.c -- .Call shim of the cellulaB200 R package.
 *
 * SEXP unpacking / packing only; every entry forwards to the plain C ABI declared in
 * include/cellula_b200.h.  No arithmetic, no Rcpp.  Rf_error() is only reached after the
 * library call has returned (the C ABI never throws or longjmps), so no C++ frame that owns
 * resources is ever skipped.  No R exists in the build image: the file is compile-checked there against a
 * stub of R's C API (tests/helpers/r_stub, tests/test_rshim_compiles.py) and built for real by R CMD INSTALL
 * on an R host; see INTEGRATION.md.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>

#include "cellula_b200.h"

/* One handle for both forms of the library: a single GPU (cb200_ctx) or several GPUs of the node driven from this
 * one R process (cb200_multi: one host thread per GPU inside the library, NCCL between them).  R stays single-threaded:
 * every .Call blocks on the R main thread and no R API is touched from the library's threads. */
typedef struct {
    cb200_ctx *one;
    cb200_multi *multi;
} cb200r_handle;

static void ctx_finalizer(SEXP ptr)
{
    cb200r_handle *h = (cb200r_handle *)R_ExternalPtrAddr(ptr);
    if (h != NULL) {
        if (h->one != NULL)
            cb200_free(h->one);
        if (h->multi != NULL)
            cb200_multi_free(h->multi);
        R_Free(h);
        R_ClearExternalPtr(ptr);
    }
}

static cb200r_handle *get_handle(SEXP ptr)
{
    cb200r_handle *h = (cb200r_handle *)R_ExternalPtrAddr(ptr);
    if (h == NULL)
        Rf_error("cellulaB200: the device context has been released");
    return h;
}

static const char *last_error(const cb200r_handle *h)
{
    return h->multi != NULL ? cb200_multi_last_error(h->multi) : cb200_last_error(h->one);
}

#define CB_OK(h, call)                                                 \
    do {                                                               \
        if ((call) != 0)                                               \
            Rf_error("%s", last_error(h));                             \
    } while (0)

/* device: GPU of the single-GPU form; n_gpus > 1: that many GPUs (0 .. n_gpus-1) from this process */
SEXP cb200r_init(SEXP device, SEXP n_gpus)
{
    cb200r_handle *h = R_Calloc(1, cb200r_handle);
    int n = Rf_asInteger(n_gpus);
    if (n > 1) {
        if (cb200_multi_init(n, &h->multi) != 0) {
            R_Free(h);
            Rf_error("%s", cb200_multi_last_error(NULL));
        }
    } else if (cb200_init(Rf_asInteger(device), NULL, &h->one) != 0) {
        R_Free(h);
        Rf_error("%s", cb200_last_error(NULL));
    }
    SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* counts: a dgCMatrix.  Its slots are read in place (never copied, never written); they are ordinary pageable R
 * memory, which the library stages through its pinned rings (cb200_stage.cu).  The row indices may still be
 * uploading when this returns: `counts` must stay referenced until cb200r_lognorm_begin has returned (the wrapper
 * holds it in a local variable). */
SEXP cb200r_load_counts(SEXP sctx, SEXP counts)
{
    cb200r_handle *h = get_handle(sctx);
    SEXP p = R_do_slot(counts, Rf_install("p")), i = R_do_slot(counts, Rf_install("i"));
    SEXP x = R_do_slot(counts, Rf_install("x")), dim = R_do_slot(counts, Rf_install("Dim"));
    int64_t n_genes = INTEGER(dim)[0], n_cells = INTEGER(dim)[1];
    if (h->multi != NULL)
        CB_OK(h, cb200_multi_load_csc_i32(h->multi, n_genes, n_cells, INTEGER(p), INTEGER(i), REAL(x)));
    else
        CB_OK(h, cb200_load_csc_i32(h->one, n_genes, n_cells, INTEGER(p), INTEGER(i), REAL(x), 0, n_cells));
    return R_NilValue;
}

/* sf_in: NULL (library-size factors) or a numeric vector that is only re-centred. */
SEXP cb200r_size_factors(SEXP sctx, SEXP sf_in, SEXP n_cells)
{
    cb200r_handle *h = get_handle(sctx);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(n_cells)));
    const double *in = Rf_isNull(sf_in) ? NULL : REAL(sf_in);
    int rc = h->multi != NULL ? cb200_multi_size_factors(h->multi, in, REAL(out)) : cb200_size_factors(h->one, in, REAL(out));
    UNPROTECT(1);
    if (rc != 0)
        Rf_error("%s", last_error(h));
    return out;
}

static SEXP named_list3(SEXP a, const char *na, SEXP b, const char *nb, SEXP c, const char *nc)
{
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3)), nm = PROTECT(Rf_allocVector(STRSXP, 3));
    SET_VECTOR_ELT(out, 0, a);
    SET_VECTOR_ELT(out, 1, b);
    SET_VECTOR_ELT(out, 2, c);
    SET_STRING_ELT(nm, 0, Rf_mkChar(na));
    SET_STRING_ELT(nm, 1, Rf_mkChar(nb));
    SET_STRING_ELT(nm, 2, Rf_mkChar(nc));
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(2);
    return out;
}

/* returns list(x = logcounts@x, mean = , var = ): everything on return (simple form) */
SEXP cb200r_lognorm_genestats(SEXP sctx, SEXP nnz, SEXP n_genes)
{
    cb200r_handle *h = get_handle(sctx);
    SEXP lx = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)Rf_asReal(nnz)));
    SEXP mean = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(n_genes)));
    SEXP var = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(n_genes)));
    int rc = h->multi != NULL ? cb200_multi_lognorm_genestats(h->multi, REAL(lx), REAL(mean), REAL(var))
                              : cb200_lognorm_genestats(h->one, REAL(lx), REAL(mean), REAL(var));
    if (rc != 0) {
        UNPROTECT(3);
        Rf_error("%s", last_error(h));
    }
    SEXP out = named_list3(lx, "x", mean, "mean", var, "var");
    UNPROTECT(3);
    return out;
}

/* Values-first / asynchronous form (single GPU): the logcounts are computed from the values and the column
 * pointers while the row indices are still uploading, and their 8 B/nnz copy into the R vector runs in the
 * background behind the gene sums, the trend fit, the PCA ...  Returns list(pending = <external pointer that keeps
 * the logcount vector alive and unreachable from R code>, mean, var); cb200r_logcounts_wait(ctx, pending) blocks
 * until the copy has landed and hands the vector out. */
SEXP cb200r_lognorm_begin(SEXP sctx, SEXP nnz, SEXP n_genes)
{
    cb200r_handle *h = get_handle(sctx);
    if (h->multi != NULL)
        Rf_error("cellulaB200: cb200r_lognorm_begin is the single-GPU form; use cb200r_lognorm_genestats");
    SEXP lx = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)Rf_asReal(nnz)));
    SEXP mean = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(n_genes)));
    SEXP var = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(n_genes)));
    int rc = cb200_lognorm_values(h->one);
    if (rc == 0)
        rc = cb200_fetch_logcounts_async(h->one, REAL(lx));
    if (rc == 0)
        rc = cb200_lognorm_genestats_local(h->one);
    if (rc == 0)
        rc = cb200_genestats_finish(h->one, REAL(mean), REAL(var));
    if (rc != 0) {
        cb200_wait_copies(h->one);       /* nothing may still write into lx once it is unprotected */
        UNPROTECT(3);
        Rf_error("%s", last_error(h));
    }
    SEXP pending = PROTECT(R_MakeExternalPtr(NULL, R_NilValue, lx));   /* `prot` slot: lx stays alive with the pointer */
    SEXP out = named_list3(pending, "pending", mean, "mean", var, "var");
    UNPROTECT(4);
    return out;
}

SEXP cb200r_logcounts_wait(SEXP sctx, SEXP pending)
{
    cb200r_handle *h = get_handle(sctx);
    if (h->one != NULL)
        CB_OK(h, cb200_wait_copies(h->one));
    return R_ExternalPtrProtected(pending);
}

/* ---- batch-aware branch (cellula R/doNormAndReduce.R:74-76, :84-86); single-GPU form ---- */
static cb200_ctx *one_ctx(cb200r_handle *h, const char *what)
{
    if (h->one == NULL)
        Rf_error("cellulaB200: %s is available with options(cellula.b200.gpus = 1)", what);
    return h->one;
}

/* codes: 0-based integer batch of every cell.  Returns list(ave = G x B matrix of the per-batch averages of
 * counts / batch-centred size factors, sfmean = mean size factor per batch, n = cells per batch). */
SEXP cb200r_block_averages(SEXP sctx, SEXP codes, SEXP n_blocks, SEXP n_genes)
{
    cb200r_handle *h = get_handle(sctx);
    cb200_ctx *ctx = one_ctx(h, "the batch-aware path");
    int B = Rf_asInteger(n_blocks), G = Rf_asInteger(n_genes);
    SEXP ave = PROTECT(Rf_allocMatrix(REALSXP, G, B)), sfm = PROTECT(Rf_allocVector(REALSXP, B));
    SEXP nn = PROTECT(Rf_allocVector(REALSXP, B));
    int64_t *nb = (int64_t *)R_alloc((size_t)B, sizeof(int64_t));
    if (cb200_block_averages(ctx, INTEGER(codes), B, REAL(ave), REAL(sfm), nb) != 0) {
        UNPROTECT(3);
        Rf_error("%s", last_error(h));
    }
    for (int b = 0; b < B; ++b)
        REAL(nn)[b] = (double)nb[b];
    SEXP out = named_list3(ave, "ave", sfm, "sfmean", nn, "n");
    UNPROTECT(3);
    return out;
}

/* size factors used as they are (multiBatchNorm's rescaled factors; no centring) */
SEXP cb200r_size_factors_given(SEXP sctx, SEXP sf)
{
    cb200r_handle *h = get_handle(sctx);
    CB_OK(h, cb200_size_factors_given(one_ctx(h, "the batch-aware path"), REAL(sf)));
    return R_NilValue;
}

/* returns list(x = logcounts@x, mean = G x B, var = G x B): statistics within every batch */
SEXP cb200r_lognorm_genestats_blocked(SEXP sctx, SEXP nnz, SEXP n_genes, SEXP n_blocks)
{
    cb200r_handle *h = get_handle(sctx);
    cb200_ctx *ctx = one_ctx(h, "the batch-aware path");
    int B = Rf_asInteger(n_blocks), G = Rf_asInteger(n_genes);
    SEXP lx = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)Rf_asReal(nnz)));
    SEXP mean = PROTECT(Rf_allocMatrix(REALSXP, G, B)), var = PROTECT(Rf_allocMatrix(REALSXP, G, B));
    if (cb200_lognorm_genestats_blocked(ctx, REAL(lx), REAL(mean), REAL(var)) != 0) {
        UNPROTECT(3);
        Rf_error("%s", last_error(h));
    }
    SEXP out = named_list3(lx, "x", mean, "mean", var, "var");
    UNPROTECT(3);
    return out;
}

/* hvg_idx: 0-based integer rows in getTopHVGs order.  returns list(scores, rotation, varExplained, totalVar) */
SEXP cb200r_pca(SEXP sctx, SEXP hvg_idx, SEXP ndims, SEXP n_cells)
{
    cb200r_handle *hd = get_handle(sctx);
    int d = Rf_asInteger(ndims), h = LENGTH(hvg_idx), n = Rf_asInteger(n_cells);
    SEXP scores = PROTECT(Rf_allocMatrix(REALSXP, n, d));
    SEXP rot = PROTECT(Rf_allocMatrix(REALSXP, h, d));
    SEXP ve = PROTECT(Rf_allocVector(REALSXP, d));
    SEXP tv = PROTECT(Rf_allocVector(REALSXP, 1));
    int rc = hd->multi != NULL ? cb200_multi_pca(hd->multi, INTEGER(hvg_idx), h, d, REAL(scores), REAL(rot), REAL(ve), REAL(tv))
                               : cb200_pca(hd->one, INTEGER(hvg_idx), h, d, REAL(scores), REAL(rot), REAL(ve), REAL(tv));
    if (rc != 0) {
        UNPROTECT(4);
        Rf_error("%s", last_error(hd));
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
    cb200r_handle *h = get_handle(sctx);
    int64_t n = Rf_nrows(space);
    int d = Rf_ncols(space), kk = Rf_asInteger(k);
    int64_t ne = 0;
    /* kNN, then: count, size the R vectors, emit straight into them (copies overlap the emission) */
    if (h->multi != NULL) {
        CB_OK(h, cb200_multi_knn(h->multi, REAL(space), n, d, kk, NULL));
        CB_OK(h, cb200_multi_snn_count(h->multi, NULL, n, kk, Rf_asInteger(type), &ne));
    } else {
        CB_OK(h, cb200_knn(h->one, REAL(space), n, d, kk, 0, n, NULL));
        CB_OK(h, cb200_snn_count_device(h->one, NULL, n, kk, Rf_asInteger(type), 0, n, &ne));
    }
    SEXP sf = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)ne)), st = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)ne));
    SEXP sw = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)ne));
    int rc = h->multi != NULL ? cb200_multi_snn_emit(h->multi, INTEGER(sf), INTEGER(st), REAL(sw))
                              : cb200_snn_emit(h->one, INTEGER(sf), INTEGER(st), REAL(sw));
    if (rc != 0) {
        UNPROTECT(3);
        Rf_error("%s", last_error(h));
    }
    SEXP out = named_list3(sf, "from", st, "to", sw, "weight");
    UNPROTECT(3);
    return out;
}

/* findKNN()$index equivalent: integer matrix N x k, 1-based */
SEXP cb200r_knn(SEXP sctx, SEXP space, SEXP k)
{
    cb200r_handle *h = get_handle(sctx);
    int64_t n = Rf_nrows(space);
    int d = Rf_ncols(space), kk = Rf_asInteger(k);
    SEXP idx = PROTECT(Rf_allocMatrix(INTSXP, (int)n, kk));
    int rc = h->multi != NULL ? cb200_multi_knn(h->multi, REAL(space), n, d, kk, INTEGER(idx))
                              : cb200_knn(h->one, REAL(space), n, d, kk, 0, n, INTEGER(idx));
    UNPROTECT(1);
    if (rc != 0)
        Rf_error("%s", last_error(h));
    return idx;
}

/* bluster::approxSilhouette(x, clusters): codes = 0-based position of every cell's label in
 * sort(unique(clusters)).  Returns list(width, other) (other: 0-based code of the closest other cluster). */
/* the cluster diagnostics are single-GPU reductions: a multi handle opens a single-GPU context beside it on demand */
static cb200_ctx *single_ctx(cb200r_handle *h)
{
    if (h->one == NULL && cb200_init(0, NULL, &h->one) != 0)
        Rf_error("%s", cb200_last_error(NULL));
    return h->one;
}

/* scran::computeSumFactors(sce, clusters = cl) (cellula R/doNormAndReduce.R:66-71) on the resident counts.
 * codes: 0-based integer cluster codes or NULL; sizes: integer pool sizes or NULL (seq(21, 101, 5)); min_mean: NA = guess.
 * Returns list(sf, nonpositive, ms).  Single-GPU form only (the clusters are arbitrary cell subsets). */
SEXP cb200r_pooled_size_factors(SEXP sctx, SEXP codes, SEXP sizes, SEXP min_mean, SEXP max_cluster_size, SEXP n_cells)
{
    cb200r_handle *h = get_handle(sctx);
    if (h->one == NULL)
        Rf_error("cellulaB200: pooled size factors need the single-GPU form (options(cellula.b200.gpus = 1L))");
    SEXP sf = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(n_cells)));
    SEXP ms = PROTECT(Rf_allocVector(REALSXP, 8));
    SEXP nn = PROTECT(Rf_allocVector(REALSXP, 1));
    int64_t nonpos = 0;
    double mm = Rf_asReal(min_mean);
    int rc = cb200_pooled_size_factors(h->one, Rf_isNull(codes) ? NULL : INTEGER(codes), Rf_isNull(sizes) ? NULL : INTEGER(sizes),
                                       Rf_isNull(sizes) ? 0 : Rf_length(sizes), ISNAN(mm) ? -1.0 : mm,
                                       (int64_t)Rf_asReal(max_cluster_size), REAL(sf), &nonpos, REAL(ms));
    REAL(nn)[0] = (double)nonpos;
    SEXP out = named_list3(sf, "sf", nn, "nonpositive", ms, "ms");
    UNPROTECT(3);
    if (rc != 0)
        Rf_error("%s", last_error(h));
    return out;
}

/* scuttle::perCellQCMetrics(sce, subsets) (cellula R/doQC.R:106-108) on the resident counts.  subset_ptr / subset_genes:
 * CSR-style 0-based gene lists.  Returns list(sum, detected, subsets = list(sum = S x N by row, detected = ...)). */
SEXP cb200r_qc_metrics(SEXP sctx, SEXP subset_ptr, SEXP subset_genes, SEXP threshold, SEXP n_cells)
{
    cb200r_handle *h = get_handle(sctx);
    if (h->one == NULL)
        Rf_error("cellulaB200: QC metrics need the single-GPU form (options(cellula.b200.gpus = 1L))");
    int n = Rf_asInteger(n_cells), S = Rf_length(subset_ptr) - 1;
    if (S < 0)
        S = 0;
    SEXP sum = PROTECT(Rf_allocVector(REALSXP, n)), det = PROTECT(Rf_allocVector(INTSXP, n));
    SEXP ssum = PROTECT(Rf_allocMatrix(REALSXP, n, S)), sdet = PROTECT(Rf_allocMatrix(INTSXP, n, S));
    int rc = cb200_qc_metrics(h->one, S, S > 0 ? INTEGER(subset_ptr) : NULL, S > 0 ? INTEGER(subset_genes) : NULL,
                              Rf_asReal(threshold), REAL(sum), INTEGER(det), S > 0 ? REAL(ssum) : NULL,
                              S > 0 ? INTEGER(sdet) : NULL);
    SEXP sub = PROTECT(named_list3(ssum, "sum", sdet, "detected", R_NilValue, "unused"));
    SEXP out = named_list3(sum, "sum", det, "detected", sub, "subsets");
    UNPROTECT(5);
    if (rc != 0)
        Rf_error("%s", last_error(h));
    return out;
}

/* findKNN(space, k, get.distance = TRUE): list(index, distance), both n x k -- what uwot::umap(nn_method = ) takes
 * (cellula R/doNormAndReduce.R:102-106 with n_neighbors = floor(sqrt(N))). */
SEXP cb200r_knn_with_distance(SEXP sctx, SEXP space, SEXP k)
{
    cb200_ctx *ctx = single_ctx(get_handle(sctx));
    int64_t n = Rf_nrows(space);
    int d = Rf_ncols(space), kk = Rf_asInteger(k);
    SEXP idx = PROTECT(Rf_allocMatrix(INTSXP, (int)n, kk)), dist = PROTECT(Rf_allocMatrix(REALSXP, (int)n, kk));
    int rc = cb200_knn(ctx, REAL(space), n, d, kk, 0, n, INTEGER(idx));
    if (rc == 0)
        rc = cb200_knn_distances(ctx, REAL(dist));
    SEXP out = named_list3(idx, "index", dist, "distance", R_NilValue, "unused");
    UNPROTECT(2);
    if (rc != 0)
        Rf_error("%s", cb200_last_error(ctx));
    return out;
}

SEXP cb200r_approx_silhouette(SEXP sctx, SEXP space, SEXP codes, SEXP nclust)
{
    cb200_ctx *ctx = single_ctx(get_handle(sctx));
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
    cb200_ctx *ctx = single_ctx(get_handle(sctx));
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
    {"cb200r_init", (DL_FUNC)&cb200r_init, 2},
    {"cb200r_load_counts", (DL_FUNC)&cb200r_load_counts, 2},
    {"cb200r_size_factors", (DL_FUNC)&cb200r_size_factors, 3},
    {"cb200r_lognorm_genestats", (DL_FUNC)&cb200r_lognorm_genestats, 3},
    {"cb200r_lognorm_begin", (DL_FUNC)&cb200r_lognorm_begin, 3},
    {"cb200r_logcounts_wait", (DL_FUNC)&cb200r_logcounts_wait, 2},
    {"cb200r_block_averages", (DL_FUNC)&cb200r_block_averages, 4},
    {"cb200r_size_factors_given", (DL_FUNC)&cb200r_size_factors_given, 2},
    {"cb200r_lognorm_genestats_blocked", (DL_FUNC)&cb200r_lognorm_genestats_blocked, 4},
    {"cb200r_pca", (DL_FUNC)&cb200r_pca, 4},
    {"cb200r_snn_graph", (DL_FUNC)&cb200r_snn_graph, 4},
    {"cb200r_knn", (DL_FUNC)&cb200r_knn, 3},
    {"cb200r_pooled_size_factors", (DL_FUNC)&cb200r_pooled_size_factors, 6},
    {"cb200r_qc_metrics", (DL_FUNC)&cb200r_qc_metrics, 5},
    {"cb200r_knn_with_distance", (DL_FUNC)&cb200r_knn_with_distance, 3},
    {"cb200r_approx_silhouette", (DL_FUNC)&cb200r_approx_silhouette, 4},
    {"cb200r_pairwise_modularity", (DL_FUNC)&cb200r_pairwise_modularity, 6},
    {NULL, NULL, 0}};

void R_init_cellulaB200(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
