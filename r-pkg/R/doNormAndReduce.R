#' Drop-in for cellula::doNormAndReduce() (cellula R/doNormAndReduce.R:37-109).
#' Same signature, same checks and messages, same returned slots: sizeFactors, assay
#' "logcounts" (dgCMatrix sharing i/p with counts), metadata(sce)$hvgs, reducedDim "PCA" with
#' attributes varExplained / percentVar / rotation, reducedDim "UMAP".  Additionally the
#' modelGeneVar table is kept in rowData(sce).
doNormAndReduce <- function(sce, batch = NULL, name = NULL,
                            ndims = 20,
                            hvg_ntop = 2000,
                            umap_seed = 11,
                            verbose = TRUE,
                            parallel_param = SerialParam()) {
  ep <- .redm("{cellula::doNormAndReduce()} - ")
  if (!is(sce, "SingleCellExperiment"))
    stop(ep, "must provide a SingleCellExperiment object")
  if (!is.null(batch)) {
    if (!batch %in% colnames(colData(sce)))
      stop(ep, "batch column not found in the colData of the object")
  }
  if (hvg_ntop > nrow(sce))
    stop(ep, "hvg_ntop cannot be higher than the number of features (nrow) in the object")
  if (!is.null(batch))
    stop(ep, "the batch-aware path is not implemented on the B200 backend")  # no CPU fallback

  ctx <- .cb200_ctx()
  cnt <- counts(sce)
  if (!is(cnt, "dgCMatrix")) cnt <- as(cnt, "CsparseMatrix")

  if (verbose) message(.bluem("[NORM] "), "Calculating size factors and normalizing.")
  .Call("cb200r_load_counts", ctx, cnt, PACKAGE = "cellulaB200")
  # Size factors.  DEFAULT = what the reference does: its own pre-clustering and pooled (deconvolution) factors
  # (cellula R/doNormAndReduce.R:57-71) run unchanged in R -- they are the only calls that receive `parallel_param`
  # -- and the GPU re-centres the vector exactly as logNormCounts would, so that a default call returns the
  # reference's sizeFactors / logcounts / HVGs.  options(cellula.b200.sizefactors = "library") is the explicit
  # fast path (library-size factors computed on the GPU, no scran pre-clustering): different numbers, announced.
  sf_in <- NULL
  if (identical(getOption("cellula.b200.sizefactors", "deconvolution"), "library")) {
    if (verbose) message(.bluem("[NORM] "), "   Library-size factors (options(cellula.b200.sizefactors = \"library\")): ",
                         "results differ from cellula's deconvolution factors.")
  } else {
    if (verbose) message(.bluem("[NORM] "), "   Preclustering.")
    cl <- scran::quickCluster(sce, min.size = floor(sqrt(ncol(sce))), BPPARAM = parallel_param)
    if (verbose) message(.bluem("[NORM] "), "   Calculating pooled factors.")
    sf_in <- sizeFactors(scran::computeSumFactors(sce, clusters = cl, BPPARAM = parallel_param))
  }
  if (verbose) message(.bluem("[NORM] "), "   Log-normalization.")
  sf <- .Call("cb200r_size_factors", ctx, sf_in, ncol(sce), PACKAGE = "cellulaB200")
  sizeFactors(sce) <- sf
  # Single GPU: the logcounts leave the GPU in the background (8 B per non-zero) while the trend fit and the PCA
  # run; the vector is handed out by cb200r_logcounts_wait below.  Several GPUs: everything on return.
  async <- .cb200_gpus() <= 1L
  st <- if (async) .Call("cb200r_lognorm_begin", ctx, as.numeric(length(cnt@x)), nrow(sce), PACKAGE = "cellulaB200")
        else .Call("cb200r_lognorm_genestats", ctx, as.numeric(length(cnt@x)), nrow(sce), PACKAGE = "cellulaB200")

  if (verbose) message(.bluem("[DR] "), "Selecting HVGs.")
  # trend fit and decomposition stay R code on G points (scran::fitTrendVar is exported)
  fit <- scran::fitTrendVar(st$mean, st$var)
  tech <- fit$trend(st$mean)
  vargenes <- DataFrame(mean = st$mean, total = st$var, tech = tech, bio = st$var - tech,
                        row.names = rownames(sce))
  vargenes$p.value <- pnorm(vargenes$total / vargenes$tech, mean = 1, sd = fit$std.dev, lower.tail = FALSE)
  vargenes$FDR <- p.adjust(vargenes$p.value, method = "BH")
  rowData(sce)[, colnames(vargenes)] <- vargenes
  hvgs <- getTopHVGs(vargenes, n = hvg_ntop)
  metadata(sce)$hvgs <- hvgs

  if (verbose) message(.bluem("[DR] "), "Running PCA.")
  hvg_idx <- if (is.character(hvgs)) match(hvgs, rownames(sce)) - 1L else as.integer(hvgs) - 1L
  pca <- .Call("cb200r_pca", ctx, as.integer(hvg_idx), as.integer(ndims), ncol(sce), PACKAGE = "cellulaB200")
  scores <- pca$scores
  dimnames(scores) <- list(colnames(sce), paste0("PC", seq_len(ncol(scores))))
  rot <- pca$rotation
  dimnames(rot) <- list(hvgs, colnames(scores))
  attr(scores, "varExplained") <- pca$varExplained
  attr(scores, "percentVar") <- 100 * pca$varExplained / pca$totalVar
  attr(scores, "rotation") <- rot
  reducedDim(sce, "PCA") <- scores

  lx <- if (async) .Call("cb200r_logcounts_wait", ctx, st$pending, PACKAGE = "cellulaB200") else st$x
  assay(sce, "logcounts") <- new("dgCMatrix", i = cnt@i, p = cnt@p, x = lx, Dim = cnt@Dim,
                                 Dimnames = cnt@Dimnames)

  if (!isTRUE(getOption("cellula.b200.skip_umap", FALSE))) {
    if (verbose) message(.bluem("[DR] "), "Running UMAP on uncorrected PCA.")
    neighbor_n <- floor(sqrt(ncol(sce)))
    reducedDim(sce, "UMAP") <- umap(reducedDim(sce, "PCA")[, seq_len(ndims)],
                                    n_neighbors = neighbor_n,
                                    seed = umap_seed,
                                    min_dist = 0.7)
  }
  sce
}
