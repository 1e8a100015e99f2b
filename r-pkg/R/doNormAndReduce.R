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
  if (!is.null(batch) && .cb200_gpus() > 1L)
    stop(ep, "the batch-aware path runs on one GPU: set options(cellula.b200.gpus = 1)")  # no CPU fallback

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
    cl <- if (!is.null(batch))
      scran::quickCluster(sce, block = colData(sce)[, batch],
                          min.size = floor(sqrt(min(table(colData(sce)[, batch])))), BPPARAM = parallel_param)
    else scran::quickCluster(sce, min.size = floor(sqrt(ncol(sce))), BPPARAM = parallel_param)
    if (verbose) message(.bluem("[NORM] "), "   Calculating pooled factors.")
    if (identical(getOption("cellula.b200.sizefactors", "deconvolution"), "deconvolution_gpu") && .cb200_gpus() <= 1L) {
      # SURVEY 8(f)-1: the pooling itself (N x 18 medians over the expressed genes + one least-squares solve per
      # cluster) on the GPU; quickCluster's labels go in as integer codes.  Opt-in: scran's own call stays the default
      # until an R host has pinned the two against each other (baseline/run_reference.R dumps both).
      r <- .Call("cb200r_pooled_size_factors", ctx, as.integer(factor(cl)) - 1L, NULL, NA_real_, 3000, ncol(sce),
                 PACKAGE = "cellulaB200")
      sf_in <- r$sf
      if (r$nonpositive > 0) {                       # scuttle: warning + cleanSizeFactors when positive = TRUE
        warning("encountered non-positive size factor estimates")
        sf_in <- scuttle::cleanSizeFactors(sf_in, num.detected = Matrix::colSums(cnt > 0))
      }
    } else {
      sf_in <- sizeFactors(scran::computeSumFactors(sce, clusters = cl, BPPARAM = parallel_param))
    }
  }
  if (verbose) message(.bluem("[NORM] "), "   Log-normalization.")
  sf <- .Call("cb200r_size_factors", ctx, sf_in, ncol(sce), PACKAGE = "cellulaB200")
  sizeFactors(sce) <- sf
  if (!is.null(batch)) {
    # multiBatchNorm (cellula R/doNormAndReduce.R:74-76): per-batch gene averages on the GPU, then the B x B
    # median-ratio rescaling of batchelor:::.rescale_size_factors in R (G-sized), factors handed back as they are
    bf <- factor(colData(sce)[, batch])
    codes <- as.integer(bf) - 1L
    av <- .Call("cb200r_block_averages", ctx, codes, nlevels(bf), nrow(sce), PACKAGE = "cellulaB200")
    sf <- .rescale_size_factors_b200(av$ave, sf, codes)
    .Call("cb200r_size_factors_given", ctx, sf, PACKAGE = "cellulaB200")
    sizeFactors(sce) <- sf
  }
  # Single GPU: the logcounts leave the GPU in the background (8 B per non-zero) while the trend fit and the PCA
  # run; the vector is handed out by cb200r_logcounts_wait below.  Several GPUs: everything on return.
  async <- .cb200_gpus() <= 1L && is.null(batch)
  st <- if (!is.null(batch))
          .Call("cb200r_lognorm_genestats_blocked", ctx, as.numeric(length(cnt@x)), nrow(sce), nlevels(bf),
                PACKAGE = "cellulaB200")
        else if (async) .Call("cb200r_lognorm_begin", ctx, as.numeric(length(cnt@x)), nrow(sce), PACKAGE = "cellulaB200")
        else .Call("cb200r_lognorm_genestats", ctx, as.numeric(length(cnt@x)), nrow(sce), PACKAGE = "cellulaB200")

  if (verbose) message(.bluem("[DR] "), "Selecting HVGs.")
  # trend fit and decomposition stay R code on G points (scran::fitTrendVar is exported)
  if (is.null(batch)) {
    fit <- scran::fitTrendVar(st$mean, st$var)
    tech <- fit$trend(st$mean)
    vargenes <- DataFrame(mean = st$mean, total = st$var, tech = tech, bio = st$var - tech,
                          row.names = rownames(sce))
    vargenes$p.value <- pnorm(vargenes$total / vargenes$tech, mean = 1, sd = fit$std.dev, lower.tail = FALSE)
    vargenes$FDR <- p.adjust(vargenes$p.value, method = "BH")
  } else {
    # modelGeneVar(block = ): one trend per batch with >= 2 cells, unweighted average of the per-batch columns
    # (equiweight = TRUE); p-values combined with Stouffer's method like scran (metapod::parallelStouffer)
    per <- lapply(which(av$n >= 2), function(b) {
      fit <- scran::fitTrendVar(st$mean[, b], st$var[, b])
      tech <- fit$trend(st$mean[, b])
      list(mean = st$mean[, b], total = st$var[, b], tech = tech, bio = st$var[, b] - tech,
           p = pnorm(st$var[, b] / tech, mean = 1, sd = fit$std.dev, lower.tail = FALSE))
    })
    avg <- function(k) rowMeans(do.call(cbind, lapply(per, `[[`, k)))
    vargenes <- DataFrame(mean = avg("mean"), total = avg("total"), tech = avg("tech"), bio = avg("bio"),
                          row.names = rownames(sce))
    z <- rowSums(do.call(cbind, lapply(per, function(p) qnorm(p$p, lower.tail = FALSE)))) / sqrt(length(per))
    vargenes$p.value <- pnorm(z, lower.tail = FALSE)
    vargenes$FDR <- p.adjust(vargenes$p.value, method = "BH")
  }
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
    space <- reducedDim(sce, "PCA")[, seq_len(ndims)]
    nn <- NULL
    if (isTRUE(getOption("cellula.b200.umap_knn", FALSE))) {
      # SURVEY 8(f)-4: uwot searches neighbor_n neighbours (self included) with Annoy; hand it the exact GPU search
      # instead (same kernel as makeSNNGraph's, k = neighbor_n - 1 others + the cell itself in column 1)
      storage.mode(space) <- "double"
      r <- .Call("cb200r_knn_with_distance", .cb200_ctx(), space, as.integer(neighbor_n - 1L), PACKAGE = "cellulaB200")
      nn <- list(idx = cbind(seq_len(nrow(space)), r$index), dist = cbind(0, r$distance))
    }
    reducedDim(sce, "UMAP") <- umap(space,
                                    n_neighbors = neighbor_n,
                                    nn_method = nn,
                                    seed = umap_seed,
                                    min_dist = 0.7)
  }
  sce
}


# batchelor:::.rescale_size_factors on the GPU-computed batch averages (ave: G x B): pairwise median ratio of the
# averages over the genes whose grand mean is >= min.mean; every batch is scaled down to the lowest-coverage one.
.rescale_size_factors_b200 <- function(ave, sf, codes, min.mean = 1) {
  nb <- ncol(ave)
  ratios <- matrix(1, nb, nb)
  for (first in seq_len(nb - 1L)) {
    fa <- ave[, first]; fs <- sum(fa)
    for (second in first + seq_len(nb - first)) {
      sa <- ave[, second]; ss <- sum(sa)
      grand <- (fa / fs + sa / ss) / 2 * (fs + ss) / 2
      keep <- grand >= min.mean
      cur <- median(sa[keep] / fa[keep])
      if (!is.finite(cur) || cur == 0) stop("median ratio of averages between batches is not finite")
      ratios[first, second] <- cur
      ratios[second, first] <- 1 / cur
    }
  }
  smallest <- which.min(apply(ratios, 2, min, na.rm = TRUE))
  out <- sf
  for (b in seq_len(nb)) {
    cells <- codes == (b - 1L)
    out[cells] <- sf[cells] / mean(sf[cells]) / ratios[b, smallest]
  }
  out
}
