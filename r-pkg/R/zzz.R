# Device context shared by the wrappers (an external pointer with a finalizer that calls cb200_free /
# cb200_multi_free).  options(cellula.b200.device = 0L) selects the GPU of the single-GPU form;
# options(cellula.b200.gpus = 8L) drives that many GPUs of the node from this one R process (cells sharded by
# column inside the library, NCCL between the GPUs; results identical to the single-GPU ones).
.cb200_env <- new.env(parent = emptyenv())

.cb200_gpus <- function() as.integer(getOption("cellula.b200.gpus", 1L))

.cb200_ctx <- function() {
  if (is.null(.cb200_env$ctx)) {
    device <- as.integer(getOption("cellula.b200.device", 0L))
    .cb200_env$ctx <- .Call("cb200r_init", device, .cb200_gpus(), PACKAGE = "cellulaB200")
  }
  .cb200_env$ctx
}

# message colouring exactly as cellula's R/utils.R:22-33
.bluem <- function(x) crayon::blue(x)
.redm <- function(x) crayon::red(x)

.snn_type_code <- function(type) match(type, c("rank", "number", "jaccard")) - 1L

#' bluster::makeSNNGraph(x, k, type) on the GPU: exact kNN + SNN edge weights, then the same
#' igraph construction neighborsToSNNGraph performs.
makeSNNGraphB200 <- function(x, k = 10, type = "rank") {
  x <- as.matrix(x)
  storage.mode(x) <- "double"
  e <- .Call("cb200r_snn_graph", .cb200_ctx(), x, as.integer(k), .snn_type_code(type),
             PACKAGE = "cellulaB200")
  g <- igraph::make_graph(rbind(e$from, e$to), n = nrow(x), directed = FALSE)
  igraph::E(g)$weight <- e$weight
  g
}

#' bluster::approxSilhouette(x, clusters) on the GPU (same DataFrame columns: cluster, other, width).
approxSilhouetteB200 <- function(x, clusters) {
  x <- as.matrix(x)
  storage.mode(x) <- "double"
  uclust <- sort(unique(clusters))
  r <- .Call("cb200r_approx_silhouette", .cb200_ctx(), x, match(clusters, uclust) - 1L, length(uclust),
             PACKAGE = "cellulaB200")
  S4Vectors::DataFrame(cluster = clusters, other = uclust[r$other + 1L], width = r$width, row.names = rownames(x))
}

#' bluster::pairwiseModularity(graph, clusters, get.weights, as.ratio) on the GPU.
pairwiseModularityB200 <- function(graph, clusters, get.weights = FALSE, as.ratio = FALSE) {
  uclust <- sort(unique(clusters))
  el <- igraph::as_edgelist(graph, names = FALSE)
  w <- igraph::E(graph)$weight
  if (is.null(w)) w <- rep(1, nrow(el))
  r <- .Call("cb200r_pairwise_modularity", .cb200_ctx(), as.integer(el[, 1]), as.integer(el[, 2]), as.numeric(w),
             match(clusters, uclust) - 1L, length(uclust), PACKAGE = "cellulaB200")
  obs <- t(r$observed_t); expd <- t(r$expected_t)            # the library fills row-major
  dimnames(obs) <- dimnames(expd) <- list(uclust, uclust)
  if (get.weights) list(observed = obs, expected = expd)
  else if (as.ratio) obs / expd
  else (obs - expd) / r$total
}

#' scuttle::perCellQCMetrics(sce, subsets = list(...)) on the GPU, as doQC() calls it (cellula R/doQC.R:106-108):
#' same DataFrame columns (sum, detected, subsets_<name>_sum / _detected / _percent, total).
perCellQCMetricsB200 <- function(sce, subsets = NULL, threshold = 0) {
  cnt <- SingleCellExperiment::counts(sce)
  if (!is(cnt, "dgCMatrix")) cnt <- as(cnt, "CsparseMatrix")
  ctx <- .cb200_ctx()
  .Call("cb200r_load_counts", ctx, cnt, PACKAGE = "cellulaB200")
  idx <- lapply(subsets, function(s) {
    if (is.character(s)) match(s, rownames(sce)) - 1L
    else if (is.logical(s)) which(s) - 1L
    else as.integer(s) - 1L
  })
  if (any(vapply(idx, anyNA, TRUE))) stop("subsets contain genes that are not rows of the object")
  ptr <- as.integer(c(0L, cumsum(lengths(idx))))
  r <- .Call("cb200r_qc_metrics", ctx, if (length(idx)) ptr else integer(0), as.integer(unlist(idx)), as.numeric(threshold),
             ncol(sce), PACKAGE = "cellulaB200")
  out <- S4Vectors::DataFrame(sum = r$sum, detected = r$detected, row.names = colnames(sce))
  for (s in seq_along(idx)) {
    nm <- names(subsets)[s]
    out[[paste0("subsets_", nm, "_sum")]] <- r$subsets$sum[, s]
    out[[paste0("subsets_", nm, "_detected")]] <- r$subsets$detected[, s]
    out[[paste0("subsets_", nm, "_percent")]] <- r$subsets$sum[, s] / r$sum * 100
  }
  out$total <- r$sum
  out
}
