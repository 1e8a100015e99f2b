#' Drop-in for cellula::makeGraphsAndClusters() (cellula R/makeGraphsAndClusters.R:46-157).
#' Only the bluster::makeSNNGraph() calls are replaced (makeSNNGraphB200); the igraph
#' Louvain/Leiden sweep, modularity and silhouette code is the reference's, unchanged.
#' `space` is defined before the branch: the reference's "SNN" sweep fails with
#' "object 'space' not found" (R/makeGraphsAndClusters.R:80 vs :125).
makeGraphsAndClusters <- function(sce,
                                  neighbors = 10L,
                                  weighting_scheme = "jaccard",
                                  sweep_on = "clustering",
                                  method = "louvain",
                                  k = seq(0.1, 1, length.out = 6),
                                  dr = "PCA",
                                  ndims = 20L,
                                  calculate_modularity = TRUE,
                                  calculate_silhouette = TRUE,
                                  leiden_iterations = 5L,
                                  save_graphs = FALSE,
                                  prefix = "SNN_",
                                  verbose = FALSE) {
  ep <- .redm("{cellula::makeGraphsAndClusters} - ")
  if (!is(sce, "SingleCellExperiment"))
    stop(ep, "must provide a SingleCellExperiment object")
  if (!method %in% c("leiden", "louvain"))
    stop(ep, "method not recognized - must be one of \"leiden\" or \"louvain\"")
  if (!weighting_scheme %in% c("jaccard", "rank", "number"))
    stop(ep, "method not recognized - must be one of \"jaccard\", \"rank\", or \"number\"")
  if (length(reducedDim(sce)) == 0) stop(ep, "there are no dimensionality reductions in the SingleCellExperiment object.")
  if (!dr %in% reducedDimNames(sce)) stop(ep, "the dr name supplied was not found in the SingleCellExperiment object")

  space <- reducedDim(sce, dr)
  if (ncol(space) > ndims) space <- space[, seq_len(ndims)]

  one_clustering <- function(g, i, resolution_given) {
    if (method == "louvain") {
      cl <- if (resolution_given) factor(cluster_louvain(g, resolution = i)$membership)
            else factor(cluster_louvain(g)$membership)
    } else {
      cl <- if (resolution_given)
        factor(cluster_leiden(g, objective_function = "modularity", n_iterations = leiden_iterations,
                              resolution_parameter = i)$membership)
      else factor(cluster_leiden(g, objective_function = "modularity", n_iterations = leiden_iterations)$membership)
    }
    gname <- paste0(prefix, i)
    colData(sce)[, gname] <<- as.character(cl)
    if (verbose) message(.bluem("[CLU]"), "Found", length(unique(cl)), "clusters.")
    if (calculate_modularity) {
      if (verbose) message(.bluem("[CLU]"), "Calculating pairwise modularity.")
      metadata(sce)[[paste0("modularity_", gname)]] <<- pairwiseModularityB200(g, clusters = cl, as.ratio = TRUE)
    }
    if (calculate_silhouette) {
      if (verbose) message(.bluem("[CLU]"), "Calculating approximate silhouette widths.")
      silhouette <- as.data.frame(approxSilhouetteB200(space, clusters = cl))
      silhouette$closest <- factor(ifelse(silhouette$width > 0, cl, silhouette$other))
      silhouette$cluster <- cl
      metadata(sce)[[paste0("silhouette_", gname)]] <<- silhouette
    }
  }

  if (sweep_on == "clustering" & !is.null(neighbors)) {
    if (verbose) message(.bluem("[CLU]"), "Creating SNN graph.")
    g <- makeSNNGraphB200(space, k = neighbors, type = weighting_scheme)
    for (i in k) {
      if (verbose) message(.bluem("[CLU]"), "Clustering at resolution ", i, ".")
      one_clustering(g, i, TRUE)
      if (save_graphs)
        metadata(sce)[[paste0("SNN_", neighbors, "_", weighting_scheme)]] <- g
    }
  } else if (sweep_on == "SNN" & !is.null(k)) {
    k <- floor(k)
    for (i in k) {
      if (verbose) message(.bluem("[CLU]"), "Creating SNN graph with k =", i, "neighbors.")
      g <- makeSNNGraphB200(space, k = i, type = weighting_scheme)
      if (save_graphs)
        metadata(sce)[[paste0("SNN_", i, "_", weighting_scheme)]] <- g
      one_clustering(g, i, FALSE)
    }
  }
  return(sce)
}

#' Drop-in for cellula::rebuildModularity() (cellula R/makeGraphsAndClusters.R:270-290).
rebuildModularity <- function(sce, dr = "PCA", neighbors = NULL, ndims = 20,
                              labels, weighting_scheme = "jaccard") {
  ep <- .redm("{cellula::rebuildModularity} - ")
  if (!dr %in% reducedDimNames(sce))
    stop(ep, "the dr name supplied was not found in the SingleCellExperiment object")
  if (!weighting_scheme %in% c("jaccard", "rank", "number"))
    stop(ep, "method not recognized - must be one of \"jaccard\", \"rank\", or \"number\"")
  space <- reducedDim(sce, dr)[, seq_len(ndims), drop = FALSE]
  if (is.null(neighbors)) neighbors <- round(sqrt(nrow(space)))
  g <- makeSNNGraphB200(space, k = neighbors, type = weighting_scheme)
  cl <- colData(sce)[, labels]
  mlab <- paste0("modularity_", labels)
  metadata(sce)[[mlab]] <- pairwiseModularityB200(g, clusters = cl, as.ratio = TRUE)
  sce
}
