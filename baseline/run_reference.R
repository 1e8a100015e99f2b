#!/usr/bin/env Rscript
# baseline/run_reference.R -- runs the UNMODIFIED reference hot path (gdagstn/cellula) on an R host and dumps
# every intermediate of the path in a language-neutral form, so that the parity of cellula_b200 (and of its CPU
# oracle) can be pinned to the real reference, and times every upstream call (the CPU baseline of BASELINE.md 3).
#
# This script cannot run in the build image or on the GPU box (no R there: SURVEY.md 8c).  It needs R >= 4.4 with
# SingleCellExperiment, scran, scuttle, scater, BiocSingular, BiocNeighbors, bluster, BiocParallel, igraph, Matrix
# (and uwot unless --skip-umap), plus a checkout of the reference.
#
#   python scripts/write_dataset.py --workload cfg1 --out data/cfg1          # the shared synthetic counts
#   Rscript baseline/run_reference.R --reference /path/to/cellula --data data/cfg1 --out dump/cfg1 \
#           --ndims 30 --k 10 --scheme jaccard --workers 16 --skip-umap
#   python scripts/compare_reference_dump.py --data data/cfg1 --dump dump/cfg1   # oracle (and GPU) vs the dump
#
# What is called (nothing of the reference is edited; its files are source()d):
#   variant A  doNormAndReduce() as is            /root/reference/R/doNormAndReduce.R:37-109 (deconvolution factors)
#   variant B  the same upstream calls with library-size factors (skips :57-71) -- the like-for-like arm of the
#              GPU path's default: logNormCounts :78, modelGeneVar :88, getTopHVGs :90, runPCA :94-97
#   both       makeGraphsAndClusters(save_graphs=TRUE, ...)  /root/reference/R/makeGraphsAndClusters.R:46-157
#              and, timed separately, the two calls makeSNNGraph makes: findKNN + neighborsToSNNGraph (:83-85)
# Dumps (little-endian raw vectors + JSON): see write_vec() calls below and scripts/compare_reference_dump.py.

suppressPackageStartupMessages({
  library(Matrix); library(SingleCellExperiment); library(scran); library(scuttle); library(scater)
  library(BiocSingular); library(BiocNeighbors); library(bluster); library(BiocParallel); library(igraph)
})

args <- commandArgs(trailingOnly = TRUE)
opt <- list(reference = "/root/reference", data = NULL, out = NULL, ndims = 50L, k = 20L, scheme = "rank",
            hvg = 2000L, workers = parallel::detectCores(), variant = "both", "skip-umap" = FALSE, irlba = FALSE,
            resolution = 0.5)
i <- 1
while (i <= length(args)) {
  key <- sub("^--", "", args[i])
  if (key %in% c("skip-umap", "irlba")) { opt[[key]] <- TRUE; i <- i + 1 }
  else { opt[[key]] <- args[i + 1]; i <- i + 2 }
}
for (n in c("ndims", "k", "hvg", "workers")) opt[[n]] <- as.integer(opt[[n]])
opt$resolution <- as.numeric(opt$resolution)
stopifnot(!is.null(opt$data), !is.null(opt$out))
dir.create(opt$out, recursive = TRUE, showWarnings = FALSE)

# ---- the shared dataset (scripts/write_dataset.py; generator spec cellula_b200/csrc/cb200_synth_core.h) ----------
man_txt <- paste(readLines(file.path(opt$data, "manifest.json")), collapse = "")
man_int <- function(key) as.numeric(sub(paste0('.*"', key, '": *([0-9]+).*'), "\\1", man_txt))
G <- man_int("n_genes"); N <- man_int("n_cells"); nnz <- man_int("nnz")
stopifnot(nnz < 2^31)                      # dgCMatrix limit (cfg5 is GPU-only)
p <- readBin(file.path(opt$data, "p.i64"), "integer", n = 2 * (N + 1), size = 4, endian = "little")
p <- p[seq(1, 2 * (N + 1), by = 2)]        # low words of the int64 column pointers (nnz < 2^31)
ii <- readBin(file.path(opt$data, "i.i32"), "integer", n = nnz, size = 4, endian = "little")
xx <- readBin(file.path(opt$data, "x.f64"), "double", n = nnz, size = 8, endian = "little")
counts <- new("dgCMatrix", i = ii, p = p, x = xx, Dim = c(as.integer(G), as.integer(N)),
              Dimnames = list(paste0("g", seq_len(G) - 1L), paste0("c", seq_len(N) - 1L)))
rm(ii, xx); gc()

write_vec <- function(v, name, what = c("double", "integer")) {
  what <- match.arg(what)
  con <- file(file.path(opt$out, name), "wb"); on.exit(close(con))
  if (what == "double") writeBin(as.double(v), con, size = 8, endian = "little")
  else writeBin(as.integer(v), con, size = 4, endian = "little")
}
timings <- list()
tick <- function(label, expr) { t <- system.time(val <- force(expr)); timings[[label]] <<- unname(t["elapsed"]); val }

# ---- the reference's own functions, sourced unchanged --------------------------------------------------------------
ref_env <- new.env()
for (f in c("utils.R", "doNormAndReduce.R", "makeGraphsAndClusters.R"))
  sys.source(file.path(opt$reference, "R", f), envir = ref_env)
if (opt[["skip-umap"]])                    # uwot::umap at n_neighbors = floor(sqrt(N)) is off the measured path (SURVEY 8d)
  assign("umap", function(X, ...) matrix(0, nrow(X), 2), envir = ref_env)
if (opt$irlba) options(BiocSingularParam.default = BiocSingular::IrlbaParam())
bp <- if (opt$workers > 1) MulticoreParam(workers = opt$workers) else SerialParam()

dump_path <- function(sce, tag, vargenes = NULL) {
  write_vec(sizeFactors(sce), paste0(tag, "_sf.f64"))
  write_vec(logcounts(sce)@x, paste0(tag, "_logcounts_x.f64"))
  stopifnot(identical(logcounts(sce)@i, counts@i), identical(logcounts(sce)@p, counts@p))
  if (is.null(vargenes)) vargenes <- modelGeneVar(sce)
  for (col in c("mean", "total", "tech", "bio")) write_vec(vargenes[[col]], paste0(tag, "_gene_", col, ".f64"))
  hv <- metadata(sce)$hvgs
  write_vec(match(hv, rownames(sce)) - 1L, paste0(tag, "_hvg_idx.i32"), "integer")
  pca <- reducedDim(sce, "PCA")
  write_vec(as.vector(pca), paste0(tag, "_pca_scores.f64"))                       # column-major N x ndims
  write_vec(as.vector(attr(pca, "rotation")), paste0(tag, "_pca_rotation.f64"))   # column-major H x ndims
  write_vec(attr(pca, "varExplained"), paste0(tag, "_pca_var_explained.f64"))
  write_vec(attr(pca, "percentVar"), paste0(tag, "_pca_percent_var.f64"))
  space <- pca[, seq_len(opt$ndims), drop = FALSE]
  knn <- tick(paste0(tag, "_findKNN"), findKNN(space, k = opt$k, BNPARAM = KmknnParam(), get.distance = FALSE))
  write_vec(as.vector(knn$index), paste0(tag, "_knn_index.i32"), "integer")       # column-major N x k, 1-based
  g0 <- tick(paste0(tag, "_neighborsToSNNGraph"), neighborsToSNNGraph(knn$index, type = opt$scheme))
  sce <- tick(paste0(tag, "_makeGraphsAndClusters"),
              ref_env$makeGraphsAndClusters(sce, neighbors = opt$k, weighting_scheme = opt$scheme, ndims = opt$ndims,
                                            k = opt$resolution, calculate_modularity = FALSE,
                                            calculate_silhouette = FALSE, save_graphs = TRUE))
  g <- metadata(sce)[[paste0("SNN_", opt$k, "_", opt$scheme)]]
  el <- as_edgelist(g, names = FALSE)
  write_vec(el[, 1], paste0(tag, "_snn_from.i32"), "integer")
  write_vec(el[, 2], paste0(tag, "_snn_to.i32"), "integer")
  write_vec(E(g)$weight, paste0(tag, "_snn_weight.f64"))
  el0 <- as_edgelist(g0, names = FALSE)
  stopifnot(identical(el, el0), identical(E(g)$weight, E(g0)$weight), is_simple(g))
  invisible(sce)
}

# ---- variant B: library-size factors, the upstream calls in the reference's order -------------------------------
if (opt$variant %in% c("both", "B")) {
  sce <- SingleCellExperiment(assays = list(counts = counts))
  sce <- tick("B_logNormCounts", logNormCounts(sce))                                  # doNormAndReduce.R:78
  vargenes <- tick("B_modelGeneVar", modelGeneVar(sce))                                # :88
  hvgs <- tick("B_getTopHVGs", getTopHVGs(vargenes, n = opt$hvg))                      # :90
  metadata(sce)$hvgs <- hvgs                                                           # :91
  sce <- tick("B_runPCA", runPCA(sce, subset_row = hvgs, exprs_values = "logcounts", ncomponents = opt$ndims))  # :94-97
  dump_path(sce, "B", vargenes)
}

# ---- variant A: doNormAndReduce() exactly as the reference runs it ----------------------------------------------
if (opt$variant %in% c("both", "A")) {
  sce <- SingleCellExperiment(assays = list(counts = counts))
  sce <- tick("A_doNormAndReduce", ref_env$doNormAndReduce(sce, ndims = opt$ndims, hvg_ntop = opt$hvg, verbose = FALSE,
                                                            parallel_param = bp))
  dump_path(sce, "A")
}

# ---- SURVEY 8(f)-1 / 8(f)-4: the calls either side of the path, on the same counts ------------------------------
# quickCluster + computeSumFactors exactly as doNormAndReduce.R:57-71 calls them (the labels are dumped too: they are
# the INPUT of cb200_pooled_size_factors, whose output is compared with F1_pooled_sf); perCellQCMetrics as
# doQC.R:106-108 calls it, with three deterministic gene subsets; findKNN(get.distance = TRUE) for cb200_knn_distances.
if (opt$variant %in% c("both", "A", "F")) {
  sce <- SingleCellExperiment(assays = list(counts = counts))
  cl <- tick("F1_quickCluster", quickCluster(sce, min.size = floor(sqrt(ncol(sce))), BPPARAM = bp))
  write_vec(as.integer(factor(cl)) - 1L, "F1_quickcluster_codes.i32", "integer")
  sce <- tick("F1_computeSumFactors", computeSumFactors(sce, clusters = cl, BPPARAM = bp))
  write_vec(sizeFactors(sce), "F1_pooled_sf.f64")
  subsets <- list(first = seq(1, G, by = 97), second = seq(5, G, by = 1013), third = seq_len(min(G, 50)))
  qc <- tick("F4_perCellQCMetrics", perCellQCMetrics(sce, subsets = subsets, BPPARAM = bp))
  write_vec(qc$sum, "F4_qc_sum.f64"); write_vec(qc$detected, "F4_qc_detected.i32", "integer")
  for (nm in names(subsets)) {
    write_vec(subsets[[nm]] - 1L, paste0("F4_subset_", nm, "_genes.i32"), "integer")
    write_vec(qc[[paste0("subsets_", nm, "_sum")]], paste0("F4_qc_", nm, "_sum.f64"))
    write_vec(qc[[paste0("subsets_", nm, "_detected")]], paste0("F4_qc_", nm, "_detected.i32"), "integer")
    write_vec(qc[[paste0("subsets_", nm, "_percent")]], paste0("F4_qc_", nm, "_percent.f64"))
  }
}

# ---- SURVEY 8c.2 checklist: the upstream behaviours the oracle assumes ---------------------------------------------
chk <- list()
lat <- as.matrix(expand.grid(0:9, 0:9, 0:4)) * 1.0                # integer lattice: exact distance ties everywhere
kl <- findKNN(lat, k = 7, BNPARAM = KmknnParam(), get.distance = TRUE)
write_vec(as.vector(lat), "check_lattice_x.f64"); write_vec(as.vector(kl$index), "check_lattice_knn.i32", "integer")
write_vec(as.vector(kl$distance), "check_lattice_dist.f64")       # 8(f)-4: cb200_knn_distances
dup <- rbind(matrix(rnorm(200 * 5), 200), matrix(1.5, 30, 5))     # 30 identical rows: self-exclusion with duplicates
kd <- findKNN(dup, k = 10, BNPARAM = KmknnParam(), get.distance = FALSE)
write_vec(as.vector(dup), "check_dup_x.f64"); write_vec(as.vector(kd$index), "check_dup_knn.i32", "integer")
kat <- matrix(c(0, 1, 3, 6, 10, 15), ncol = 1)                     # SURVEY 8c.3 KAT-SNN-1
for (ty in c("rank", "number", "jaccard")) {
  gk <- makeSNNGraph(kat, k = 2, type = ty)
  chk[[paste0("kat_", ty)]] <- list(edges = as_edgelist(gk, names = FALSE), weight = E(gk)$weight)
}
chk$bsparam <- capture.output(print(bsparam()))
chk$versions <- sapply(c("scran", "scuttle", "scater", "BiocSingular", "BiocNeighbors", "bluster", "batchelor",
                         "igraph", "Matrix", "irlba"), function(p) tryCatch(as.character(packageVersion(p)), error = function(e) NA))

# ---- report --------------------------------------------------------------------------------------------------------
json_num <- function(l) paste0("{", paste(sprintf('"%s": %s', names(l), format(unlist(l), digits = 10)), collapse = ", "), "}")
writeLines(c("{", sprintf('"n_genes": %d, "n_cells": %d, "nnz": %.0f, "ndims": %d, "k": %d, "scheme": "%s", "hvg": %d,',
                           G, N, nnz, opt$ndims, opt$k, opt$scheme, opt$hvg),
             sprintf('"workers": %d, "skip_umap": %s, "irlba": %s,', opt$workers, tolower(opt[["skip-umap"]]), tolower(opt$irlba)),
             sprintf('"timings_s": %s', json_num(timings)), "}"), file.path(opt$out, "report.json"))
capture.output(print(chk), file = file.path(opt$out, "checklist.txt"))
capture.output(sessionInfo(), file = file.path(opt$out, "sessionInfo.txt"))
message("reference dump written to ", opt$out)
