# Device context shared by the wrappers (an external pointer with a finalizer that calls
# cb200_free).  options(cellula.b200.device = 0L) selects the GPU.
.cb200_env <- new.env(parent = emptyenv())

.cb200_ctx <- function() {
  if (is.null(.cb200_env$ctx))
    .cb200_env$ctx <- .Call("cb200r_init", as.integer(getOption("cellula.b200.device", 0L)),
                            PACKAGE = "cellulaB200")
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
