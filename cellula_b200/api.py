"""Python mirror of the two reference entry points of the hot path, same names, argument
meaning and error behaviour:

    doNormAndReduce()        /root/reference/R/doNormAndReduce.R:37-109
    makeGraphsAndClusters()  /root/reference/R/makeGraphsAndClusters.R:46-157
    makeSNNGraph()           bluster::makeSNNGraph as called at makeGraphsAndClusters.R:83-85
    rebuildModularity()      /root/reference/R/makeGraphsAndClusters.R:270-290 (graph part)

The drop-in for R users is r-pkg/ (identical R signatures over the same C ABI); this module is
the host side for the image the library is built and tested in, which has no R.  All arithmetic
of the path runs in libcellula_b200.so on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import math
import sys

import numpy as np

from . import _lib, trendvar
from .sce import CSC, ReducedDim, SingleCellExperiment, SNNGraph

# R: options(cellula.b200.*) -- knobs that must not change the reference signatures
options = dict(device=0, skip_umap=False, size_factors="library", keep_context=True)

_context = None


def _blue(s):
    return f"\033[34m{s}\033[39m"


def _red(s):
    return f"\033[31m{s}\033[39m"


def _message(*parts):
    print("".join(str(p) for p in parts), file=sys.stderr)


def get_context():
    """The process-wide device context (R: an external pointer with a finalizer)."""
    global _context
    if _context is None:
        _context = _lib.Context(options["device"])
    return _context


def release_context():
    global _context
    if _context is not None:
        _context.close()
        _context = None


class SerialParam:
    """Placeholder for BiocParallel::SerialParam() (the GPU path has no use for it)."""


def doNormAndReduce(sce, batch=None, name=None, ndims=20, hvg_ntop=2000, umap_seed=11, verbose=True,
                    parallel_param=None):
    ep = _red("{cellula::doNormAndReduce()} - ")
    if not isinstance(sce, SingleCellExperiment):
        raise TypeError(ep + "must provide a SingleCellExperiment object")
    if batch is not None:
        if batch not in sce.colData:
            raise ValueError(ep + "batch column not found in the colData of the object")
    if hvg_ntop > sce.nrow():
        raise ValueError(ep + "hvg_ntop cannot be higher than the number of features (nrow) in the object")
    ctx = get_context()
    counts = sce.counts
    n_genes, n_cells = counts.shape
    if verbose:
        _message(_blue("[NORM] "), "Calculating size factors and normalizing.")
    ctx.load_csc(n_genes, counts.p, counts.i, counts.x)
    # Size factors.  The reference runs scran deconvolution first (R/doNormAndReduce.R:57-71);
    # here library-size factors are the default and a pre-set sizeFactor column (e.g. deconvolution
    # factors computed elsewhere) is re-centred exactly as logNormCounts would.
    sf_in = sce.colData.get("sizeFactor") if options["size_factors"] == "given" else None
    if options["size_factors"] == "deconvolution":
        # scran::computeSumFactors on the GPU (R/doNormAndReduce.R:66-71).  quickCluster's walktrap is
        # igraph code that stays outside this library: its labels are read from colData["quickCluster"]; without them
        # the pooling runs on one cluster dealt into parts of <= 3000 cells (computeSumFactors(clusters = NULL)).
        if verbose:
            _message(_blue("[NORM] "), "   Calculating pooled factors.")
        cl = sce.colData.get("quickCluster")
        codes = None if cl is None else np.unique(np.asarray(cl), return_inverse=True)[1].astype(np.int32)
        sf_in, n_bad = ctx.pooled_size_factors(codes)
        if n_bad:
            raise ValueError(ep + f"{n_bad} non-positive size factor estimates (scuttle::cleanSizeFactors is not part of "
                             "this mirror)")
    if verbose and sf_in is None:
        # the R drop-in (r-pkg/) runs the reference's own quickCluster + computeSumFactors by default; this mirror
        # has no scran, so its default is the library-size fast path -- announced, because the numbers differ
        _message(_blue("[NORM] "), "   Library-size factors: cellula's default is scran deconvolution (options['size_factors'] = "
                 "'deconvolution' computes it on the GPU, 'given' takes colData['sizeFactor']); results differ from that default.")
    if verbose:
        _message(_blue("[NORM] "), "   Log-normalization.")
    sf = ctx.size_factors(sf_in)
    if batch is None:
        logx, mean, var = ctx.lognorm_genestats(want_logx=True, want_stats=True)
    else:
        # batch-aware branch (R/doNormAndReduce.R:74-76, :84-86): multiBatchNorm rescales the size factors so that the
        # batches are comparable (per-batch gene averages on the GPU, the B x B median ratios on the host), then
        # modelGeneVar(block=) fits one trend per batch on the per-batch statistics
        levels, codes = np.unique(np.asarray(sce.colData[batch]), return_inverse=True)
        codes = codes.astype(np.int32)
        ave, _, n_per = ctx.block_averages(codes, levels.size)
        sf = trendvar.rescale_size_factors(ave, sf, codes)
        ctx.size_factors_given(sf)
        logx, mean_b, var_b = ctx.lognorm_genestats_blocked(want_logx=True)
    sce.colData["sizeFactor"] = sf
    sce.assays["logcounts"] = CSC(counts.p, counts.i, logx, counts.shape)  # shares i/p like R

    if verbose:
        _message(_blue("[DR] "), "Selecting HVGs.")
    stats = trendvar.model_gene_var(mean, var) if batch is None else trendvar.model_gene_var_blocked(mean_b, var_b, n_per)
    for key in ("mean", "total", "tech", "bio", "p_value", "FDR"):
        if key in stats:
            sce.rowData[key] = stats[key]
    hvg_idx = trendvar.get_top_hvgs(stats, hvg_ntop)
    hvgs = [sce.rownames[i] for i in hvg_idx] if sce.rownames is not None else hvg_idx.copy()
    sce.metadata["hvgs"] = hvgs

    if verbose:
        _message(_blue("[DR] "), "Running PCA.")
    pca = ctx.pca(hvg_idx, int(ndims))
    sce.reducedDims["PCA"] = ReducedDim(pca["scores"], varExplained=pca["var_explained"],
                                        percentVar=pca["percent_var"], rotation=pca["rotation"])

    if not options["skip_umap"]:
        try:
            import umap  # noqa: F401  (umap-learn; the reference calls uwot::umap, R/doNormAndReduce.R:102-106)
        except ImportError:
            if verbose:
                _message(_blue("[DR] "), "UMAP skipped: no UMAP implementation available in this environment.")
        else:
            if verbose:
                _message(_blue("[DR] "), "Running UMAP on uncorrected PCA.")
            reducer = umap.UMAP(n_neighbors=int(math.floor(math.sqrt(n_cells))), min_dist=0.7,
                                random_state=umap_seed)
            sce.reducedDims["UMAP"] = reducer.fit_transform(np.asarray(sce.reducedDims["PCA"])[:, :ndims])
    return sce


def makeSNNGraph(space, k=10, type="rank"):
    """bluster::makeSNNGraph(x, k, type): exact kNN (findKNN, KmknnParam) + SNN edge weights."""
    ctx = get_context()
    space = np.asarray(space, dtype=np.float64)
    ctx.knn(space, int(k), want=False)
    frm, to, w = ctx.snn(None, type, n_total=space.shape[0], k=int(k))
    return SNNGraph(n=space.shape[0], frm=frm, to=to, weight=w, k=int(k), type=type)


def _label_codes(clusters):
    """sort(unique(clusters)) and the 0-based position of every cell's label in it."""
    clusters = np.asarray(clusters)
    uclust = np.unique(clusters)
    return uclust, np.searchsorted(uclust, clusters).astype(np.int32)


def approxSilhouette(x, clusters):
    """bluster::approxSilhouette(x, clusters) on the GPU (R/makeGraphsAndClusters.R:110, :149).
    Returns dict(cluster, other, width) -- the columns of bluster's DataFrame."""
    uclust, codes = _label_codes(clusters)
    width, other = get_context().approx_silhouette(np.asarray(x, dtype=np.float64), codes, uclust.size)
    return dict(cluster=np.asarray(clusters), other=np.where(other >= 0, uclust[np.maximum(other, 0)], uclust[0]),
                width=width)


def pairwiseModularity(graph, clusters, get_weights=False, as_ratio=False):
    """bluster::pairwiseModularity(graph, clusters, get.weights, as.ratio) on the GPU
    (R/makeGraphsAndClusters.R:105, :144, :288).  graph: what makeSNNGraph returns (SNNGraph, or a dict with
    from / to / weight)."""
    uclust, codes = _label_codes(clusters)
    edges = (graph.frm, graph.to, graph.weight) if hasattr(graph, "frm") else (graph["from"], graph["to"], graph["weight"])
    obs, expd, total = get_context().pairwise_modularity(edges, codes, uclust.size)
    if get_weights:
        return dict(observed=obs, expected=expd)
    with np.errstate(invalid="ignore", divide="ignore"):
        return obs / expd if as_ratio else (obs - expd) / total


def _igraph_of(g):
    import igraph
    ig = igraph.Graph(n=g.n, edges=np.column_stack((g.frm - 1, g.to - 1)).tolist(), directed=False)
    ig.es["weight"] = g.weight.tolist()
    return ig


def makeGraphsAndClusters(sce, neighbors=10, weighting_scheme="jaccard", sweep_on="clustering", method="louvain",
                          k=None, dr="PCA", ndims=20, calculate_modularity=True, calculate_silhouette=True,
                          leiden_iterations=5, save_graphs=False, prefix="SNN_", verbose=False):
    ep = _red("{cellula::makeGraphsAndClusters} - ")
    if k is None:
        k = np.linspace(0.1, 1, 6)
    if not isinstance(sce, SingleCellExperiment):
        raise TypeError(ep + "must provide a SingleCellExperiment object")
    if method not in ("leiden", "louvain"):
        raise ValueError(ep + 'method not recognized - must be one of "leiden" or "louvain"')
    if weighting_scheme not in ("jaccard", "rank", "number"):
        raise ValueError(ep + 'method not recognized - must be one of "jaccard", "rank", or "number"')
    if len(sce.reducedDims) == 0:
        raise ValueError(ep + "there are no dimensionality reductions in the SingleCellExperiment object.")
    if dr not in sce.reducedDims:
        raise ValueError(ep + "the dr name supplied was not found in the SingleCellExperiment object")

    # `space` is defined before branching: the reference only assigns it in the first branch and
    # its "SNN" sweep fails with "object 'space' not found" (R/makeGraphsAndClusters.R:80 vs :125).
    space = np.asarray(sce.reducedDims[dr])
    if space.shape[1] > ndims:
        space = space[:, :ndims]

    try:
        import igraph  # noqa: F401
        have_igraph = True
    except ImportError:
        have_igraph = False

    def cluster(g, resolution):
        ig = _igraph_of(g)
        if method == "louvain":
            return ig.community_multilevel(weights="weight", resolution=resolution).membership
        return ig.community_leiden(objective_function="modularity", weights="weight",
                                   n_iterations=leiden_iterations, resolution=resolution).membership

    def record(g, cl, gname):
        """colData labels + the per-resolution diagnostics of R/makeGraphsAndClusters.R:98-114 (both sweeps)."""
        sce.colData[gname] = [str(c + 1) for c in cl]
        if calculate_modularity:
            sce.metadata[f"modularity_{gname}"] = pairwiseModularity(g, cl, as_ratio=True)
        if calculate_silhouette:
            sil = approxSilhouette(space, cl)
            sil["closest"] = np.where(sil["width"] > 0, np.asarray(cl), sil["other"])
            sce.metadata[f"silhouette_{gname}"] = sil

    if sweep_on == "clustering" and neighbors is not None:
        if verbose:
            _message(_blue("[CLU]"), "Creating SNN graph.")
        g = makeSNNGraph(space, k=neighbors, type=weighting_scheme)
        key = f"SNN_{neighbors}_{weighting_scheme}"
        if save_graphs or not have_igraph:
            sce.metadata[key] = g
        if not have_igraph:
            if verbose:
                _message(_blue("[CLU]"), "python-igraph is not installed: graph stored in metadata['", key,
                         "'], clustering sweep skipped (in R the graph goes to igraph unchanged).")
            return sce
        for res in k:
            if verbose:
                _message(_blue("[CLU]"), "Clustering at resolution ", res, ".")
            record(g, cluster(g, float(res)), f"{prefix}{res:g}")
    elif sweep_on == "SNN" and k is not None:
        for kk in np.floor(np.asarray(k)).astype(int):
            if verbose:
                _message(_blue("[CLU]"), "Creating SNN graph with k =", kk, "neighbors.")
            g = makeSNNGraph(space, k=int(kk), type=weighting_scheme)
            if save_graphs or not have_igraph:
                sce.metadata[f"SNN_{kk}_{weighting_scheme}"] = g
            if have_igraph:
                record(g, cluster(g, 1.0), f"{prefix}{kk}")
    return sce


def rebuildModularity(sce, dr="PCA", neighbors=None, ndims=20, labels=None, weighting_scheme="jaccard"):
    """cellula::rebuildModularity: the same makeSNNGraph call with k = round(sqrt(N)) by default, then the pairwise
    modularity of colData[labels] on that graph (as.ratio = TRUE), stored as metadata['modularity_<labels>']."""
    ep = _red("{cellula::rebuildModularity} - ")
    if dr not in sce.reducedDims:
        raise ValueError(ep + "the dr name supplied was not found in the SingleCellExperiment object")
    if weighting_scheme not in ("jaccard", "rank", "number"):
        raise ValueError(ep + 'method not recognized - must be one of "jaccard", "rank", or "number"')
    space = np.asarray(sce.reducedDims[dr])[:, :ndims]
    if neighbors is None:
        neighbors = int(round(math.sqrt(space.shape[0])))
    g = makeSNNGraph(space, k=neighbors, type=weighting_scheme)
    sce.metadata[f"SNN_{neighbors}_{weighting_scheme}"] = g
    if labels is not None:     # R/makeGraphsAndClusters.R:286-288
        sce.metadata[f"modularity_{labels}"] = pairwiseModularity(g, np.asarray(sce.colData[labels]), as_ratio=True)
    return sce
