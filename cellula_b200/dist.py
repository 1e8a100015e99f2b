"""Cell-sharded execution of the hot path: one process per GPU, `torch.distributed` for the
plumbing.

Every rank holds a contiguous block of cells (dgCMatrix columns).  The arithmetic of a stage is
rank-local (libcellula_b200.so through `CudaShard`); between stages this module performs the
small exchanges the path needs:

    size factors   all-reduce  [sum of library sizes, #cells]           (2 doubles)
    gene stats     all-reduce  [sum y, sum y^2] per gene                (2 G int64, fixed point: exact)
    PCA            all-reduce  H x H cross-product                      (2 H^2 int64, fixed point: exact)
                   all-gather  scores (N x d doubles)
    kNN            all-gather  neighbour index (N x k int32)
    SNN            none: every rank emits the edges of a contiguous range of cells j (ascending j; the ranges
                   are balanced by work, `snn_bounds`), so the concatenation in rank order is the single-GPU
                   edge list.

`ShardedPipeline` only talks to a *shard backend* (the methods of `CudaShard`) and to
`torch.distributed`, so the same driver runs on CPU with the gloo backend and a test double as
backend (tests/test_dist_gloo.py) -- that is how the N>1 host logic is covered without GPUs.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist

from . import trendvar


def shard_bounds(n_total: int, world: int):
    """Contiguous, as-equal-as-possible split of [0, n_total) into `world` blocks."""
    base, rem = divmod(int(n_total), int(world))
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


def snn_bounds(n_total: int, world: int):
    """Contiguous ranges of cells j for the SNN stage with about equal WORK per rank.  A cell j is joined to
    partners h < j only, so its edge count (and its share of the list scanning) grows like j: equal work
    means boundaries at N sqrt(r / R), not N r / R.  Any contiguous split keeps the contract that the
    concatenation in rank order is the single-GPU edge list."""
    b = [int(round(int(n_total) * math.sqrt(r / world))) for r in range(world)] + [int(n_total)]
    for r in range(1, world + 1):          # strictly increasing even for tiny inputs
        b[r] = min(max(b[r], b[r - 1] + 1), int(n_total) - (world - r))
    return b


def slice_csc(colptr, rowidx, x, c0, c1):
    """Columns [c0, c1) of a CSC matrix with the column pointers rebased to 0."""
    colptr = np.asarray(colptr)
    lo, hi = int(colptr[c0]), int(colptr[c1])
    return (colptr[c0:c1 + 1] - lo).astype(np.int64), rowidx[lo:hi], x[lo:hi]


class _DevView:
    """Zero-copy torch view of a device buffer owned by the library."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False),
                                             version=3, strides=None)


def device_tensor(ptr, shape, dtype, device):
    typestr = {torch.float64: "<f8", torch.int32: "<i4", torch.int64: "<i8"}[dtype]
    return torch.as_tensor(_DevView(ptr, shape, typestr), device=device)


class CudaShard:
    """Shard backend over one cb200 context (one GPU)."""

    def __init__(self, ctx, device):
        self.ctx = ctx
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        self.check_stream()

    def check_stream(self):
        """STREAM CONTRACT.  torch.distributed orders its collectives against torch's *current* stream only, while the
        library launches its kernels on the context's stream.  The collectives of ShardedPipeline run directly on the
        library's buffers, so the two must be the same stream: build the context with
        ``Context(device, stream=torch.cuda.current_stream(device).cuda_stream)`` on a ``torch.cuda.Stream`` that is
        current (``torch.cuda.set_stream``) whenever the pipeline is driven, as bench.py and tests/dist_worker.py do.
        Anything else (a context with its own private stream, or torch's legacy default stream, handle 0, which the C
        ABI cannot be asked to use) would let a collective read a buffer before the kernel that fills it has finished:
        refused here instead of producing silently wrong output."""
        cur = torch.cuda.current_stream(self.device).cuda_stream
        mine = getattr(self.ctx, "stream_handle", None)
        if mine is None or int(mine) != int(cur):
            raise RuntimeError(
                "CudaShard: the cb200 context does not run on torch's current CUDA stream (context stream "
                f"{mine}, torch current stream {cur}); create it as Context(device, stream=s.cuda_stream) for a "
                "torch.cuda.Stream s and make s current (torch.cuda.set_stream(s)) while the pipeline runs")

    def _view(self, which, shape, dtype):
        ptr, _ = self.ctx.device_ptr(which)
        return device_tensor(ptr, shape, dtype, self.device)

    # K1
    def libsize_partial(self, sf_in=None):
        self.ctx.size_factors_local(sf_in)
        return self._view("libsize_partial", (2,), torch.float64)

    def finish_size_factors(self, total, n_total, want=True):
        return self.ctx.size_factors_finish(total, n_total, want=want)

    # K2 + K3
    def gene_acc(self):
        self.ctx.lognorm_genestats_local()
        # 64-bit fixed-point sums: the all-reduce is an integer sum, exact for any number of ranks
        return self._view("gene_acc", (2 * self.ctx.n_genes,), torch.int64)

    def finish_genestats(self):
        return self.ctx.genestats_finish(want=True)

    def lognorm_values(self):
        self.ctx.lognorm_values()

    def logcounts(self, out=None):
        return self.ctx.fetch_logcounts(out)

    def logcounts_async(self, out):
        self.ctx.fetch_logcounts_async(out)

    def wait_copies(self):
        self.ctx.wait_copies()

    # K5
    def gram(self, hvg_idx):
        self.ctx.pca_gram_local(hvg_idx)
        ptr, nbytes = self.ctx.device_ptr("gram")
        h = len(hvg_idx)
        if nbytes == 2 * h * h * 8:     # tensor-core engine: 64-bit fixed-point planes, exact integer all-reduce
            return device_tensor(ptr, (2 * h * h,), torch.int64, self.device)
        return device_tensor(ptr, (h * h,), torch.float64, self.device)

    def finish_pca(self, d, n_hvg, want_scores=True, scores_out=None):
        res = self.ctx.pca_finish(d, n_hvg, want_scores=want_scores, scores_out=scores_out)
        res["scores_dev"] = self._view("scores", (self.ctx.n_cells, d), torch.float64)
        return res

    # K6
    def knn(self, emb_full, k, q_begin, q_end, want=False):
        n, d = emb_full.shape
        self.ctx.knn_device(emb_full.data_ptr(), n, d, k, q_begin, q_end, want=want)
        return self._view("knn", (q_end - q_begin, k), torch.int32)

    # K7
    def snn(self, knn_full, snn_type, j_begin, j_end, fetch=True, edge_buffers=None):
        n, k = knn_full.shape
        if fetch and edge_buffers is not None:
            # count first, then emit straight into the caller's (pinned) arrays, copies behind the emission
            ne = self.ctx.snn_count_device(knn_full.data_ptr(), n, k, snn_type, j_begin, j_end)
            f, t, w = edge_buffers(ne)
            return self.ctx.snn_emit(ne, f, t, w)
        return self.ctx.snn_device(knn_full.data_ptr(), n, k, snn_type, j_begin, j_end, fetch=fetch)


class ShardedPipeline:
    """counts shard -> SNN edges of the shard's cells, with the exchanges listed above."""

    def __init__(self, shard, n_total, group=None):
        self.shard = shard
        self.n_total = int(n_total)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bounds = shard_bounds(self.n_total, self.world)
        self.c0, self.c1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        sb = snn_bounds(self.n_total, self.world)
        self.j0, self.j1 = sb[self.rank], sb[self.rank + 1]     # cells whose SNN edges this rank emits

    # -- collectives ---------------------------------------------------------------
    def _allreduce(self, t):
        if self.world > 1:
            if hasattr(self.shard, "check_stream"):
                self.shard.check_stream()
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def _allgather_rows(self, local):
        """Concatenate row blocks of unequal height in rank order."""
        if self.world == 1:
            return local
        if hasattr(self.shard, "check_stream"):
            self.shard.check_stream()
        sizes = [self.bounds[r + 1] - self.bounds[r] for r in range(self.world)]
        mx = max(sizes)
        pad = local
        if local.shape[0] < mx:
            pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            pad[:local.shape[0]] = local
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(parts, pad.contiguous(), group=self.group)
        return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0).contiguous()

    # -- stages ----------------------------------------------------------------------
    def size_factors(self, sf_in=None, want=True):
        part = self._allreduce(self.shard.libsize_partial(sf_in))
        if part.is_cuda:   # pinned landing zone: a pageable read would queue behind background copies
            if getattr(self, "_pin2", None) is None:
                self._pin2 = torch.empty(2, dtype=torch.float64).pin_memory()
            self._pin2.copy_(part, non_blocking=True)
            torch.cuda.current_stream(part.device).synchronize()
            tot = self._pin2.numpy().copy()
        else:
            tot = part.cpu().numpy()
        return self.shard.finish_size_factors(float(tot[0]), int(round(float(tot[1]))), want=want)

    def gene_stats(self):
        self._allreduce(self.shard.gene_acc())
        return self.shard.finish_genestats()

    def pca(self, hvg_idx, d, want_scores=True, scores_out=None):
        self._allreduce(self.shard.gram(hvg_idx))
        if scores_out is not None:
            res = self.shard.finish_pca(d, len(hvg_idx), want_scores=True, scores_out=scores_out)
        else:
            res = self.shard.finish_pca(d, len(hvg_idx), want_scores=want_scores)
        res["scores_full_dev"] = self._allgather_rows(res["scores_dev"])
        return res

    def knn(self, emb_full, k):
        local = self.shard.knn(emb_full, k, self.c0, self.c1)
        return self._allgather_rows(local)

    def snn(self, knn_full, snn_type, fetch=True, edge_buffers=None):
        if fetch and edge_buffers is not None:
            return self.shard.snn(knn_full, snn_type, self.j0, self.j1, fetch=True, edge_buffers=edge_buffers)
        if fetch:
            return self.shard.snn(knn_full, snn_type, self.j0, self.j1)
        return self.shard.snn(knn_full, snn_type, self.j0, self.j1, fetch=False)

    def run(self, ndims=50, hvg_ntop=2000, neighbors=20, snn_type="rank", sf_in=None, want_host=True,
            scores_out=None, edge_buffers=None, logx_out=None):
        """The whole path for this rank's shard.  Returns host results of the shard.
        scores_out: optional column-major (n_local x ndims) host array (e.g. pinned) for the scores;
        edge_buffers: optional callable n_edges -> (from, to, weight) host arrays to fill;
        logx_out: optional pinned array for the logcounts, copied out behind the following stages."""
        import time
        marks = [("start", time.perf_counter())]
        sf = self.size_factors(sf_in, want=want_host)
        marks.append(("size_factors", time.perf_counter()))
        if logx_out is not None and hasattr(self.shard, "lognorm_values"):
            # logcounts first (values + column pointers only) and on their way to the host while the row
            # indices are still being uploaded; the gene sums follow
            self.shard.lognorm_values()
            self.shard.logcounts_async(logx_out)            # overlaps everything that follows
        mean, var = self.gene_stats()
        marks.append(("gene_stats", time.perf_counter()))
        if logx_out is not None and not hasattr(self.shard, "lognorm_values"):
            self.shard.logcounts_async(logx_out)            # overlaps the trend fit, PCA, kNN and SNN
        stats = trendvar.model_gene_var(mean, var, want_pvalues=False)   # replicated host step (R: scran::fitTrendVar)
        hvgs = trendvar.get_top_hvgs(stats, hvg_ntop)
        marks.append(("trend_hvg", time.perf_counter()))
        pca = self.pca(hvgs, ndims, want_scores=want_host, scores_out=scores_out if want_host else None)
        marks.append(("pca", time.perf_counter()))
        knn_full = self.knn(pca["scores_full_dev"], neighbors)
        marks.append(("knn", time.perf_counter()))
        # (from, to, w), or the count if resident
        edges = self.snn(knn_full, snn_type, fetch=want_host, edge_buffers=edge_buffers)
        marks.append(("snn", time.perf_counter()))
        if logx_out is not None:
            self.shard.wait_copies()
        marks.append(("wait_copies", time.perf_counter()))
        # host wall-clock split of this call (ms since entry; device work is asynchronous between syncs)
        self.last_marks = {k: 1e3 * (t - marks[0][1]) for k, t in marks[1:]}
        return dict(sf=sf, mean=mean, var=var, stats=stats, hvgs=hvgs, pca=pca, knn_full=knn_full,
                    edges=edges)
