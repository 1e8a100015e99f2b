"""ctypes binding of libcellula_b200.so (the plain C ABI of include/cellula_b200.h).

This is the Python counterpart of the R `.Call` shim (r-pkg/src/rshim.c): it only converts
arguments.  There is no CPU fallback anywhere below: if the shared library is missing or no
B200 is present, the calls raise.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcellula_b200.so")
CSRC = os.path.join(_HERE, "csrc")

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_f64p = ctypes.POINTER(ctypes.c_double)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_vp = ctypes.c_void_p

SNN_TYPES = {"rank": 0, "number": 1, "jaccard": 2}
BUF = dict(libsize_partial=0, gene_acc=1, gram=2, scores=3, knn=4, size_factors=5, logcounts=6, block_part=7, block_acc=8)


class Timing(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in (
        "h2d_ms", "colsum_ms", "lognorm_ms", "hvg_extract_ms", "gram_ms", "eig_ms", "project_ms",
        "knn_prep_ms", "knn_tile_ms", "knn_rerank_ms", "snn_revlist_ms", "snn_edges_ms", "d2h_ms")] + [
        ("kernel_launches", c_i64), ("knn_fallback_queries", c_i64), ("eig_iterations", c_i64),
        ("gram_engine", c_i64), ("knn_tiles_scanned", c_i64), ("knn_tiles_total", c_i64),
        ("knn_engine", c_i64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/cellula_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "cb200_init": (c_int, [c_int, c_vp, ctypes.POINTER(c_vp)]),
    "cb200_free": (c_int, [c_vp]),
    "cb200_last_error": (ctypes.c_char_p, [c_vp]),
    "cb200_version": (ctypes.c_char_p, []),
    "cb200_synchronize": (c_int, [c_vp]),
    "cb200_stage_timings": (c_int, [c_vp, ctypes.POINTER(Timing)]),
    "cb200_reset_timings": (c_int, [c_vp]),
    "cb200_device_ptr": (c_int, [c_vp, c_int, ctypes.POINTER(c_vp), c_i64p]),
    "cb200_load_csc": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64]),
    "cb200_load_csc_i32": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64]),
    "cb200_size_factors_local": (c_int, [c_vp, c_vp]),
    "cb200_size_factors_finish": (c_int, [c_vp, ctypes.c_double, c_i64, c_vp]),
    "cb200_size_factors": (c_int, [c_vp, c_vp, c_vp]),
    "cb200_lognorm_genestats_local": (c_int, [c_vp]),
    "cb200_genestats_finish": (c_int, [c_vp, c_vp, c_vp]),
    "cb200_fetch_logcounts": (c_int, [c_vp, c_vp]),
    "cb200_fetch_logcounts_async": (c_int, [c_vp, c_vp]),
    "cb200_wait_copies": (c_int, [c_vp]),
    "cb200_lognorm_genestats": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "cb200_lognorm_values": (c_int, [c_vp]),
    "cb200_set_blocks": (c_int, [c_vp, c_vp, c_int]),
    "cb200_block_prepare_local": (c_int, [c_vp]),
    "cb200_block_averages_local": (c_int, [c_vp, ctypes.c_double]),
    "cb200_block_averages_finish": (c_int, [c_vp, c_vp, c_vp]),
    "cb200_block_averages": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "cb200_size_factors_given": (c_int, [c_vp, c_vp]),
    "cb200_lognorm_genestats_blocked_local": (c_int, [c_vp]),
    "cb200_genestats_blocked_finish": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "cb200_lognorm_genestats_blocked": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "cb200_pca_gram_local": (c_int, [c_vp, c_vp, c_int]),
    "cb200_pca_finish": (c_int, [c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "cb200_pca": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "cb200_knn": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_i64, c_i64, c_vp]),
    "cb200_knn_device": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_i64, c_i64, c_vp]),
    "cb200_snn": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_i64, c_i64, c_i64p,
                          ctypes.POINTER(c_i32p), ctypes.POINTER(c_i32p), ctypes.POINTER(c_f64p)]),
    "cb200_snn_device": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_i64, c_i64, c_i64p,
                                 ctypes.POINTER(c_i32p), ctypes.POINTER(c_i32p), ctypes.POINTER(c_f64p)]),
    "cb200_snn_resident": (c_int, [c_vp, c_int, c_int, c_i64p]),
    "cb200_snn_count_device": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_i64, c_i64, c_i64p]),
    "cb200_snn_emit": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "cb200_free_edges": (c_int, [c_i32p, c_i32p, c_f64p]),
    "cb200_approx_silhouette": (c_int, [c_vp, c_vp, c_i64, c_int, c_vp, c_int, c_vp, c_vp]),
    "cb200_pairwise_modularity": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_vp, c_vp]),
    "cb200_fetch_edges": (c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "cb200_edges_checksum": (c_int, [c_vp, c_i64p, ctypes.POINTER(ctypes.c_uint64)]),
    "cb200_pooled_size_factors": (c_int, [c_vp, c_vp, c_vp, c_int, ctypes.c_double, c_i64, c_vp, c_i64p, c_vp]),
    "cb200_qc_metrics": (c_int, [c_vp, c_int, c_vp, c_vp, ctypes.c_double, c_vp, c_vp, c_vp, c_vp]),
    "cb200_knn_distances": (c_int, [c_vp, c_vp]),
    "cb200_multi_init": (c_int, [c_int, ctypes.POINTER(c_vp)]),
    "cb200_multi_free": (c_int, [c_vp]),
    "cb200_multi_last_error": (ctypes.c_char_p, [c_vp]),
    "cb200_multi_n_gpus": (c_int, [c_vp]),
    "cb200_multi_stage_timings": (c_int, [c_vp, c_int, ctypes.POINTER(Timing)]),
    "cb200_multi_load_csc": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "cb200_multi_load_csc_i32": (c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "cb200_multi_size_factors": (c_int, [c_vp, c_vp, c_vp]),
    "cb200_multi_lognorm_genestats": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "cb200_multi_pca": (c_int, [c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "cb200_multi_knn": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp]),
    "cb200_multi_snn_count": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_i64p]),
    "cb200_multi_snn_emit": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "cb200_multi_edges_checksum": (c_int, [c_vp, c_i64p, ctypes.POINTER(ctypes.c_uint64)]),
    "cb200_synth_counts": (c_int, [c_vp, c_i64, c_i64, ctypes.c_double, ctypes.c_uint64, c_i64, c_i64, c_i64p]),
    "cb200_export_csc": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "cb200_synth_host": (c_int, [c_i64, c_i64, ctypes.c_double, ctypes.c_uint64, c_i64, c_i64, c_int, c_i64p,
                                 ctypes.POINTER(c_i64p), ctypes.POINTER(c_i32p), ctypes.POINTER(c_f64p)]),
    "cb200_synth_host_free": (None, [c_i64p, c_i32p, c_f64p]),
    "cb200_synth_embedding_host": (c_int, [c_i64, c_int, ctypes.c_uint64, c_int, c_vp]),
    "cb200_host_weighted_lowess": (c_int, [c_i64, c_vp, c_vp, c_vp, ctypes.c_double, c_int, c_int, c_vp]),
    "cb200_host_nls_trend": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, ctypes.c_double, c_int]),
    "cb200_host_nls_grid": (c_int, [c_i64, c_vp, c_vp, ctypes.c_double, c_int, c_vp, c_int, c_vp, c_vp, c_vp]),
    "cb200_host_pool_layout": (c_int, [c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cb200_host_gauss_kde": (c_int, [c_i64, c_vp, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_int, c_vp]),
}

_lib = None


class Cb200Error(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into cellula_b200/libcellula_b200.so (nvcc, in-tree)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise Cb200Error("building libcellula_b200.so failed")
    return LIB_PATH


def lib():
    """The loaded shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Cb200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). cellula_b200 has no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return c_vp(a)
    return c_vp(a.ctypes.data)


class Context:
    """One cb200_ctx = one GPU (one process per GPU in a multi-GPU job)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = c_vp()
        L = lib()
        rc = L.cb200_init(device, c_vp(stream) if stream else None, ctypes.byref(self._h))
        if rc != 0:
            raise Cb200Error(L.cb200_last_error(None).decode())
        self.device = device
        # the cudaStream_t the context runs on when the caller supplied one (None: a private non-blocking stream made by
        # the library -- note that handle 0, the legacy default stream, cannot be requested through this ABI)
        self.stream_handle = int(stream) if stream else None
        self.n_genes = self.n_cells = self.nnz = 0
        self.cell_offset = 0
        self.n_cells_total = 0

    # -- plumbing -------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise Cb200Error(lib().cb200_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value is not None:
            lib().cb200_free(self._h)
            self._h = c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def synchronize(self):
        self._ck(lib().cb200_synchronize(self._h))

    def timings(self) -> dict:
        t = Timing()
        self._ck(lib().cb200_stage_timings(self._h, ctypes.byref(t)))
        return t.as_dict()

    def reset_timings(self):
        self._ck(lib().cb200_reset_timings(self._h))

    def device_ptr(self, which: str):
        p, n = c_vp(), c_i64()
        self._ck(lib().cb200_device_ptr(self._h, BUF[which], ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    # -- a1 -------------------------------------------------------------------------
    def load_csc(self, n_genes, colptr, rowidx, values, cell_offset=0, n_cells_total=None):
        colptr = np.ascontiguousarray(colptr)
        rowidx = np.ascontiguousarray(rowidx, dtype=np.int32)
        values = np.ascontiguousarray(values, dtype=np.float64)
        n_cells = colptr.size - 1
        if n_cells_total is None:
            n_cells_total = n_cells
        if colptr.dtype == np.int32:
            fn = lib().cb200_load_csc_i32
        else:
            colptr = np.ascontiguousarray(colptr, dtype=np.int64)
            fn = lib().cb200_load_csc
        self._ck(fn(self._h, n_genes, n_cells, _ptr(colptr), _ptr(rowidx), _ptr(values), cell_offset,
                    n_cells_total))
        self.n_genes, self.n_cells, self.nnz = int(n_genes), int(n_cells), int(values.size)
        self.cell_offset, self.n_cells_total = int(cell_offset), int(n_cells_total)

    def load_csc_ptrs(self, n_genes, n_cells, nnz, colptr_ptr, rowidx_ptr, values_ptr, cell_offset=0,
                      n_cells_total=None):
        """Same, from raw host addresses (pinned torch buffers in bench.py)."""
        if n_cells_total is None:
            n_cells_total = n_cells
        self._ck(lib().cb200_load_csc(self._h, n_genes, n_cells, c_vp(colptr_ptr), c_vp(rowidx_ptr),
                                      c_vp(values_ptr), cell_offset, n_cells_total))
        self.n_genes, self.n_cells, self.nnz = int(n_genes), int(n_cells), int(nnz)
        self.cell_offset, self.n_cells_total = int(cell_offset), int(n_cells_total)

    def synth_counts(self, n_genes, n_cells, nnz_per_cell, seed, cell_offset=0, n_cells_total=None):
        if n_cells_total is None:
            n_cells_total = n_cells
        nnz = c_i64()
        self._ck(lib().cb200_synth_counts(self._h, n_genes, n_cells, float(nnz_per_cell), seed, cell_offset,
                                          n_cells_total, ctypes.byref(nnz)))
        self.n_genes, self.n_cells, self.nnz = int(n_genes), int(n_cells), int(nnz.value)
        self.cell_offset, self.n_cells_total = int(cell_offset), int(n_cells_total)
        return self.nnz

    def export_csc(self, colptr=None, rowidx=None, values=None):
        colptr = np.empty(self.n_cells + 1, dtype=np.int64) if colptr is None else colptr
        rowidx = np.empty(self.nnz, dtype=np.int32) if rowidx is None else rowidx
        values = np.empty(self.nnz, dtype=np.float64) if values is None else values
        self._ck(lib().cb200_export_csc(self._h, _ptr(colptr), _ptr(rowidx), _ptr(values)))
        return colptr, rowidx, values

    # -- a2 -------------------------------------------------------------------------
    def size_factors(self, sf_in=None, want=True):
        sf_in = None if sf_in is None else np.ascontiguousarray(sf_in, dtype=np.float64)
        out = np.empty(self.n_cells) if want else None
        self._ck(lib().cb200_size_factors(self._h, _ptr(sf_in), _ptr(out)))
        return out

    def size_factors_local(self, sf_in=None):
        sf_in = None if sf_in is None else np.ascontiguousarray(sf_in, dtype=np.float64)
        self._ck(lib().cb200_size_factors_local(self._h, _ptr(sf_in)))

    def size_factors_finish(self, total, n_cells_total, want=True):
        out = np.empty(self.n_cells) if want else None
        self._ck(lib().cb200_size_factors_finish(self._h, float(total), int(n_cells_total), _ptr(out)))
        return out

    # -- a3 + a4 ----------------------------------------------------------------------
    def lognorm_genestats(self, want_logx=True, want_stats=True, logx_out=None):
        logx = logx_out if logx_out is not None else (np.empty(self.nnz) if want_logx else None)
        mean = np.empty(self.n_genes) if want_stats else None
        var = np.empty(self.n_genes) if want_stats else None
        self._ck(lib().cb200_lognorm_genestats(self._h, _ptr(logx), _ptr(mean), _ptr(var)))
        return logx, mean, var

    def lognorm_values(self):
        """Logcounts only (no gene sums yet): runs while the row indices are still being uploaded."""
        self._ck(lib().cb200_lognorm_values(self._h))

    def lognorm_genestats_local(self):
        self._ck(lib().cb200_lognorm_genestats_local(self._h))

    def genestats_finish(self, want=True):
        mean = np.empty(self.n_genes) if want else None
        var = np.empty(self.n_genes) if want else None
        self._ck(lib().cb200_genestats_finish(self._h, _ptr(mean), _ptr(var)))
        return mean, var

    def fetch_logcounts(self, out=None):
        out = np.empty(self.nnz) if out is None else out
        self._ck(lib().cb200_fetch_logcounts(self._h, _ptr(out)))
        return out

    def fetch_logcounts_async(self, out):
        self._ck(lib().cb200_fetch_logcounts_async(self._h, _ptr(out)))

    def wait_copies(self):
        self._ck(lib().cb200_wait_copies(self._h))

    # -- 8(f)-1: deconvolution size factors ----------------------------------------------
    def pooled_size_factors(self, clusters=None, sizes=None, min_mean=None, max_cluster_size=3000):
        """scran::computeSumFactors(clusters=...) on the resident matrix.  Returns (sf scaled to unit mean over the
        positive factors, number of non-positive factors); stage times in self.pool_ms."""
        cl = None if clusters is None else np.ascontiguousarray(clusters, dtype=np.int32)
        assert cl is None or cl.size == self.n_cells
        sz = None if sizes is None else np.ascontiguousarray(sorted(sizes), dtype=np.int32)
        out = np.empty(self.n_cells)
        nneg = c_i64()
        ms = np.zeros(8)
        self._ck(lib().cb200_pooled_size_factors(self._h, _ptr(cl), _ptr(sz), 0 if sz is None else int(sz.size),
                                                 -1.0 if min_mean is None else float(min_mean),
                                                 0 if max_cluster_size is None else int(max_cluster_size), _ptr(out),
                                                 ctypes.byref(nneg), _ptr(ms)))
        self.pool_ms = dict(zip(("pseudo_cells_ms", "pool_medians_ms", "solve_ms", "host_rescale_ms", "total_host_ms", "kept_genes_max",
                                 "clusters", "kept_entries"),
                                ms.tolist()))
        return out, int(nneg.value)

    # -- 8(f)-4: per-cell QC metrics, kNN distances ---------------------------------------
    def qc_metrics(self, subsets=(), threshold=0.0):
        """scuttle::perCellQCMetrics: dict(sum, detected, subsets=[dict(sum, detected, percent), ...])."""
        subsets = [np.ascontiguousarray(g, dtype=np.int32) for g in subsets]
        S = len(subsets)
        ptr = np.zeros(S + 1, dtype=np.int32)
        if S:
            ptr[1:] = np.cumsum([g.size for g in subsets])
        genes = np.concatenate(subsets).astype(np.int32) if S and ptr[-1] > 0 else np.zeros(1, dtype=np.int32)
        n = self.n_cells
        tot, det = np.empty(n), np.empty(n, dtype=np.int32)
        ssum = np.empty((max(S, 1), n))
        sdet = np.empty((max(S, 1), n), dtype=np.int32)
        self._ck(lib().cb200_qc_metrics(self._h, S, _ptr(ptr), _ptr(genes), float(threshold), _ptr(tot), _ptr(det),
                                        _ptr(ssum), _ptr(sdet)))
        with np.errstate(divide="ignore", invalid="ignore"):
            subs = [dict(sum=ssum[s], detected=sdet[s], percent=100.0 * ssum[s] / tot) for s in range(S)]
        return dict(sum=tot, detected=det, subsets=subs)

    def knn_distances(self):
        """Distances of the resident kNN result (n_queries x k, findKNN(get.distance=TRUE)$distance)."""
        nq, k = self._knn_shape
        out = np.empty((nq, k), order="F")
        self._ck(lib().cb200_knn_distances(self._h, _ptr(out)))
        return out

    # -- 8(f)-3: batch-aware branch ----------------------------------------------------
    def block_averages(self, block, n_blocks):
        """Per-batch gene averages of x / (sf centred within the batch): (ave B x G, mean sf per batch, cells per batch)."""
        block = np.ascontiguousarray(block, dtype=np.int32)
        assert block.size == self.n_cells
        ave = np.empty((n_blocks, self.n_genes))
        sfm = np.empty(n_blocks)
        nb = np.empty(n_blocks, dtype=np.int64)
        self._ck(lib().cb200_block_averages(self._h, _ptr(block), int(n_blocks), _ptr(ave), _ptr(sfm), _ptr(nb)))
        self.n_blocks = int(n_blocks)
        return ave, sfm, nb

    def set_blocks(self, block, n_blocks):
        block = np.ascontiguousarray(block, dtype=np.int32)
        assert block.size == self.n_cells
        self._ck(lib().cb200_set_blocks(self._h, _ptr(block), int(n_blocks)))
        self.n_blocks = int(n_blocks)

    def size_factors_given(self, sf):
        """Size factors used as they are (multiBatchNorm's rescaled factors): no centring."""
        sf = np.ascontiguousarray(sf, dtype=np.float64)
        assert sf.size == self.n_cells
        self._ck(lib().cb200_size_factors_given(self._h, _ptr(sf)))

    def lognorm_genestats_blocked(self, want_logx=True):
        """(logcounts, mean B x G, var B x G): statistics within every block set by set_blocks / block_averages."""
        logx = np.empty(self.nnz) if want_logx else None
        mean = np.empty((self.n_blocks, self.n_genes))
        var = np.empty((self.n_blocks, self.n_genes))
        self._ck(lib().cb200_lognorm_genestats_blocked(self._h, _ptr(logx), _ptr(mean), _ptr(var)))
        return logx, mean, var

    # -- a7 -------------------------------------------------------------------------
    def pca(self, hvg_idx, d, want_scores=True, scores_out=None):
        hvg_idx = np.ascontiguousarray(hvg_idx, dtype=np.int32)
        self.pca_gram_local(hvg_idx)
        return self.pca_finish(d, len(hvg_idx), want_scores=want_scores, scores_out=scores_out)

    def pca_gram_local(self, hvg_idx):
        hvg_idx = np.ascontiguousarray(hvg_idx, dtype=np.int32)
        self._ck(lib().cb200_pca_gram_local(self._h, _ptr(hvg_idx), len(hvg_idx)))

    def pca_finish(self, d, n_hvg, want_scores=True, scores_out=None):
        scores = scores_out if scores_out is not None else (
            np.empty((self.n_cells, d), order="F") if want_scores else None)
        rotation = np.empty((n_hvg, d), order="F")
        varexp = np.empty(d)
        total = ctypes.c_double()
        self._ck(lib().cb200_pca_finish(self._h, d, _ptr(scores), _ptr(rotation), _ptr(varexp),
                                        ctypes.cast(ctypes.byref(total), c_vp)))
        return dict(scores=scores, rotation=rotation, var_explained=varexp, total_var=total.value,
                    percent_var=100.0 * varexp / total.value)

    # -- a9 -------------------------------------------------------------------------
    def knn(self, emb, k, q_begin=0, q_end=None, want=True, n_total=None, d=None, out=None):
        """emb: N x d array (any layout; passed column-major) or None for the resident PCA scores."""
        if emb is not None:
            emb = np.asfortranarray(emb, dtype=np.float64)
            n_total, d = emb.shape
        q_end = n_total if q_end is None else q_end
        idx = out if out is not None else (np.empty((q_end - q_begin, k), dtype=np.int32, order="F") if want else None)
        self._ck(lib().cb200_knn(self._h, _ptr(emb), n_total, d, k, q_begin, q_end, _ptr(idx)))
        self._knn_shape = (q_end - q_begin, k)
        return idx

    def knn_device(self, d_emb_ptr, n_total, d, k, q_begin, q_end, want=True):
        idx = np.empty((q_end - q_begin, k), dtype=np.int32, order="F") if want else None
        self._ck(lib().cb200_knn_device(self._h, c_vp(d_emb_ptr), n_total, d, k, q_begin, q_end, _ptr(idx)))
        self._knn_shape = (q_end - q_begin, k)
        return idx

    # -- a10 ------------------------------------------------------------------------
    def _take_edges(self, ne, pf, pt, pw):
        m = ne.value
        if m > 0:
            f = np.ctypeslib.as_array(pf, shape=(m,)).copy()
            t = np.ctypeslib.as_array(pt, shape=(m,)).copy()
            w = np.ctypeslib.as_array(pw, shape=(m,)).copy()
        else:
            f, t, w = np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0)
        lib().cb200_free_edges(pf, pt, pw)
        return f, t, w

    def snn(self, index, snn_type="rank", j_begin=0, j_end=None, n_total=None, k=None):
        """index: N x k 1-based (findKNN layout) or None for the resident kNN result."""
        if index is not None:
            index = np.asfortranarray(index, dtype=np.int32)
            n_total, k = index.shape
        j_end = n_total if j_end is None else j_end
        ne = c_i64()
        pf, pt, pw = c_i32p(), c_i32p(), c_f64p()
        self._ck(lib().cb200_snn(self._h, _ptr(index), n_total, k, SNN_TYPES[snn_type], j_begin, j_end,
                                 ctypes.byref(ne), ctypes.byref(pf), ctypes.byref(pt), ctypes.byref(pw)))
        return self._take_edges(ne, pf, pt, pw)

    def snn_device(self, d_index_ptr, n_total, k, snn_type, j_begin, j_end, fetch=True):
        ne = c_i64()
        if not fetch:  # edges stay in HBM, only the count comes back
            self._ck(lib().cb200_snn_device(self._h, c_vp(d_index_ptr), n_total, k, SNN_TYPES[snn_type], j_begin,
                                            j_end, ctypes.byref(ne), None, None, None))
            return ne.value
        pf, pt, pw = c_i32p(), c_i32p(), c_f64p()
        self._ck(lib().cb200_snn_device(self._h, c_vp(d_index_ptr), n_total, k, SNN_TYPES[snn_type], j_begin,
                                        j_end, ctypes.byref(ne), ctypes.byref(pf), ctypes.byref(pt),
                                        ctypes.byref(pw)))
        return self._take_edges(ne, pf, pt, pw)

    def snn_count_device(self, d_index_ptr, n_total, k, snn_type, j_begin, j_end):
        """Phase 1 of the two-phase SNN call: number of edges the cells [j_begin, j_end) will emit."""
        ne = c_i64()
        self._ck(lib().cb200_snn_count_device(self._h, c_vp(d_index_ptr), n_total, k, SNN_TYPES[snn_type], j_begin,
                                              j_end, ctypes.byref(ne)))
        return ne.value

    def snn_emit(self, n_edges, frm=None, to=None, weight=None, resident=False):
        """Phase 2: emit the counted edges into caller-owned arrays (copies overlap the emission)."""
        if resident:
            self._ck(lib().cb200_snn_emit(self._h, None, None, None))
            return n_edges
        frm = np.empty(n_edges, np.int32) if frm is None else frm
        to = np.empty(n_edges, np.int32) if to is None else to
        weight = np.empty(n_edges, np.float64) if weight is None else weight
        assert frm.size >= n_edges and to.size >= n_edges and weight.size >= n_edges
        self._ck(lib().cb200_snn_emit(self._h, _ptr(frm), _ptr(to), _ptr(weight)))
        return frm[:n_edges], to[:n_edges], weight[:n_edges]

    def fetch_edges(self, n_edges, frm=None, to=None, weight=None):
        """Edges of the last snn* call into caller-owned (e.g. pinned) arrays."""
        frm = np.empty(n_edges, np.int32) if frm is None else frm
        to = np.empty(n_edges, np.int32) if to is None else to
        weight = np.empty(n_edges, np.float64) if weight is None else weight
        self._ck(lib().cb200_fetch_edges(self._h, n_edges, _ptr(frm), _ptr(to), _ptr(weight)))
        return frm[:n_edges], to[:n_edges], weight[:n_edges]

    def edges_checksum(self):
        """(n_edges, 64-bit order-independent digest) of the edge list resident in HBM."""
        ne, cs = c_i64(), ctypes.c_uint64()
        self._ck(lib().cb200_edges_checksum(self._h, ctypes.byref(ne), ctypes.byref(cs)))
        return ne.value, cs.value

    # -- 8(f) rank 2: cluster diagnostics -------------------------------------------
    def approx_silhouette(self, emb, labels, n_clusters, d=None):
        """emb: N x d array or None for the resident PCA scores (first d components); labels: 0-based int32[N].
        Returns (width float64[N], other int32[N])."""
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        n = labels.size
        if emb is not None:
            emb = np.asfortranarray(emb, dtype=np.float64)
            assert emb.shape[0] == n
            d = emb.shape[1]
        width = np.empty(n)
        other = np.empty(n, dtype=np.int32)
        self._ck(lib().cb200_approx_silhouette(self._h, _ptr(emb), n, int(d), _ptr(labels), int(n_clusters),
                                               _ptr(width), _ptr(other)))
        return width, other

    def pairwise_modularity(self, edges, labels, n_clusters):
        """edges: (from, to, weight) 1-based arrays, or None for the edge list resident in HBM.
        Returns (observed K x K, expected K x K, total weight); upper triangles filled."""
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        K = int(n_clusters)
        obs, exp = np.empty((K, K)), np.empty((K, K))
        tot = ctypes.c_double()
        if edges is None:
            f = t = w = None
            ne = 0
        else:
            f = np.ascontiguousarray(edges[0], dtype=np.int32)
            t = np.ascontiguousarray(edges[1], dtype=np.int32)
            w = np.ascontiguousarray(edges[2], dtype=np.float64)
            ne = f.size
        self._ck(lib().cb200_pairwise_modularity(self._h, _ptr(f), _ptr(t), _ptr(w), ne, _ptr(labels), labels.size, K,
                                                 _ptr(obs), _ptr(exp), ctypes.cast(ctypes.byref(tot), c_vp)))
        return obs, exp, tot.value

    def snn_resident(self, k, snn_type="rank"):
        ne = c_i64()
        self._ck(lib().cb200_snn_resident(self._h, k, SNN_TYPES[snn_type], ctypes.byref(ne)))
        return ne.value


class MultiContext:
    """All GPUs of one node from this one process: the binding of cb200_multi_* (what the R shim calls when
    options(cellula.b200.gpus = n) asks for more than one GPU).  Same array layouts as `Context`, covering all cells."""

    def __init__(self, n_gpus: int = 0):
        self._h = c_vp()
        L = lib()
        if L.cb200_multi_init(int(n_gpus), ctypes.byref(self._h)) != 0:
            raise Cb200Error(L.cb200_multi_last_error(None).decode())
        self.n_gpus = L.cb200_multi_n_gpus(self._h)
        self.n_genes = self.n_cells = self.nnz = 0

    def _ck(self, rc):
        if rc != 0:
            raise Cb200Error(lib().cb200_multi_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value is not None:
            lib().cb200_multi_free(self._h)
            self._h = c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def timings(self, rank=0) -> dict:
        t = Timing()
        self._ck(lib().cb200_multi_stage_timings(self._h, rank, ctypes.byref(t)))
        return t.as_dict()

    def load_csc(self, n_genes, colptr, rowidx, values):
        colptr = np.ascontiguousarray(colptr)
        rowidx = np.ascontiguousarray(rowidx, dtype=np.int32)
        values = np.ascontiguousarray(values, dtype=np.float64)
        n_cells = colptr.size - 1
        if colptr.dtype == np.int32:
            fn = lib().cb200_multi_load_csc_i32
        else:
            colptr = np.ascontiguousarray(colptr, dtype=np.int64)
            fn = lib().cb200_multi_load_csc
        self._ck(fn(self._h, n_genes, n_cells, _ptr(colptr), _ptr(rowidx), _ptr(values)))
        self._keep = (colptr, rowidx, values)    # the upload of the row indices may still be running
        self.n_genes, self.n_cells, self.nnz = int(n_genes), int(n_cells), int(values.size)

    def size_factors(self, sf_in=None, want=True):
        sf_in = None if sf_in is None else np.ascontiguousarray(sf_in, dtype=np.float64)
        out = np.empty(self.n_cells) if want else None
        self._ck(lib().cb200_multi_size_factors(self._h, _ptr(sf_in), _ptr(out)))
        return out

    def lognorm_genestats(self, want_logx=True):
        logx = np.empty(self.nnz) if want_logx else None
        mean, var = np.empty(self.n_genes), np.empty(self.n_genes)
        self._ck(lib().cb200_multi_lognorm_genestats(self._h, _ptr(logx), _ptr(mean), _ptr(var)))
        return logx, mean, var

    def pca(self, hvg_idx, d, want_scores=True):
        hvg_idx = np.ascontiguousarray(hvg_idx, dtype=np.int32)
        h = len(hvg_idx)
        scores = np.empty((self.n_cells, d), order="F") if want_scores else None
        rotation = np.empty((h, d), order="F")
        varexp = np.empty(d)
        total = ctypes.c_double()
        self._ck(lib().cb200_multi_pca(self._h, _ptr(hvg_idx), h, d, _ptr(scores), _ptr(rotation), _ptr(varexp),
                                       ctypes.cast(ctypes.byref(total), c_vp)))
        return dict(scores=scores, rotation=rotation, var_explained=varexp, total_var=total.value,
                    percent_var=100.0 * varexp / total.value)

    def knn(self, emb, k, want=True, n_total=None, d=None):
        if emb is not None:
            emb = np.asfortranarray(emb, dtype=np.float64)
            n_total, d = emb.shape
        idx = np.empty((n_total, k), dtype=np.int32, order="F") if want else None
        self._ck(lib().cb200_multi_knn(self._h, _ptr(emb), n_total, d, k, _ptr(idx)))
        return idx

    def snn(self, index, snn_type="rank", n_total=None, k=None, fetch=True):
        if index is not None:
            index = np.asfortranarray(index, dtype=np.int32)
            n_total, k = index.shape
        ne = c_i64()
        self._ck(lib().cb200_multi_snn_count(self._h, _ptr(index), n_total, k, SNN_TYPES[snn_type], ctypes.byref(ne)))
        if not fetch:
            self._ck(lib().cb200_multi_snn_emit(self._h, None, None, None))
            return ne.value
        f, t, w = np.empty(ne.value, np.int32), np.empty(ne.value, np.int32), np.empty(ne.value)
        self._ck(lib().cb200_multi_snn_emit(self._h, _ptr(f), _ptr(t), _ptr(w)))
        return f, t, w

    def edges_checksum(self):
        ne, cs = c_i64(), ctypes.c_uint64()
        self._ck(lib().cb200_multi_edges_checksum(self._h, ctypes.byref(ne), ctypes.byref(cs)))
        return ne.value, cs.value
