"""N > 1 host logic on CPU: world_size 2 over gloo.  `ShardedPipeline` (cellula_b200/dist.py) is
driven with a test double as shard backend that does the rank-local arithmetic with the oracle,
so the sharding, the order and content of the collectives, and the per-rank edge emission are
exercised exactly as on GPUs; the result must equal the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cellula_b200 import synth
from cellula_b200.dist import ShardedPipeline, shard_bounds, slice_csc
from oracle import oracle as O


class OracleShard:
    """Same surface as cellula_b200.dist.CudaShard, oracle arithmetic, CPU tensors."""

    def __init__(self, n_genes, colptr, rowidx, x):
        self.G, self.colptr, self.rowidx, self.x = n_genes, colptr, rowidx, x
        self.n = len(colptr) - 1

    def libsize_partial(self, sf_in=None):
        ls = np.zeros(self.n)
        nz = np.diff(self.colptr) > 0
        ls[nz] = np.add.reduceat(self.x, self.colptr[:-1][nz])
        self.ls = ls if sf_in is None else np.asarray(sf_in, dtype=float)
        return torch.tensor([self.ls.sum(), float(self.n)], dtype=torch.float64)

    def finish_size_factors(self, total, n_total, want=True):
        self.n_total = n_total
        self.sf = self.ls / (total / n_total)
        if np.any(self.sf <= 0):
            raise ValueError("size factors should be positive")
        return self.sf

    # values-first order of the e2e path: logcounts before (and independently of) the gene sums
    def lognorm_values(self):
        self.logx = O.log_norm_counts(self.colptr, self.x, self.sf)
        self.values_first = True

    def logcounts_async(self, out):
        assert getattr(self, "values_first", False), "the logcounts must be requested after lognorm_values"
        out[:] = self.logx

    def wait_copies(self):
        pass

    def gene_acc(self):
        if not getattr(self, "values_first", False):
            self.logx = O.log_norm_counts(self.colptr, self.x, self.sf)
        s = np.bincount(self.rowidx, weights=self.logx, minlength=self.G)
        q = np.bincount(self.rowidx, weights=self.logx ** 2, minlength=self.G)
        self._acc = torch.from_numpy(np.concatenate([s, q]))
        return self._acc

    def finish_genestats(self):
        a = self._acc.numpy()
        s, q = a[:self.G], a[self.G:]
        n = float(self.n_total)
        self.mean = s / n
        self.var = np.maximum(q - s * s / n, 0) / (n - 1)
        return self.mean, self.var

    def gram(self, hvg_idx):
        import scipy.sparse as sp
        m = sp.csc_matrix((self.logx, self.rowidx, self.colptr), shape=(self.G, self.n))
        self.X = np.asarray(m.tocsr()[np.asarray(hvg_idx, dtype=np.int64), :].T.todense())
        self.hvg = np.asarray(hvg_idx)
        self._gram = torch.from_numpy((self.X.T @ self.X).ravel().copy())
        return self._gram

    def finish_pca(self, d, n_hvg, want_scores=True):
        n = float(self.n_total)
        mu = self.mean[self.hvg]
        cov = (self._gram.numpy().reshape(n_hvg, n_hvg) - n * np.outer(mu, mu)) / (n - 1)
        ev, evec = np.linalg.eigh(cov)
        v = evec[:, ::-1][:, :d].copy()
        for j in range(d):
            if v[np.argmax(np.abs(v[:, j])), j] < 0:
                v[:, j] *= -1
        scores = (self.X - mu) @ v
        return dict(scores=scores, rotation=v, var_explained=ev[::-1][:d], total_var=float(self.var[self.hvg].sum()),
                    scores_dev=torch.from_numpy(np.ascontiguousarray(scores)))

    def knn(self, emb_full, k, q_begin, q_end, want=False):
        idx = O.find_knn(emb_full.numpy(), k)[q_begin:q_end] - 1
        return torch.from_numpy(np.ascontiguousarray(idx.astype(np.int32)))

    def snn(self, knn_full, snn_type, j_begin, j_end):
        f, t, w = O.snn_edges(knn_full.numpy() + 1, snn_type)
        keep = (f > j_begin) & (f <= j_end)
        return f[keep], t[keep], w[keep]


def _worker(rank, world, port, n_cells, n_genes, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = synth.make_counts(n_cells, n_genes, nnz_per_cell=200, seed=5)
        b = shard_bounds(n_cells, world)
        p, i, x = slice_csc(d["colptr"], d["rowidx"], d["x"], b[rank], b[rank + 1])
        pipe = ShardedPipeline(OracleShard(n_genes, p, i, x), n_cells)
        assert (pipe.c0, pipe.c1) == (b[rank], b[rank + 1])
        logx_out = np.empty(x.size)
        res = pipe.run(ndims=8, hvg_ntop=300, neighbors=5, snn_type="rank", logx_out=logx_out)
        assert np.array_equal(logx_out, pipe.shard.logx)       # copied out before the gene statistics
        assert pipe.j0 < pipe.j1                                # SNN range of this rank (balanced by work)
        gathered = [None] * world
        dist.all_gather_object(gathered, dict(sf=res["sf"], edges=res["edges"], hvgs=res["hvgs"],
                                              mean=res["mean"], var=res["var"],
                                              scores=res["pca"]["scores_full_dev"].numpy(),
                                              knn=res["knn_full"].numpy()))
        if rank == 0:
            out.put(gathered)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(600)
def test_sharded_pipeline_world2_equals_unsharded_oracle():
    n_cells, n_genes, world = 901, 1500, 2  # odd cell count: unequal shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_cells, n_genes, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=500)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    d = synth.make_counts(n_cells, n_genes, nnz_per_cell=200, seed=5)
    ref = O.pipeline(n_genes, d["colptr"], d["rowidx"], d["x"], ndims=8, hvg_ntop=300, neighbors=5, snn_type="rank")
    sf = np.concatenate([g["sf"] for g in gathered])
    assert np.allclose(sf, ref["sf"], rtol=1e-13)
    for g in gathered:  # replicated results identical on every rank
        assert np.allclose(g["mean"], ref["mean"], rtol=1e-11, atol=1e-300)
        assert np.allclose(g["var"], ref["var"], rtol=1e-9, atol=1e-300)
        assert np.array_equal(g["hvgs"], ref["hvgs"])
        assert np.array_equal(g["scores"], gathered[0]["scores"])
        assert np.array_equal(g["knn"], gathered[0]["knn"])
    # PCA: same subspace as the oracle's SVD, sign-aligned scores
    s, r = gathered[0]["scores"], ref["pca"]["scores"]
    sign = np.sign(np.sum(s * r, axis=0))
    assert np.allclose(s * sign, r, rtol=0, atol=1e-6 * np.abs(r).max())
    # kNN / SNN: exact against the oracle fed the same embedding
    assert np.array_equal(gathered[0]["knn"] + 1, O.find_knn(s, 5))
    f = np.concatenate([g["edges"][0] for g in gathered])
    t = np.concatenate([g["edges"][1] for g in gathered])
    w = np.concatenate([g["edges"][2] for g in gathered])
    rf, rt, rw = O.snn_edges(gathered[0]["knn"] + 1, "rank")
    assert np.array_equal(f, rf) and np.array_equal(t, rt) and np.array_equal(w, rw)  # rank order == emission order
