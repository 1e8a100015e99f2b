"""Worker for tests/test_gpu_dist.py: launched by torch.distributed.run, one rank per GPU.
Each rank loads its shard of the same host-generated count matrix, runs the sharded pipeline
(NCCL collectives between the stages) and rank 0 checks the assembled result against the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import cellula_b200  # noqa: E402
from cellula_b200 import synth  # noqa: E402
from cellula_b200.dist import CudaShard, ShardedPipeline, shard_bounds, slice_csc  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_cells, n_genes = 5003, 3000
    d = synth.make_counts(n_cells, n_genes, nnz_per_cell=250, seed=13)
    b = shard_bounds(n_cells, world)
    p, i, x = slice_csc(d["colptr"], d["rowidx"], d["x"], b[rank], b[rank + 1])
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = cellula_b200.Context(local, stream=stream.cuda_stream)
    ctx.load_csc(n_genes, p, i, x, cell_offset=b[rank], n_cells_total=n_cells)
    pipe = ShardedPipeline(CudaShard(ctx, local), n_cells)
    res = pipe.run(ndims=15, hvg_ntop=600, neighbors=12, snn_type="rank", want_host=True)
    logx = ctx.fetch_logcounts()
    torch.cuda.synchronize()
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(sf=res["sf"], logx=logx, edges=res["edges"], hvgs=res["hvgs"],
                                          mean=res["mean"], var=res["var"], scores=res["pca"]["scores"],
                                          rotation=res["pca"]["rotation"], varexp=res["pca"]["var_explained"],
                                          knn=res["knn_full"].cpu().numpy()))
    ok = True
    if rank == 0:
        from oracle import oracle as O
        ref = O.pipeline(n_genes, d["colptr"], d["rowidx"], d["x"], ndims=15, hvg_ntop=600, neighbors=12,
                         snn_type="rank")
        sf = np.concatenate([g["sf"] for g in gathered])
        lx = np.concatenate([g["logx"] for g in gathered])
        assert np.allclose(sf, ref["sf"], rtol=1e-13)
        assert np.allclose(lx, ref["logx"], rtol=1e-12)
        for g in gathered:
            assert np.allclose(g["mean"], ref["mean"], rtol=1e-11, atol=1e-300)
            assert np.allclose(g["var"], ref["var"], rtol=1e-9, atol=1e-18)
            assert np.array_equal(g["hvgs"], ref["hvgs"])
            assert np.array_equal(g["knn"], gathered[0]["knn"])           # identical on every rank
            assert np.allclose(g["rotation"], gathered[0]["rotation"], atol=1e-12)
        scores = np.concatenate([g["scores"] for g in gathered])
        want = ref["pca"]["scores"]
        sign = np.sign(np.sum(scores * want, axis=0))
        err = np.abs(scores * sign - want).max(axis=0) / np.abs(want).max(axis=0)
        assert err.max() <= 1e-4, err
        assert np.allclose(gathered[0]["varexp"], ref["pca"]["var_explained"], rtol=1e-6)
        knn = gathered[0]["knn"] + 1
        assert np.array_equal(knn, O.find_knn(scores, 12))               # oracle fed the same embedding
        f = np.concatenate([g["edges"][0] for g in gathered])
        t = np.concatenate([g["edges"][1] for g in gathered])
        w = np.concatenate([g["edges"][2] for g in gathered])
        rf, rt, rw = O.snn_edges(knn, "rank")
        assert np.array_equal(f, rf) and np.array_equal(t, rt) and np.array_equal(w, rw)
        print(f"DIST_OK world={world} edges={len(f)}")
    ctx.close()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
