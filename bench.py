#!/usr/bin/env python
"""bench.py -- cells/sec, counts -> SNN graph (BASELINE.json metric), on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3                 # cfg3 (1 306 127 x 27 998), the headline
    python bench.py --workload cfg1|cfg2|cfg5                     # the other count-matrix configs
    python bench.py --workload cfg4 [--k 30 --scheme rank]        # kNN/SNN-only sweep on the fixed 1M x 50 embedding
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU path (oracle port) on the host cores

A step is one pass of the hot path over the synthetic count matrix of the chosen workload:
size factors -> logcounts + gene mean/var -> (host) trend + HVGs -> PCA -> exact kNN -> SNN
edges.  `value` is measured with the counts resident in HBM, `e2e` through the host-facing C ABI
with host buffers (H2D of the dgCMatrix slots and D2H of every result inside the timed region; pinned
buffers for the headline, pageable ones -- what R vectors are -- reported beside it).  After the timed
regions the results are CHECKED (sampled kNN rows against fp64 brute force, the SNN edges of a range of
cells against the oracle on the GPU's own kNN, an order-independent digest of the whole edge list against
the single-GPU digest) and the outcome is printed as `verified`.  One JSON line is printed by rank 0.

All inputs come from ONE generator specification (cellula_b200/csrc/cb200_synth_core.h): the GPU arm
generates its shard on the device, the CPU arms generate the first cells of the SAME matrix on the host.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cells/sec counts->SNN graph"
WORKLOADS = {
    # BASELINE.json configs: cells, genes, nnz/cell, PCs, k, scheme
    "cfg1": dict(cells=2700, genes=32738, nnz_per_cell=850.0, ndims=30, k=10, scheme="jaccard"),
    "cfg2": dict(cells=68579, genes=32738, nnz_per_cell=525.0, ndims=50, k=20, scheme="rank"),
    "cfg3": dict(cells=1306127, genes=27998, nnz_per_cell=1500.0, ndims=50, k=20, scheme="rank"),
    # kNN/SNN-only sweep on a fixed 1 000 000 x 50 PC matrix (k and scheme from --k / --scheme)
    "cfg4": dict(cells=1000000, ndims=50, k=30, scheme="rank", kind="embedding"),
    "cfg5": dict(cells=4000000, genes=30000, nnz_per_cell=1000.0, ndims=50, k=20, scheme="rank"),
}
HVG_NTOP = 2000
SEED = 20261019
EMB_SEED = 20261020
GOLDEN_DIGESTS = os.path.join(ROOT, "tests", "golden", "edge_digests.json")


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sustained=j["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (there is no R on any box: DESIGN.md section 2)
# ---------------------------------------------------------------------------------------------
def host_sample(w, n_sample):
    """The first n_sample cells of the workload's matrix -- the same cells the GPU arm holds (one generator)."""
    from cellula_b200 import synth
    return synth.make_counts(n_sample, w["genes"], nnz_per_cell=w["nnz_per_cell"], seed=SEED, n_cells_total=w["cells"])


def cpu_counts_once(d, w):
    from oracle import oracle
    t0 = time.perf_counter()
    oracle.pipeline(d["n_genes"], d["colptr"], d["rowidx"], d["x"], ndims=w["ndims"], hvg_ntop=HVG_NTOP,
                    neighbors=w["k"], snn_type=w["scheme"], fast_pca=True)
    return time.perf_counter() - t0


def cpu_embedding_once(emb, w):
    from oracle import oracle
    t0 = time.perf_counter()
    knn = oracle.find_knn(emb, w["k"])
    oracle.snn_edges(knn, w["scheme"])
    return time.perf_counter() - t0


def cpu_sample_and_runner(w, n_sample):
    if w.get("kind") == "embedding":
        from cellula_b200 import synth
        emb = synth.make_embedding(w["cells"], w["ndims"], seed=EMB_SEED)[:n_sample]
        return (lambda: cpu_embedding_once(emb, w)), "C brute-force kNN + SNN (OpenMP)"
    d = host_sample(w, n_sample)
    return (lambda: cpu_counts_once(d, w)), "numpy/scipy ARPACK PCA + C brute-force kNN/SNN (OpenMP)"


def run_reference_arm(args, w):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n_sample = min(w["cells"], args.cpu_sample)
    run, what = cpu_sample_and_runner(w, n_sample)
    warm = min(args.warmup, 1)
    for _ in range(warm):
        run()
    t = float(np.mean([run() for _ in range(args.steps)]))
    val = n_sample / t
    # `config` states what this arm really ran: `cells` is the sample, `cells_full` the workload's size
    cfg = dict(workload=args.workload, **{k: v for k, v in w.items() if k != "cells"}, cells=n_sample, cells_full=w["cells"],
               hvg_ntop=HVG_NTOP, data_seed=SEED, generator="cb200_synth_core.h (host): first cells of the GPU arm's matrix")
    line = dict(impl="reference", metric=METRIC, value=val, unit="cells/s", n_gpus=args.gpus, steps=args.steps,
                warmup=warm, ms_per_step=1e3 * t, higher_is_better=True, scaling="strong",
                vs_baseline=None, dtype="f64", data="synthetic", config=cfg,
                cpu_baseline=dict(value=val, unit="cells/s", cores=os.cpu_count(), kind="port",
                                  sample=f"first {n_sample} of {w['cells']} cells; oracle port ({what}); the kNN cost per "
                                         "cell grows with N, so the full-size CPU rate is lower than this"),
                e2e=dict(value=val, unit="cells/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# checks after the timed region (the oracle is the checker, never the thing measured)
# ---------------------------------------------------------------------------------------------
def verify_results(ctx, emb_rowmajor, knn0, k, scheme, j0, j1, n_rows=1024, n_cells=2048):
    """emb_rowmajor: N x d host array the kNN was computed from; knn0: N x k 0-based host array (the GPU result).
    Checks n_rows sampled kNN rows against brute force in the reference's arithmetic (sequential fp64, sqrt,
    (distance, index) order) and the SNN edges of n_cells cells of [j0, j1) against the oracle on knn0."""
    from oracle import oracle as O
    n = emb_rowmajor.shape[0]
    rng = np.random.default_rng(12345)
    rows = np.sort(rng.choice(n, min(n_rows, n), replace=False))
    want = O.find_knn_queries(emb_rowmajor, rows, k)
    knn_bad = int(((knn0[rows] + 1) != want).any(axis=1).sum())
    ja = j0 + (j1 - j0) // 2
    jb = min(j1, ja + n_cells)
    got = ctx.snn(knn0 + 1, scheme, j_begin=ja, j_end=jb)
    ref = O.snn_edges(knn0 + 1, scheme, j_begin=ja, j_end=jb)
    snn_ok = all(a.shape == b.shape and np.array_equal(a, b) for a, b in zip(got, ref))
    return dict(knn_rows_checked=int(rows.size), knn_rows_wrong=knn_bad, snn_cells_checked=int(jb - ja),
                snn_edges_checked=int(ref[0].size), snn_edges_identical=bool(snn_ok))


def golden_digest(workload, k, scheme):
    try:
        return json.load(open(GOLDEN_DIGESTS)).get(f"{workload}:k{k}:{scheme}")
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cellula_b200", choices=["cellula_b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("CB200_WORKLOAD", "cfg3"), choices=sorted(WORKLOADS))
    ap.add_argument("--k", type=int, default=None, help="neighbours (default: the workload's)")
    ap.add_argument("--scheme", default=None, choices=["rank", "number", "jaccard"])
    ap.add_argument("--cpu-sample", type=int, default=20000, help="cells in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--write-golden", action="store_true", help="store this run's edge digest as the single-GPU golden")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.k:
        w["k"] = args.k
    if args.scheme:
        w["scheme"] = args.scheme
    if args.impl == "reference":
        return run_reference_arm(args, w)
    if args.warmup < 3:
        args.warmup = max(args.warmup, 1)  # the driver passes W >= 3; smaller only for debugging

    import torch
    import torch.distributed as dist

    import cellula_b200
    from cellula_b200 import synth
    from cellula_b200.dist import CudaShard, ShardedPipeline, shard_bounds, snn_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = cellula_b200.Context(local_rank, stream=stream.cuda_stream)
    embedding_only = w.get("kind") == "embedding"
    n_total = w["cells"]
    bounds = shard_bounds(n_total, world)
    c0, c1 = bounds[rank], bounds[rank + 1]
    n_local = c1 - c0
    shard = CudaShard(ctx, local_rank)
    pipe = ShardedPipeline(shard, n_total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stage_keys = ["colsum_ms", "lognorm_ms", "hvg_extract_ms", "gram_ms", "eig_ms", "project_ms", "knn_prep_ms",
                  "knn_tile_ms", "knn_rerank_ms", "snn_revlist_ms", "snn_edges_ms"]

    if embedding_only:
        # the fixed PC matrix: generated on the host (every rank the same), resident row-major for the `value` leg
        emb_host = synth.make_embedding(n_total, w["ndims"], seed=EMB_SEED)            # column-major, what R holds
        emb_dev = torch.from_numpy(np.ascontiguousarray(emb_host)).to(dev)             # row-major N x d
        nnz = 0

        def resident_step():
            knn_full = pipe.knn(emb_dev, w["k"])
            return dict(knn_full=knn_full, edges=pipe.snn(knn_full, w["scheme"], fetch=False), emb=emb_dev)
    else:
        # synthetic shard, generated on the device, then exported to host buffers for the e2e legs
        nnz = ctx.synth_counts(w["genes"], n_local, w["nnz_per_cell"], SEED, cell_offset=c0, n_cells_total=n_total)

        def resident_step():
            r = pipe.run(ndims=w["ndims"], hvg_ntop=HVG_NTOP, neighbors=w["k"], snn_type=w["scheme"], want_host=False)
            r["emb"] = r["pca"]["scores_full_dev"]
            return r

    for _ in range(args.warmup):
        res = resident_step()
    barrier()
    ctx.reset_timings()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_sum = {k: 0.0 for k in stage_keys}
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        res = resident_step()
        tm = ctx.timings()
        for k in stage_keys:
            stage_sum[k] += tm[k]
    ev1.record(stream)
    barrier()
    sampler.stop_flag.set()
    sampler.join()
    launches = tm["kernel_launches"]
    resident_marks = dict(getattr(pipe, "last_marks", {}))   # host wall-clock split of the last resident step
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    n_edges_local = int(res["edges"])
    fallback_q = tm["knn_fallback_queries"]
    eig_iters = tm["eig_iterations"]
    tiles_scanned, tiles_total = tm["knn_tiles_scanned"], tm["knn_tiles_total"]
    engines = dict(gram={1: "fp64-l2-red", 2: "fp64-smem-tiles", 3: "tcgen05-int8-digits"}.get(tm["gram_engine"]),
                   knn={1: "fp32-cuda-cores", 2: "tcgen05-fp16-split"}.get(tm["knn_engine"]))

    # ---- checks on the results of the last resident step (edges still in HBM) ----
    verified = None
    if not args.no_verify:
        ne_chk, digest = ctx.edges_checksum()
        dg = torch.tensor([digest - (1 << 64) if digest >= (1 << 63) else digest, ne_chk], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(dg, op=dist.ReduceOp.SUM)       # int64 wrap-around = sum modulo 2^64
        digest_all = int(dg[0].item()) & ((1 << 64) - 1)
        verified = dict(n_edges=int(dg[1].item()), edge_digest=f"{digest_all:016x}")
        if rank == 0:
            emb_h = res["emb"].cpu().numpy()
            knn_h = res["knn_full"].cpu().numpy()
            sb = snn_bounds(n_total, world)
            try:
                verified.update(verify_results(ctx, emb_h, knn_h, w["k"], w["scheme"], sb[0], sb[1]))
            except cellula_b200.Cb200Error as e:   # e.g. cfg5 on ONE GPU: no HBM left for the checker's own upload
                verified.update(knn_rows_wrong=None, snn_edges_identical=None, check_error=str(e)[:200])
            gold = golden_digest(args.workload, w["k"], w["scheme"])
            verified["single_gpu_digest"] = gold["edge_digest"] if gold else None
            verified["matches_single_gpu"] = (gold["edge_digest"] == verified["edge_digest"] and
                                              gold["n_edges"] == verified["n_edges"]) if gold else None
            verified["ok"] = None if "check_error" in verified else bool(
                verified["knn_rows_wrong"] == 0 and verified["snn_edges_identical"] and
                verified["matches_single_gpu"] is not False)
            if args.write_golden and world == 1:
                allg = json.load(open(GOLDEN_DIGESTS)) if os.path.exists(GOLDEN_DIGESTS) else {}
                allg[f"{args.workload}:k{w['k']}:{w['scheme']}"] = dict(n_edges=verified["n_edges"],
                                                                        edge_digest=verified["edge_digest"])
                json.dump(allg, open(GOLDEN_DIGESTS, "w"), indent=1, sort_keys=True)
            del emb_h, knn_h
        barrier()

    # ---- e2e legs: host buffers in, host results out, through the C ABI ----
    e2e = None
    e2e_pageable = None
    if not args.no_e2e:
        def make_host(shape, dtype, pinned):
            t = torch.empty(shape, dtype=dtype)
            return t.pin_memory() if pinned else t

        def run_e2e(pinned):
            cap = int(1.05 * n_edges_local) + 1024
            ebuf = [make_host(cap, torch.int32, pinned), make_host(cap, torch.int32, pinned),
                    make_host(cap, torch.float64, pinned)]

            def edge_buffers(ne):
                if ne > ebuf[0].numel():  # grow (not expected: same data every step)
                    ebuf[0], ebuf[1] = make_host(ne, torch.int32, pinned), make_host(ne, torch.int32, pinned)
                    ebuf[2] = make_host(ne, torch.float64, pinned)
                return ebuf[0].numpy(), ebuf[1].numpy(), ebuf[2].numpy()

            knn_h = make_host((n_local, w["k"]), torch.int32, pinned)
            d2h_box = [0]
            if embedding_only:
                emb_in = make_host((w["ndims"], n_total), torch.float64, pinned)      # column-major N x d
                emb_in.numpy()[:] = emb_host.T
                emb_np = emb_in.numpy().T                                               # F-ordered view, no copy
                h2d = emb_in.numel() * 8

                def step():
                    ctx.knn(emb_np, w["k"], q_begin=c0, q_end=c1, want=False)
                    knn_full = pipe._allgather_rows(shard._view("knn", (n_local, w["k"]), torch.int32))
                    knn_h.copy_(knn_full[c0:c1], non_blocking=False)
                    f, t, wt = pipe.snn(knn_full, w["scheme"], fetch=True, edge_buffers=edge_buffers)
                    d2h_box[0] = knn_h.numel() * 4 + f.size * 4 + t.size * 4 + wt.size * 8
            else:
                colptr_h = make_host(n_local + 1, torch.int64, pinned)
                rowidx_h = make_host(nnz, torch.int32, pinned)
                x_h = make_host(nnz, torch.float64, pinned)
                ctx.export_csc(colptr_h.numpy(), rowidx_h.numpy(), x_h.numpy())
                logx_h = make_host(nnz, torch.float64, pinned)
                scores_h = make_host(n_local * w["ndims"], torch.float64, pinned)
                scores_np = scores_h.numpy().reshape((n_local, w["ndims"]), order="F")
                h2d = colptr_h.numel() * 8 + rowidx_h.numel() * 4 + x_h.numel() * 8 + HVG_NTOP * 4

                def step():
                    ctx.load_csc_ptrs(w["genes"], n_local, nnz, colptr_h.data_ptr(), rowidx_h.data_ptr(), x_h.data_ptr(),
                                      cell_offset=c0, n_cells_total=n_total)
                    r = pipe.run(ndims=w["ndims"], hvg_ntop=HVG_NTOP, neighbors=w["k"], snn_type=w["scheme"],
                                 want_host=True, scores_out=scores_np, edge_buffers=edge_buffers, logx_out=logx_h.numpy())
                    knn_h.copy_(r["knn_full"][c0:c1], non_blocking=False)
                    f, t, wt = r["edges"]
                    d2h_box[0] = (n_local * 8 + nnz * 8 + 2 * w["genes"] * 8 + n_local * w["ndims"] * 8 +
                                  HVG_NTOP * w["ndims"] * 8 + knn_h.numel() * 4 + f.size * 4 + t.size * 4 + wt.size * 8)

            for _ in range(min(args.warmup, 2)):
                step()
            barrier()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                step()
            e1.record(stream)
            barrier()
            ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)) / args.steps
            tme = ctx.timings()
            return dict(ms=ms, h2d=h2d, d2h=d2h_box[0],
                        stage_ms={k: tme[k] for k in ["h2d_ms", "d2h_ms"] + stage_keys} |
                                 {"host_marks_ms": dict(getattr(pipe, "last_marks", {}))})

        e2e = run_e2e(pinned=True)
        e2e_pageable = run_e2e(pinned=False)

    # ---- max over ranks ----
    vals = torch.tensor([dev_ms, e2e["ms"] if e2e else 0.0, float(n_edges_local), float(launches),
                         float(e2e["h2d"] if e2e else 0), float(e2e["d2h"] if e2e else 0),
                         e2e_pageable["ms"] if e2e_pageable else 0.0],
                        dtype=torch.float64, device=dev)
    if world > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms, e2e_pg_ms = float(mx[0]), float(mx[1]), float(mx[6])
        n_edges, launches_all = int(sm[2]), int(sm[3])
        h2d_all, d2h_all = int(sm[4]), int(sm[5])
    else:
        e2e_ms, e2e_pg_ms = float(vals[1]), float(vals[6])
        n_edges, launches_all = n_edges_local, int(launches)
        h2d_all, d2h_all = int(vals[4]), int(vals[5])

    if rank == 0:
        peaks = read_peaks()
        stage_ms = {k: v / args.steps for k, v in stage_sum.items()}
        kk, nd = w["k"], w["ndims"]
        # dominant kernel: the kNN distance-tile kernel (k_knn_tiles_tc).  `achieved` / `frac` follow SURVEY 8d: the
        # ALGORITHMIC work 2 * n_query * N * d flop per launch (the whole distance matrix; tiles skipped by the exact
        # triangle-inequality bound count as answered), so frac can exceed 1 -- it measures the algorithm.  What the
        # tensor pipe really does is `executed_*`: 3 fp16 products of 128 x 128 x 64 per SCANNED tile.
        flop = 2.0 * n_local * n_total * nd
        knn_ms = stage_ms["knn_tile_ms"]
        achieved_tf = flop / (knn_ms * 1e-3) / 1e12 if knn_ms > 0 else 0.0
        executed_tf = (tiles_scanned * 3 * 2.0 * 128 * 128 * 64 / (knn_ms * 1e-3) / 1e12) if knn_ms > 0 else None
        traffic = traffic_note = None
        try:  # DRAM bytes per launch from the committed ncu capture of this kernel, when it matches this run
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_knn_tiles_tc"]
            if tj["workload"] == args.workload and tj["n_gpus"] == world:
                traffic, traffic_note = tj["bytes_per_launch"], tj.get("capture")
        except Exception:
            pass
        roofline = dict(bound="tensor", kernel="k_knn_tiles_tc", achieved=achieved_tf, peak=peaks["tf_sustained"],
                        unit="TFLOP/s", frac=achieved_tf / peaks["tf_sustained"], traffic=traffic, traffic_capture=traffic_note,
                        peak_source=peaks["source"] + " bf16 sustained",
                        share_of_step=knn_ms / dev_ms if dev_ms > 0 else None,
                        frac_is="algorithmic (SURVEY 8d: 2 N^2 d / t); pruned tiles count as done",
                        tiles_scanned_frac=(tiles_scanned / tiles_total) if tiles_total else None,
                        executed_tflops=executed_tf,
                        executed_frac=(executed_tf / peaks["tf_sustained"]) if executed_tf else None,
                        executed_is="3 fp16 products x 2*128*128*64 flop per scanned tile / kernel time")
        # every other stage against its own roofline (algorithmic work of DESIGN.md section 4 / SURVEY 8d)
        stages = {}

        def add_stage(key, kind, work, note=None):
            ms = stage_ms.get(key, 0.0)
            if ms <= 0:
                return
            if kind == "hbm":
                ach = work / (ms * 1e-3) / 1e9
                stages[key.replace("_ms", "")] = dict(bound="hbm", ms=ms, achieved_gbs=ach, frac=ach / peaks["hbm"], note=note)
            else:
                ach = work / (ms * 1e-3) / 1e12
                stages[key.replace("_ms", "")] = dict(bound="tensor", ms=ms, achieved_tflops=ach,
                                                      frac=ach / peaks["tf_sustained"], note=note)
        if not embedding_only:
            H = HVG_NTOP
            add_stage("colsum_ms", "hbm", 8.0 * nnz + 16.0 * n_local)
            add_stage("lognorm_ms", "hbm", 20.0 * nnz + 8.0 * n_local + 16.0 * w["genes"])
            add_stage("hvg_extract_ms", "hbm", 8.0 * nnz, "2 reads of the row indices; the values of non-HVG rows are not read")
            add_stage("gram_ms", "tensor", 1.0 * n_local * H * H, "N H^2 (symmetric half, un-split fp32-equivalent flops)")
            add_stage("project_ms", "tensor", 2.0 * n_local * H * nd, "2 N H d; runs on the fp64 CUDA cores over the sparse HVG matrix")
        add_stage("snn_edges_ms", "hbm", 12.0 * n_total * kk + 8.0 * n_local * (kk + 1) ** 2 + 16.0 * n_edges_local,
                  "kNN read + reverse lists + list gathers + edge write (SURVEY 8d K7), rank 0's share")
        replicated = {k: stage_ms[k] for k in ("eig_ms", "knn_prep_ms", "snn_revlist_ms")}
        replicated["host_trend_hvg_ms"] = (resident_marks.get("trend_hvg", 0.0) - resident_marks.get("gene_stats", 0.0)) \
            if resident_marks else None
        cfg = dict(workload=args.workload, **{k: v for k, v in w.items() if k != "kind"}, hvg_ntop=HVG_NTOP, nnz_local=nnz,
                   n_edges=n_edges, l2="inputs larger than L2 (no flush)" if args.workload in ("cfg3", "cfg5", "cfg4")
                   else "inputs smaller than L2: parity case, not a bench line",
                   parallelism=f"cells sharded x{world}", knn_fallback_queries=int(fallback_q), eig_iterations=int(eig_iters),
                   knn_tiles_scanned_rank0=int(tiles_scanned), knn_tiles_total_rank0=int(tiles_total),
                   engines=engines, size_factors="library", data_seed=SEED,
                   generator="cb200_synth_core.h (device): bit-identical to the host compilation the CPU arm uses")
        line = dict(metric=METRIC, value=n_total / (dev_ms * 1e-3), unit="cells/s", n_gpus=world,
                    steps=args.steps, warmup=args.warmup, ms_per_step=dev_ms, higher_is_better=True,
                    scaling="strong", vs_baseline=None, dtype="f64", data="synthetic", config=cfg,
                    clocks=sampler.summary(), gpu_launches=launches_all // max(args.steps, 1) * args.steps,
                    stage_ms=stage_ms, host_marks_ms_last_step_rank0=resident_marks, stages=stages,
                    replicated_ms=replicated, roofline=roofline, verified=verified)
        if e2e:
            line["e2e"] = dict(value=n_total / (e2e_ms * 1e-3), unit="cells/s", ms_per_step=e2e_ms,
                               h2d_bytes_per_step=h2d_all, d2h_bytes_per_step=d2h_all, host_memory="pinned",
                               last_step_stage_ms_rank0=e2e.get("stage_ms"))
            line["e2e_pageable"] = dict(value=n_total / (e2e_pg_ms * 1e-3), unit="cells/s", ms_per_step=e2e_pg_ms,
                                        host_memory="pageable (what R vectors are): the library stages the copies through "
                                                    "its pinned ring", last_step_stage_ms_rank0=e2e_pageable.get("stage_ms"))
        if world == 1 and not args.no_cpu_baseline:
            n_sample = min(n_total, args.cpu_sample)
            run, what = cpu_sample_and_runner(w, n_sample)
            t = run()
            line["cpu_baseline"] = dict(value=n_sample / t, unit="cells/s", cores=os.cpu_count(), kind="port",
                                        sample=f"first {n_sample} of {n_total} cells of the same matrix (one pass, {t:.1f} s): "
                                               f"oracle port, {what}")
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
