#!/usr/bin/env python
"""bench.py -- cells/sec, counts -> SNN graph (BASELINE.json metric), on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU path (oracle port) on the host cores

A step is one pass of the hot path over the synthetic count matrix of the chosen workload:
size factors -> logcounts + gene mean/var -> (host) trend + HVGs -> PCA -> exact kNN -> SNN
edges.  `value` is measured with the counts resident in HBM, `e2e` through the host-facing C ABI
with pinned host buffers (H2D of the dgCMatrix slots and D2H of every result inside the timed
region).  One JSON line is printed by rank 0.
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
    # cfg4 (kNN/SNN-only sweep on a fixed 1M x 50 embedding) is scripts/knn_only.py
    "cfg5": dict(cells=4000000, genes=30000, nnz_per_cell=1000.0, ndims=50, k=20, scheme="rank"),
}
HVG_NTOP = 2000
SEED = 20261019


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
# CPU arm: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_pipeline_once(n_genes, colptr, rowidx, x, w):
    from oracle import oracle
    t0 = time.perf_counter()
    oracle.pipeline(n_genes, colptr, rowidx, x, ndims=w["ndims"], hvg_ntop=HVG_NTOP, neighbors=w["k"],
                    snn_type=w["scheme"], fast_pca=True)
    return time.perf_counter() - t0


def host_sample(w, n_sample, seed=SEED):
    """A bounded sample of the workload for the CPU arm, generated on the host (numpy)."""
    from cellula_b200 import synth
    return synth.make_counts(n_sample, w["genes"], nnz_per_cell=w["nnz_per_cell"], seed=seed, chunk=2048)


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = min(w["cells"], args.cpu_sample)
    d = host_sample(w, n_sample)
    for _ in range(min(args.warmup, 1)):
        cpu_pipeline_once(d["n_genes"], d["colptr"], d["rowidx"], d["x"], w)
    times = [cpu_pipeline_once(d["n_genes"], d["colptr"], d["rowidx"], d["x"], w) for _ in range(args.steps)]
    t = float(np.mean(times))
    cores = os.cpu_count()
    val = n_sample / t
    line = dict(impl="reference", metric=METRIC, value=val, unit="cells/s", n_gpus=args.gpus, steps=args.steps,
                warmup=min(args.warmup, 1), ms_per_step=1e3 * t, higher_is_better=True, scaling="strong",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=args.workload, **w, hvg_ntop=HVG_NTOP),
                cpu_baseline=dict(value=val, unit="cells/s", cores=cores, kind="port",
                                  sample=f"{n_sample} of {w['cells']} cells x {w['genes']} genes, same generator spec; "
                                         "oracle port (numpy/scipy ARPACK PCA, C brute-force kNN + SNN with OpenMP); "
                                         "kNN cost grows with N^2 so full-size cells/s is lower"),
                e2e=dict(value=val, unit="cells/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


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
    ap.add_argument("--cpu-sample", type=int, default=20000, help="cells in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args, w)
    if args.warmup < 3:
        args.warmup = max(args.warmup, 1)  # the driver passes W >= 3; smaller only for debugging

    import torch
    import torch.distributed as dist

    import cellula_b200
    from cellula_b200.dist import CudaShard, ShardedPipeline, shard_bounds

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
    bounds = shard_bounds(w["cells"], world)
    c0, c1 = bounds[rank], bounds[rank + 1]
    n_local = c1 - c0

    # synthetic shard, generated on the device, then exported to pinned host buffers for the e2e leg
    nnz = ctx.synth_counts(w["genes"], n_local, w["nnz_per_cell"], SEED, cell_offset=c0, n_cells_total=w["cells"])
    shard = CudaShard(ctx, local_rank)
    pipe = ShardedPipeline(shard, w["cells"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        return pipe.run(ndims=w["ndims"], hvg_ntop=HVG_NTOP, neighbors=w["k"], snn_type=w["scheme"], want_host=False)

    stage_keys = ["colsum_ms", "lognorm_ms", "hvg_extract_ms", "gram_ms", "eig_ms", "project_ms", "knn_prep_ms",
                  "knn_tile_ms", "knn_rerank_ms", "snn_revlist_ms", "snn_edges_ms"]

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
    resident_marks = dict(pipe.last_marks)   # host wall-clock split of the last resident step (rank 0's is printed)
    dev_ms = ev0.elapsed_time(ev1) / args.steps
    n_edges_local = int(res["edges"])
    fallback_q = tm["knn_fallback_queries"]
    eig_iters = tm["eig_iterations"]
    tiles_scanned, tiles_total = tm["knn_tiles_scanned"], tm["knn_tiles_total"]
    engines = dict(gram={1: "fp64-l2-red", 2: "fp64-smem-tiles", 3: "tcgen05-int8-digits"}.get(tm["gram_engine"]),
                   knn={1: "fp32-cuda-cores", 2: "tcgen05-fp16-split"}.get(tm["knn_engine"]))

    # ---- e2e leg: host buffers in, host results out, through the C ABI ----
    e2e = None
    if not args.no_e2e:
        colptr_h = torch.empty(n_local + 1, dtype=torch.int64).pin_memory()
        rowidx_h = torch.empty(nnz, dtype=torch.int32).pin_memory()
        x_h = torch.empty(nnz, dtype=torch.float64).pin_memory()
        ctx.export_csc(colptr_h.numpy(), rowidx_h.numpy(), x_h.numpy())
        logx_h = torch.empty(nnz, dtype=torch.float64).pin_memory()
        scores_h = torch.empty(n_local * w["ndims"], dtype=torch.float64).pin_memory()
        knn_h = torch.empty((n_local, w["k"]), dtype=torch.int32).pin_memory()
        cap = int(1.05 * n_edges_local) + 1024
        ebuf = [torch.empty(cap, dtype=torch.int32).pin_memory(), torch.empty(cap, dtype=torch.int32).pin_memory(),
                torch.empty(cap, dtype=torch.float64).pin_memory()]

        def edge_buffers(ne):
            if ne > ebuf[0].numel():  # grow (not expected: same data every step)
                ebuf[0] = torch.empty(ne, dtype=torch.int32).pin_memory()
                ebuf[1] = torch.empty(ne, dtype=torch.int32).pin_memory()
                ebuf[2] = torch.empty(ne, dtype=torch.float64).pin_memory()
            return ebuf[0].numpy(), ebuf[1].numpy(), ebuf[2].numpy()

        scores_np = scores_h.numpy().reshape((n_local, w["ndims"]), order="F")
        h2d = colptr_h.numel() * 8 + rowidx_h.numel() * 4 + x_h.numel() * 8 + HVG_NTOP * 4
        d2h_box = [0]

        def e2e_step():
            ctx.load_csc_ptrs(w["genes"], n_local, nnz, colptr_h.data_ptr(), rowidx_h.data_ptr(), x_h.data_ptr(),
                              cell_offset=c0, n_cells_total=w["cells"])
            r = pipe.run(ndims=w["ndims"], hvg_ntop=HVG_NTOP, neighbors=w["k"], snn_type=w["scheme"], want_host=True,
                         scores_out=scores_np, edge_buffers=edge_buffers, logx_out=logx_h.numpy())
            knn_h.copy_(r["knn_full"][c0:c1], non_blocking=False)
            knn_local = knn_h
            f, t, wt = r["edges"]
            d2h_box[0] = (n_local * 8 + nnz * 8 + 2 * w["genes"] * 8 + n_local * w["ndims"] * 8 +
                          HVG_NTOP * w["ndims"] * 8 + knn_local.numel() * 4 + f.size * 4 + t.size * 4 + wt.size * 8)
            return r

        for _ in range(min(args.warmup, 2)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            e2e_step()
        e1.record(stream)
        barrier()
        e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)) / args.steps
        tme = ctx.timings()
        e2e = dict(ms=e2e_ms, h2d=h2d, d2h=d2h_box[0],
                   stage_ms={k: tme[k] for k in ["h2d_ms", "d2h_ms"] + stage_keys} | {"host_marks_ms": pipe.last_marks})

    # ---- max over ranks ----
    vals = torch.tensor([dev_ms, e2e["ms"] if e2e else 0.0, float(n_edges_local), float(launches),
                         float(e2e["h2d"] if e2e else 0), float(e2e["d2h"] if e2e else 0)],
                        dtype=torch.float64, device=dev)
    if world > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms = float(mx[0]), float(mx[1])
        n_edges, launches_all = int(sm[2]), int(sm[3])
        h2d_all, d2h_all = int(sm[4]), int(sm[5])
    else:
        e2e_ms = float(vals[1])
        n_edges, launches_all = n_edges_local, int(launches)
        h2d_all, d2h_all = int(vals[4]), int(vals[5])

    if rank == 0:
        peaks = read_peaks()
        stage_ms = {k: v / args.steps for k, v in stage_sum.items()}
        # dominant kernel: the kNN distance-tile kernel (k_knn_tiles_tc): ALGORITHMIC work 2 * n_query * N * d flop
        # per launch (the full distance matrix; tiles skipped by the exact triangle-inequality bound still count as answered)
        flop = 2.0 * n_local * w["cells"] * w["ndims"]
        knn_ms = stage_ms["knn_tile_ms"]
        achieved_tf = flop / (knn_ms * 1e-3) / 1e12 if knn_ms > 0 else 0.0
        traffic = None
        try:  # DRAM bytes per launch from the committed ncu capture of this kernel, when it matches this run
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_knn_tiles_tc"]
            if tj["workload"] == args.workload and tj["n_gpus"] == world:
                traffic = tj["bytes_per_launch"]
        except Exception:
            traffic = None
        roofline = dict(bound="tensor", kernel="k_knn_tiles_tc", achieved=achieved_tf, peak=peaks["tf_sustained"],
                        unit="TFLOP/s", frac=achieved_tf / peaks["tf_sustained"], traffic=traffic,
                        peak_source=peaks["source"] + " bf16 sustained",
                        share_of_step=knn_ms / dev_ms if dev_ms > 0 else None,
                        # transparency: `achieved` counts the whole N x N distance matrix; the exact
                        # triangle-inequality pruning skips most tiles, so the flops actually EXECUTED on the
                        # tensor cores (3 fp16 products of 128 x 128 x 64 per scanned tile) are reported too
                        tiles_scanned_frac=(tiles_scanned / tiles_total) if tiles_total else None,
                        executed_tflops=(tiles_scanned * 3 * 2.0 * 128 * 128 * 64 / (knn_ms * 1e-3) / 1e12) if knn_ms > 0 else None)
        # memory-bound stages against the measured HBM copy bandwidth (algorithmic bytes)
        hbm = {}
        kk = w["k"]
        for key, nbytes in (("colsum_ms", 8.0 * nnz + 16.0 * n_local),
                            ("lognorm_ms", 20.0 * nnz + 8.0 * n_local + 16.0 * w["genes"]),
                            # kNN read + reverse-list build + host-list gathers + edge write (rank 0's share)
                            ("snn_edges_ms", 12.0 * w["cells"] * kk + 8.0 * n_local * (kk + 1) ** 2 + 16.0 * n_edges_local)):
            if stage_ms[key] > 0:
                gbs = nbytes / (stage_ms[key] * 1e-3) / 1e9
                hbm[key.replace("_ms", "")] = dict(achieved_gbs=gbs, frac=gbs / peaks["hbm"])
        line = dict(metric=METRIC, value=w["cells"] / (dev_ms * 1e-3), unit="cells/s", n_gpus=world,
                    steps=args.steps, warmup=args.warmup, ms_per_step=dev_ms, higher_is_better=True,
                    scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(workload=args.workload, **w, hvg_ntop=HVG_NTOP, nnz_local=nnz, n_edges=n_edges,
                                l2="inputs larger than L2 (no flush)", parallelism=f"cells sharded x{world}",
                                knn_fallback_queries=int(fallback_q), eig_iterations=int(eig_iters),
                                knn_tiles_scanned_rank0=int(tiles_scanned), knn_tiles_total_rank0=int(tiles_total),
                                engines=engines, size_factors="library"),
                    clocks=sampler.summary(), gpu_launches=launches_all // max(args.steps, 1) * args.steps,
                    stage_ms=stage_ms, host_marks_ms_last_step_rank0=resident_marks, hbm_stages=hbm, roofline=roofline)
        if e2e:
            line["e2e"] = dict(value=w["cells"] / (e2e_ms * 1e-3), unit="cells/s", ms_per_step=e2e_ms,
                               h2d_bytes_per_step=h2d_all, d2h_bytes_per_step=d2h_all,
                               last_step_stage_ms_rank0=e2e.get("stage_ms"))
        if world == 1 and not args.no_cpu_baseline:
            n_sample = min(w["cells"], args.cpu_sample)
            d = host_sample(w, n_sample)
            t = cpu_pipeline_once(d["n_genes"], d["colptr"], d["rowidx"], d["x"], w)
            line["cpu_baseline"] = dict(value=n_sample / t, unit="cells/s", cores=os.cpu_count(), kind="port",
                                        sample=f"{n_sample} of {w['cells']} cells x {w['genes']} genes "
                                               f"(one pass, {t:.1f} s): oracle port, numpy/scipy ARPACK PCA + C "
                                               "brute-force kNN/SNN (OpenMP)")
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
