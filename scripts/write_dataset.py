#!/usr/bin/env python
"""Writes the shared synthetic dataset of a BASELINE.json workload in the language-neutral on-disk form
(SURVEY.md 8d): p.i64 / i.i32 / x.f64 little-endian + manifest.json (shape, seed, sha256).  The matrix is the
one bench.py generates on the device and the parity tests generate on the host (one generator specification,
cellula_b200/csrc/cb200_synth_core.h; no GPU needed here).  baseline/run_reference.R reads it on an R host.

    python scripts/write_dataset.py --workload cfg1 --out data/cfg1
    python scripts/write_dataset.py --workload cfg4 --out data/cfg4        # emb.f64, column-major 1 000 000 x 50
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cellula_b200 import synth  # noqa: E402

WORKLOADS = {   # BASELINE.json configs (same table as bench.py)
    "cfg1": dict(cells=2700, genes=32738, nnz_per_cell=850.0),
    "cfg2": dict(cells=68579, genes=32738, nnz_per_cell=525.0),
    "cfg3": dict(cells=1306127, genes=27998, nnz_per_cell=1500.0),
    "cfg4": dict(cells=1000000, d=50),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", required=True, choices=sorted(WORKLOADS))
    ap.add_argument("--out", required=True)
    ap.add_argument("--cells", type=int, default=None, help="first CELLS cells of the workload's matrix only")
    a = ap.parse_args()
    w = WORKLOADS[a.workload]
    if a.workload == "cfg4":
        os.makedirs(a.out, exist_ok=True)
        emb = synth.make_embedding(w["cells"], w["d"], seed=synth.EMB_SEED)
        emb.T.ravel().astype("<f8").tofile(os.path.join(a.out, "emb.f64"))      # column-major
        json.dump(dict(format="cellula_b200 embedding v1", n=w["cells"], d=w["d"], seed=synth.EMB_SEED,
                       layout="column-major", sha256=synth._sha256(emb.T.ravel())),
                  open(os.path.join(a.out, "manifest.json"), "w"), indent=1)
    else:
        n = a.cells or w["cells"]
        d = synth.make_counts(n, w["genes"], nnz_per_cell=w["nnz_per_cell"], seed=synth.SEED, n_cells_total=w["cells"])
        man = synth.write_dataset(a.out, d, seed=synth.SEED, workload=a.workload, nnz_per_cell=w["nnz_per_cell"],
                                  n_cells_total=w["cells"])
        print(json.dumps({k: man[k] for k in ("n_genes", "n_cells", "nnz")}))
