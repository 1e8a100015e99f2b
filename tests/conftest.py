import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def small_counts():
    from cellula_b200 import synth
    return synth.make_counts(2000, 4000, nnz_per_cell=300, seed=11)


@pytest.fixture(scope="session")
def small_oracle(small_counts):
    """Oracle results for `small_counts` (HVG subset 800, 20 PCs, k = 10)."""
    from oracle import oracle
    d = small_counts
    return oracle.pipeline(d["n_genes"], d["colptr"], d["rowidx"], d["x"], ndims=20, hvg_ntop=800,
                           neighbors=10, snn_type="rank")


@pytest.fixture(scope="session")
def gpu_ctx():
    import cellula_b200
    ctx = cellula_b200.Context(0)
    yield ctx
    ctx.close()


def canonical(frm, to, w):
    a = np.minimum(frm, to).astype(np.int64)
    b = np.maximum(frm, to).astype(np.int64)
    o = np.lexsort((b, a))
    return a[o], b[o], np.asarray(w)[o]
