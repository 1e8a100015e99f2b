"""Pins the oracle (and, with -m gpu, the CUDA path) to the REAL reference when a dump made by
baseline/run_reference.R on an R host has been committed under tests/golden/ref_dump_<cfg>/ (with its dataset
under tests/golden/ref_data_<cfg>/, or regenerable from the manifest's workload).  No R exists in the build
image or on the GPU box, so until someone runs that script these tests skip -- and parity stays "unpinned"
(DESIGN.md section 2)."""
import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
DUMPS = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ref_dump_*")))


def _run(dump, compute):
    import compare_reference_dump as C
    from cellula_b200 import synth
    tag = os.path.basename(dump)[len("ref_dump_"):]
    data = os.path.join(ROOT, "tests", "golden", f"ref_data_{tag}")
    d, _ = synth.read_dataset(data)
    ref = C.load_dump(dump, "B", d["n_cells"])
    rep = ref["report"]
    rows = C.compare(compute.__name__, compute(d, ref, rep["ndims"], rep["k"], rep["scheme"], None), ref, d, True)
    rows += C.compare_side_calls(dump, d, compute is C.gpu_results)    # SURVEY 8(f)-1 / 8(f)-4 dumps, when present
    bad = [r for r in rows if not r[3]]
    assert not bad, bad


@pytest.mark.skipif(not DUMPS, reason="parity unpinned: no baseline/run_reference.R dump committed (needs an R host)")
@pytest.mark.parametrize("dump", DUMPS)
def test_oracle_matches_the_reference_dump(dump):
    import compare_reference_dump as C
    _run(dump, C.oracle_results)


@pytest.mark.gpu
@pytest.mark.skipif(not DUMPS, reason="parity unpinned: no baseline/run_reference.R dump committed (needs an R host)")
@pytest.mark.parametrize("dump", DUMPS)
def test_gpu_matches_the_reference_dump(dump):
    import compare_reference_dump as C
    _run(dump, C.gpu_results)
