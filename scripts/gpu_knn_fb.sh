#!/bin/bash
# exact-scan fallback (batched over queries): kNN tests + the cfg4 k = 100 bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout -k 10 600 python -m pytest tests/test_gpu_knn.py -m gpu -q --timeout 400 -x > gpurun_out/test_gpu_knn.log 2>&1; echo "knn tests exit $?"; tail -4 gpurun_out/test_gpu_knn.log
timeout 600 python bench.py --workload cfg4 --k 100 --scheme rank --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02d_bench_cfg4_k100_rank_1gpu.json 2> gpurun_out/cfg4k100.err
python - <<PY
import json
j=json.loads([x for x in open("gpurun_out/r02d_bench_cfg4_k100_rank_1gpu.json") if x.startswith("{")][-1])
print("k=100", round(j["ms_per_step"],1), {a:round(b,1) for a,b in j["stage_ms"].items() if b>0.5}, j["verified"]["ok"], j["verified"]["n_edges"], j["config"]["knn_fallback_queries"])
PY
