#!/bin/bash
# SNN tests + the cfg4 k = 100 bench line (CTA-per-cell hash kernels: 8 Ki / 16 Ki tiers and the long tier)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
TAG=${TAG:-r02e}
timeout 600 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout -k 10 600 python -m pytest tests/test_gpu_snn.py -m gpu -q --timeout 400 -x > gpurun_out/test_gpu_snn.log 2>&1; echo "snn tests exit $?"; tail -4 gpurun_out/test_gpu_snn.log
for k in ${KS:-100}; do
CB200_TRACE=1 timeout 600 python bench.py --workload cfg4 --k $k --scheme rank --steps 3 --warmup 3 --no-cpu-baseline ${E2E:---no-e2e} > gpurun_out/${TAG}_bench_cfg4_k${k}_rank_1gpu.json 2> gpurun_out/cfg4k$k.err
grep "snn:" gpurun_out/cfg4k$k.err | tail -2
python - <<PY
import json
j=json.loads([x for x in open("gpurun_out/${TAG}_bench_cfg4_k${k}_rank_1gpu.json") if x.startswith("{")][-1])
print("k=$k", round(j["ms_per_step"],1), {a:round(b,1) for a,b in j["stage_ms"].items() if b>0.5}, j["verified"]["ok"], j["verified"]["n_edges"])
PY
done
