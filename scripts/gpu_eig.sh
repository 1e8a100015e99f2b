#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
for f in tests/test_gpu_pca.py tests/test_gpu_snn.py tests/test_gpu_pipeline.py tests/test_gpu_baseline_shapes.py; do
  n=$(basename $f .py); timeout -k 10 900 python -m pytest $f -m gpu -q --timeout 600 > gpurun_out/$n.log 2>&1; echo "$n exit $?"; tail -4 gpurun_out/$n.log
done
for dg in 3 6 8; do
  CB200_EIG_DEGREE=$dg timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/eig_cfg2_d$dg.json 2>&1
  CB200_EIG_DEGREE=$dg timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/eig_cfg3_d$dg.json 2>&1
  python - <<PY
import json
for w in ("cfg2","cfg3"):
    try:
        j=json.loads([x for x in open(f"gpurun_out/eig_{w}_d$dg.json") if x.startswith("{")][-1])
        print("degree $dg", w, "step", round(j["ms_per_step"],2), "eig", round(j["stage_ms"]["eig_ms"],2), "iters", j["config"]["eig_iterations"], "verified", j["verified"]["ok"], j["verified"]["edge_digest"])
    except Exception as e:
        print("degree $dg", w, "failed", e)
PY
done
