#!/bin/bash
# kNN tile-kernel sweep over an environment variable: VAR=CB200_KNN_SEED VALUES="64 -32 0" bash scripts/gpu_knn_sweep.sh
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
TAG=${TAG:-sweep}
if [ "${RUN_TESTS:-1}" = "1" ]; then
timeout -k 10 900 python -m pytest ${TEST_FILES:-tests/test_gpu_knn.py} -m gpu -q -x --timeout 600 --timeout-method=thread > gpurun_out/${TAG}_tests.log 2>&1
echo "tests exit $?"; tail -${TAIL:-4} gpurun_out/${TAG}_tests.log
fi
for v in ${VALUES}; do
  export ${VAR}=$v
  timeout -k 10 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e ${BENCH_EXTRA} > gpurun_out/${TAG}_${v}_bench.json 2> gpurun_out/${TAG}_${v}_bench.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_${v}_bench.json").read().strip().splitlines()[-1])
    print("${VAR}=$v", "ms_per_step", round(j["ms_per_step"], 2), "knn_tile", round(j["stage_ms"]["knn_tile_ms"], 2), "rerank", round(j["stage_ms"]["knn_rerank_ms"], 2),
          "tiles", j["config"]["knn_tiles_scanned_rank0"], "fallback", j["config"]["knn_fallback_queries"], "ok", (j.get("verified") or {}).get("ok"), (j.get("verified") or {}).get("edge_digest"))
except Exception as e:
    print("${VAR}=$v no bench line", e); print(open("gpurun_out/${TAG}_${v}_bench.err").read()[-1500:])
PY
done
if [ -n "${TL_VALUE}" ]; then
export ${VAR}=${TL_VALUE}
CB200_KNN_TIMELINE=gpurun_out/${TAG}_timeline.bin timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${TAG}_tl.log 2>&1
python scripts/timeline_summary.py gpurun_out/${TAG}_timeline.bin
python scripts/cta_log_summary.py gpurun_out/${TAG}_timeline.bin 2>/dev/null | head -4
fi
