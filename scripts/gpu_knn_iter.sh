#!/bin/bash
# kNN iteration loop on the GPU box: parity tests of the kNN stage, a short bench, the tile-kernel timeline
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
TAG=${TAG:-knn}
timeout -k 10 900 python -m pytest ${TEST_FILES:-tests/test_gpu_knn.py} -m gpu -q -x --timeout 600 --timeout-method=thread > gpurun_out/${TAG}_tests.log 2>&1
echo "tests exit $?"; tail -${TAIL:-6} gpurun_out/${TAG}_tests.log
timeout -k 10 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e ${BENCH_EXTRA} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("ms_per_step", j["ms_per_step"], {k: round(v, 2) for k, v in j["stage_ms"].items()})
    print("verified", j["verified"]); print("roofline", {k: j["roofline"][k] for k in ("executed_frac", "tiles_scanned_frac", "share_of_step")}, "fallback", j["config"]["knn_fallback_queries"])
except Exception as e:
    print("no bench line", e); print(open("gpurun_out/${TAG}_bench.err").read()[-2000:])
PY
if [ "${TIMELINE:-1}" = "1" ]; then
CB200_KNN_TIMELINE=gpurun_out/${TAG}_timeline.bin timeout 300 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${TAG}_tl.log 2>&1
python scripts/timeline_summary.py gpurun_out/${TAG}_timeline.bin
fi
if [ "${TIMELINE:-1}" = "1" ]; then python scripts/cta_log_summary.py gpurun_out/${TAG}_timeline.bin | head -4; fi
