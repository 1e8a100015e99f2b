#!/bin/bash
# Round-end style run on the GPU box: the full `-m gpu` suite, smoke, default bench (both arms), then the profile capture.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
TAG=${TAG:-r02}
timeout 600 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/${TAG}_gputests.log 2>&1
echo "gpu tests exit $?"; tail -5 gpurun_out/${TAG}_gputests.log
timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${TAG}_smoke.log | cut -c1-300
timeout -k 10 900 python bench.py > gpurun_out/${TAG}_bench_cfg3_1gpu.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"; tail -1 gpurun_out/${TAG}_bench_cfg3_1gpu.json | cut -c1-400
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref exit $?"; tail -1 gpurun_out/${TAG}_bench_reference.json | cut -c1-400
if [ "${SKIP_PROFILE:-0}" != "1" ]; then TAG=$TAG bash scripts/gpu_profile_r02.sh; fi
