#!/bin/bash
# Runs on the GPU box (via gpurun): environment probe, GPU parity tests file by file, smoke, bench.
# Every step is wrapped in its own timeout so a hung kernel cannot hold the box.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
{
  echo "== env"; nvidia-smi -L; nproc; free -g | head -2; which Rscript R || echo "no R"
  nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
} > gpurun_out/env.log 2>&1
cat gpurun_out/env.log
echo "== build"; timeout 600 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
for f in tests/test_gpu_norm.py tests/test_gpu_snn.py tests/test_gpu_knn.py tests/test_gpu_pca.py tests/test_gpu_diag.py tests/test_gpu_pipeline.py; do
  n=$(basename $f .py)
  echo "== $n"
  timeout -k 10 ${TEST_TIMEOUT:-700} python -m pytest $f -m gpu -q --timeout 400 --timeout-method=thread > gpurun_out/$n.log 2>&1
  echo "exit $?"; tail -${TAIL:-25} gpurun_out/$n.log
done
echo "== smoke"; timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -4 gpurun_out/smoke.log
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  echo "== bench"; timeout -k 10 900 python bench.py --steps 3 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.log 2>&1; echo "exit $?"; tail -3 gpurun_out/bench.log
fi
