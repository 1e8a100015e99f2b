#!/bin/bash
# quick GPU iteration: selected test files + bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
for f in ${TEST_FILES}; do
  n=$(basename $f .py); echo "== $n"
  timeout -k 10 900 python -m pytest $f -m gpu -q --timeout 600 --timeout-method=thread > gpurun_out/$n.log 2>&1
  echo "exit $?"; tail -${TAIL:-25} gpurun_out/$n.log
done
if [ -n "${BENCH_ARGS+x}" ]; then
  echo "== bench ${BENCH_ARGS}"; timeout -k 10 1200 python bench.py ${BENCH_ARGS} > gpurun_out/bench.log 2>&1; echo "exit $?"; tail -3 gpurun_out/bench.log
fi
