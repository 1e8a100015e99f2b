#!/bin/bash
# bench lines of the BASELINE configs other than the headline (cfg3): cfg1, cfg2, cfg4 (k x scheme sweep), cfg5 on one GPU
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
TAG=${TAG:-r02b}
timeout 600 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
run() { # name, args...
  local name=$1; shift
  timeout -k 10 900 python bench.py "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
  echo "$name exit $?"; tail -1 gpurun_out/${TAG}_bench_${name}.json | cut -c1-260
}
run cfg1_1gpu --workload cfg1 --steps 5 --warmup 3
run cfg2_1gpu --workload cfg2 --steps 5 --warmup 3
for k in 10 30 100; do for s in rank jaccard; do
  run cfg4_k${k}_${s}_1gpu --workload cfg4 --k $k --scheme $s --steps 3 --warmup 3 --no-cpu-baseline
done; done
if [ "${SKIP_CFG5:-0}" != "1" ]; then run cfg5_1gpu --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e; fi
