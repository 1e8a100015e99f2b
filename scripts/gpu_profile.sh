#!/bin/bash
# Round profile capture (B200_PROFILING.md recipe): plain bench, launch list, one full capture of the top kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e ${BENCH_WORKLOAD:+--workload $BENCH_WORKLOAD}"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"; tail -1 gpurun_out/plain.log | cut -c1-400
python bench.py $ARGS > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_knn_tiles_tc -s 1 -c 1 -o gpurun_out/knn_tiles_full \
    python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"; tail -2 gpurun_out/ncu_full.log | cut -c1-300
