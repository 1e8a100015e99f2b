#!/bin/bash
# ncu --set full capture of the pool-median kernel at the cfg2 shape (one launch), summarised on the box
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
TAG=${TAG:-r02}
timeout 600 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
python scripts/pool_bench.py ${POOL_ARGS:---cells 68579 --genes 32738 --nnz 525} --reps 1 > gpurun_out/${TAG}_pool_plain.json 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_pool_plain.json; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_pool_medians -c 1 -o gpurun_out/${TAG}_pool_medians -f \
    python scripts/pool_bench.py ${POOL_ARGS:---cells 68579 --genes 32738 --nnz 525} --reps 1 > gpurun_out/${TAG}_pool_ncu.log 2>&1
echo "capture exit $?"
python scripts/summarize_ncu.py gpurun_out/${TAG}_pool_medians.ncu-rep gpurun_out/${TAG}_pool_medians_ncu.txt \
  "ncu --set full --clock-control none -k k_pool_medians -c 1 python scripts/pool_bench.py ${POOL_ARGS:---cells 68579 --genes 32738 --nnz 525} (cfg2 shape, 1 B200)" > /dev/null
ncu -i gpurun_out/${TAG}_pool_medians.ncu-rep --page source --csv > gpurun_out/${TAG}_pool_medians_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_pool_medians*
tail -40 gpurun_out/${TAG}_pool_medians_ncu.txt
