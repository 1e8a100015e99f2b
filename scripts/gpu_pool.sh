#!/bin/bash
# pooled size factors: tests + scale runs (cfg2 and cfg3 shapes)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
timeout -k 10 600 python -m pytest tests/test_gpu_pool.py -m gpu -q --timeout 300 > gpurun_out/test_gpu_pool.log 2>&1; echo "exit $?"; tail -5 gpurun_out/test_gpu_pool.log
timeout -k 10 600 python scripts/pool_bench.py --cells 68579 --genes 32738 --nnz 525 --cpu-sample 400 > gpurun_out/pool_cfg2.json 2> gpurun_out/pool_cfg2.err; echo "cfg2 exit $?"; tail -1 gpurun_out/pool_cfg2.json | cut -c1-1500; tail -3 gpurun_out/pool_cfg2.err
timeout -k 10 900 python scripts/pool_bench.py --cells 1306127 --genes 27998 --nnz 1500 > gpurun_out/pool_cfg3.json 2> gpurun_out/pool_cfg3.err; echo "cfg3 exit $?"; tail -1 gpurun_out/pool_cfg3.json | cut -c1-1500; tail -3 gpurun_out/pool_cfg3.err
