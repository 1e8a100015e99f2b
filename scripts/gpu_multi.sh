#!/bin/bash
# N-GPU run: multi-GPU parity tests, then the bench at N GPUs (torchrun, as the driver launches it)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
N=${NGPU:-2}; TAG=${TAG:-r02b}
timeout 600 python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
nvidia-smi -L | head -8
timeout -k 10 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_multi.py -m gpu -q --timeout 600 > gpurun_out/${TAG}_multi_tests_${N}gpu.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/${TAG}_multi_tests_${N}gpu.log
timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 \
  > gpurun_out/${TAG}_bench_cfg3_${N}gpu.json 2> gpurun_out/${TAG}_bench_cfg3_${N}gpu.err; echo "bench exit $?"; tail -1 gpurun_out/${TAG}_bench_cfg3_${N}gpu.json | cut -c1-300
if [ "$N" = "8" ]; then
  timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload cfg5 --steps 3 --warmup 3 \
    > gpurun_out/${TAG}_bench_cfg5_8gpu.json 2> gpurun_out/${TAG}_bench_cfg5_8gpu.err; echo "cfg5 exit $?"; tail -1 gpurun_out/${TAG}_bench_cfg5_8gpu.json | cut -c1-300
fi
