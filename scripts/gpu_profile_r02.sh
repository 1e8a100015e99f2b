#!/bin/bash
# Round-2 profile capture (B200_PROFILING.md recipe): plain bench, launch list, `--set full` captures of the
# kernels that carry the step (one launch each, warm-up step), and one source-level capture of the top kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
TAG=${TAG:-r02}
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e ${BENCH_WORKLOAD:+--workload $BENCH_WORKLOAD}"
python bench.py $ARGS > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list exit $?"; tail -1 gpurun_out/${TAG}_plain.log | cut -c1-300
KREGEX='regex:^(k_genestats_slab|k_lognorm_values|k_hvg_fill|k_gram_img|k_gram_tc|k_gram_tc_cl|k_project_slab|k_knn_tiles_tc|k_knn_rerank|k_snn_flat|k_snn_compact|k_knn_tc_assign)'
ncu --set full --clock-control none -k "$KREGEX" -c ${NCAP:-24} -o gpurun_out/${TAG}_stages_full -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "stage captures exit $?"; tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-300
# summarise on the box (the report itself is too large to bring back), then drop it
python scripts/summarize_ncu.py gpurun_out/${TAG}_stages_full.ncu-rep gpurun_out/${TAG}_stages_ncu.txt \
  "ncu --set full --clock-control none -k <stage kernels> -c ${NCAP:-24} python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-verify (cfg3, 1 B200; first launches = warm-up step)" > /dev/null
rm -f gpurun_out/${TAG}_stages_full.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:k_knn_tiles_tc -c 1 -o gpurun_out/${TAG}_knn_tiles_full -f \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${TAG}_ncu_knn.log 2>&1
echo "knn capture exit $?"; ls -la gpurun_out/*.ncu-rep
python scripts/summarize_ncu.py gpurun_out/${TAG}_knn_tiles_full.ncu-rep gpurun_out/${TAG}_knn_tiles_tc_ncu.txt \
  "ncu --set full --clock-control none --import-source on -k regex:k_knn_tiles_tc -c 1 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-verify (cfg3, 1 B200)" > /dev/null
CB200_KNN_TIMELINE=gpurun_out/${TAG}_timeline.bin python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${TAG}_tl.log 2>&1
python scripts/timeline_summary.py gpurun_out/${TAG}_timeline.bin
