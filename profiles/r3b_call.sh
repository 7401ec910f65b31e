#!/bin/bash
# ncu evidence for the state after the sparse first layer: launch list of one eager step (same command as the bench,
# --no-cuda-graph so that every kernel is a separate launch), --set full of the aggregation kernel and of the two
# tile-packed GEMMs inside that step
mkdir -p gpurun_out
rm -f gpurun_out/launches.csv gpurun_out/*_full.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --no-cuda-graph --no-other-configs > gpurun_out/launches.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_spmm_unw -s 10 -c 3 -f -o gpurun_out/spmm_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --no-cuda-graph --no-other-configs > gpurun_out/spmm_full.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_gemm_xt -c 4 -f -o gpurun_out/gemm_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --no-cuda-graph --no-other-configs > gpurun_out/gemm_full.log 2>&1
tail -2 gpurun_out/gemm_full.log; ls -la gpurun_out | head -30
