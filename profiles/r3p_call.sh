#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_fma.sum --clock-control none --profile-from-start off -k regex:k_mmd --csv --log-file gpurun_out/r3p_mmd.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --no-cuda-graph --no-other-configs > gpurun_out/r3p.log 2>&1
grep -v "^==" gpurun_out/r3p_mmd.csv | cut -d, -f5,13,15 | tail -16
