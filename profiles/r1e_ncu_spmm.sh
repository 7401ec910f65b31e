#!/bin/bash
# ncu --set full of the aggregation kernel alone (micro-benchmark process), default and 8-deep variants
mkdir -p gpurun_out
for v in 4 8; do
WIDE_ONLY=1 GDA_SPMM_TASKS=$v ncu --set full --clock-control none --import-source on -k regex:k_spmm_tasks -s 20 -c 2 -f \
    -o gpurun_out/r1e_tasks_u$v python profiles/bench_spmm.py > gpurun_out/r1e_ncu_u$v.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
