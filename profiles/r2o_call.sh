#!/bin/bash
mkdir -p gpurun_out
export ROWS=50000
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_gemm_xt -c 1 -f -o gpurun_out/r2o_xt_fwd python profiles/bench_gemm_xt.py > gpurun_out/r2o_ncu_fwd.log 2>&1
tail -3 gpurun_out/r2o_ncu_fwd.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_gemm_xt --launch-skip 13 -c 1 -f -o gpurun_out/r2o_xt_dw python profiles/bench_gemm_xt.py > gpurun_out/r2o_ncu_dw.log 2>&1
tail -3 gpurun_out/r2o_ncu_dw.log
ls -la gpurun_out/*.ncu-rep
