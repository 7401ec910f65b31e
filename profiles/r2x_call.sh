#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_xt.py -q -p no:cacheprovider > gpurun_out/r2x_xt_tests.log 2>&1
tail -25 gpurun_out/r2x_xt_tests.log
timeout 300 python profiles/bench_gemm_xt.py > gpurun_out/r2x_xt_bench.json 2> gpurun_out/r2x_xt_bench.err
cat gpurun_out/r2x_xt_bench.json; tail -3 gpurun_out/r2x_xt_bench.err
