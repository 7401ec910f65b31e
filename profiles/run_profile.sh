#!/bin/bash
# Run on the GPU box (via gpurun): bench line, ncu launch list, ncu --set full of the two top kernels.
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e > gpurun_out/launches.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_spmm -s 10 -c 3 -f -o gpurun_out/spmm_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e > gpurun_out/spmm_full.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_gemm_bf16x3 -c 12 -f -o gpurun_out/gemm_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e > gpurun_out/gemm_full.log 2>&1
ls -la gpurun_out
