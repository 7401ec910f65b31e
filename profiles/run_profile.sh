#!/bin/bash
# Run on the GPU box (via gpurun): full GPU test suite, smoke(), the bench line, the ncu launch list of one eager step,
# ncu --set full of the two top kernels, and a compute-sanitizer memcheck pass over the small kernel tests.
# Outputs land in gpurun_out/ (profiles/summarize.py turns them into the committed summaries).
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 4000 gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --no-cuda-graph > gpurun_out/launches.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_spmm_tasks -s 10 -c 3 -f -o gpurun_out/spmm_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --no-cuda-graph > gpurun_out/spmm_full.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_gemm_bf16x3 -c 4 -f -o gpurun_out/gemm_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e --no-cuda-graph > gpurun_out/gemm_full.log 2>&1
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pair.py tests/test_gpu_loader.py \
    tests/test_gpu_dgsda.py -q -m gpu -x -k "not forward_model_step and not at_scale and not fit" 2>&1 | tail -8 | tee gpurun_out/memcheck.log
ls -la gpurun_out
