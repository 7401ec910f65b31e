#!/bin/bash
# Round-1 session-3 GPU check: paired bottleneck evaluations (batched aggregation, fused epilogue),
# vectorised elementwise kernels.  Full GPU suite, then bench A/B (paired vs GDA_NO_PAIR=1).
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r1f_tests.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-e2e 2>gpurun_out/r1f_bench_pair.err | tee gpurun_out/r1f_bench_pair.json
GDA_NO_PAIR=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-e2e 2>gpurun_out/r1f_bench_nopair.err | tee gpurun_out/r1f_bench_nopair.json
tail -3 gpurun_out/r1f_bench_pair.err
